#!/bin/bash
O=gpurun_out/s6f; mkdir -p $O
{
for st in 0 1 0 1; do echo "== STAGGER=$st"; SAIS_GEMM_STAGGER=$st timeout 100 python tools/gemm_bench.py 256 fc1+lnin,fc1 2>&1 | grep -v "^frames"; done
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_STAGGER=1 SAIS_GEMM_TIMELINE=$O/tl_fc1_stagger.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
for sh in fc2+lnout proj+lnout fc2 qkv+lnin; do SAIS_GEMM_TIMELINE=$O/tl_${sh}.txt timeout 120 python tools/gemm_bench.py 256 $sh > /dev/null 2>&1; done
SAIS_GEMM_STAGGER=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm_bias_act or folded_consumer" > $O/pytest_stagger.log 2>&1; echo "pytest stagger rc=$?"; tail -2 $O/pytest_stagger.log
