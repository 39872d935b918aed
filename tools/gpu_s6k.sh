#!/bin/bash
O=gpurun_out/s6k; mkdir -p $O
SAIS_GEMM_ASTAT=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_astat.log 2>&1; echo "pytest astat rc=$?"; tail -5 $O/pytest_astat.log
{
for as in 0 1; do echo "== ASTAT=$as"; SAIS_GEMM_ASTAT=$as timeout 60 python tools/gemm_bench.py 256 qkv,qkv+lnin,fc1,fc1+lnin,fc1-noact 2>&1 | grep -v "^frames"; done
echo "== ASTAT=1 EW=16 (qkv lean)"; SAIS_GEMM_EW=16 SAIS_GEMM_ASTAT=1 timeout 60 python tools/gemm_bench.py 256 qkv+lnin 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_ASTAT=1 SAIS_GEMM_TIMELINE=$O/tl_qkv_astat.txt timeout 60 python tools/gemm_bench.py 256 qkv+lnin > /dev/null 2>&1
SAIS_GEMM_ASTAT=1 SAIS_GEMM_TIMELINE=$O/tl_fc1_astat.txt timeout 60 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
