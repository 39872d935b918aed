#!/bin/bash
O=gpurun_out/s6e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -3 $O/pytest_gemm.log
{
echo "== default"; timeout 100 python tools/gemm_bench.py 256 qkv+lnin,fc1+lnin,fc1 2>&1 | grep -v "^frames"
echo "== EW=16 all"; SAIS_GEMM_EW=16 timeout 100 python tools/gemm_bench.py 256 qkv+lnin,qkv 2>&1 | grep -v "^frames"
echo "== NBUF=2"; SAIS_GEMM_NBUF=2 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_TIMELINE=$O/tl_fc1.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
echo "== bench"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
