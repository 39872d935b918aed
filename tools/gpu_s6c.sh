#!/bin/bash
# session 6, run c: lean 16-warp GELU epilogue as the fc1 default — parity (all GEMM tests + variant bit-identity), timing, bench
O=gpurun_out/s6c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -5 $O/pytest_gemm.log
S="qkv+lnin,fc1+lnin,fc1"
{
echo "== default"; timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
for nb in 2; do echo "== NBUF=$nb"; SAIS_GEMM_NBUF=$nb timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"; done
for st in 3 4; do echo "== STAGES=$st"; SAIS_GEMM_STAGES=$st timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"; done
echo "== NOSTORE=1 (no global store)"; SAIS_GEMM_DEBUG_NOSTORE=1 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_TIMELINE=$O/tl_fc1.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/alu_rate tools/alu_rate_bench.cu && timeout 60 /tmp/alu_rate > $O/alu_rate.log 2>&1; cat $O/alu_rate.log
echo "== bench"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
