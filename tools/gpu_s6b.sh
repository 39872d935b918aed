#!/bin/bash
# session 6, run b: lean 16-warp bf16 epilogue (SAIS_GEMM_EW=16) — parity, then timing
O=gpurun_out/s6b; mkdir -p $O
SAIS_GEMM_EW=16 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_ew16.log 2>&1; echo "pytest ew16 rc=$?"; tail -5 $O/pytest_ew16.log
S="qkv,fc1,qkv+lnin,fc1+lnin,fc1-noact"
{
for ew in 8 16; do echo "== EW=$ew"; SAIS_GEMM_EW=$ew timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
for nb in 2 3; do echo "== EW=16 NBUF=$nb"; SAIS_GEMM_EW=16 SAIS_GEMM_NBUF=$nb timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
echo "== EW=16 BN=128"; SAIS_GEMM_EW=16 SAIS_GEMM_FORCE_BN=128 timeout 100 python tools/gemm_bench.py 256 qkv+lnin,fc1+lnin 2>&1 | grep -v "^frames"
echo "== EW=16 BN=192 fc1"; SAIS_GEMM_EW=16 SAIS_GEMM_FORCE_BN=192 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_EW=16 SAIS_GEMM_TIMELINE=$O/tl_fc1_ew16.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
SAIS_GEMM_EW=16 SAIS_GEMM_TIMELINE=$O/tl_qkv_ew16.txt timeout 120 python tools/gemm_bench.py 256 qkv+lnin > /dev/null 2>&1
echo "== bench EW=16"; SAIS_GEMM_EW=16 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_ew16.json 2> $O/bench.err; cut -c1-250 $O/bench_ew16.json
