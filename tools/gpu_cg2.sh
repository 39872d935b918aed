#!/bin/bash
OUT=gpurun_out/cg2; mkdir -p $OUT
echo "== gemm tests, CG=2 forced"; SAIS_GEMM_CG=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" 2>&1 | tail -15
echo "== gemm bench default"; timeout 200 python tools/gemm_bench.py 256 2>&1 | tee $OUT/gemm_bench_cg2.log
echo "== gemm bench CG=1"; SAIS_GEMM_CG=1 timeout 200 python tools/gemm_bench.py 256 2>&1 | tee $OUT/gemm_bench_cg1.log
echo "== full gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee $OUT/bench.json | cut -c1-600
