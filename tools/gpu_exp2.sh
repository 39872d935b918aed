#!/bin/bash
mkdir -p gpurun_out/tl2
for cg in 2; do for ns in 0 1; do
echo "=== CG=$cg NOSTORE=$ns"
SAIS_GEMM_CG=$cg SAIS_GEMM_DEBUG_NOSTORE=$ns timeout 120 python tools/gemm_bench.py 256 qkv,fc1,fc1-noact,fc2,fc2-bf16out,proj,tmp-ff1-split 2>&1 | grep -v "^frames"
done; done
for sh in fc1-noact qkv; do
SAIS_GEMM_TIMELINE=gpurun_out/tl2/${sh}_cg2.txt SAIS_GEMM_CG=2 timeout 120 python tools/gemm_bench.py 256 $sh 2>&1 | grep -v "^frames"
done
SAIS_GEMM_CG=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" 2>&1 | tail -3
