#!/bin/bash
mkdir -p gpurun_out/tl
for cg in 1 2; do for sh in fc1-noact qkv fc2-bf16out fc1; do
SAIS_GEMM_TIMELINE=gpurun_out/tl/${sh}_cg${cg}.txt SAIS_GEMM_CG=$cg timeout 120 python tools/gemm_bench.py 256 $sh 2>&1 | grep -v "^frames"
done; done
