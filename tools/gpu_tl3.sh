#!/bin/bash
mkdir -p gpurun_out/tl3
for sh in fc1 fc1+lnin qkv fc2 fc2+lnout proj; do
SAIS_GEMM_TIMELINE=gpurun_out/tl3/${sh}.txt timeout 120 python tools/gemm_bench.py 256 $sh 2>&1 | grep -v "^frames"
done
