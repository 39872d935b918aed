#!/bin/bash
for cg in 1 2; do for ns in 0 1; do for nb in 2 4; do
echo "=== CG=$cg NOSTORE=$ns NBUF=$nb"
SAIS_GEMM_CG=$cg SAIS_GEMM_DEBUG_NOSTORE=$ns SAIS_GEMM_NBUF=$nb timeout 120 python tools/gemm_bench.py 256 qkv,fc1,fc1-noact,fc2,fc2-bf16out,proj 2>&1 | grep -v "^frames"
done; done; done
