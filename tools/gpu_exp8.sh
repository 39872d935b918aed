#!/bin/bash
S="qkv,fc1-noact"
for k in 48 56 112 0 8 64; do echo "== NOSTORE=$k"; SAIS_GEMM_DEBUG_NOSTORE=$k timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
for st in 3 4 5; do echo "== stages $st"; SAIS_GEMM_STAGES=$st timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
for st in 3 4 5; do echo "== stages $st NOSTORE=48"; SAIS_GEMM_DEBUG_NOSTORE=48 SAIS_GEMM_STAGES=$st timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
