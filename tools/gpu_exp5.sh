#!/bin/bash
# epilogue interference decomposition: 16 = no TMEM reads, 32 = no staging/TMA store, 48 = neither (pure MMA + operand pipeline)
for k in 0 16 32 48; do echo "== NOSTORE=$k"; SAIS_GEMM_DEBUG_NOSTORE=$k timeout 120 python tools/gemm_bench.py 256 qkv,fc1-noact,fc1,fc2-bf16out 2>&1 | grep -v "^frames"; done
