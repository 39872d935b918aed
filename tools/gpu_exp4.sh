#!/bin/bash
# (1) operand-ingress experiment: A loaded only for the first n-tile of each m-tile (results wrong, timing valid)
# (2) 12 epilogue warps
for k in 0 8 9; do echo "== NOSTORE=$k"; SAIS_GEMM_DEBUG_NOSTORE=$k timeout 120 python tools/gemm_bench.py 256 qkv+lnin,fc1+lnin,fc1-noact,qkv 2>&1 | grep -v "^frames"; done
for e in 8 12 16; do echo "== EW=$e"; SAIS_GEMM_EW=$e timeout 120 python tools/gemm_bench.py 256 qkv+lnin,fc1+lnin,fc1-noact 2>&1 | grep -v "^frames"; done
