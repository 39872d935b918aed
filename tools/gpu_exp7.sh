#!/bin/bash
S="qkv,fc1-noact"
echo "== base (no epilogue work)"; SAIS_GEMM_DEBUG_NOSTORE=48 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
echo "== no operand loads"; SAIS_GEMM_DEBUG_NOSTORE=112 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
for st in 3 4 5; do echo "== stages $st"; SAIS_GEMM_STAGES=$st SAIS_GEMM_DEBUG_NOSTORE=48 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
echo "== CG=1"; SAIS_GEMM_CG=1 SAIS_GEMM_DEBUG_NOSTORE=48 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
echo "== CG=1 no operand loads"; SAIS_GEMM_CG=1 SAIS_GEMM_DEBUG_NOSTORE=112 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
echo "== BN=128"; SAIS_GEMM_FORCE_BN=128 SAIS_GEMM_DEBUG_NOSTORE=48 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
