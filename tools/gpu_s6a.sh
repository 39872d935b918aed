#!/bin/bash
# session 6, run a: epilogue-warp / staging-buffer knobs after the warp-uniform issue fix
O=gpurun_out/s6a; mkdir -p $O
S="qkv+lnin,fc1+lnin,fc1-noact"
{
for ew in 8 12 16; do echo "== EW=$ew"; SAIS_GEMM_EW=$ew timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
for nb in 3 4; do echo "== NBUF=$nb"; SAIS_GEMM_NBUF=$nb timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
for nb in 3; do echo "== EW=12 NBUF=$nb"; SAIS_GEMM_EW=12 SAIS_GEMM_NBUF=$nb timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
echo "== BN=128 fc1"; SAIS_GEMM_FORCE_BN=128 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
echo "== BN=128 EW=16 fc1"; SAIS_GEMM_EW=16 SAIS_GEMM_FORCE_BN=128 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
echo "== BN=192 EW=12 fc1"; SAIS_GEMM_EW=12 SAIS_GEMM_FORCE_BN=192 timeout 100 python tools/gemm_bench.py 256 fc1+lnin 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_EW=16 SAIS_GEMM_TIMELINE=$O/tl_fc1_ew16.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
SAIS_GEMM_TIMELINE=$O/tl_fc1_ew8.txt timeout 120 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
