#!/bin/bash
O=gpurun_out/tl4; mkdir -p $O
for sh in qkv fc1-noact fc1 qkv+lnin; do
  SAIS_GEMM_TIMELINE=$O/tl_${sh}.txt timeout 120 python tools/gemm_bench.py 256 $sh > /dev/null 2>&1
  SAIS_GEMM_DEBUG_NOSTORE=48 SAIS_GEMM_TIMELINE=$O/tl_${sh}_k48.txt timeout 120 python tools/gemm_bench.py 256 $sh > /dev/null 2>&1
done
ls $O
