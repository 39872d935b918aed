#!/bin/bash
O=gpurun_out/s6p; mkdir -p $O
{
for ns in 0 200 400 800; do echo "== PHASE_NS=$ns"; SAIS_GEMM_PHASE_NS=$ns timeout 60 python tools/gemm_bench.py 256 fc1+lnin,fc1 2>&1 | grep -v "^frames"; done
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_PHASE_NS=300 SAIS_GEMM_TIMELINE=$O/tl_fc1_p300.txt timeout 60 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
SAIS_GEMM_TIMELINE=$O/tl_fc1_p0.txt timeout 60 python tools/gemm_bench.py 256 fc1+lnin > /dev/null 2>&1
