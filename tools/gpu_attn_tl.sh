#!/bin/bash
mkdir -p gpurun_out/attn
SAIS_ATTN_TIMELINE=gpurun_out/attn/timeline.txt timeout 120 python tools/kernel_bench.py 256 2>&1 | tail -4
cat gpurun_out/attn/timeline.txt
