#!/bin/bash
OUT=gpurun_out/mlp; mkdir -p $OUT
SAIS_MLP_TIMELINE=$OUT/timeline.txt timeout 120 python tools/mlp_bench.py 256 2>&1 | tail -2; grep -v " -1       -1       -1       -1" $OUT/timeline.txt
