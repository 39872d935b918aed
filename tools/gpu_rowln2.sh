#!/bin/bash
OUT=gpurun_out/rowln; mkdir -p $OUT
timeout 120 python tools/rowln_bench.py 256 2>&1 | tee $OUT/rowln_bench.log
SAIS_ROWLN_TIMELINE=$OUT/timeline.txt timeout 120 python tools/rowln_bench.py 256 > /dev/null 2>&1; cat $OUT/timeline.txt
