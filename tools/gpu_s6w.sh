#!/bin/bash
O=gpurun_out/s6w; mkdir -p $O
SAIS_ATTN_TIMELINE=$O/tl_attn.txt timeout 120 python tools/kernel_bench.py 256 > /dev/null 2>&1
SAIS_MLP_TIMELINE=$O/tl_mlp.txt timeout 100 python tools/mlp_bench.py 256 > $O/mlp_bench.log 2>&1; cat $O/mlp_bench.log
SAIS_MLP_TAIL_SPLIT=0 timeout 100 python tools/mlp_bench.py 256 2>&1 | tail -2
ls $O
