#!/bin/bash
O=gpurun_out/s6d; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gelu_mix tools/gelu_mix_bench.cu && timeout 60 /tmp/gelu_mix > $O/gelu_mix.log 2>&1; cat $O/gelu_mix.log
