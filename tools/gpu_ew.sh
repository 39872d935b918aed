#!/bin/bash
echo "== EW=12"; SAIS_GEMM_EW=12 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or fold" 2>&1 | tail -2
SAIS_GEMM_EW=12 timeout 200 python tools/gemm_bench.py 256 qkv,fc1,fc1-noact,qkv+lnin,fc1+lnin,tmp-ff1-split
echo "== EW=8"; SAIS_GEMM_EW=8 timeout 200 python tools/gemm_bench.py 256 qkv,fc1,fc1-noact,qkv+lnin,fc1+lnin,tmp-ff1-split
