#!/bin/bash
echo "BN=192 (default)"; timeout 200 python tools/gemm_bench.py 256 proj,fc2,fc2-bf16out
echo "BN=128"; SAIS_GEMM_FORCE_BN=128 timeout 200 python tools/gemm_bench.py 256 proj,fc2,fc2-bf16out
echo "frames=240 default"; timeout 200 python tools/gemm_bench.py 240 qkv,proj,fc1,fc2
