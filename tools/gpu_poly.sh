#!/bin/bash
# A/B of the FMA-pipe exp2 share in the ViT attention softmax
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -2
for v in 0 2 3 4; do
  echo "== SAIS_ATTN_POLY=$v"
  SAIS_ATTN_POLY=$v timeout 100 python tools/kernel_bench.py 256 vit_attn
  SAIS_ATTN_POLY=$v timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -k "attention or vit" 2>&1 | tail -1
done
