#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -3
timeout 120 python tools/kernel_bench.py 256 vit_attn
SAIS_ATTN_NOTURNS=1 timeout 120 python tools/kernel_bench.py 256 vit_attn
