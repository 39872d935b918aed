#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or fold or producer" 2>&1 | tail -2
timeout 200 python tools/gemm_bench.py 256
