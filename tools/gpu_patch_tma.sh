#!/bin/bash
O=gpurun_out/patch_tma; mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_default.log 2>&1; echo "default gemm rc=$?"; tail -1 $O/pytest_default.log
SAIS_PATCH_TMA=1 timeout 100 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "patch_embed or vit" > $O/pytest_tma.log 2>&1; echo "patch_tma rc=$?"; tail -3 $O/pytest_tma.log
