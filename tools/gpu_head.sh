#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "accumulate or layernorm or temporal" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -3
timeout 120 python tools/head_bench.py
SAIS_TMP_SPLITK=0 timeout 120 python tools/head_bench.py
