#!/bin/bash
O=gpurun_out/s6x; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mlp" > $O/pytest_mlp.log 2>&1; echo "pytest mlp rc=$?"; tail -2 $O/pytest_mlp.log
timeout 100 python tools/mlp_bench.py 256 > $O/mlp_bench.log 2>&1; cat $O/mlp_bench.log
SAIS_MLP_TIMELINE=$O/tl_mlp.txt timeout 100 python tools/mlp_bench.py 256 > /dev/null 2>&1; grep -n "unit(" -A 3 $O/tl_mlp.txt
