#!/bin/bash
O=gpurun_out/s6q; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mlp_fused" > $O/pytest_mlp.log 2>&1; echo "pytest mlp rc=$?"; tail -3 $O/pytest_mlp.log
timeout 100 python tools/mlp_bench.py 256 > $O/mlp_bench.log 2>&1; cat $O/mlp_bench.log
SAIS_MLP_TIMELINE=$O/tl_mlp.txt timeout 100 python tools/mlp_bench.py 256 > /dev/null 2>&1; ls $O
