#!/bin/bash
OUT=gpurun_out/mlp; mkdir -p $OUT
echo "== mlp tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "mlp" 2>&1 | tail -15
echo "== mlp bench"; timeout 120 python tools/mlp_bench.py 256 2>&1 | tee $OUT/mlp_bench.log
SAIS_MLP_TAIL_SPLIT=1 timeout 120 python tools/mlp_bench.py 256 2>&1 | tail -2
SAIS_MLP_TIMELINE=$OUT/timeline.txt timeout 120 python tools/mlp_bench.py 256 > /dev/null 2>&1; head -80 $OUT/timeline.txt
echo "== full gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee $OUT/bench.json | cut -c1-400
