#!/bin/bash
OUT=gpurun_out/fold; mkdir -p $OUT
echo "== fold tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "fold or producer or rowstats or gemm" 2>&1 | tail -15
echo "== full gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench fold"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee $OUT/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['kernel_classes']))"
echo "== bench nofold"; SAIS_LN_FOLD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(json.dumps(d['kernel_classes']))"
timeout 200 python tools/gemm_bench.py 256
