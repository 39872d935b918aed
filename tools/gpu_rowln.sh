#!/bin/bash
OUT=gpurun_out/rowln; mkdir -p $OUT
echo "== rowln tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "residual_layernorm or mlp" 2>&1 | tail -15
echo "== full gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench rowln"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee $OUT/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['kernel_classes']))"
echo "== bench rowln+mlpfused"; SAIS_MLP_FUSED=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(json.dumps(d['kernel_classes']))"
echo "== bench old"; SAIS_ROWLN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(json.dumps(d['kernel_classes']))"
