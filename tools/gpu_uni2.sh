#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q 2>&1 | tail -3
timeout 200 python tools/kernel_bench.py 256 vit_attn,vit_cls_attn
timeout 200 python tools/gemm_bench.py 256 2>&1 | grep -v "^frames"
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_ghz_in_loop_median'])"; done
