#!/bin/bash
# bf16 copy of the LayerNorm-producer GEMMs through a staging tile + TMA store (SAIS_GEMM_XB=1, default) vs direct stores (0)
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q 2>&1 | tail -3
for v in 1 0; do echo "== SAIS_GEMM_XB=$v"; SAIS_GEMM_XB=$v timeout 120 python tools/gemm_bench.py 256 proj+lnout,fc2+lnout,proj,fc2 2>&1 | grep -v "^frames"; done
for v in 1 0 1; do echo "== bench SAIS_GEMM_XB=$v"; SAIS_GEMM_XB=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; done
