#!/bin/bash
O=gpurun_out/s6l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for as in 0 1 0 1; do
  SAIS_GEMM_ASTAT=$as timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_astat$as.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench_astat$as.json')); print('astat=$as', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
