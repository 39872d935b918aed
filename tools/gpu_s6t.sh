#!/bin/bash
O=gpurun_out/s6t; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for sn in 0 1 0 1; do
  SAIS_SNAKE=$sn timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_sn$sn.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench_sn$sn.json')); print('snake=$sn', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks']['sm_ghz_in_loop_median'])"
done
timeout 200 python tools/kernel_bench.py 256 2>&1 | tail -5
