#!/bin/bash
O=gpurun_out/s6v; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rowstats or folded or producer" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 200 python tools/kernel_bench.py 256 2>&1 | tail -6
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench$i.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"; done
