#!/bin/bash
O=gpurun_out/chunk; mkdir -p $O
for c in 256 128 64 256; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --chunk $c 2>/dev/null > $O/bench_c$c.json
  python -c "import sys,json; d=json.load(open('$O/bench_c$c.json')); print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_ghz_in_loop_median']); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"
done
