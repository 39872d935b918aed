#!/bin/bash
O=gpurun_out/s6s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 100 python tools/mlp_bench.py 256 > $O/mlp_bench.log 2>&1; cat $O/mlp_bench.log
for mf in 0 1 0 1; do
  SAIS_MLP_FOLD=$mf timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_mf$mf.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench_mf$mf.json')); print('mlp_fold=$mf', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks']['sm_ghz_in_loop_median'])"
done
SAIS_MLP_TAIL_SPLIT=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_split.json 2> $O/bench.err
python -c "import json; d=json.load(open('$O/bench_split.json')); print('mlp_fold=1 tail split', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
