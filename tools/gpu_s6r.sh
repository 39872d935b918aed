#!/bin/bash
O=gpurun_out/s6r; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mlp" > $O/pytest_mlp.log 2>&1; echo "pytest mlp rc=$?"; tail -4 $O/pytest_mlp.log
SAIS_MLP_FOLD=1 timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/pytest_models_mlpfold.log 2>&1; echo "pytest models (mlp fold) rc=$?"; tail -3 $O/pytest_models_mlpfold.log
timeout 100 python tools/mlp_bench.py 256 > $O/mlp_bench.log 2>&1; cat $O/mlp_bench.log
for mf in 0 1 0 1; do
  SAIS_MLP_FOLD=$mf timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_mf$mf.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench_mf$mf.json')); print('mlp_fold=$mf', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks']['sm_ghz_in_loop_median'])"
done
