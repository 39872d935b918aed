#!/bin/bash
O=gpurun_out/s6m; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -3 $O/pytest_gemm.log
timeout 200 python tools/gemm_bench.py 256 > $O/gemm_bench.log 2>&1; cat $O/gemm_bench.log
SAIS_GEMM_TIMELINE=$O/tl_qkv.txt timeout 60 python tools/gemm_bench.py 256 qkv+lnin > /dev/null 2>&1
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench$i.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"; done
