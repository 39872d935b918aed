#!/bin/bash
O=gpurun_out/s6n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 200 python tools/gemm_bench.py 256 > $O/gemm_bench.log 2>&1; cat $O/gemm_bench.log
SAIS_GEMM_TIMELINE=$O/tl_qkv.txt timeout 60 python tools/gemm_bench.py 256 qkv+lnin > /dev/null 2>&1
SAIS_GEMM_TIMELINE=$O/tl_proj.txt timeout 60 python tools/gemm_bench.py 256 proj+lnout > /dev/null 2>&1
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench$i.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"; done
