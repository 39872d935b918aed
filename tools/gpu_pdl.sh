#!/bin/bash
# A/B of programmatic dependent launch (SAIS_PDL=1 default vs 0) + the GPU parity suite under PDL.
OUT=gpurun_out/pdl; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for v in 1 0 1 0; do
  echo "== bench SAIS_PDL=$v"
  SAIS_PDL=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> $OUT/bench_$v.err | tee $OUT/bench_pdl$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
done
for v in 1 0; do echo "== head SAIS_PDL=$v"; SAIS_PDL=$v timeout 200 python tools/head_bench.py 2>&1 | tail -4; done
