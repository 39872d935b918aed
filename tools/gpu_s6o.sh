#!/bin/bash
O=gpurun_out/s6o; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for h in 0 1 0 1 0 1; do
  SAIS_L2_HINTS=$h timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_h$h.json 2> $O/bench.err
  python -c "import json; d=json.load(open('$O/bench_h$h.json')); print('hints=$h', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
for h in 0 1; do
SAIS_L2_HINTS=$h timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --csv --log-file $O/insitu_h$h.csv python tools/profile_step.py > $O/ncu_$h.log 2>&1
python tools/ncu_insitu.py $O/insitu_h$h.csv > $O/insitu_h$h.txt 2>&1; tail -1 $O/insitu_h$h.txt
done
