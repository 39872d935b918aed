#!/bin/bash
O=gpurun_out/s6u; mkdir -p $O
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --csv --log-file $O/insitu.csv python tools/profile_step.py > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_insitu.py $O/insitu.csv > $O/insitu.txt 2>&1; tail -1 $O/insitu.txt
