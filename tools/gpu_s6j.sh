#!/bin/bash
O=gpurun_out/s6j; mkdir -p $O
for sn in 0 1; do
SAIS_SNAKE=$sn timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --csv --log-file $O/insitu_snake$sn.csv python tools/profile_step.py > $O/ncu_$sn.log 2>&1; echo "ncu snake=$sn rc=$?"
python tools/ncu_insitu.py $O/insitu_snake$sn.csv > $O/insitu_snake$sn.txt 2>&1; tail -1 $O/insitu_snake$sn.txt
done
sed -n 1,30p $O/insitu_snake1.txt
