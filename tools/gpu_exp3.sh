#!/bin/bash
# effective SM clock inside the GEMMs (clock64 vs globaltimer), power / clock trace, cost of the LayerNorm-producer extras
O=gpurun_out/exp3; mkdir -p $O
for sh in qkv+lnin fc1+lnin proj+lnout fc2+lnout; do
  SAIS_GEMM_TIMELINE=$O/tl_${sh}.txt timeout 120 python tools/gemm_bench.py 256 $sh 2>&1 | grep -v "^frames"
  head -2 $O/tl_${sh}.txt
done
echo "== power trace during gemm_bench"
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 20 > $O/smi_gemm.csv &
SMI=$!
timeout 200 python tools/gemm_bench.py 256 2>&1
kill $SMI
sort $O/smi_gemm.csv | uniq -c | sort -rn | head -12
echo "== knobs"
for k in 0 2 4 6; do echo "NOSTORE=$k"; SAIS_GEMM_DEBUG_NOSTORE=$k timeout 120 python tools/gemm_bench.py 256 proj+lnout,fc2+lnout 2>&1 | grep -v "^frames"; done
echo "== power trace during bench"
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 20 > $O/smi_bench.csv &
SMI=$!
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-200
kill $SMI
awk -F, '{print $1}' $O/smi_bench.csv | sort | uniq -c | sort -rn | head -8
awk -F, '{s+=$2; n++} END {print "mean power", s/n}' $O/smi_bench.csv
