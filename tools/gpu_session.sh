#!/bin/bash
# One gpurun call = one session: bash tools/gpu_session.sh <tag> <stage> [<stage> ...]   (from the repo root, on the GPU box)
# Every stage runs under its own `timeout` and logs to gpurun_out/<tag>/.  Stages:
#   quick:<pytest -k expr>   a subset of the GPU tests first (a hang in a new kernel is caught in minutes, not at the end)
#   test                     python -m pytest tests -m gpu -x -q
#   smoke                    __graft_entry__.smoke()
#   bench[:extra args]       python bench.py --steps 30 --warmup 5 [extra args]       -> bench.json
#   ref                      python bench.py --impl reference --steps 5 --warmup 1      -> bench_ref.json
#   ab:<ENV>=<v>[,<ENV>=<v>] short bench with the environment knobs set                 -> bench_<ENV>=<v>.json
#   abarg:<bench args>       short bench with extra command-line arguments              -> bench_<args>.json
#   kernels                  tools/gemm_bench.py, mlp_bench.py, kernel_bench.py, head_bench.py, frames_bench.py
#   timeline                 SAIS_MLP_TIMELINE / SAIS_ATTN_TIMELINE dumps of CTA 0 (tools/mlp_bench.py, kernel_bench.py)
#   ncu_list                 launch list of two bench steps (gpu__time_duration)        -> launches.csv
#   ncu_full:<regex>[:skip]  ncu --set full of three launches of the matching kernels (after `skip` of them, default 12)
#                            of tools/profile_step.py (two ViT forwards at batch 256)   -> prof_<n>.ncu-rep
#   py:<script> [args]       any other tool script under tools/
#   sh:<command>             any shell command (logged to sh_<n>.log)
TAG=${1:-r02}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
NFULL=0
NSH=0
for STAGE in "$@"; do
  KIND=${STAGE%%:*}; ARG=""; [[ "$STAGE" == *:* ]] && ARG=${STAGE#*:}
  echo "== $STAGE"
  case $KIND in
    quick)  timeout 420 python -m pytest tests -m gpu -x -q -k "$ARG" > $OUT/pytest_quick.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_quick.log ;;
    test)   timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "deviation|passed|failed|error" $OUT/pytest_gpu.log | tail -12 ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/smoke.log ;;
    bench)  timeout 900 python bench.py --steps 30 --warmup 5 $ARG > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-400 $OUT/bench.json; tail -3 $OUT/bench.err ;;
    ref)    timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_ref.json ;;
    ab)     ( IFS=','; for kv in $ARG; do export "$kv"; done
              timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > "$OUT/bench_$ARG.json" 2>> $OUT/bench.err
              python - "$OUT/bench_$ARG.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
      {k: round(v["ms_per_step"] * 1e3) for k, v in d["kernel_classes"].items()})
PY
            ) ;;
    abarg)  NAME=$(echo "$ARG" | tr ' =' '__')
            timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra $ARG > "$OUT/bench_$NAME.json" 2>> $OUT/bench.err
            python - "$OUT/bench_$NAME.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
            ;;
    kernels) for t in gemm_bench.py mlp_bench.py kernel_bench.py head_bench.py "frames_bench.py 32"; do
               timeout 300 python tools/$t > $OUT/${t%%.*}.log 2>&1; cat $OUT/${t%%.*}.log; done ;;
    timeline) SAIS_MLP_TIMELINE=$OUT/timeline_mlp.txt timeout 300 python tools/mlp_bench.py 256 > /dev/null 2>&1
              SAIS_ATTN_TIMELINE=$OUT/timeline_attention.txt timeout 300 python tools/kernel_bench.py 256 > /dev/null 2>&1
              ls -la $OUT/timeline_* ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
                python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --no-extra > $OUT/bench_under_ncu.log 2>&1; echo "rc=$?" ;;
    ncu_full) NFULL=$((NFULL+1)); RX=${ARG%%:*}; SKIP=12; [[ "$ARG" == *:* ]] && SKIP=${ARG#*:}
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c 3 -f -o $OUT/prof_$NFULL \
                python tools/profile_step.py > $OUT/prof_$NFULL.log 2>&1; echo "rc=$?" ;;
    py)     timeout 600 python tools/$ARG > "$OUT/$(echo $ARG | tr ' /' '__').log" 2>&1; echo "rc=$?"; tail -30 "$OUT/$(echo $ARG | tr ' /' '__').log" ;;
    sh)     NSH=$((NSH+1)); timeout 600 bash -c "$ARG" > $OUT/sh_$NSH.log 2>&1; echo "rc=$?"; tail -12 $OUT/sh_$NSH.log ;;
    *)      echo "unknown stage $STAGE" ;;
  esac
done
ls $OUT
