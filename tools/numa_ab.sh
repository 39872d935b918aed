for NB in 0 1 0 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $1 --steps 40 --warmup 5 --no-cpu-baseline --no-extra --numa-bind $NB > gpurun_out/numa_$1_$NB.json 2> gpurun_out/numa_$1_$NB.err
  python - gpurun_out/numa_$1_$NB.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("numa_bound_cpus"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
