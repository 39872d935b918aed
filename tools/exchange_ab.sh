for X in nccl peer nccl peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 --steps 40 --warmup 5 --no-cpu-baseline --no-extra --exchange $X > gpurun_out/ab_$1_$X.json 2> gpurun_out/ab_$1_$X.err
  python - gpurun_out/ab_$1_$X.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("exchange"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
