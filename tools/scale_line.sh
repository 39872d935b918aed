N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
python - gpurun_out/scale_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("exchange"))
    for k, v in d.get("extra_configs", {}).items():
        print("  ", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
tail -3 gpurun_out/scale_n$N.err
