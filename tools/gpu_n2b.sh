#!/bin/bash
OUT=gpurun_out/n2h; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; cat $OUT/bench_n2.json | cut -c1-400; tail -3 $OUT/bench_n2.err
