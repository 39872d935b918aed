#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (+reference arm), per-shape kernel timings, ncu launch list + full captures.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_ref.json
echo "== per-kernel timings"
timeout 300 python tools/gemm_bench.py 256 > $OUT/gemm_bench.log 2>&1; cat $OUT/gemm_bench.log
timeout 300 python tools/kernel_bench.py 256 > $OUT/kernel_bench.log 2>&1; cat $OUT/kernel_bench.log
timeout 300 python tools/head_bench.py > $OUT/head_bench.log 2>&1; cat $OUT/head_bench.log
timeout 300 python tools/frames_bench.py 32 > $OUT/frames_bench.log 2>&1; cat $OUT/frames_bench.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: gemm (one block of the second forward) / attention, layernorm, patchify, rowstats"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|mlp_fused' -s 68 -c 6 -f -o $OUT/prof_gemm \
    python tools/profile_step.py > $OUT/prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vit_attention_tc|layernorm384|normalize_patchify|rowstats_cast|vit_cls_attention' -s 8 -c 5 -f -o $OUT/prof_other \
    python tools/profile_step.py > $OUT/prof_other.log 2>&1; echo "ncu other rc=$?"
ls -la $OUT
