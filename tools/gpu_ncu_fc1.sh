#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 1 -f -o gpurun_out/ncu/fc1 python tools/gemm_bench.py 256 fc1 > gpurun_out/ncu/fc1.log 2>&1; tail -3 gpurun_out/ncu/fc1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 1 -f -o gpurun_out/ncu/fc2ln python tools/gemm_bench.py 256 fc2+lnout > gpurun_out/ncu/fc2ln.log 2>&1; tail -3 gpurun_out/ncu/fc2ln.log
ls -la gpurun_out/ncu
