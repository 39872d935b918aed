#!/bin/bash
O=gpurun_out/s6h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -3 $O/pytest_gemm.log
S="proj,proj+lnout,fc2,fc2+lnout"
{
echo "== default"; timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
echo "== NBUF_RES=2"; SAIS_GEMM_NBUF_RES=2 timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"
echo "== NBUF_RES=2 XB=2"; SAIS_GEMM_XB=2 SAIS_GEMM_NBUF_RES=2 timeout 100 python tools/gemm_bench.py 256 proj+lnout,fc2+lnout 2>&1 | grep -v "^frames"
echo "== NBUF_RES=3 XB=1"; SAIS_GEMM_XB=1 SAIS_GEMM_NBUF_RES=3 timeout 100 python tools/gemm_bench.py 256 proj+lnout 2>&1 | grep -v "^frames"
echo "== NBUF_RES=3 XB=2"; SAIS_GEMM_XB=2 SAIS_GEMM_NBUF_RES=3 timeout 100 python tools/gemm_bench.py 256 proj+lnout 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
SAIS_GEMM_XB=2 SAIS_GEMM_NBUF_RES=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "layernorm_producer" > $O/pytest_xb2.log 2>&1; echo "pytest xb2 rc=$?"; tail -2 $O/pytest_xb2.log
SAIS_GEMM_TIMELINE=$O/tl_proj.txt timeout 120 python tools/gemm_bench.py 256 proj+lnout > /dev/null 2>&1
echo "== bench"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
