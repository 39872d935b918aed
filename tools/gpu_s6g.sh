#!/bin/bash
O=gpurun_out/s6g; mkdir -p $O
S="proj,proj+lnout,fc2,fc2+lnout"
{
for nb in 2 3 4; do echo "== NBUF_RES=$nb"; SAIS_GEMM_NBUF_RES=$nb timeout 100 python tools/gemm_bench.py 256 $S 2>&1 | grep -v "^frames"; done
echo "== NBUF_RES=3 XB=0"; SAIS_GEMM_XB=0 SAIS_GEMM_NBUF_RES=3 timeout 100 python tools/gemm_bench.py 256 proj+lnout,fc2+lnout 2>&1 | grep -v "^frames"
} > $O/knobs.log 2>&1
cat $O/knobs.log
for nb in 2 3 4; do
SAIS_GEMM_NBUF_RES=$nb timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/pytest_nb$nb.log 2>&1; echo "pytest nbuf_res=$nb rc=$?"; tail -2 $O/pytest_nb$nb.log
done
SAIS_GEMM_NBUF_RES=3 SAIS_GEMM_TIMELINE=$O/tl_proj_nb3.txt timeout 120 python tools/gemm_bench.py 256 proj+lnout > /dev/null 2>&1
SAIS_GEMM_NBUF_RES=3 SAIS_GEMM_TIMELINE=$O/tl_fc2_nb3.txt timeout 120 python tools/gemm_bench.py 256 fc2+lnout > /dev/null 2>&1
