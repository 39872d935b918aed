"""Timing of the fused MLP kernel vs the two-GEMM path (CUDA events).  Dev tool, GPU only.
usage: python tools/mlp_bench.py [frames]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = frames * 197
xn = torch.randn(M, 384, device=dev).bfloat16()
w1, b1 = (torch.randn(1536, 384, device=dev) * 0.05).bfloat16(), torch.randn(1536, device=dev)
w2, b2 = (torch.randn(384, 1536, device=dev) * 0.03).bfloat16(), torch.randn(384, device=dev)
x = torch.randn(M, 384, device=dev)
hid = torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def unfused():
    ops.gemm_bias_act(xn, w1, b1, act=_lib.ACT_GELU_ERF, out=hid)
    ops.gemm_bias_act(hid, w2, b2, residual=x, out=x)


flops = 4.0 * M * 384 * 1536
gamma, beta = 1 + 0.1 * torch.randn(384, device=dev), 0.1 * torch.randn(384, device=dev)
wg, c, d = ops.fold_layernorm(gamma, beta, w1.float(), b1)
xb, stats = ops.rowstats_cast(x)
xb2, stats2 = xb.clone(), stats.clone()


def folded_then_cast():
    ops.vit_mlp_ln(xb, stats, wg, c, d, w2, b2, x)
    ops.rowstats_cast(x)


ms_u = timeit(unfused)
ms_f = timeit(lambda: ops.vit_mlp(xn, w1, b1, w2, b2, x))
ms_l = timeit(lambda: ops.vit_mlp_ln(xb, stats, wg, c, d, w2, b2, x))
ms_s = timeit(folded_then_cast)
ms_c = timeit(lambda: ops.vit_mlp_ln(xb2, stats2, wg, c, d, w2, b2, x, xb_out=xb2, stats_out=stats2))  # (timeline: last call)
print(f"frames={frames} rows={M}")
print(f"mlp 2-gemm                   {ms_u*1e3:8.1f} us  {flops/ms_u/1e9:8.1f} TFLOP/s")
print(f"mlp fused                    {ms_f*1e3:8.1f} us  {flops/ms_f/1e9:8.1f} TFLOP/s")
print(f"mlp fused + LN fold          {ms_l*1e3:8.1f} us  {flops/ms_l/1e9:8.1f} TFLOP/s")
print(f"  ... + rowstats_cast kernel {ms_s*1e3:8.1f} us")
print(f"mlp fused + LN fold + cast warps {ms_c*1e3:8.1f} us  {flops/ms_c/1e9:8.1f} TFLOP/s")
