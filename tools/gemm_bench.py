"""Per-shape timing of the tcgen05 GEMM (CUDA events) for the ViT / temporal shapes.  Dev tool, GPU only.
usage: python tools/gemm_bench.py [frames] """
import sys
import os
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 96
M = frames * 197
SHAPES = [  # name, M, N, K, act, residual, f32 out, split3
    ("patch", frames * 196, 384, 768, 0, False, True, False),
    ("qkv", M, 1152, 384, 0, False, False, False),
    ("proj", M, 384, 384, 0, True, True, False),
    ("fc1", M, 1536, 384, 1, False, False, False),
    ("fc1-noact", M, 1536, 384, 0, False, False, False),
    ("fc2", M, 384, 1536, 0, True, True, False),
    ("fc2-bf16out", M, 384, 1536, 0, False, False, False),
    ("tmp-ff1-split", 16384, 2048, 384, 2, False, False, True),
    # LayerNorm-folded variants (name suffix selects the mode)
    ("qkv+lnin", M, 1152, 384, 0, False, False, False),
    ("fc1+lnin", M, 1536, 384, 1, False, False, False),
    ("proj+lnout", M, 384, 384, 0, True, True, False),
    ("fc2+lnout", M, 384, 1536, 0, True, True, False),
]
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
if only:
    SHAPES = [s for s in SHAPES if s[0] in only]
print(f"frames={frames} M={M} env NOSTORE={os.environ.get('SAIS_GEMM_DEBUG_NOSTORE')} FORCE_BN={os.environ.get('SAIS_GEMM_FORCE_BN')}")
for name, m, n, k, act, res, f32, split in SHAPES:
    kk = 2 * k if split else k
    a = torch.randn(m, kk, device=dev).bfloat16()
    w = (torch.randn(n, kk, device=dev) / k ** 0.5).bfloat16()
    bias = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev) if res else None
    out = torch.empty(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    kw = {}
    if name.endswith("+lnin"):
        kw = dict(ln_stats_in=torch.rand(m, 8, device=dev) * 384, ln_colsum=torch.randn(n, device=dev))
    elif name.endswith("+lnout"):
        kw = dict(ln_stats_out=torch.empty(m, 8, device=dev), out2=torch.empty(m, n, device=dev, dtype=torch.bfloat16))
    for _ in range(3):
        ops.gemm_bias_act(a, w, bias, act=act, residual=r, out=out, split3=split, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        ops.gemm_bias_act(a, w, bias, act=act, residual=r, out=out, split3=split, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * m * n * k * (3 if split else 1)
    # effective SM clock while this shape runs back to back (probe kernel between launches)
    probe = torch.zeros((iters, 4), dtype=torch.int64, device=dev)
    for i in range(iters):
        ops.gemm_bias_act(a, w, bias, act=act, residual=r, out=out, split3=split, **kw)
        _lib.check(_lib.lib().sais_clock_probe(probe[i].data_ptr(), 4000, torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    pr = probe.cpu().double()
    ghz = sorted(((pr[:, 3] - pr[:, 1]) / (pr[:, 2] - pr[:, 0]).clamp(min=1.0)).tolist())
    print(f"{name:14s} M={m:6d} N={n:5d} K={k:5d}  {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s   SM clock {ghz[len(ghz)//2]:.2f} GHz")
