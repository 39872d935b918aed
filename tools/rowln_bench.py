"""Timing of the GEMM + residual + LayerNorm kernel vs the generic GEMM + separate LayerNorm (CUDA events).
Dev tool, GPU only.  usage: python tools/rowln_bench.py [frames]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = frames * 197


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


print(f"frames={frames} rows={M}")
for name, K in (("proj", 384), ("fc2", 1536)):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(384, K, device=dev) / K ** 0.5).bfloat16()
    bias, gamma, beta = torch.randn(384, device=dev), torch.rand(384, device=dev) + 0.5, torch.randn(384, device=dev)
    x = torch.randn(M, 384, device=dev)
    flops = 2.0 * M * 384 * K
    t_f = timeit(lambda: ops.gemm_residual_layernorm(a, w, bias, x, gamma, beta))
    t_n = timeit(lambda: ops.gemm_residual_layernorm(a, w, bias, x, want_ln=False))

    def unfused():
        ops.gemm_bias_act(a, w, bias, residual=x, out=x)
        ops.layernorm(x, gamma, beta, 1e-6)

    t_u = timeit(unfused)
    print(f"{name:5s} K={K:5d}  fused+LN {t_f*1e3:7.1f} us ({flops/t_f/1e9:6.1f} TF)   fused noLN {t_n*1e3:7.1f} us   "
          f"gemm + layernorm {t_u*1e3:7.1f} us")
