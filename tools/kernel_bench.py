"""Timing of the non-GEMM ViT kernels (LayerNorm, attention, patchify) at the bench batch size.  Dev tool, GPU only."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
M = B * 197


def timeit(name, fn, bytes_=0, flops=0, iters=20):
    if only and name not in only:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:12s} {ms*1e3:8.1f} us  {bytes_/ms/1e6:8.1f} GB/s  {flops/ms/1e9:8.1f} TFLOP/s")


x = torch.randn(M, 384, device=dev)
g, b = torch.ones(384, device=dev), torch.zeros(384, device=dev)
timeit("layernorm", lambda: ops.layernorm(x, g, b, 1e-6), bytes_=M * 384 * 6)
timeit("rowstats_cast", lambda: ops.rowstats_cast(x), bytes_=M * (384 * 6 + 32))  # (allocates its outputs: a little pessimistic)
qkv = torch.randn(M, 1152, device=dev).bfloat16()
timeit("vit_attn", lambda: ops.vit_attention(qkv, B), bytes_=M * 1536 * 2, flops=4.0 * B * 6 * 197 * 197 * 64)
fr = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
timeit("patchify_u8", lambda: ops.normalize_patchify_u8(fr), bytes_=B * 224 * 224 * 3 * 3)
timeit("vit_cls_attn", lambda: ops.vit_cls_attention(qkv, B), bytes_=B * 197 * 768 * 2)
