"""Timing of the SAIS temporal head + scoring alone (bench.py's per-step head workload and the C3 shape).
Dev tool, GPU only.  usage: python tools/head_bench.py"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import _lib, pipeline, scoring  # noqa: E402
from sais_b200.prepare_model import fullModel  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
torch.manual_seed(0)
head = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT',
                 modalities='RGB-Flow').to(dev).eval()
protos = torch.randn(2, 256, device=dev)
for clips, T in ((8, 16), (512, 30)):
    x = torch.randn(clips, 1, T, 384, device=dev)
    f = torch.randn(clips, 1, T, 384, device=dev)
    pad = pipeline.full_mask(clips, T, dev)

    def step():
        out, attn = head(x, f, None, None, 'Prototypes', pad, pad, None)
        return scoring.predict(out, protos)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ncls = 10
    ms_c, work_c, n_c = (C.c_double * ncls)(), (C.c_double * ncls)(), (C.c_int64 * ncls)()
    lib.sais_profile_begin()
    for _ in range(5):
        step()
    lib.sais_profile_end(ms_c, work_c, n_c, ncls)
    names = ["gemm", "vit_attn", "layernorm", "patchify", "temporal_attn", "misc", "gemm_split3", "mlp_fused", "gemm_qkv", "gemm_proj"]
    per = ", ".join(f"{n} {ms_c[i]/5*1e3:.0f}us/{n_c[i]//5}" for i, n in enumerate(names) if n_c[i])
    print(f"head {clips} clips x {T} frames (RGB+flow): {ms*1e3:8.1f} us/step  {clips/ms*1e3:9.0f} clips/s | {per}")
