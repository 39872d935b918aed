"""Latency of small ViT batches (+ the C1 clip step) — dev tool, GPU only.  usage: python tools/c1_bench.py [B ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import pipeline, scoring  # noqa: E402
from sais_b200 import vision_transformer as vits  # noqa: E402
from sais_b200.prepare_model import fullModel  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
vit = vits.vit_small(16).to(dev).eval()
head = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT',
                 modalities='RGB-Flow').to(dev).eval()
protos = torch.randn(2, 256, device=dev)


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for B in [int(a) for a in sys.argv[1:]] or [8, 20, 32, 64, 96, 128, 192, 256]:
    fr = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
    us = timeit(lambda: vit.forward_u8(fr))
    g = pipeline.CapturedStep(lambda f: vit.forward_u8(f), fr)
    us_g = timeit(lambda: g(fr))
    print(f"ViT B={B:4d}: {us:8.1f} us eager  {us_g:8.1f} us graph  ({B / us_g * 1e6:8.0f} frames/s)")

pad = pipeline.full_mask(1, 10, dev)
fr = torch.randint(0, 256, (20, 224, 224, 3), dtype=torch.uint8, device=dev)


def c1(frames):
    e = vit.forward_u8(frames)
    o, _ = head(e[:10].view(1, 1, 10, 384), e[10:].view(1, 1, 10, 384), None, None, 'Prototypes', pad, pad, None)
    return scoring.predict(o, protos)


us = timeit(lambda: c1(fr))
g = pipeline.CapturedStep(c1, fr)
print(f"C1 clip 10+10: {us:8.1f} us eager  {timeit(lambda: g(fr)):8.1f} us graph")
