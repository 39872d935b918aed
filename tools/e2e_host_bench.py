"""extract_features from pinned vs pageable host frames (the call a user makes) — dev tool, GPU only."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sais_b200 import pipeline  # noqa: E402
from sais_b200 import vision_transformer as vits  # noqa: E402

dev = torch.device("cuda:0")
vit = vits.vit_small(16).to(dev).eval()
n = 1024
g = torch.Generator().manual_seed(0)
pageable = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, generator=g)
pinned = pageable.pin_memory()
resident = pageable.to(dev)
for name, fr in (("device-resident", resident), ("pinned host", pinned), ("pageable host", pageable), ("numpy (pageable)", pageable.numpy())):
    for _ in range(2):
        pipeline.extract_features(vit, fr, 256, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = pipeline.extract_features(vit, fr, 256, device=dev)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"extract_features {n} frames from {name:18s}: {dt * 1e3:8.1f} ms  {n / dt:9.0f} frames/s")
