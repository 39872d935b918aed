"""Frame front-end timing: crop 0.8 + Pillow-exact resize of decoded 1080p / 720p frames, GB/s against the
algorithmic bytes (crop window read + 224x224x3 written).  usage: python tools/frames_bench.py [batch]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from sais_b200 import frames as F  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
peak = 6554.9
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
for (h, w) in [(1080, 1920), (720, 1280), (480, 854)]:
    x = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device=dev)
    top, left, ch, cw = F.center_crop_box(h, w)
    for _ in range(3):
        F.crop_resize(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F.crop_resize(x)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    alg = B * (ch * cw * 3 + 224 * 224 * 3)
    print(f"crop_resize {h}x{w} batch {B}: {us:8.1f} us  {us / B:6.2f} us/frame  {alg / us / 1e3:7.1f} GB/s algorithmic "
          f"({alg / us / 1e3 / peak:.1%} of {peak:.0f} GB/s)")
