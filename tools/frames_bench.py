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

# ---- the whole front-end from JPEG bytes: batched nvJPEG decode -> crop 0.8 -> resize 224 (frames.load_frames), next to
# the reference's host path for the same streams (PIL decode + CenterCrop + Resize, one frame at a time, main_dino.py:295-313)
import io
import time

import numpy as np
from PIL import Image

rng = np.random.default_rng(0)
for (h, w) in [(1080, 1920), (480, 854)]:
    # smooth synthetic frames (random low-frequency field + noise) so the entropy-coded size is video-like, 4:2:0, quality 90
    base = rng.integers(0, 256, (h // 40 + 1, w // 40 + 1, 3), dtype=np.uint8)
    img = np.asarray(Image.fromarray(base).resize((w, h), Image.BILINEAR)).astype(np.int16)
    streams = []
    for i in range(B):
        fr = np.clip(img + rng.integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
        buf = io.BytesIO()
        Image.fromarray(fr).save(buf, format="JPEG", quality=90, subsampling="4:2:0")
        streams.append(buf.getvalue())
    kb = sum(len(s) for s in streams) / B / 1e3
    for _ in range(2):
        F.load_frames(streams, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        F.load_frames(streams, dev)
    torch.cuda.synchronize()
    us_gpu = (time.perf_counter() - t0) / 5 / B * 1e6
    top, left, ch, cw = F.center_crop_box(h, w)
    t0 = time.perf_counter()
    for s in streams[:8]:
        im = Image.open(io.BytesIO(s)).convert("RGB").crop((left, top, left + cw, top + ch)).resize((224, 224), Image.BILINEAR)
        np.asarray(im)
    us_cpu = (time.perf_counter() - t0) / 8 * 1e6
    from sais_b200 import _lib
    print(f"  [nvJPEG decoder: {['none', 'hardware engines', 'batched API', 'threaded (T host threads)'][_lib.lib().sais_jpeg_last_backend()]}]")
    print(f"jpeg front-end {h}x{w} batch {B} ({kb:.0f} KB/frame): nvJPEG decode + crop + resize {us_gpu:8.1f} us/frame "
          f"({1e6 / us_gpu:7.0f} frames/s, wall clock incl. host Huffman stage) | PIL on one host core {us_cpu:8.1f} us/frame")
