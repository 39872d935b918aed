"""Writes tests/golden/frames_golden.npz from the REAL libraries the reference calls (torchvision CenterCrop + Resize on
PIL images) — run in the build container; the GPU box only reads the fixture.

    python -m oracle.make_golden_frames

Inputs are regenerated from seeds by the tests (``make_frame``); per geometry the fixture stores the crop box PIL
produced, the SHA-256 of the full 224x224x3 result and its first 16 rows.
"""
import hashlib
import os

import numpy as np

GEOMETRIES = [(1080, 1920, 0.8, 0.8), (720, 1280, 0.8, 0.7), (480, 854, 0.8, 0.8), (281, 501, 0.8, 0.8),
              (100, 130, 0.8, 0.8), (280, 280, 0.8, 0.8), (224, 224, 0.8, 0.8), (1001, 777, 0.8, 0.8)]


def make_frame(h, w, seed):
    """Seeded test frame: smooth gradients + noise + hard edges (exercises rounding in both passes)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1).astype(np.int64)
    noise = rng.integers(-40, 41, (h, w, 3))
    img = np.clip(base + noise, 0, 255)
    img[h // 3: h // 3 + 5, :, :] = 255
    img[:, w // 2: w // 2 + 3, :] = 0
    return img.astype(np.uint8)


def main():
    from PIL import Image
    import torchvision.transforms as T

    out = {"geometries": np.array(GEOMETRIES, np.float64)}
    for i, (h, w, hf, wf) in enumerate(GEOMETRIES):
        h, w = int(h), int(w)
        img = make_frame(h, w, 100 + i)
        pil = Image.fromarray(img)
        cc = T.CenterCrop((hf * h, wf * w))(pil)  # main_dino.py:301
        crop = np.asarray(cc)
        # locate the crop inside the frame (the fixture pins PIL's box, not ours)
        ch, cw = crop.shape[:2]
        top = int(round((h - hf * h) / 2.0))
        left = int(round((w - wf * w) / 2.0))
        assert np.array_equal(crop, img[top:top + ch, left:left + cw])
        res = np.asarray(T.Resize((224, 224))(cc))  # extract_representations.py:158-159
        out[f"box_{i}"] = np.array([top, left, ch, cw], np.int32)
        out[f"sha_{i}"] = np.frombuffer(hashlib.sha256(res.tobytes()).digest(), np.uint8)
        out[f"head_{i}"] = res[:16].copy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frames_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
