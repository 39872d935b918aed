"""TEST INFRASTRUCTURE ONLY — CPU restatement of the frame front-end that precedes the ViT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product path
(``sais_b200``) never does.

What the reference does to every decoded frame before ``ToTensor`` / ``Normalize``:

* ``transforms.CenterCrop((height_frac*height, width_frac*width))`` with ``height_frac, width_frac = 0.8, 0.8``
  (0.8, 0.7 for the two ``*_Gronau`` datasets) — ``SAIS/scripts/dino-main/main_dino.py:298-301, 317-322``.  The crop
  size is a pair of *floats*; torchvision's ``center_crop`` turns them into a PIL box through two rounds of Python
  ``round`` (banker's rounding): ``top = int(round((H - ch) / 2.))``, then ``Image.crop`` rounds
  ``(left, top, left + cw, top + ch)`` again (torchvision ``functional.center_crop`` / ``functional_pil.crop``,
  Pillow ``Image.crop``; third-party, pinned ``torchvision==0.9.0`` / ``Pillow==9.1.1`` in the reference's
  requirements, not vendored).  :func:`center_crop_box` restates that arithmetic.
* ``transforms.Resize((224, 224))`` — ``SAIS/scripts/extract_representations.py:158-162`` — which for a PIL image is
  ``Image.resize((224, 224), BILINEAR)``: Pillow's two-pass separable *antialiased* triangle filter in 8-bit fixed
  point (``libImaging/Resample.c``: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
  ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``; ``PRECISION_BITS = 32 - 8 - 2``).  The horizontal pass runs
  first and rounds to uint8, the vertical pass reads that uint8 intermediate.  :func:`resample_coeffs` and
  :func:`resize_bilinear_u8` restate the published algorithm; integer arithmetic, so parity is BIT-EXACT.

Pinned: ``tests/test_frames.py`` checks both functions against the Pillow / torchvision installed in the build
container (Pillow 12.2, torchvision 0.26 — same algorithm as the pinned versions) and against the committed
fixture ``tests/golden/frames_golden.npz`` written by ``oracle/make_golden_frames.py`` from those libraries.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def center_crop_box(height: int, width: int, height_frac: float = 0.8, width_frac: float = 0.8):
    """(top, left, crop_h, crop_w) of ``CenterCrop((height_frac*height, width_frac*width))`` on a PIL image.

    main_dino.py:298-301 (``getCropDims`` :317-322).  torchvision: ``crop_top = int(round((H - ch) / 2.))`` with the
    float ``ch``; PIL ``crop`` then rounds the float box ``(left, top, left + cw, top + ch)`` corner by corner."""
    ch, cw = height_frac * height, width_frac * width
    top = int(round((height - ch) / 2.0))
    left = int(round((width - cw) / 2.0))
    x0, y0, x1, y1 = (int(round(v)) for v in (left, top, left + cw, top + ch))
    return y0, x0, y1 - y0, x1 - x0


def resample_coeffs(in_size: int, out_size: int):
    """``precompute_coeffs`` + ``normalize_coeffs_8bpc`` of Pillow's Resample.c for the bilinear (triangle, support 1)
    filter over the full input range.  Returns ``(bounds int32 [out,2] = (xmin, count), kk int32 [out,ksize], ksize)``."""
    in0, in1 = 0.0, float(in_size)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(xmax, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit resampling pass along ``axis`` (0 = rows / vertical, 1 = columns / horizontal) of ``[H,W,C]`` uint8."""
    bounds, kk, _ = resample_coeffs(img.shape[axis], out_size)
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, np.uint8)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    dst = np.moveaxis(out, axis, 0)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0))
        dst[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img: np.ndarray, out_h: int = 224, out_w: int = 224) -> np.ndarray:
    """``Image.resize((out_w, out_h), BILINEAR)`` on ``[H,W,C]`` uint8: horizontal pass, uint8 rounding, vertical pass
    (``ImagingResample``; a pass whose size does not change is skipped, as in Pillow)."""
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return np.ascontiguousarray(img)


def crop_resize_frames(frames: np.ndarray, height_frac: float = 0.8, width_frac: float = 0.8, out: int = 224) -> np.ndarray:
    """The whole front-end on ``[N,H,W,3]`` uint8 decoded frames -> ``[N,out,out,3]`` uint8 (what ``ToTensor`` sees)."""
    n, h, w, _ = frames.shape
    top, left, ch, cw = center_crop_box(h, w, height_frac, width_frac)
    res = np.empty((n, out, out, frames.shape[3]), np.uint8)
    for i in range(n):
        res[i] = resize_bilinear_u8(frames[i, top:top + ch, left:left + cw], out, out)
    return res
