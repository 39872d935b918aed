"""Frame front-end: the crop + resize the reference applies to every decoded frame before ``ToTensor``/``Normalize``.

Reference call sites:

* ``transforms.CenterCrop((height_frac*height, width_frac*width))`` in the dataset's ``__getitem__`` —
  ``SAIS/scripts/dino-main/main_dino.py:298-301``; fractions from ``getCropDims`` (:317-322): 0.8 / 0.8, or
  0.8 / 0.7 for the ``*_Gronau`` datasets;
* ``transforms.Resize((224, 224))`` — ``SAIS/scripts/extract_representations.py:158-162`` — PIL bilinear (antialiased).

Here both run on the GPU over a whole batch of decoded ``uint8 [N,H,W,3]`` frames (``sais_crop_resize_u8``,
``csrc/frames.cu``), bit-identical to Pillow; the result feeds ``VisionTransformer.forward_u8`` (which fuses
ToTensor + Normalize + patch layout).  JPEG decoding itself is plumbing (``torchvision.io.decode_jpeg(device='cuda')``
= nvJPEG, or any decoder that yields uint8 HWC frames).  No CPU path: CPU tensors raise ``SaisError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import check, current_stream, lib, ptr, require_cuda

_GRONAU = ("NS_Gronau", "VUA_Gronau")


def get_crop_dims(dataset: str = "") -> Tuple[float, float]:
    """``getCropDims`` (main_dino.py:317-322): (height_frac, width_frac)."""
    return (0.8, 0.7) if dataset in _GRONAU else (0.8, 0.8)


def center_crop_box(height: int, width: int, height_frac: float = 0.8, width_frac: float = 0.8):
    """(top, left, crop_h, crop_w) of the reference's float-sized ``CenterCrop`` on a PIL image (host arithmetic)."""
    box = (C.c_int32 * 4)()
    check(lib().sais_center_crop_box(int(height), int(width), float(height_frac), float(width_frac), box),
          "sais_center_crop_box")
    return tuple(int(v) for v in box)


def resize_table(in_size: int) -> np.ndarray:
    """Pillow's 8-bit bilinear coefficient table for ``in_size -> 224`` (host): int32 ``[224*2 + ksize*224]``."""
    n = int(lib().sais_resize_table_ints(int(in_size)))
    if n <= 0:
        raise ValueError(f"bad source size {in_size}")
    tab = np.empty(n, np.int32)
    check(lib().sais_resize_build_table(int(in_size), tab.ctypes.data_as(C.POINTER(C.c_int32))), "sais_resize_build_table")
    return tab


_tables: Dict[Tuple[int, str], torch.Tensor] = {}


def _device_table(in_size: int, device) -> torch.Tensor:
    key = (in_size, str(device))
    t = _tables.get(key)
    if t is None:
        t = torch.from_numpy(resize_table(in_size)).to(device)
        _tables[key] = t
    return t


@torch.no_grad()
def crop_resize(frames: torch.Tensor, height_frac: float = 0.8, width_frac: float = 0.8, box=None) -> torch.Tensor:
    """Decoded frames ``uint8 [N,H,W,3]`` (device) -> ``uint8 [N,224,224,3]``: centre crop (or an explicit
    ``box = (top, left, h, w)``) + Pillow-exact bilinear resize."""
    require_cuda(frames, "frames")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_contiguous():
        raise ValueError("frames must be a contiguous uint8 [N,H,W,3] tensor")
    n, h, w, _ = frames.shape
    top, left, ch, cw = box if box is not None else center_crop_box(h, w, height_frac, width_frac)
    out = torch.empty((n, 224, 224, 3), dtype=torch.uint8, device=frames.device)
    if n == 0:
        return out
    tmp = torch.empty((n, ch, 224, 3), dtype=torch.uint8, device=frames.device)
    th, tv = _device_table(cw, frames.device), _device_table(ch, frames.device)
    with torch.cuda.device(frames.device):
        check(lib().sais_crop_resize_u8(ptr(frames), n, h, w, top, left, ch, cw, ptr(th), ptr(tv), ptr(tmp), ptr(out),
                                        current_stream()), "sais_crop_resize_u8")
    return out
