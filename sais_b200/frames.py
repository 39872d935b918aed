"""Frame front-end: the crop + resize the reference applies to every decoded frame before ``ToTensor``/``Normalize``.

Reference call sites:

* ``transforms.CenterCrop((height_frac*height, width_frac*width))`` in the dataset's ``__getitem__`` —
  ``SAIS/scripts/dino-main/main_dino.py:298-301``; fractions from ``getCropDims`` (:317-322): 0.8 / 0.8, or
  0.8 / 0.7 for the ``*_Gronau`` datasets;
* ``transforms.Resize((224, 224))`` — ``SAIS/scripts/extract_representations.py:158-162`` — PIL bilinear (antialiased).

Here both run on the GPU over a whole batch of decoded ``uint8 [N,H,W,3]`` frames (``sais_crop_resize_u8``,
``csrc/frames.cu``), bit-identical to Pillow; the result feeds ``VisionTransformer.forward_u8`` (which fuses
ToTensor + Normalize + patch layout).  :func:`decode_jpegs` / :func:`load_frames` put the JPEG decode in front of it
(one batched nvJPEG call straight into the ``[N,H,W,3]`` buffer the crop-resize kernel reads; the reference decodes with
Pillow on the host, ``main_dino.py:313`` via ``ImageFolder``).  The optical-flow stream takes the same route: RAFT's output
is stored by the reference as ``flows_%08d.jpg`` images (``extract_representations.py:246-261``), so externally supplied
flow frames are JPEG inputs like the RGB ones — :func:`flow_frame_name` gives the reference's file name for frame ``n``.
No CPU path: CPU tensors raise ``SaisError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple

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


# --------------------------------------------------------------------------------------------- JPEG front-end
def jpeg_size(data: bytes) -> Tuple[int, int]:
    """(height, width) from the frame header of one JPEG stream (host-only parse, no CUDA context)."""
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    hw = (C.c_int32 * 2)()
    check(lib().sais_jpeg_info(C.cast(buf, C.c_void_p), len(data), hw), "sais_jpeg_info")
    return int(hw[0]), int(hw[1])


@torch.no_grad()
def decode_jpegs(streams: Sequence[bytes], device, out: torch.Tensor = None) -> torch.Tensor:
    """Same-sized JPEG streams (``bytes``, as read from the frame files) -> ``uint8 [N,H,W,3]`` RGB on ``device``."""
    device = torch.device(device)
    if device.type != "cuda":
        from ._lib import SaisError
        raise SaisError("decode_jpegs decodes on a CUDA device (there is no CPU path)")
    n = len(streams)
    if n == 0:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8, device=device)
    h, w = jpeg_size(streams[0])
    if out is None:
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=device)
    elif tuple(out.shape) != (n, h, w, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != device:
        raise ValueError(f"out must be a contiguous uint8 [{n},{h},{w},3] tensor on {device}")
    # pointers into the bytes objects themselves (no copy of the bit streams; `streams` keeps them alive past the sync below)
    bufs = [d if isinstance(d, bytes) else bytes(d) for d in streams]
    ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in bufs])
    lens = (C.c_size_t * n)(*[len(b) for b in bufs])
    with torch.cuda.device(device):
        check(lib().sais_jpeg_decode_batch(C.cast(ptrs, C.c_void_p), C.cast(lens, C.c_void_p), n, h, w, ptr(out),
                                           current_stream()), "sais_jpeg_decode_batch")
        if lib().sais_jpeg_last_backend() != 3:
            # the batched API may read the host bit streams while the decode is in flight: keep them alive until the stream
            # has passed it.  (The threaded decoder entropy-decodes every stream on the host before its call returns; only
            # the IDCT / colour kernels are still in flight then, ordered before anything the caller enqueues next.)
            torch.cuda.current_stream(device).synchronize()
    return out


@torch.no_grad()
def load_frames(streams: Sequence[bytes], device, dataset: str = "") -> torch.Tensor:
    """JPEG streams -> decoded -> centre-cropped (``getCropDims``) -> resized ``uint8 [N,224,224,3]``: everything the
    reference's dataset + transform pipeline does to a frame before ``ToTensor`` (main_dino.py:295-313)."""
    frames = decode_jpegs(streams, device)
    hf, wf = get_crop_dims(dataset)
    return crop_resize(frames, hf, wf)


def flow_frame_name(nflow: int) -> str:
    """File name the reference gives the optical-flow image of frame pair ``nflow`` (saveFlows,
    extract_representations.py:253-254): ``flows_`` + the index zero-padded to 8 digits + ``.jpg``."""
    return "flows_" + "0" * (8 - len(str(int(nflow)))) + str(int(nflow)) + ".jpg"
