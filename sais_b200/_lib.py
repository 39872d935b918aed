"""ctypes binding of ``libsais_b200.so`` (C ABI declared in ``include/sais_b200.h``).

PyTorch is used only for device memory and streams: every entry point receives raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream handle.  There is NO CPU fallback — if the shared
library is missing this module raises, and the kernels themselves fail without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
_CSRC = _PKG_DIR / "csrc"
# (dev: SAIS_B200_LIB points at an alternative build of the same C ABI, e.g. for A/B timing of a kernel rewrite)
LIB_PATH = Path(os.environ["SAIS_B200_LIB"]).resolve() if os.environ.get("SAIS_B200_LIB") else _PKG_DIR / "libsais_b200.so"

VIT_DEPTH = 12
TMP_LAYERS = 4

ACT_NONE, ACT_GELU_ERF, ACT_RELU = 0, 1, 2
INPUT_F32_CHW, INPUT_U8_HWC = 0, 1

_p = C.c_void_p


class SaisGemmArgs(C.Structure):
    _fields_ = [
        ("a", _p), ("w", _p), ("bias", _p), ("residual", _p), ("out_f32", _p), ("out_bf16", _p), ("row_add", _p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldr", C.c_int64), ("ldo32", C.c_int64), ("ldo16", C.c_int64),
        ("act", C.c_int32), ("remap_group", C.c_int32), ("split3", C.c_int32), ("split_out", C.c_int32),
        ("ln_stats_in", _p), ("ln_colsum", _p), ("ln_stats_out", _p), ("out2_bf16", _p), ("ldo2", C.c_int64),
        ("ln_eps", C.c_float), ("k_slices", C.c_int32),
    ]


class SaisVitBlockWeights(C.Structure):
    _fields_ = [(n, _p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
        "qkv_wg", "qkv_c", "qkv_d", "fc1_wg", "fc1_c", "fc1_d")]


class SaisVitWeights(C.Structure):
    _fields_ = [
        ("patch_w", _p), ("patch_b", _p), ("cls_pos0", _p), ("pos_patch", _p),
        ("blocks", SaisVitBlockWeights * VIT_DEPTH),
        ("norm_w", _p), ("norm_b", _p),
    ]


MAX_PEERS = 15


class SaisFanout(C.Structure):
    """include/sais_b200.h SaisFanout: where this rank's output slice lives in the other GPUs' mappings."""
    _fields_ = [("multicast", C.c_void_p), ("peers", C.c_void_p * MAX_PEERS), ("n_peers", C.c_int32)]


def fanout_shifted(f, byte_offset):
    """Copy of a SaisFanout with every address advanced by ``byte_offset`` (a sub-range of the output slice)."""
    g = SaisFanout()
    g.multicast = (f.multicast + byte_offset) if f.multicast else None
    g.n_peers = f.n_peers
    for i in range(f.n_peers):
        g.peers[i] = f.peers[i] + byte_offset
    return g


class SaisTemporalLayerWeights(C.Structure):
    _fields_ = [(n, _p) for n in (
        "in_w", "in_b", "out_w", "out_b", "n1_w", "n1_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "n2_w", "n2_b")]


class SaisTemporalWeights(C.Structure):
    _fields_ = [
        ("frame_cls", _p), ("frame_pos", _p), ("n_pos", C.c_int32),
        ("layers", SaisTemporalLayerWeights * TMP_LAYERS),
    ]


# name -> (restype, argtypes); must list every symbol include/sais_b200.h declares
SIGNATURES = {
    "sais_version": (C.c_int, []),
    "sais_last_error": (C.c_char_p, []),
    "sais_launch_count": (C.c_int64, []),
    "sais_set_sm_limit": (C.c_int, [C.c_int32]),
    "sais_clock_probe": (C.c_int, [_p, C.c_int32, _p]),
    "sais_profile_begin": (None, []),
    "sais_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "sais_gemm_bias_act": (C.c_int, [C.POINTER(SaisGemmArgs), _p]),
    "sais_vit_mlp": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int64, _p]),
    "sais_vit_mlp_ln": (C.c_int, [_p, _p, C.c_float, _p, _p, _p, _p, _p, _p, C.c_int64, _p, _p, _p]),
    "sais_rowstats_cast": (C.c_int, [_p, C.c_int64, _p, _p, _p]),
    "sais_layernorm": (C.c_int, [_p, C.c_int64, _p, _p, C.c_float, C.c_int64, C.c_int32, _p, _p, C.c_int32, _p]),
    "sais_jpeg_info": (C.c_int, [_p, C.c_size_t, C.POINTER(C.c_int32)]),
    "sais_jpeg_last_backend": (C.c_int, []),
    "sais_jpeg_decode_batch": (C.c_int, [_p, _p, C.c_int32, C.c_int32, C.c_int32, _p, _p]),
    "sais_center_crop_box": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int32)]),
    "sais_resize_table_ints": (C.c_int64, [C.c_int32]),
    "sais_resize_build_table": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "sais_crop_resize_u8": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      _p, _p, _p, _p, _p]),
    "sais_normalize_patchify_u8": (C.c_int, [_p, C.c_int32, _p, _p, _p, C.c_int32, _p]),
    "sais_patchify_f32": (C.c_int, [_p, C.c_int32, _p, C.c_int32, _p]),
    "sais_vit_attention": (C.c_int, [_p, C.c_int32, _p, _p, _p]),
    "sais_vit_cls_attention": (C.c_int, [_p, C.c_int32, _p, _p]),
    "sais_vit_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "sais_vit_forward": (C.c_int, [C.POINTER(SaisVitWeights), _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p,
                                   C.c_size_t, _p, _p, _p, _p]),
    "sais_set_mlp_policy": (C.c_int, [C.c_int32]),
    "sais_vit_forward_layers": (C.c_int, [C.POINTER(SaisVitWeights), _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p,
                                          C.c_size_t, _p, C.c_int32, _p, _p]),
    "sais_vit_forward_fanout": (C.c_int, [C.POINTER(SaisVitWeights), _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p,
                                          C.c_size_t, _p, _p, _p, C.POINTER(SaisFanout), _p]),
    "sais_temporal_prep": (C.c_int, [_p, _p, C.c_int32, C.c_int32, _p, _p, C.c_int32, _p, _p, _p]),
    "sais_temporal_attention": (C.c_int, [_p, _p, _p, _p, C.c_int32, C.c_int32, _p, _p, _p]),
    "sais_temporal_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "sais_temporal_forward": (C.c_int, [C.POINTER(SaisTemporalWeights), _p, _p, _p, _p, C.c_int32, C.c_int32,
                                        C.c_int32, _p, C.c_size_t, _p, _p, _p, _p]),
    "sais_clip_head": (C.c_int, [_p, _p, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "sais_add_pos_rows": (C.c_int, [_p, _p, C.c_int64, C.c_int32, _p, _p]),
    "sais_mil_head": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "sais_prototype_score": (C.c_int, [_p, _p, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p]),
}

_lib = None


class SaisError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile the CUDA sources in-tree for sm_100a (``make -C sais_b200/csrc``)."""
    cmd = ["make", "-C", str(_CSRC), "../libsais_b200.so"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise SaisError("building libsais_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("SAIS_B200_NO_BUILD") or not (_CSRC / "Makefile").exists():
            raise SaisError(f"{LIB_PATH} is missing and there is no CPU fallback; run __graft_entry__.build()")
        build()
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


class sm_limit:
    """``with sm_limit(144): vit.forward_u8(...)`` — kernels launched inside occupy at most that many SMs
    (``sais_set_sm_limit``); the rest stay free for kernels launched on another stream outside the block."""

    def __init__(self, n_sms: int):
        self.n, self.prev = int(n_sms), 0

    def __enter__(self):
        self.prev = lib().sais_set_sm_limit(self.n)
        if self.prev < 0:
            check(self.prev, "sais_set_sm_limit")
        return self

    def __exit__(self, *exc):
        lib().sais_set_sm_limit(self.prev)
        return False


class mlp_policy:
    """``with mlp_policy(1): ...`` — ``sais_set_mlp_policy`` for the forwards launched inside: 0 = fused MLP kernel always
    (default: embeddings independent of the batch size, bit for bit), 1 = fc1 / fc2 GEMM pair below 48 frames (single-clip
    latency), 2 = GEMM pair always."""

    def __init__(self, policy: int):
        self.p, self.prev = int(policy), 0

    def __enter__(self):
        self.prev = lib().sais_set_mlp_policy(self.p)
        if self.prev < 0:
            check(self.prev, "sais_set_mlp_policy")
        return self

    def __exit__(self, *exc):
        lib().sais_set_mlp_policy(self.prev)
        return False


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().sais_last_error().decode("utf-8", "replace")
        raise SaisError(f"{what or 'sais call'} failed with code {rc}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str) -> None:
    if not t.is_cuda:
        raise SaisError(f"{name} must be a CUDA tensor: sais_b200 has no CPU path (got device {t.device})")
