"""Thin Python wrappers over the individual C-ABI kernels (used by the parity tests and by callers that want a
single op).  Each wrapper allocates outputs with torch, passes raw device pointers and the current stream, and
raises :class:`sais_b200._lib.SaisError` on any non-zero return code.  No CPU fallbacks."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import SaisGemmArgs, check, current_stream, lib, ptr, require_cuda

import functools


def _on_device_of_first_tensor(fn):
    """Run ``fn`` with the CUDA device of its first tensor argument current: the library launches on
    ``torch.cuda.current_stream()`` of the CURRENT device, so a model living on cuda:1 (``loadModel(rank=1)``) must not
    launch on cuda:0's stream with cuda:1 pointers."""

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.Tensor):
                if a.is_cuda:
                    with torch.cuda.device(a.device):
                        return fn(*args, **kw)
                break
        return fn(*args, **kw)  # CPU tensor: the op itself raises SaisError (no CPU path)

    return wrapper


_IMAGENET_MEAN = (0.485, 0.456, 0.406)  # extract_representations.py:161
_IMAGENET_STD = (0.229, 0.224, 0.225)


def split_bf16(t):
    """fp32 [R,C] -> bf16 [R,2C] = [hi | lo] with hi = bf16(t), lo = bf16(t - hi): t == hi + lo to ~2^-17."""
    t = t.float()
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


@_on_device_of_first_tensor
def gemm_bias_act(a, w, bias=None, act=_lib.ACT_NONE, residual=None, out_dtype=torch.bfloat16, out=None,
                  row_add=None, remap_group=0, split3=False, split_out=False, ln_stats_in=None, ln_colsum=None,
                  ln_eps=1e-6, ln_stats_out=None, out2=None, k_slices=0):
    """out = act(a @ w.T + bias) [+ residual].  a: bf16 [M,K]; w: bf16 [N,K]; bias/residual fp32.
    split3: a and w are [hi | lo] halves ([M,2K], [N,2K], see split_bf16) and the product is fp32-equivalent;
    split_out: the bf16 output is written as [hi | lo] ([M,2N])."""
    require_cuda(a, "a")
    require_cuda(w, "w")
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    M, K = a.shape
    if split3:
        K //= 2
    N = w.shape[0]
    rows_out = M if remap_group == 0 else (M // remap_group) * (remap_group + 1)
    if out is None:
        out = torch.empty((rows_out, N * (2 if split_out else 1)), device=a.device, dtype=out_dtype)
    g = SaisGemmArgs()
    g.a, g.w, g.bias = ptr(a), ptr(w), ptr(bias)
    g.residual = ptr(residual)
    g.row_add = ptr(row_add)
    if out.dtype == torch.float32:
        g.out_f32, g.ldo32 = ptr(out), out.stride(0)
    else:
        g.out_bf16, g.ldo16 = ptr(out), out.stride(0)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw = a.stride(0), w.stride(0)
    g.ldr = residual.stride(0) if residual is not None else 0
    g.act, g.remap_group = act, remap_group
    g.split3, g.split_out = int(split3), int(split_out)
    # LayerNorm folding (see include/sais_b200.h): consumer side / producer side
    g.ln_stats_in, g.ln_colsum, g.ln_eps = ptr(ln_stats_in), ptr(ln_colsum), float(ln_eps)
    g.ln_stats_out, g.out2_bf16 = ptr(ln_stats_out), ptr(out2)
    g.ldo2 = out2.stride(0) if out2 is not None else 0
    g.k_slices = int(k_slices)  # >= 1: accumulate mode, out += a @ w.T in that many K slices (TMA reduce-add)
    check(lib().sais_gemm_bias_act(C.byref(g), current_stream()), "sais_gemm_bias_act")
    return out


@_on_device_of_first_tensor
def rowstats_cast(x):
    """fp32 [rows,384] -> (bf16 copy [rows,384], stats fp32 [rows,8] = {sum, sum of squares, 0...})."""
    require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 384
    xb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty((x.shape[0], 8), device=x.device, dtype=torch.float32)
    check(lib().sais_rowstats_cast(ptr(x), x.shape[0], ptr(xb), ptr(stats), current_stream()), "sais_rowstats_cast")
    return xb, stats


def fold_layernorm(gamma, beta, weight, bias):
    """LN(x) W^T + b = rstd (x W'^T - mean c) + d:  returns (W' bf16 [N,K], c fp32 [N], d fp32 [N])."""
    wg = (weight.float() * gamma.float()[None, :]).to(torch.bfloat16).contiguous()
    return wg, wg.float().sum(dim=1).contiguous(), (weight.float() @ beta.float() + bias.float()).contiguous()


@_on_device_of_first_tensor
def vit_mlp(xn, fc1_w, fc1_b, fc2_w, fc2_b, x):
    """x += fc2(GELU(fc1(xn) + fc1_b)) + fc2_b in one kernel.  xn bf16 [rows,384]; weights bf16; x fp32 [rows,384],
    updated IN PLACE and returned."""
    require_cuda(xn, "xn")
    require_cuda(x, "x")
    assert xn.dtype == torch.bfloat16 and x.dtype == torch.float32 and xn.is_contiguous() and x.is_contiguous()
    assert fc1_w.dtype == torch.bfloat16 and fc2_w.dtype == torch.bfloat16
    assert tuple(fc1_w.shape) == (1536, 384) and tuple(fc2_w.shape) == (384, 1536) and xn.shape[1] == 384
    assert fc1_w.is_contiguous() and fc2_w.is_contiguous() and x.shape == xn.shape
    check(lib().sais_vit_mlp(ptr(xn), ptr(fc1_w), ptr(fc1_b), ptr(fc2_w), ptr(fc2_b), ptr(x), xn.shape[0],
                             current_stream()), "sais_vit_mlp")
    return x


@_on_device_of_first_tensor
def vit_mlp_ln(xb, ln_stats, fc1_wg, fc1_c, fc1_d, fc2_w, fc2_b, x, ln_eps=1e-6, xb_out=None, stats_out=None):
    """vit_mlp with norm2 folded in: xb = raw bf16 rows + ln_stats [rows,8] (rowstats_cast / a LayerNorm-producer GEMM),
    (fc1_wg, fc1_c, fc1_d) = fold_layernorm(gamma, beta, fc1_w, fc1_b).  x fp32 [rows,384] updated IN PLACE.
    xb_out bf16 [rows,384] / stats_out fp32 [rows,8] (both or neither; may be xb / ln_stats themselves): receive
    rowstats_cast(x_updated), produced inside the kernel."""
    require_cuda(xb, "xb")
    require_cuda(x, "x")
    assert xb.dtype == torch.bfloat16 and x.dtype == torch.float32 and xb.is_contiguous() and x.is_contiguous()
    assert fc1_wg.dtype == torch.bfloat16 and fc2_w.dtype == torch.bfloat16 and ln_stats.dtype == torch.float32
    assert tuple(fc1_wg.shape) == (1536, 384) and tuple(fc2_w.shape) == (384, 1536) and xb.shape[1] == 384
    assert tuple(ln_stats.shape) == (xb.shape[0], 8) and ln_stats.is_contiguous() and x.shape == xb.shape
    if (xb_out is None) != (stats_out is None):
        raise ValueError("xb_out and stats_out go together")
    if xb_out is not None:
        assert xb_out.dtype == torch.bfloat16 and xb_out.shape == xb.shape and xb_out.is_contiguous()
        assert stats_out.dtype == torch.float32 and tuple(stats_out.shape) == (xb.shape[0], 8) and stats_out.is_contiguous()
    check(lib().sais_vit_mlp_ln(ptr(xb), ptr(ln_stats), float(ln_eps), ptr(fc1_wg), ptr(fc1_c), ptr(fc1_d), ptr(fc2_w),
                                ptr(fc2_b), ptr(x), xb.shape[0], ptr(xb_out), ptr(stats_out), current_stream()),
          "sais_vit_mlp_ln")
    return x


@_on_device_of_first_tensor
def layernorm(x, gamma, beta, eps, out_f32=False, out_bf16=True, in_pitch=None, rows=None, split_out=False):
    require_cuda(x, "x")
    assert x.dtype == torch.float32
    if in_pitch is None:
        x2 = x.reshape(-1, x.shape[-1])
        in_pitch, rows = x2.stride(0), x2.shape[0]
    cols = gamma.numel()
    of = torch.empty((rows, cols), device=x.device, dtype=torch.float32) if out_f32 else None
    ob = torch.empty((rows, cols * (2 if split_out else 1)), device=x.device,
                     dtype=torch.bfloat16) if out_bf16 else None
    check(lib().sais_layernorm(ptr(x), in_pitch, ptr(gamma), ptr(beta), float(eps), rows, cols, ptr(of), ptr(ob),
                               int(split_out), current_stream()), "sais_layernorm")
    return of, ob


@_on_device_of_first_tensor
def normalize_patchify_u8(frames, split_out=False):
    """u8 [B,224,224,3] -> bf16 patches [B*196,768] with ImageNet normalisation ([B*196,1536] = [hi|lo] if split)."""
    require_cuda(frames, "frames")
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and tuple(frames.shape[1:]) == (224, 224, 3)
    B = frames.shape[0]
    out = torch.empty((B * 196, 768 * (2 if split_out else 1)), device=frames.device, dtype=torch.bfloat16)
    mean = (C.c_float * 3)(*_IMAGENET_MEAN)
    std = (C.c_float * 3)(*_IMAGENET_STD)
    check(lib().sais_normalize_patchify_u8(ptr(frames), B, C.cast(mean, C.c_void_p), C.cast(std, C.c_void_p),
                                           ptr(out), int(split_out), current_stream()), "sais_normalize_patchify_u8")
    return out


@_on_device_of_first_tensor
def patchify_f32(frames, split_out=False):
    """fp32 [B,3,224,224] (already normalised) -> bf16 patches [B*196,768] ([B*196,1536] = [hi|lo] if split)."""
    require_cuda(frames, "frames")
    assert frames.dtype == torch.float32 and frames.is_contiguous() and tuple(frames.shape[1:]) == (3, 224, 224)
    B = frames.shape[0]
    out = torch.empty((B * 196, 768 * (2 if split_out else 1)), device=frames.device, dtype=torch.bfloat16)
    check(lib().sais_patchify_f32(ptr(frames), B, ptr(out), int(split_out), current_stream()), "sais_patchify_f32")
    return out


@_on_device_of_first_tensor
def vit_attention(qkv, B, emit_probs=False):
    """qkv bf16 [B*197,1152] -> (out bf16 [B*197,384], probs fp32 [B,6,197,197] or None)."""
    require_cuda(qkv, "qkv")
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * 197, 1152)
    out = torch.empty((B * 197, 384), device=qkv.device, dtype=torch.bfloat16)
    probs = torch.empty((B, 6, 197, 197), device=qkv.device, dtype=torch.float32) if emit_probs else None
    check(lib().sais_vit_attention(ptr(qkv), B, ptr(out), ptr(probs), current_stream()), "sais_vit_attention")
    return out, probs


@_on_device_of_first_tensor
def vit_cls_attention(qkv, B):
    """qkv bf16 [B*197,1152] -> attention output of the CLS query only, bf16 [B,384] (last-block shortcut)."""
    require_cuda(qkv, "qkv")
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * 197, 1152)
    out = torch.empty((B, 384), device=qkv.device, dtype=torch.bfloat16)
    check(lib().sais_vit_cls_attention(ptr(qkv), B, ptr(out), current_stream()), "sais_vit_cls_attention")
    return out


@_on_device_of_first_tensor
def temporal_attention(qkv, seq_offsets, key_pad=None, attn_offsets=None, max_S=None, attn_numel=0):
    """qkv fp32 [tokens,1152]; seq_offsets int32 [nseq+1] (device); returns (out bf16 [tokens,768] = [hi|lo] halves
    of the fp32 attention output, head-averaged attention maps flat fp32 or None)."""
    require_cuda(qkv, "qkv")
    assert qkv.dtype == torch.float32 and qkv.is_contiguous()
    nseq = seq_offsets.numel() - 1
    out = torch.empty((qkv.shape[0], 768), device=qkv.device, dtype=torch.bfloat16)
    attn = torch.zeros(attn_numel, device=qkv.device, dtype=torch.float32) if attn_numel else None
    check(lib().sais_temporal_attention(ptr(qkv), ptr(seq_offsets), ptr(key_pad), ptr(attn_offsets), nseq,
                                        int(max_S), ptr(out), ptr(attn), current_stream()),
          "sais_temporal_attention")
    return out, attn


@_on_device_of_first_tensor
def clip_head(cls_a, cls_b, B, nsnip, lin_w, lin_b):
    require_cuda(cls_a, "cls_a")
    out = torch.empty((B, 256), device=cls_a.device, dtype=torch.float32)
    check(lib().sais_clip_head(ptr(cls_a), ptr(cls_b), B, nsnip, ptr(lin_w), ptr(lin_b), ptr(out),
                               current_stream()), "sais_clip_head")
    return out


@_on_device_of_first_tensor
def prototype_score(reps, protos, want_sims=False):
    """reps fp32 [B,D], protos fp32 [P,D] -> (probs [B,P], sims or None, pred int32 [B])."""
    require_cuda(reps, "reps")
    reps = reps.contiguous().float()
    protos = protos.contiguous().float().to(reps.device)
    B, D = reps.shape
    P = protos.shape[0]
    probs = torch.empty((B, P), device=reps.device, dtype=torch.float32)
    sims = torch.empty((B, P), device=reps.device, dtype=torch.float32) if want_sims else None
    pred = torch.empty((B,), device=reps.device, dtype=torch.int32)
    check(lib().sais_prototype_score(ptr(reps), ptr(protos), B, P, D, ptr(probs), ptr(sims), ptr(pred),
                                     current_stream()), "sais_prototype_score")
    return probs, sims, pred


@_on_device_of_first_tensor
def add_pos_rows(x, pos):
    """x fp32 [..., S, 384] + pos fp32 [S, 384] broadcast over the leading dims (clip positional embeddings)."""
    require_cuda(x, "x")
    x = x.contiguous().float()
    pos = pos.contiguous().float()
    assert x.shape[-1] == 384 and pos.shape[-1] == 384 and x.shape[-2] == pos.shape[0]
    out = torch.empty_like(x)
    check(lib().sais_add_pos_rows(ptr(x), ptr(pos), x.numel() // 384, pos.shape[0], ptr(out), current_stream()),
          "sais_add_pos_rows")
    return out


@_on_device_of_first_tensor
def mil_head(enc_out, att_a, att_b, att_c_w, att_c_b, final_w, final_b):
    """enc_out fp32 [B,nsnip,384] (clip encoder output before the ReLU); att_a / att_b: (weight [256,384], bias [256]);
    att_c_w [ncls,256], att_c_b [ncls], final_w [ncls,384], final_b [ncls].
    Returns (reps [B,nsnip,384], logits [B,ncls], attention [ncls,B,nsnip])."""
    require_cuda(enc_out, "enc_out")
    enc_out = enc_out.contiguous().float()
    B, ns, E = enc_out.shape
    ncls = att_c_w.shape[0]
    f = lambda t: t.detach().contiguous().float()
    ws = [f(att_a[0]), f(att_a[1]), f(att_b[0]), f(att_b[1]), f(att_c_w), f(att_c_b), f(final_w), f(final_b)]
    reps = torch.empty_like(enc_out)
    logits = torch.empty((B, ncls), device=enc_out.device, dtype=torch.float32)
    attn = torch.empty((ncls, B, ns), device=enc_out.device, dtype=torch.float32)
    check(lib().sais_mil_head(ptr(enc_out), B, ns, ncls, *[ptr(w) for w in ws], ptr(reps), ptr(logits), ptr(attn),
                              current_stream()), "sais_mil_head")
    return reps, logits, attn
