"""GPU parity tests, one per C-ABI kernel, against plain fp32 CPU restatements on the same seeded inputs.
Operands that the kernels consume as bf16 are rounded to bf16 on the CPU side too, so the comparison isolates the
kernel arithmetic (fp32 accumulation) from the format conversion; tolerances are written next to each check."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sais_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def bf(t):
    return t.to(torch.bfloat16).float()


def rnd(*shape, seed=0, std=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * std


# ------------------------------------------------------------------------------------------------ layernorm
@pytest.mark.parametrize("rows", [1, 7, 8, 197 * 3, 4099])
@pytest.mark.parametrize("eps", [1e-6, 1e-5])
def test_layernorm(dev, rows, eps):
    from sais_b200 import ops
    x = rnd(rows, 384, seed=rows) * 3 + 0.5
    w, b = 1 + 0.1 * rnd(384, seed=1), 0.1 * rnd(384, seed=2)
    of, ob = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), eps, out_f32=True, out_bf16=True)
    ref = O.layer_norm(x, w, b, eps)
    assert torch.allclose(of.cpu(), ref, atol=2e-5, rtol=1e-5)
    assert torch.equal(ob.cpu().float(), bf(of.cpu()))  # bf16 output is the rounding of the fp32 one


def test_layernorm_strided_rows(dev):
    """final norm reads only the CLS rows: row pitch 197*384."""
    from sais_b200 import ops
    x = rnd(5, 197, 384, seed=3)
    w, b = 1 + 0.1 * rnd(384, seed=1), 0.1 * rnd(384, seed=2)
    of, _ = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-6, out_f32=True, out_bf16=False,
                          in_pitch=197 * 384, rows=5)
    assert torch.allclose(of.cpu(), O.layer_norm(x[:, 0], w, b, 1e-6), atol=2e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ patchify
def _patches_ref(x):  # x fp32 [B,3,224,224] -> [B*196,768], k = c*256 + ky*16 + kx
    B = x.shape[0]
    return x.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * 196, 768)


@pytest.mark.parametrize("B", [1, 3])
def test_patchify_f32(dev, B):
    from sais_b200 import ops
    x = rnd(B, 3, 224, 224, seed=B)
    got = ops.patchify_f32(x.to(dev)).cpu().float()
    assert torch.equal(got, bf(_patches_ref(x)))  # pure layout + round-to-nearest-even: bit exact


@pytest.mark.parametrize("B", [1, 5])
def test_normalize_patchify_u8(dev, B):
    from sais_b200 import ops
    fr = O.make_frames_u8(B, seed=B)
    fr[0, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    got = ops.normalize_patchify_u8(fr.to(dev)).cpu().float()
    ref = _patches_ref(O.normalize_frames(fr))
    # fused (u8*scale + shift) vs ((u8/255 - mean)/std): <= 1 bf16 ulp (2^-8 relative) from double rounding
    assert torch.allclose(got, ref, atol=1e-2, rtol=2 ** -7)
    assert (got - bf(ref)).abs().max() <= 2 ** -6  # values are < 2.7 in magnitude: at most one ulp apart
    assert ((got - bf(ref)) != 0).float().mean() < 0.02


# ------------------------------------------------------------------------------------------------ GEMM
def _gemm_ref(a, w, bias, act, residual):
    y = bf(a).double() @ bf(w).double().t()
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2)))
    elif act == 2:
        y = torch.relu(y)
    if residual is not None:
        y = y + residual.double()
    return y.float()


GEMM_CASES = [
    # M, N, K, act, residual, out fp32?
    (300, 384, 384, 0, False, False),
    (1000, 1152, 384, 0, False, False),    # qkv
    (257, 1536, 384, 1, False, False),     # fc1 + GELU(erf)
    (777, 384, 1536, 0, True, True),       # fc2 + residual
    (50, 2048, 384, 2, False, False),      # temporal FF1 + ReLU
    (333, 384, 2048, 0, True, True),       # temporal FF2 + residual
    (11, 1152, 384, 0, False, False),      # C1-sized temporal in-proj (single partial tile)
    (128, 256, 384, 0, False, True),
    (197 * 96, 384, 384, 0, True, True),   # one full wave of 148 m-tiles x 2-3 n-tiles
    (5000, 1536, 384, 1, False, False),    # multi-tile persistent loop with the GELU epilogue
]


@pytest.mark.parametrize("M,N,K,act,use_res,f32_out", GEMM_CASES)
def test_gemm_bias_act(dev, M, N, K, act, use_res, f32_out):
    from sais_b200 import ops
    a = rnd(M, K, seed=M + N)
    w = rnd(N, K, seed=K, std=1 / math.sqrt(K))
    bias = rnd(N, seed=7, std=0.5)
    res = rnd(M, N, seed=9) if use_res else None
    out = ops.gemm_bias_act(a.to(dev).bfloat16(), w.to(dev).bfloat16(), bias.to(dev), act=act,
                            residual=None if res is None else res.to(dev),
                            out_dtype=torch.float32 if f32_out else torch.bfloat16)
    ref = _gemm_ref(a, w, bias, act, res)
    got = out.cpu().float()
    if f32_out:  # fp32 accumulate of K products of O(1): summation-order noise only
        assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), (got - ref).abs().max()
    else:        # plus one bf16 rounding (2^-9 relative)
        assert torch.allclose(got, ref, atol=1e-3, rtol=2 ** -8), (got - ref).abs().max()


_EPILOGUE_VARIANT_SCRIPT = r"""
import math, sys, torch
from sais_b200 import ops
dev = torch.device("cuda:0")
def rnd(*shape, seed=0, std=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * std
outs = []
for M, N, act, fold in [(5000, 1536, 1, False), (197 * 160, 1536, 1, True), (1000, 1152, 0, False), (300, 1152, 0, True),
                        (50, 2048, 2, False), (257, 1536, 1, True)]:
    x = rnd(M, 384, seed=M) * 1.7 + 0.4
    w, b = rnd(N, 384, seed=3, std=1 / math.sqrt(384)), rnd(N, seed=4, std=0.3)
    if fold:
        gamma, beta = 1 + 0.2 * rnd(384, seed=1), 0.2 * rnd(384, seed=2)
        wg, c, d = ops.fold_layernorm(gamma, beta, w, b)
        xb, stats = ops.rowstats_cast(x.to(dev))
        out = ops.gemm_bias_act(xb, wg.to(dev), d.to(dev), act=act, ln_stats_in=stats, ln_colsum=c.to(dev), ln_eps=1e-6)
    else:
        out = ops.gemm_bias_act(x.to(dev).bfloat16(), w.to(dev).bfloat16(), b.to(dev), act=act)
    outs.append(out.cpu())
torch.save(outs, sys.argv[1])
"""


def test_gemm_a_stationary_bit_identical(dev, tmp_path):
    """The A-stationary K = 384 variant (contiguous tile ranges per CTA pair, A rows resident across n-tiles; SAIS_GEMM_ASTAT,
    read once per process) accumulates the same products in the same order as the ring variant: outputs must agree bit for
    bit.  The script's shapes include CTA-pair-sized qkv / fc1 problems (197 * 160 rows) where the variant is active."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    script = _EPILOGUE_VARIANT_SCRIPT.replace("(1000, 1152, 0, False)", "(197 * 160, 1152, 0, True), (197 * 128 + 77, 1152, 0, False)")
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ, SAIS_GEMM_ASTAT=flag, PYTHONPATH=str(root))
        out = tmp_path / f"astat{flag}.pt"
        subprocess.run([sys.executable, "-c", script, str(out)], check=True, env=env, cwd=root, timeout=300)
        res[flag] = torch.load(out)
    for a, b in zip(res["0"], res["1"]):
        assert a.shape == b.shape and torch.equal(a.view(torch.int16), b.view(torch.int16)), \
            (a.float() - b.float()).abs().max()


def test_gemm_epilogue_warp_variants_bit_identical(dev, tmp_path):
    """The 8-warp and the lean 16-warp bf16 epilogues (SAIS_GEMM_EW, read once per process) must agree bit for bit: same
    accumulators, and the 16-warp GELU works on x / 2 with constants scaled by exact powers of two (common.cuh
    gelu_erf_fast2_half).  Shapes: fc1 / qkv / temporal FF1, with and without the folded LayerNorm, partial tiles."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    res = {}
    for ew in ("8", "16"):
        env = dict(os.environ, SAIS_GEMM_EW=ew, PYTHONPATH=str(root))
        out = tmp_path / f"ew{ew}.pt"
        subprocess.run([sys.executable, "-c", _EPILOGUE_VARIANT_SCRIPT, str(out)], check=True, env=env, cwd=root, timeout=300)
        res[ew] = torch.load(out)
    for a, b in zip(res["8"], res["16"]):
        assert a.shape == b.shape and torch.equal(a.view(torch.int16), b.view(torch.int16)), \
            (a.float() - b.float()).abs().max()


@pytest.mark.parametrize("block_rows", [1, 2])
def test_gemm_patch_embed_remap(dev, block_rows):
    """patch-embed epilogue: row b*196+p -> row b*197+1+p, + pos_embed[1+p]; CLS rows untouched."""
    from sais_b200 import ops
    B = 2 * block_rows
    a = rnd(B * 196, 768, seed=1)
    w = rnd(384, 768, seed=2, std=1 / math.sqrt(768))
    bias, pos = rnd(384, seed=3), rnd(196, 384, seed=4)
    out = torch.full((B * 197, 384), -7.0, device=dev)
    ops.gemm_bias_act(a.to(dev).bfloat16(), w.to(dev).bfloat16(), bias.to(dev), out=out, row_add=pos.to(dev),
                      remap_group=196)
    got = out.cpu().view(B, 197, 384)
    ref = _gemm_ref(a, w, bias, 0, None).view(B, 196, 384) + pos
    assert torch.all(got[:, 0] == -7.0)
    assert torch.allclose(got[:, 1:], ref, atol=2e-4, rtol=1e-4)


def test_gemm_in_place_residual(dev):
    """x = x + A W^T + b with residual == out (how proj / fc2 update the fp32 residual stream)."""
    from sais_b200 import ops
    a, w, bias, x = rnd(500, 384, seed=1), rnd(384, 384, seed=2, std=0.05), rnd(384, seed=3), rnd(500, 384, seed=4)
    xd = x.to(dev).clone()
    ops.gemm_bias_act(a.to(dev).bfloat16(), w.to(dev).bfloat16(), bias.to(dev), residual=xd, out=xd)
    assert torch.allclose(xd.cpu(), _gemm_ref(a, w, bias, 0, x), atol=2e-4, rtol=1e-4)


def test_gemm_rejects_bad_shapes(dev):
    from sais_b200 import SaisError, ops
    a, w = torch.zeros(8, 100, device=dev).bfloat16(), torch.zeros(384, 100, device=dev).bfloat16()
    with pytest.raises(SaisError):
        ops.gemm_bias_act(a, w)  # K not a multiple of 64 (and pitch not 16-byte aligned)


@pytest.mark.parametrize("M,K,slices,split3", [(272, 2048, 12, True), (272, 384, 3, True), (50, 2048, 1, False),
                                               (700, 1536, 5, False), (272, 2048, 200, True)])
def test_gemm_accumulate_split_k(dev, M, K, slices, split3):
    """accumulate mode: out += A W^T with the K range split over CTAs (partials meet in L2 by TMA reduce-add)."""
    from sais_b200 import ops
    a, w = rnd(M, K, seed=M + K), rnd(384, K, seed=K, std=1 / math.sqrt(K))
    y0 = rnd(M, 384, seed=5)
    y = y0.to(dev).clone()
    if split3:
        ops.gemm_bias_act(ops.split_bf16(a.to(dev)), ops.split_bf16(w.to(dev)), out=y, split3=True, k_slices=slices)
        ref = (y0.double() + a.double() @ w.double().t()).float()
        tol = dict(atol=3e-4, rtol=1e-4)  # fp32-equivalent product
    else:
        ops.gemm_bias_act(a.to(dev).bfloat16(), w.to(dev).bfloat16(), out=y, k_slices=slices)
        ref = (y0.double() + bf(a).double() @ bf(w).double().t()).float()
        tol = dict(atol=2e-4, rtol=1e-4)
    assert torch.allclose(y.cpu(), ref, **tol), (y.cpu() - ref).abs().max()


# ------------------------------------------------------------------------------------------------ LayerNorm folding
@pytest.mark.parametrize("rows", [1, 33, 4099])
def test_rowstats_cast(dev, rows):
    from sais_b200 import ops
    x = rnd(rows, 384, seed=rows) * 2 + 0.7
    xb, stats = ops.rowstats_cast(x.to(dev))
    assert torch.equal(xb.cpu().float(), bf(x))
    st = stats.cpu()
    assert torch.allclose(st[:, 0], x.sum(1), atol=1e-3, rtol=1e-5)
    assert torch.allclose(st[:, 1], (x * x).sum(1), atol=1e-3, rtol=1e-5)
    assert torch.all(st[:, 2:] == 0)


@pytest.mark.parametrize("M,N,act", [(300, 1152, 0), (5000, 1536, 1), (11, 1152, 0), (197 * 160, 1536, 1)])
def test_gemm_layernorm_folded_consumer(dev, M, N, act):
    """qkv / fc1 with norm1 / norm2 folded in: raw bf16 rows in, mean / rstd applied in the epilogue."""
    from sais_b200 import ops
    x = rnd(M, 384, seed=M) * 1.7 + 0.4
    x[:, 5] *= 8.0  # an outlier channel, as trained ViTs have
    gamma, beta = 1 + 0.2 * rnd(384, seed=1), 0.2 * rnd(384, seed=2)
    w, b = rnd(N, 384, seed=3, std=1 / math.sqrt(384)), rnd(N, seed=4, std=0.3)
    wg, c, d = ops.fold_layernorm(gamma, beta, w, b)
    xb, stats = ops.rowstats_cast(x.to(dev))
    out = ops.gemm_bias_act(xb, wg.to(dev), d.to(dev), act=act, ln_stats_in=stats, ln_colsum=c.to(dev), ln_eps=1e-6)
    got = out.cpu().float()
    # (a) the algebra, on exactly the operands the kernel sees: rstd (bf16(x) W'^T - mean c) + d
    mean, var = x.double().mean(1, keepdim=True), x.double().var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-6).rsqrt()
    y = rstd * (bf(x).double() @ wg.double().t() - mean * c.double()) + d.double()
    if act == 1:
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2)))
    assert torch.allclose(got, y.float(), atol=2e-3, rtol=2 ** -8), (got - y.float()).abs().max()
    # (b) against the reference formulation LayerNorm -> Linear in fp32: differs by the bf16 rounding of x and W'
    ref = O.layer_norm(x, gamma, beta, 1e-6) @ w.t() + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    assert torch.allclose(got, ref, atol=3e-2, rtol=2e-2), (got - ref).abs().max()
    # same relative error as the unfolded bf16 path (bf16(LN(x)) @ bf16(W)): ~0.25 % in L2 (+ bf16 output rounding)
    assert float((got.double() - ref.double()).norm() / ref.double().norm()) < 5e-3


@pytest.mark.parametrize("M,K", [(300, 384), (777, 1536), (197 * 96, 384), (256 * 80 - 57, 1536)])
def test_gemm_layernorm_producer(dev, M, K):
    """proj / fc2: fp32 residual update + bf16 copy + per-row (sum, sumsq) partials in four slots."""
    from sais_b200 import ops
    a, w, bias = rnd(M, K, seed=M), rnd(384, K, seed=K, std=1 / math.sqrt(K)), rnd(384, seed=7, std=0.5)
    x = rnd(M, 384, seed=9) * 2 + 0.3
    xd = x.to(dev).clone()
    stats = torch.full((M, 8), float("nan"), device=dev)
    xb = torch.empty((M, 384), device=dev, dtype=torch.bfloat16)
    ops.gemm_bias_act(a.to(dev).bfloat16(), w.to(dev).bfloat16(), bias.to(dev), residual=xd, out=xd,
                      ln_stats_out=stats, out2=xb)
    got = xd.cpu()
    assert torch.allclose(got, _gemm_ref(a, w, bias, 0, x), atol=2e-4, rtol=1e-4)
    assert torch.equal(xb.cpu().float(), bf(got))
    st = stats.cpu().view(M, 4, 2).sum(1)
    assert torch.allclose(st[:, 0], got.sum(1), atol=2e-3, rtol=1e-5)
    assert torch.allclose(st[:, 1], (got * got).sum(1), atol=2e-3, rtol=1e-5)


@pytest.mark.parametrize("B,scale", [(1, 1.0), (5, 3.0)])
def test_vit_cls_attention(dev, B, scale):
    """last-block shortcut: attention output of the CLS query row only == row 0 of the full attention."""
    from sais_b200 import ops
    qkv = rnd(B * 197, 1152, seed=B + 40)
    qkv[:, :768] *= scale
    got = ops.vit_cls_attention(qkv.to(dev).bfloat16(), B).cpu().float()
    ref_o, _ = _vit_attn_ref(qkv, B)
    ref = ref_o.view(B, 197, 384)[:, 0]
    assert torch.allclose(got, ref, atol=2e-3 * scale, rtol=2 ** -7), (got - ref).abs().max()
    full, _ = ops.vit_attention(qkv.to(dev).bfloat16(), B)
    full0 = full.cpu().float().view(B, 197, 384)[:, 0]
    assert (got - full0).abs().max() <= 2e-2 * scale  # the tcgen05 kernel rounds P to bf16 before PV; this one keeps fp32


# ------------------------------------------------------------------------------------------------ fused MLP
def _mlp_ref(xn, w1, b1, w2, b2, x):
    """fc1 -> erf-GELU -> (hidden rounded to bf16, as the kernel hands it to the second MMA) -> fc2 -> + residual."""
    h = bf(xn).double() @ bf(w1).double().t() + b1.double()
    h = 0.5 * h * (1 + torch.erf(h / math.sqrt(2)))
    h = h.float().to(torch.bfloat16).double()
    return (x.double() + h @ bf(w2).double().t() + b2.double()).float()


# rows: single partial tile; ragged; one pair tile; several tiles (tail split along the hidden dim on 74 pairs);
# exactly 74 tiles (no split); 80 tiles (one full round + a split tail)
@pytest.mark.parametrize("rows", [1, 130, 256, 1000, 197 * 96, 256 * 80 - 57])
def test_vit_mlp_fused(dev, rows):
    from sais_b200 import ops
    xn = rnd(rows, 384, seed=rows)
    w1, b1 = rnd(1536, 384, seed=1, std=1 / math.sqrt(384)), rnd(1536, seed=2, std=0.5)
    w2, b2 = rnd(384, 1536, seed=3, std=1 / math.sqrt(1536)), rnd(384, seed=4, std=0.5)
    x = rnd(rows, 384, seed=5)
    xd = x.to(dev).clone()
    ops.vit_mlp(xn.to(dev).bfloat16(), w1.to(dev).bfloat16(), b1.to(dev), w2.to(dev).bfloat16(), b2.to(dev), xd)
    ref = _mlp_ref(xn, w1, b1, w2, b2, x)
    got = xd.cpu()
    # fp32 accumulation over 1536 bf16-rounded hidden values of O(1): the fast erf-GELU (|err| <= 3e-5 before the
    # bf16 rounding) can flip a hidden value by one bf16 ulp (2^-8 relative) now and then -> ~1e-3 absolute
    assert torch.allclose(got, ref, atol=3e-3, rtol=1e-3), (got - ref).abs().max()
    assert (got - ref).abs().mean() < 2e-4


def test_vit_mlp_fused_matches_unfused_gemms(dev):
    """same arithmetic as the two-GEMM path (fc1+GELU -> bf16 hidden -> fc2 + residual) up to summation order."""
    from sais_b200 import _lib, ops
    rows = 3000
    xn = rnd(rows, 384, seed=11).to(dev).bfloat16()
    w1, b1 = rnd(1536, 384, seed=1, std=0.05).to(dev).bfloat16(), rnd(1536, seed=2, std=0.5).to(dev)
    w2, b2 = rnd(384, 1536, seed=3, std=0.03).to(dev).bfloat16(), rnd(384, seed=4, std=0.5).to(dev)
    x = rnd(rows, 384, seed=5).to(dev)
    hid = ops.gemm_bias_act(xn, w1, b1, act=_lib.ACT_GELU_ERF)
    ref = ops.gemm_bias_act(hid, w2, b2, residual=x, out_dtype=torch.float32)
    got = ops.vit_mlp(xn, w1, b1, w2, b2, x.clone())
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


@pytest.mark.parametrize("rows", [130, 1000, 197 * 96, 256 * 80 - 57])
def test_vit_mlp_fused_layernorm_folded(dev, rows):
    """fused MLP with norm2 folded in (sais_vit_mlp_ln): raw bf16 rows + row statistics in, against the same GEMM-pair
    arithmetic (folded fc1 + GELU -> bf16 hidden -> fc2 + residual) and against LayerNorm -> Mlp in fp32."""
    from sais_b200 import _lib, ops
    x0 = rnd(rows, 384, seed=rows) * 1.7 + 0.4
    x0[:, 5] *= 8.0  # an outlier channel, as trained ViTs have
    gamma, beta = 1 + 0.2 * rnd(384, seed=1), 0.2 * rnd(384, seed=2)
    w1, b1 = rnd(1536, 384, seed=3, std=1 / math.sqrt(384)), rnd(1536, seed=4, std=0.3)
    w2, b2 = rnd(384, 1536, seed=5, std=1 / math.sqrt(1536)), rnd(384, seed=6, std=0.5)
    wg, c, d = ops.fold_layernorm(gamma, beta, w1, b1)
    xb, stats = ops.rowstats_cast(x0.to(dev))
    w2d, b2d = w2.to(dev).bfloat16(), b2.to(dev)
    res = rnd(rows, 384, seed=7).to(dev)
    got = ops.vit_mlp_ln(xb, stats, wg.to(dev), c.to(dev), d.to(dev), w2d, b2d, res.clone())
    # (a) the two-GEMM path on the same operands
    hid = ops.gemm_bias_act(xb, wg.to(dev), d.to(dev), act=_lib.ACT_GELU_ERF, ln_stats_in=stats, ln_colsum=c.to(dev), ln_eps=1e-6)
    ref2 = ops.gemm_bias_act(hid, w2d, b2d, residual=res, out_dtype=torch.float32)
    assert torch.allclose(got, ref2, atol=1e-4, rtol=1e-4), (got - ref2).abs().max()
    # (b) LayerNorm -> fc1 -> GELU -> fc2 + residual in fp32 / fp64 (differs by the bf16 roundings of x, W', hidden)
    h = torch.nn.functional.gelu(O.layer_norm(x0, gamma, beta, 1e-6) @ w1.t() + b1)
    ref = res.cpu() + h @ w2.t() + b2
    assert float((got.cpu().double() - ref.double()).norm() / ref.double().norm()) < 5e-3


@pytest.mark.parametrize("rows", [1, 130, 256, 257, 1000, 197 * 96, 256 * 80 - 57, 197 * 256])
@pytest.mark.parametrize("alias", [False, True])
def test_vit_mlp_fused_cast_warps(dev, rows, alias):
    """sais_vit_mlp_ln with xb_out / stats_out: the residual update itself is unchanged bit for bit, and the bf16 copy +
    row statistics written by the kernel's cast warps equal sais_rowstats_cast of the UPDATED stream bit for bit — in
    separate buffers and in place over the kernel's own inputs (how sais_vit_forward uses it).  Row counts cover one unit,
    ragged last tiles, several units per CTA pair (197*256 rows = 197 tiles on 74 pairs) and the reverse walk."""
    from sais_b200 import ops
    x0 = rnd(rows, 384, seed=rows % 1000) * 1.7 + 0.4
    gamma, beta = 1 + 0.2 * rnd(384, seed=1), 0.2 * rnd(384, seed=2)
    w1, b1 = rnd(1536, 384, seed=3, std=1 / math.sqrt(384)), rnd(1536, seed=4, std=0.3)
    w2, b2 = rnd(384, 1536, seed=5, std=1 / math.sqrt(1536)), rnd(384, seed=6, std=0.5)
    wg, c, d = (t.to(dev) for t in ops.fold_layernorm(gamma, beta, w1, b1))
    w2d, b2d = w2.to(dev).bfloat16(), b2.to(dev)
    x_in = x0.to(dev)
    xb, stats = ops.rowstats_cast(x_in)
    want_x = ops.vit_mlp_ln(xb, stats, wg, c, d, w2d, b2d, x_in.clone())
    want_xb, want_stats = ops.rowstats_cast(want_x)
    for rep in range(2):  # twice: barrier phases / leftovers of a previous launch must not matter
        xg = x_in.clone()
        if alias:
            xb_io, st_io = xb.clone(), stats.clone()
            ops.vit_mlp_ln(xb_io, st_io, wg, c, d, w2d, b2d, xg, xb_out=xb_io, stats_out=st_io)
        else:
            xb_io = torch.full_like(xb, float("nan"))
            st_io = torch.full_like(stats, float("nan"))
            ops.vit_mlp_ln(xb, stats, wg, c, d, w2d, b2d, xg, xb_out=xb_io, stats_out=st_io)
        assert torch.equal(xg, want_x), rep
        assert torch.equal(xb_io.view(torch.int16), want_xb.view(torch.int16)), rep
        assert torch.equal(st_io, want_stats), rep
    with pytest.raises(ValueError):
        ops.vit_mlp_ln(xb, stats, wg, c, d, w2d, b2d, x_in.clone(), xb_out=xb.clone())


# ------------------------------------------------------------------------------------------------ ViT attention
def _vit_attn_ref(qkv, B):
    q, k, v = bf(qkv).view(B, 197, 3, 6, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-2, -1)) * 0.125, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * 197, 384), p


@pytest.mark.parametrize("B,scale", [(1, 1.0), (3, 3.0)])
def test_vit_attention(dev, B, scale):
    from sais_b200 import ops
    qkv = rnd(B * 197, 1152, seed=B)
    qkv[:, :768] *= scale  # sharper logits
    out, probs = ops.vit_attention(qkv.to(dev).bfloat16(), B, emit_probs=True)
    ref_o, ref_p = _vit_attn_ref(qkv, B)
    # probabilities: fp32 softmax of fp32-accumulated logits
    assert torch.allclose(probs.cpu(), ref_p, atol=2e-5, rtol=2e-4), (probs.cpu() - ref_p).abs().max()
    assert torch.allclose(probs.sum(-1).cpu(), torch.ones(B, 6, 197), atol=1e-5)
    # outputs: P is rounded to bf16 before PV (2^-9 relative per term) and the result to bf16
    assert torch.allclose(out.cpu().float(), ref_o, atol=1.5e-2, rtol=2e-2), (out.cpu().float() - ref_o).abs().max()
    # without probabilities the tcgen05 kernel runs (S and O in TMEM, P through swizzled smem)
    out2, none = ops.vit_attention(qkv.to(dev).bfloat16(), B, emit_probs=False)
    assert none is None
    assert torch.allclose(out2.cpu().float(), ref_o, atol=1.5e-2, rtol=2e-2), (out2.cpu().float() - ref_o).abs().max()


@pytest.mark.parametrize("B", [7, 150])
def test_vit_attention_tc_many_items(dev, B):
    """persistent loop: more (frame, head) items than SMs, odd counts, double-buffered slots."""
    from sais_b200 import ops
    qkv = rnd(B * 197, 1152, seed=B) * 1.5
    out, _ = ops.vit_attention(qkv.to(dev).bfloat16(), B)
    ref_o, _ = _vit_attn_ref(qkv, B)
    assert torch.allclose(out.cpu().float(), ref_o, atol=2e-2, rtol=2e-2), (out.cpu().float() - ref_o).abs().max()


# ------------------------------------------------------------------------------------------------ temporal attention
def _unsplit(t):
    """bf16 [R,2C] = [hi | lo] -> fp32 [R,C]"""
    c = t.shape[1] // 2
    return t[:, :c].float() + t[:, c:].float()


def _tmp_attn_ref(qkv, lens_S, pads):
    outs, attns, t0 = [], [], 0
    for i, S in enumerate(lens_S):
        blk = qkv[t0:t0 + S].double()
        q, k, v = blk.view(S, 3, 4, 96).permute(1, 2, 0, 3)
        logit = (q @ k.transpose(-2, -1)) * 96 ** -0.5
        if pads is not None:
            logit = logit.masked_fill(pads[t0:t0 + S].bool().view(1, 1, S), float("-inf"))
        p = torch.softmax(logit, -1)
        outs.append((p @ v).permute(1, 0, 2).reshape(S, 384))
        attns.append(p.mean(0))
        t0 += S
    return torch.cat(outs), attns


@pytest.mark.parametrize("lens_S", [[11], [31] * 5, [16, 13, 10, 3, 3, 2], [65, 40], [200, 31], [48, 47, 33, 1],
                                    [80, 79, 64, 17], [32, 16, 17, 8, 9], [81, 5]])
def test_temporal_attention(dev, lens_S):
    from sais_b200 import ops
    total = sum(lens_S)
    qkv = rnd(total, 1152, seed=total) * 1.5
    g = torch.Generator().manual_seed(5)
    pads = torch.zeros(total, dtype=torch.uint8)
    t0 = 0
    for S in lens_S:  # pad a suffix of the keys, never the CLS key
        npad = int(torch.randint(0, max(1, S // 2), (1,), generator=g))
        if npad:
            pads[t0 + S - npad:t0 + S] = 1
        t0 += S
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens_S)]), dtype=torch.int32)
    # emit maps for every other sequence only
    emit = [i % 2 == 0 for i in range(len(lens_S))]
    a_offs, cur = [], 0
    for S, e in zip(lens_S, emit):
        a_offs.append(cur if e else -1)
        cur += S * S if e else 0
    out, attn = ops.temporal_attention(qkv.to(dev), offs.to(dev), pads.to(dev),
                                       torch.tensor(a_offs, dtype=torch.int64).to(dev), max(lens_S), attn_numel=cur)
    ref_o, ref_a = _tmp_attn_ref(qkv, lens_S, pads)
    got_o = _unsplit(out.cpu())  # exact fp32 attention, output carried as [hi | lo] bf16 halves (~2^-17 relative)
    assert torch.allclose(got_o, ref_o.float(), atol=2e-5, rtol=2e-5), (got_o - ref_o.float()).abs().max()
    attn = attn.cpu()
    t0 = 0
    for S, e, ao, ra in zip(lens_S, emit, a_offs, ref_a):
        if e:
            got = attn[ao:ao + S * S].view(S, S)
            assert torch.allclose(got, ra.float(), atol=2e-6, rtol=1e-4), (got - ra.float()).abs().max()
            assert float(got[:, pads[t0:t0 + S].bool()].abs().max() if pads[t0:t0 + S].any() else 0) == 0.0
        t0 += S


def test_temporal_attention_no_mask_no_maps(dev):
    from sais_b200 import ops
    lens_S = [31] * 3
    qkv = rnd(93, 1152, seed=1)
    offs = torch.tensor([0, 31, 62, 93], dtype=torch.int32)
    out, attn = ops.temporal_attention(qkv.to(dev), offs.to(dev), None, None, 31)
    assert attn is None
    assert torch.allclose(_unsplit(out.cpu()), _tmp_attn_ref(qkv, lens_S, None)[0].float(), atol=2e-5, rtol=2e-5)


# ------------------------------------------------------------------------------------------------ head + scoring
@pytest.mark.parametrize("B,nsnip,flow", [(1, 1, True), (5, 1, True), (6, 3, True), (4, 2, False)])
def test_clip_head(dev, B, nsnip, flow):
    from sais_b200 import ops
    a = torch.relu(rnd(B * nsnip, 384, seed=1))
    b = torch.relu(rnd(B * nsnip, 384, seed=2)) if flow else None
    W, bias = rnd(256, 384, seed=3, std=0.05), rnd(256, seed=4)
    got = ops.clip_head(a.to(dev), None if b is None else b.to(dev), B, nsnip, W.to(dev), bias.to(dev)).cpu()
    v = a.view(B, nsnip, 384).mean(1) + (b.view(B, nsnip, 384).mean(1) if flow else 0)
    ref = torch.relu(v) @ W.t() + bias
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("P", [1, 2, 3, 6, 33, 64])
def test_prototype_score(dev, P, golden_dir):
    from sais_b200 import ops
    reps = rnd(37, 256, seed=P)
    protos = O.make_prototypes(P)
    probs, sims, pred = ops.prototype_score(reps.to(dev), protos.to(dev), want_sims=True)
    rp, rs = O.prototype_probs(reps, protos)
    assert torch.allclose(sims.cpu(), rs, atol=1e-6)
    assert torch.allclose(probs.cpu(), rp, atol=1e-6)
    safe = O.top2_margin(rs) > 1e-5 if P > 1 else torch.ones(37, dtype=torch.bool)
    assert torch.equal(pred.cpu().long()[safe], rp.argmax(1)[safe])


def test_prototype_score_golden(dev, golden_dir):
    """against the reference's own calcProbs outputs (tests/golden/scoring.npz)."""
    from sais_b200 import scoring
    g = np.load(golden_dir / "scoring.npz")
    reps = torch.from_numpy(g["reps"]).to(dev)
    for P in (2, 6):
        protos = O.make_prototypes(P)
        pdict = {str(i): protos[i:i + 1] for i in range(P)}
        _, sim, probs = scoring.calcProbs(reps, pdict)
        np.testing.assert_allclose(sim.cpu().numpy(), g[f"sim_P{P}"], atol=1e-6)
        np.testing.assert_allclose(probs.cpu().numpy(), g[f"probs_P{P}"], atol=1e-6)
        pred, _ = scoring.predict(reps, pdict)
        assert np.array_equal(pred.cpu().numpy(), g[f"probs_P{P}"].argmax(1))


# ------------------------------------------------------------------------------------------------ split precision
def test_split_outputs_of_layernorm_and_patchify(dev):
    """[hi | lo] outputs: hi is the plain bf16 result, hi + lo reproduces the fp32 value to ~2^-17."""
    from sais_b200 import ops
    x = rnd(301, 384, seed=5) * 2
    w, b = 1 + 0.1 * rnd(384, seed=1), 0.1 * rnd(384, seed=2)
    of, ob = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-5, out_f32=True, out_bf16=True, split_out=True)
    assert ob.shape == (301, 768)
    assert torch.equal(ob[:, :384].float(), bf(of))
    assert torch.allclose(_unsplit(ob), of, atol=1e-6, rtol=2 ** -16)
    fr = rnd(2, 3, 224, 224, seed=3)
    p1, p2 = ops.patchify_f32(fr.to(dev)), ops.patchify_f32(fr.to(dev), split_out=True)
    assert torch.equal(p2[:, :768], p1)
    assert torch.allclose(_unsplit(p2).cpu(), _patches_ref(fr), atol=1e-6, rtol=2 ** -16)
    u8 = O.make_frames_u8(2, seed=4)
    q1, q2 = ops.normalize_patchify_u8(u8.to(dev)), ops.normalize_patchify_u8(u8.to(dev), split_out=True)
    assert torch.equal(q2[:, :768], q1)
    assert torch.allclose(_unsplit(q2).cpu(), _patches_ref(O.normalize_frames(u8)), atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("M,N,K,act,use_res", [(300, 384, 384, 0, True), (1000, 1152, 384, 0, False),
                                               (257, 1536, 384, 1, False), (64, 384, 2048, 0, True),
                                               (197 * 40, 2048, 384, 2, False), (11, 1152, 384, 0, False)])
def test_gemm_split_precision(dev, M, N, K, act, use_res):
    """3-pass split-bf16 GEMM is fp32-equivalent: compare with an fp64 product of the UNROUNDED fp32 operands."""
    from sais_b200 import ops
    a = rnd(M, K, seed=M + 1)
    w = rnd(N, K, seed=K + 1, std=1 / math.sqrt(K))
    bias = rnd(N, seed=7, std=0.5)
    res = rnd(M, N, seed=9) if use_res else None
    out = ops.gemm_bias_act(ops.split_bf16(a.to(dev)), ops.split_bf16(w.to(dev)), bias.to(dev), act=act,
                            residual=None if res is None else res.to(dev), out_dtype=torch.float32, split3=True)
    y = a.double() @ w.double().t() + bias.double()
    if act == 1:
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2)))
    elif act == 2:
        y = torch.relu(y)
    if res is not None:
        y = y + res.double()
    err = (out.cpu().double() - y).abs().max().item()
    assert err < 1e-4, err  # plain bf16 operands would be ~1e-2 here
    if not use_res:  # [hi | lo] output feeds the next split GEMM
        o2 = ops.gemm_bias_act(ops.split_bf16(a.to(dev)), ops.split_bf16(w.to(dev)), bias.to(dev), act=act,
                               out_dtype=torch.bfloat16, split3=True, split_out=True)
        assert o2.shape == (M, 2 * N)
        assert torch.allclose(_unsplit(o2).cpu(), out.cpu(), atol=1e-5, rtol=2 ** -15)
