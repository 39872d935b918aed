"""``nn.TransformerEncoder`` / ``nn.TransformerEncoderLayer`` as SAIS needs them: ``forward`` returns
``(output, attn)`` where ``attn`` is the LAST layer's head-averaged attention map — i.e. the behaviour the
reference obtains by hand-editing ``site-packages/torch/nn/modules/transformer.py`` (reference README.md:43-48;
call sites ``prepare_model.py:213,464``).  With this module no edit of PyTorch is needed.

State-dict keys equal torch's (``layers.{i}.self_attn.in_proj_weight`` ... ``layers.{i}.norm2.bias``).
The modules only hold parameters; arithmetic is ``sais_temporal_forward`` in ``libsais_b200.so``
(post-norm layers, ReLU FF 2048, LayerNorm eps 1e-5, dropout inactive: inference only).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import SaisTemporalWeights, check, current_stream, lib, ptr, require_cuda

D_MODEL = 384
N_HEAD = 4
N_LAYERS = 4
D_FF = 2048


class _SelfAttnParams(nn.Module):
    """Parameter holder with nn.MultiheadAttention's names and default init."""

    def __init__(self, d_model):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=D_MODEL, nhead=N_HEAD, dim_feedforward=D_FF, dropout=0.1, **kwargs):
        super().__init__()
        if (d_model, nhead, dim_feedforward) != (D_MODEL, N_HEAD, D_FF):
            raise NotImplementedError("sais_b200 implements SAIS's layer only: d_model=384, nhead=4, FF=2048")
        self.self_attn = _SelfAttnParams(d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-5)
        self.norm2 = nn.LayerNorm(d_model, eps=1e-5)


class TransformerEncoder(nn.Module):
    """4-layer post-norm encoder over packed variable-length sequences."""

    def __init__(self, encoder_layer=None, num_layers=N_LAYERS, norm=None):
        super().__init__()
        if num_layers != N_LAYERS or norm is not None:
            raise NotImplementedError("sais_b200 implements SAIS's encoder only: 4 layers, no final norm")
        self.layers = nn.ModuleList([TransformerEncoderLayer() for _ in range(num_layers)])
        self.num_layers = num_layers
        self._packed = None
        self._packed_key = None
        self._ws = None
        self._geom = {}

    # ------------------------------------------------------------------ packing
    def _pack_key(self, extra):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(extra))

    def pack_weights(self, frame_cls, frame_pos):
        """Device buffers in kernel layout. ``frame_cls`` [384]/[1,384], ``frame_pos`` [n_pos,384] fp32."""
        key = self._pack_key((frame_cls, frame_pos))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        dev = frame_cls.device
        if dev.type != "cuda":
            raise _lib.SaisError("temporal encoder parameters must live on a CUDA device (no CPU path)")
        keep = []

        def f32(t):
            t = t.detach().to(dev, torch.float32).contiguous()
            keep.append(t)
            return t

        def bf(t):  # [N,K] fp32 -> [N,2K] bf16 = [hi | lo]: the temporal head always runs split-precision
            t = t.detach().to(dev, torch.float32).contiguous()
            hi = t.to(torch.bfloat16)
            t = torch.cat([hi, (t - hi.float()).to(torch.bfloat16)], dim=1).contiguous()
            keep.append(t)
            return t

        w = SaisTemporalWeights()
        w.frame_cls = ptr(f32(frame_cls.reshape(-1)))
        pos = f32(frame_pos.reshape(-1, D_MODEL))
        w.frame_pos = ptr(pos)
        w.n_pos = pos.shape[0]
        for i, layer in enumerate(self.layers):
            lw = w.layers[i]
            lw.in_w, lw.in_b = ptr(bf(layer.self_attn.in_proj_weight)), ptr(f32(layer.self_attn.in_proj_bias))
            lw.out_w, lw.out_b = ptr(bf(layer.self_attn.out_proj.weight)), ptr(f32(layer.self_attn.out_proj.bias))
            lw.n1_w, lw.n1_b = ptr(f32(layer.norm1.weight)), ptr(f32(layer.norm1.bias))
            lw.ff1_w, lw.ff1_b = ptr(bf(layer.linear1.weight)), ptr(f32(layer.linear1.bias))
            lw.ff2_w, lw.ff2_b = ptr(bf(layer.linear2.weight)), ptr(f32(layer.linear2.bias))
            lw.n2_w, lw.n2_b = ptr(f32(layer.norm2.weight)), ptr(f32(layer.norm2.bias))
        self._packed = (w, keep)
        self._packed_key = key
        return self._packed

    def _workspace(self, total_tokens, device):
        """One workspace per (device, CUDA stream) — see VisionTransformer._workspace."""
        need = lib().sais_temporal_workspace_bytes(total_tokens)
        if self._ws is None:
            self._ws = {}
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws, need

    # ------------------------------------------------------------------ packed entry point
    @torch.no_grad()
    def run_packed(self, x_frames, seq_lens, key_pad, emit_attn, frame_cls, frame_pos, want_tokens=False):
        """Run the encoder over packed sequences.

        x_frames : fp32 [sum(T_i), 384] frame embeddings of all sequences back to back (positional embedding and
                   CLS are added by the kernel);  seq_lens : host list of token counts S_i = T_i + 1;
        key_pad  : bool/uint8 [sum(S_i)] (True = padded key) or None;  emit_attn : host bool list, which sequences
                   get their [S_i,S_i] head-averaged last-layer map.
        Returns (cls fp32 [nseq,384] = relu(out[CLS]), tokens fp32 [sum S_i,384] or None, attn flat fp32, attn_offsets).
        """
        require_cuda(x_frames, "x_frames")
        dev = x_frames.device
        seq_lens = np.asarray(seq_lens, dtype=np.int64)
        nseq = int(seq_lens.shape[0])
        offs = np.zeros(nseq + 1, dtype=np.int64)
        np.cumsum(seq_lens, out=offs[1:])
        total = int(offs[-1])
        if total >= 2 ** 31:
            raise _lib.SaisError("too many tokens for int32 offsets")
        max_S = int(seq_lens.max()) if nseq else 1
        emit = np.asarray(emit_attn, dtype=bool)
        a_sizes = np.where(emit, seq_lens * seq_lens, 0)
        a_offs = np.full(nseq, -1, dtype=np.int64)
        a_cum = np.cumsum(a_sizes) - a_sizes
        a_offs[emit] = a_cum[emit]
        a_total = int(a_sizes.sum())

        # the offset tables depend on the batch geometry only: keep the device copies of the last few geometries
        # (a fixed-shape inference loop then issues no host->device copy per call)
        gkey = (str(dev), seq_lens.tobytes(), emit.tobytes())
        cached = self._geom.get(gkey)
        if cached is None:
            if len(self._geom) >= 16:
                self._geom.clear()
            cached = (torch.from_numpy(offs.astype(np.int32)).to(dev),
                      torch.from_numpy(a_offs).to(dev) if a_total else None)
            self._geom[gkey] = cached
        seq_offsets, attn_offsets = cached
        attn = torch.empty(a_total, device=dev, dtype=torch.float32) if a_total else None
        x_frames = x_frames.contiguous().float()
        if key_pad is not None:
            key_pad = key_pad.contiguous().to(torch.uint8)
            assert key_pad.numel() == total
        w, _ = self.pack_weights(frame_cls, frame_pos)
        ws, need = self._workspace(max(total, 1), dev)
        cls = torch.empty((nseq, D_MODEL), device=dev, dtype=torch.float32)
        toks = torch.empty((total, D_MODEL), device=dev, dtype=torch.float32) if want_tokens else None
        with torch.cuda.device(dev):
            check(lib().sais_temporal_forward(C.byref(w), ptr(x_frames), ptr(seq_offsets), ptr(key_pad),
                                              ptr(attn_offsets), nseq, total, max_S, ptr(ws), need, ptr(cls),
                                              ptr(toks), ptr(attn), current_stream()), "sais_temporal_forward")
        return cls, toks, attn, a_offs

    # ------------------------------------------------------------------ torch-compatible entry point
    @torch.no_grad()
    def forward(self, src, mask=None, src_key_padding_mask=None):
        """``(src[S,N,E], src_key_padding_mask=bool[N,S]) -> (out[S,N,E], attn[N,S,S])`` — the patched-encoder
        contract of prepare_model.py:213.  ``src`` must already contain CLS / positional embeddings (this entry
        point adds nothing), so it is routed through the packed kernel path with a zero CLS/pos table."""
        if mask is not None:
            raise NotImplementedError("attention masks other than key padding are not used by SAIS")
        if self.training:
            raise _lib.SaisError("sais_b200.TransformerEncoder is inference-only; call .eval()")
        require_cuda(src, "src")
        S, N, E = src.shape
        if E != D_MODEL:
            raise NotImplementedError("d_model must be 384")
        # Tokens arrive fully formed (token 0 may differ per sequence), so the fused prep kernel of the packed
        # path does not apply; the layers are sequenced here kernel by kernel instead.
        return self._forward_tokens(src.permute(1, 0, 2).contiguous().float(), src_key_padding_mask)

    def _forward_tokens(self, tokens_nse, key_padding_mask):
        """tokens [N,S,E] fp32 (batch-major) that already include CLS/pos -> (out[S,N,E], attn[N,S,S])."""
        from . import ops  # local import to avoid a cycle at package import time

        with torch.cuda.device(tokens_nse.device):  # launches go to the current device's stream
            return self._forward_tokens_on_device(tokens_nse, key_padding_mask)

    def _forward_tokens_on_device(self, tokens_nse, key_padding_mask):
        from . import ops

        N, S, E = tokens_nse.shape
        dev = tokens_nse.device
        zero_cls = torch.zeros(E, device=dev)
        zero_pos = torch.zeros((1, E), device=dev)
        w, _ = self.pack_weights(zero_cls, zero_pos)
        x = tokens_nse.reshape(N * S, E).contiguous()
        xb = ops.split_bf16(x)
        offs = torch.arange(0, (N + 1) * S, S, device=dev, dtype=torch.int32)
        a_offs = torch.arange(0, N, device=dev, dtype=torch.int64) * (S * S)
        pad = None
        if key_padding_mask is not None:
            pad = key_padding_mask.reshape(N * S).to(torch.uint8).contiguous()
        attn = None
        for li, layer in enumerate(self.layers):
            lw = w.layers[li]
            last = li == len(self.layers) - 1
            qkv = _gemm_raw(xb, lw.in_w, lw.in_b, 3 * E, E, out_dtype=torch.float32)
            ao, a = ops.temporal_attention(qkv, offs, pad, a_offs if last else None, S,
                                           attn_numel=N * S * S if last else 0)
            if last:
                attn = a.view(N, S, S)
            y = _gemm_raw(ao, lw.out_w, lw.out_b, E, E, out_dtype=torch.float32, residual=x)
            x, xb = _ln_raw(y, lw.n1_w, lw.n1_b, 1e-5)
            h = _gemm_raw(xb, lw.ff1_w, lw.ff1_b, D_FF, E, out_dtype=torch.bfloat16, act=_lib.ACT_RELU,
                          split_out=True)
            y = _gemm_raw(h, lw.ff2_w, lw.ff2_b, E, D_FF, out_dtype=torch.float32, residual=x)
            x, xb = _ln_raw(y, lw.n2_w, lw.n2_b, 1e-5)
        return x.view(N, S, E).permute(1, 0, 2).contiguous(), attn

    def train(self, mode=True):
        if mode:
            raise _lib.SaisError("sais_b200.TransformerEncoder is inference-only")
        return super().train(False)


def _gemm_raw(a, w_ptr, b_ptr, N, K, out_dtype, residual=None, act=_lib.ACT_NONE, split_out=False):
    """Split-precision GEMM against a packed [N,2K] weight given by raw pointer (per-layer torch-compatible path)."""
    g = _lib.SaisGemmArgs()
    M = a.shape[0]
    out = torch.empty((M, N * (2 if split_out else 1)), device=a.device, dtype=out_dtype)
    g.a, g.w, g.bias = ptr(a), w_ptr, b_ptr
    g.residual = ptr(residual)
    if out_dtype == torch.float32:
        g.out_f32, g.ldo32 = ptr(out), out.stride(0)
    else:
        g.out_bf16, g.ldo16 = ptr(out), out.stride(0)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw = a.stride(0), 2 * K
    g.split3, g.split_out = 1, int(split_out)
    g.ldr = residual.stride(0) if residual is not None else 0
    g.act = act
    check(lib().sais_gemm_bias_act(C.byref(g), current_stream()), "sais_gemm_bias_act")
    return out


def _ln_raw(y, w_ptr, b_ptr, eps):
    rows, cols = y.shape
    of = torch.empty_like(y)
    ob = torch.empty((rows, 2 * cols), device=y.device, dtype=torch.bfloat16)
    check(lib().sais_layernorm(ptr(y), y.stride(0), w_ptr, b_ptr, float(eps), rows, cols, ptr(of), ptr(ob), 1,
                               current_stream()), "sais_layernorm")
    return of, ob
