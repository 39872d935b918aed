"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference classes from /root/reference (build container only;
the mount does not exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may import this).

Two shims, neither touching reference files:
  1. a stub ``timm`` module — prepare_model.py:7,40 builds a pretrained ViT-B that the 'reps' path never calls;
  2. the README §1a edit ("return attn") restated for torch 2.x as subclasses swapped into ``transEncoderFrame``.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REF_ROOT = Path("/root/reference/SAIS/scripts")


def available() -> bool:
    return (REF_ROOT / "prepare_model.py").exists()


def _load(name, path, extra_sys_path=()):
    sys.dont_write_bytecode = True
    old = list(sys.path)
    sys.path[:0] = [str(p) for p in extra_sys_path]
    try:
        spec = importlib.util.spec_from_file_location(name, str(path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path[:] = old


def load_vits():
    """reference dino-main/vision_transformer.py (its only local import is utils.trunc_normal_)."""
    dino = REF_ROOT / "dino-main"
    saved = sys.modules.pop("utils", None)
    try:
        return _load("ref_vision_transformer", dino / "vision_transformer.py", extra_sys_path=[dino])
    finally:
        sys.modules.pop("utils", None)
        if saved is not None:
            sys.modules["utils"] = saved


class PatchedLayer(nn.TransformerEncoderLayer):
    """README.md:43-48 (a): the layer returns (src, attn); attn = head-averaged weights (need_weights=True)."""

    def forward(self, src, src_mask=None, src_key_padding_mask=None, is_causal=False):
        src2, attn = self.self_attn(src, src, src, attn_mask=src_mask, key_padding_mask=src_key_padding_mask,
                                    need_weights=True)
        src = self.norm1(src + self.dropout1(src2))
        src2 = self.linear2(self.dropout(self.activation(self.linear1(src))))
        src = self.norm2(src + self.dropout2(src2))
        return src, attn


class PatchedEncoder(nn.Module):
    """README.md:43-48 (b): the encoder returns (output, attn of the LAST layer)."""

    def __init__(self, stock: nn.TransformerEncoder):
        super().__init__()
        layers = []
        for l in stock.layers:
            pl = PatchedLayer(d_model=384, nhead=4)
            pl.load_state_dict(l.state_dict())
            layers.append(pl)
        self.layers = nn.ModuleList(layers)
        self.norm = stock.norm

    def forward(self, src, mask=None, src_key_padding_mask=None):
        out, attn = src, None
        for mod in self.layers:
            out, attn = mod(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
        if self.norm is not None:
            out = self.norm(out)
        return out, attn


def load_prepare_model():
    timm = types.ModuleType("timm")
    timm.create_model = lambda *a, **k: nn.Identity()
    had = sys.modules.get("timm")
    sys.modules["timm"] = timm
    try:
        return _load("ref_prepare_model", REF_ROOT / "prepare_model.py")
    finally:
        if had is None:
            sys.modules.pop("timm", None)
        else:
            sys.modules["timm"] = had


def build_full_model(pm, modalities="RGB-Flow", nclasses=2):
    """reference fullModel on CPU with the README patch applied to both encoders."""
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        m = pm.fullModel(data_type='reps', nclasses=nclasses, domain='NH_02', rep_dim=384, encoder_type='ViT',
                         modalities=modalities, freeze_encoder_params=True, self_attention=True,
                         importance_loss=False)
    m = m.to("cpu")
    m.device = torch.device("cpu")
    m.transEncoderFrame = PatchedEncoder(m.transEncoderFrame)
    m.transEncoderClip = PatchedEncoder(m.transEncoderClip)
    return m.eval()


def load_head_weights(model, sd):
    """oracle-style head weights (one pos table) -> reference fullModel parameters."""
    own = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k == "frame_pos_table":
            for i in range(v.shape[0]):
                new[f"frame_pos_embeddings.{i}"] = v[i:i + 1].clone()
        else:
            new[k] = v.clone()
    missing = [k for k in new if k not in own]
    assert not missing, missing
    own.update(new)
    model.load_state_dict(own)
    return model


def load_process_inference_results():
    argv = sys.argv
    sys.argv = ["process_inference_results.py", "-p", "/tmp"]
    try:
        return _load("ref_process_inference_results", REF_ROOT / "process_inference_results.py")
    finally:
        sys.argv = argv
