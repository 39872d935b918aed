"""GPU parity tests of the drop-in modules (through the C ABI) against (a) the CPU oracle on the same seeded inputs
and (b) the committed golden vectors produced by the unmodified reference.  Tolerances are BASELINE.json's:
embeddings cosine >= 0.999 and max|d|/max|ref| <= 1e-2 (bf16 path), attention maps within 1e-3 absolute, predicted
prototype class identical wherever the top-2 cosine margin exceeds 1e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import make_golden as MG  # noqa: E402
from oracle import sais_oracle as O  # noqa: E402

COS_MIN, REL_MAX, ATTN_ABS, MARGIN = 0.999, 1e-2, 1e-3, 1e-3
# fp32-equivalent (split-precision) mode: BASELINE.json's "<= 1e-4 in the fp32/tf32 mode"
REL_MAX_FP32 = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _vit(sd, dev, **kw):
    import sais_b200.vision_transformer as vits
    m = vits.vit_small(patch_size=16, **kw)
    m.load_state_dict(sd, strict=True)  # reference key names
    return m.to(dev).eval()


@pytest.mark.parametrize("name,style,wseed,n,iseed", MG.VIT_CASES)
def test_vit_matches_oracle_and_golden(dev, golden_dir, name, style, wseed, n, iseed):
    g = np.load(golden_dir / f"{name}.npz")
    sd = O.make_vit_weights(wseed, style)
    model = _vit(sd, dev)
    fr = O.make_frames_u8(n, iseed)
    x = O.normalize_frames(fr)
    reps = model(x.to(dev)).cpu()
    assert reps.shape == (n, 384) and reps.dtype == torch.float32
    for ref in (O.vit_forward(sd, x), torch.from_numpy(g["reps"])):
        cos, rel = O.embedding_errors(reps, ref)
        assert cos >= COS_MIN and rel <= REL_MAX, (name, cos, rel)
    # fused u8 path (ToTensor+Normalize inside the patch kernel) gives the same embeddings
    reps_u8 = model.forward_u8(fr.to(dev)).cpu()
    cos, rel = O.embedding_errors(reps_u8, torch.from_numpy(g["reps"]))
    assert cos >= COS_MIN and rel <= REL_MAX, (name, "u8", cos, rel)
    # last-block attention probabilities: the DEFAULT call of a bf16 module meets BASELINE.json's 1e-3 on every weight
    # style (get_last_selfattention runs the fp32-equivalent path unless told otherwise)
    attn = model.get_last_selfattention(x.to(dev)).cpu()
    assert attn.shape == (n, 6, 197, 197)
    assert np.abs(attn[:, :, 0, :].numpy() - g["attn_cls"]).max() <= ATTN_ABS
    assert np.abs(attn[:, :, 100, :].numpy() - g["attn_row100"]).max() <= ATTN_ABS
    assert np.abs(attn[0, 3].numpy() - g["attn_frame0_head3"]).max() <= ATTN_ABS
    assert torch.allclose(attn.sum(-1), torch.ones(n, 6, 197), atol=1e-4)
    # explicit opt-in to the bf16 fast path: rows still sum to one; its deviation is REPORTED, and
    # asserted against the 1e-3 bound only where the bf16 residual stream can meet it (the reference's own init)
    attn_b = model.get_last_selfattention(x.to(dev), precision="bf16").cpu()
    assert torch.allclose(attn_b.sum(-1), torch.ones(n, 6, 197), atol=1e-4)
    dev_b = max(np.abs(attn_b[:, :, 0, :].numpy() - g["attn_cls"]).max(),
                np.abs(attn_b[:, :, 100, :].numpy() - g["attn_row100"]).max(),
                np.abs(attn_b[0, 3].numpy() - g["attn_frame0_head3"]).max())
    print(f"[bf16 attention-map deviation] {name} ({style}): max abs {dev_b:.3e}")
    if style == "init":
        assert dev_b <= ATTN_ABS
    toks = model.get_intermediate_layers(x.to(dev), 1)[0].cpu()
    cos, rel = O.embedding_errors(toks[:, :8], torch.from_numpy(g["tokens_first8"]))
    assert cos >= COS_MIN and rel <= 2 * REL_MAX, (name, "tokens", cos, rel)


@pytest.mark.parametrize("name,style,wseed,n,iseed", MG.VIT_CASES)
def test_vit_fp32_mode_matches_oracle_and_golden(dev, golden_dir, name, style, wseed, n, iseed):
    """precision='fp32' (split-precision GEMMs + exact attention): embeddings <= 1e-4, attention maps <= 1e-3 even
    with sharp attention."""
    g = np.load(golden_dir / f"{name}.npz")
    sd = O.make_vit_weights(wseed, style)
    model = _vit(sd, dev, precision="fp32")
    fr = O.make_frames_u8(n, iseed)
    x = O.normalize_frames(fr)
    reps = model(x.to(dev)).cpu()
    for ref in (O.vit_forward(sd, x), torch.from_numpy(g["reps"])):
        cos, rel = O.embedding_errors(reps, ref)
        assert cos >= 1 - 1e-6 and rel <= REL_MAX_FP32, (name, cos, rel)
    cos, rel = O.embedding_errors(model.forward_u8(fr.to(dev)).cpu(), torch.from_numpy(g["reps"]))
    assert rel <= REL_MAX_FP32, (name, "u8", cos, rel)
    attn = model.get_last_selfattention(x.to(dev)).cpu()
    assert np.abs(attn[:, :, 0, :].numpy() - g["attn_cls"]).max() <= 1e-4
    assert np.abs(attn[:, :, 100, :].numpy() - g["attn_row100"]).max() <= 1e-4
    assert np.abs(attn[0, 3].numpy() - g["attn_frame0_head3"]).max() <= 1e-4
    # per-call override on a bf16-default module gives the same numbers
    m2 = _vit(sd, dev)
    assert torch.equal(m2(x.to(dev), precision="fp32").cpu(), reps)
    toks = model.get_intermediate_layers(x.to(dev), 1)[0].cpu()
    cos, rel = O.embedding_errors(toks[:, :8], torch.from_numpy(g["tokens_first8"]))
    assert rel <= 2 * REL_MAX_FP32, (name, "tokens", cos, rel)


def test_vit_intermediate_layers_n4(dev, golden_dir):
    """get_intermediate_layers(x, 4) — eval_linear.py's call (vision_transformer.py:225-233): four final-norm'd token tensors,
    earliest first, against the executed reference (tests/golden/vit_inter4.npz) and the oracle; chunked through the
    workspace (2 frames per chunk), bf16 and fp32-equivalent mode; the last entry equals the n = 1 call bit for bit."""
    g = np.load(golden_dir / "vit_inter4.npz")["layers_first8"]
    sd = O.make_vit_weights(0, "stress")
    x = O.normalize_frames(O.make_frames_u8(3, 1))
    ref = O.vit_intermediate_layers(sd, x, 4)
    for precision, tol in (("bf16", 2 * REL_MAX), ("fp32", 2 * REL_MAX_FP32)):
        model = _vit(sd, dev, precision=precision, chunk_frames=2)
        outs = model.get_intermediate_layers(x.to(dev), 4)
        assert len(outs) == 4 and all(tuple(t.shape) == (3, 197, 384) for t in outs)
        for j, t in enumerate(outs):
            cos, rel = O.embedding_errors(t[:, :8].cpu(), torch.from_numpy(g[j]))
            assert cos >= COS_MIN and rel <= tol, (precision, j, cos, rel)
            cos, rel = O.embedding_errors(t.cpu().reshape(-1, 384), ref[j].reshape(-1, 384))
            assert cos >= COS_MIN and rel <= tol, (precision, j, "all tokens", cos, rel)
        assert torch.equal(outs[-1], model.get_intermediate_layers(x.to(dev), 1)[0])
        assert torch.equal(model.get_intermediate_layers(x.to(dev), 12)[-4], outs[0])
    with pytest.raises(ValueError):
        model.get_intermediate_layers(x.to(dev), 13)


@pytest.mark.parametrize("B,chunk", [(1, 96), (5, 2), (9, 4), (33, 96)])
def test_vit_chunking_is_invisible(dev, B, chunk):
    """results must not depend on how the batch is chunked through the workspace (ragged last chunk included)."""
    sd = O.make_vit_weights(3, "stress")
    x = O.normalize_frames(O.make_frames_u8(B, 11)).to(dev)
    a = _vit(sd, dev, chunk_frames=chunk)(x)
    b = _vit(sd, dev, chunk_frames=256)(x)
    assert torch.equal(a, b)
    if B <= 9:
        cos, rel = O.embedding_errors(a.cpu(), O.vit_forward(sd, x.cpu()))
        assert cos >= COS_MIN and rel <= REL_MAX


def test_vit_empty_batch_and_bad_inputs(dev):
    from sais_b200 import SaisError
    model = _vit(O.make_vit_weights(0, "init"), dev)
    assert model(torch.zeros(0, 3, 224, 224, device=dev)).shape == (0, 384)
    with pytest.raises(NotImplementedError):
        model(torch.zeros(1, 3, 256, 256, device=dev))
    with pytest.raises(SaisError):
        model(torch.zeros(1, 3, 224, 224))  # CPU tensor: there is no CPU path
    with pytest.raises(SaisError):
        model.train()


def _head(sd, dev, mods):
    from sais_b200.prepare_model import fullModel
    m = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT', modalities=mods)
    own = m.state_dict()
    for k, v in sd.items():
        if k == "frame_pos_table":
            for i in range(v.shape[0]):
                own[f"frame_pos_embeddings.{i}"] = v[i:i + 1]
        else:
            own[k] = v
    m.load_state_dict(own, strict=True)  # reference key names incl. frame_pos_embeddings.{i}
    return m.to(dev).eval()


@pytest.mark.parametrize("name,style,seed,mods,B,t_rgb,t_flow,ragged", MG.HEAD_CASES)
def test_head_matches_oracle_and_golden(dev, golden_dir, name, style, seed, mods, B, t_rgb, t_flow, ragged):
    g = np.load(golden_dir / f"{name}.npz")
    sd = O.make_head_weights(seed, style)
    model = _head(sd, dev, mods)
    xs, fs, xps, fps = MG.head_inputs(B, t_rgb, t_flow, seed, ragged)
    to = lambda ts: [t.to(dev) for t in ts]
    is_list = len(t_rgb) > 1
    x_in = to(xs)
    keep = [t.clone() for t in x_in]
    if is_list:
        out, attn = model(x_in, to(fs), [None] * 3, [None] * 3, 'Prototypes', to(xps), to(fps), None)
        ref_out, ref_attn = O.full_model_forward(sd, xs, fs, xps, fps, mods)
    else:
        out, attn = model(x_in[0], to(fs)[0], None, None, 'Prototypes', to(xps)[0], to(fps)[0], None)
        ref_out, ref_attn = O.full_model_forward(sd, xs[0], fs[0], xps[0], fps[0], mods)
        out, ref_out = [out], [ref_out]
    assert all(torch.equal(a, b) for a, b in zip(x_in, keep)), "inputs must not be mutated"
    for v, (o, r) in enumerate(zip(out, ref_out)):
        assert o.shape == (B, 256)
        for ref in (r, torch.from_numpy(g[f"out{v}"])):
            cos, rel = O.embedding_errors(o.cpu(), ref)
            # the temporal head always runs the fp32-equivalent path
            assert cos >= 1 - 1e-6 and rel <= REL_MAX_FP32, (name, v, cos, rel)
    assert attn.shape == ref_attn.shape
    assert (attn.cpu() - ref_attn).abs().max() <= 1e-4
    assert np.abs(attn.cpu().numpy() - g["attn"]).max() <= 1e-4
    pad = xps[0].reshape(-1, xps[0].shape[-1])
    if pad.any():  # padded keys carry exactly zero probability
        assert float(attn.cpu()[pad.unsqueeze(1).expand_as(ref_attn)].abs().max()) == 0.0


def test_patched_encoder_contract(dev):
    """TransformerEncoder.forward(src[S,N,E], src_key_padding_mask=bool[N,S]) -> (out[S,N,E], attn[N,S,S])."""
    from sais_b200.transformer import TransformerEncoder
    sd = O.make_head_weights(4, "stress")
    enc = TransformerEncoder()
    enc.load_state_dict({k[len("transEncoderFrame."):]: v for k, v in sd.items() if k.startswith("transEncoderFrame.")})
    enc = enc.to(dev).eval()
    x, pad, _ = O.make_clip_batch(6, 12, seed=9)
    tokens = O.temporal_prepare(sd, x).reshape(6, 13, 384)
    out, attn = enc(tokens.permute(1, 0, 2).to(dev), src_key_padding_mask=pad.reshape(6, 13).to(dev))
    ref_out, ref_attn = O.temporal_encoder(sd, tokens, pad.reshape(6, 13))
    assert out.shape == (13, 6, 384) and attn.shape == (6, 13, 13)
    cos, rel = O.embedding_errors(out.permute(1, 0, 2).cpu(), ref_out)
    assert cos >= 1 - 1e-6 and rel <= 2 * REL_MAX_FP32, (cos, rel)
    assert (attn.cpu() - ref_attn).abs().max() <= 1e-4


def test_class_identity_end_to_end(dev):
    """C1-style pipeline on several clips: u8 frames -> ViT -> temporal head -> prototypes; the predicted class
    equals the oracle's on every clip whose top-2 cosine margin exceeds the stated tolerance."""
    from sais_b200 import scoring
    vsd, hsd = O.make_vit_weights(0, "stress"), O.make_head_weights(0, "stress")
    vit, head = _vit(vsd, dev), _head(hsd, dev, "RGB-Flow")
    nclips, T = 6, 10
    rgb, flow = O.make_frames_u8(nclips * T, 21), O.make_frames_u8(nclips * T, 22)
    er = vit.forward_u8(rgb.to(dev)).view(nclips, 1, T, 384)
    ef = vit.forward_u8(flow.to(dev)).view(nclips, 1, T, 384)
    pad = O.padding_mask([T] * nclips, T).to(dev)
    out, attn = head(er, ef, None, None, 'Prototypes', pad, pad, None)
    r_er = O.vit_forward(vsd, O.normalize_frames(rgb)).view(nclips, 1, T, 384)
    r_ef = O.vit_forward(vsd, O.normalize_frames(flow)).view(nclips, 1, T, 384)
    r_out, r_attn = O.full_model_forward(hsd, r_er, r_ef, pad.cpu(), pad.cpu())
    cos, rel = O.embedding_errors(out.cpu(), r_out)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    # here the head consumes bf16-path ViT embeddings (0.7% off), so its attention map moves accordingly
    assert (attn.cpu() - r_attn).abs().max() <= 5 * ATTN_ABS
    # ... and matches to 1e-3 when the ViT runs in fp32 mode
    er32 = vit.forward_u8(rgb.to(dev), precision="fp32").view(nclips, 1, T, 384)
    ef32 = vit.forward_u8(flow.to(dev), precision="fp32").view(nclips, 1, T, 384)
    out32, attn32 = head(er32, ef32, None, None, 'Prototypes', pad, pad, None)
    assert (attn32.cpu() - r_attn).abs().max() <= ATTN_ABS
    assert O.embedding_errors(out32.cpu(), r_out)[1] <= 10 * REL_MAX_FP32
    for P in (2, 6):
        # prototypes = a few reference clip vectors + noise, so classes are balanced
        protos = r_out[:P] + 0.5 * O.make_prototypes(P, seed=P)
        pred, probs = scoring.predict(out, protos.to(dev))
        r_probs, r_sim = O.prototype_probs(r_out, protos)
        safe = O.top2_margin(r_sim) > MARGIN
        assert safe.any()
        assert torch.equal(pred.cpu()[safe], r_probs.argmax(1)[safe])
        assert (probs.cpu() - r_probs).abs().max() < 5e-3


def test_loadmodel_inference_checkpoint_on_gpu(dev, tmp_path):
    """prepare_model.loadModel(..., inference=True) (reference :517-570, caller train.py:38): reads params.zip with the
    DDP 'module.' prefix and prototypes.zip holding an nn.ParameterDict, places the model on cuda:rank, and the loaded model
    reproduces the oracle on the same weights."""
    import copy
    import torch.nn as nn
    from sais_b200 import scoring
    from sais_b200.prepare_model import loadModel
    sd = O.make_head_weights(6, "stress")
    from sais_b200.prepare_model import fullModel
    m = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT')
    own = m.state_dict()
    for k, v in sd.items():
        if k == "frame_pos_table":
            for i in range(v.shape[0]):
                own[f"frame_pos_embeddings.{i}"] = v[i:i + 1]
        else:
            own[k] = v
    torch.save(copy.deepcopy({"module." + k: v for k, v in own.items()}), tmp_path / "params.zip")
    protos = nn.ParameterDict({str(i): nn.Parameter(O.make_prototypes(2, seed=9)[i:i + 1].clone()) for i in range(2)})
    torch.save(copy.deepcopy(protos), tmp_path / "prototypes.zip")
    md, opt, device = loadModel(0, 1, str(tmp_path), 'reps', 2, 'NH_02', 384, 'ViT', 'Prototypes', 0, inference=True)
    assert device == torch.device("cuda:0") and set(md) == {"model", "prototypes"}
    x, pad, _ = O.make_clip_batch(5, 9, seed=12)
    f, fpad, _ = O.make_clip_batch(5, 4, seed=13)
    out, attn = md["model"](x.to(dev), f.to(dev), None, None, 'Prototypes', pad.to(dev), fpad.to(dev), None)
    r_out, r_attn = O.full_model_forward(sd, x, f, pad, fpad)
    cos, rel = O.embedding_errors(out.cpu(), r_out)
    assert cos >= 1 - 1e-6 and rel <= REL_MAX_FP32 and float((attn.cpu() - r_attn).abs().max()) <= 1e-4
    pred, probs = scoring.predict(out, md["prototypes"])
    r_probs, _ = O.prototype_probs(r_out, O.make_prototypes(2, seed=9))
    assert float((probs.cpu() - r_probs).abs().max()) <= 1e-5


_VIT_VARIANT_SCRIPT = """
import sys, torch
sys.path.insert(0, '.')
from oracle import sais_oracle as O
import sais_b200.vision_transformer as vits
dev = torch.device('cuda:0')
m = vits.vit_small(patch_size=16)
m.load_state_dict(O.make_vit_weights(0, 'stress'))
m = m.to(dev).eval()
g = torch.Generator().manual_seed(9)
outs = []
for n in (3, 70, 256):
    fr = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev)
    outs.append(m.forward_u8(fr).cpu())
    outs.append(m.get_intermediate_layers(O.normalize_frames(fr[:2].cpu()).to(dev), 1)[0].cpu())
torch.save(outs, sys.argv[1])
"""


@pytest.mark.parametrize("knob", ["SAIS_MLP_CAST", "SAIS_SNAKE"])
def test_vit_forward_variants_bit_identical(dev, tmp_path, knob):
    """Whole-backbone A/B of two scheduling choices that must not change a single bit: the fused MLP's cast warps vs the
    stand-alone rowstats_cast pass (SAIS_MLP_CAST=0), and the snake row order vs forward walks (SAIS_SNAKE=0) — at 3, 70
    and 256 frames (one to several row tiles per CTA pair), CLS-only and all-token variants of the last block."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ, PYTHONPATH=str(root))
        env[knob] = flag
        out = tmp_path / f"{knob}{flag}.pt"
        subprocess.run([sys.executable, "-c", _VIT_VARIANT_SCRIPT, str(out)], check=True, env=env, cwd=root, timeout=600)
        res[flag] = torch.load(out)
    for a, b in zip(res["0"], res["1"]):
        assert a.shape == b.shape and torch.equal(a, b), (a - b).abs().max()


def test_mil_task_matches_oracle_and_golden(dev, golden_dir):
    """fullModel.forward(task='MIL') — SURVEY §8f row 4 "other heads": frame-level encoder per snippet, clip-level
    transEncoderClip over the snippet CLS vectors, gated-attention MIL head — against the oracle and the executed
    reference's outputs (tests/golden/mil.npz)."""
    from oracle import make_golden_mil as MM
    from sais_b200.prepare_model import fullModel
    g = np.load(golden_dir / "mil.npz")
    for name, wseed, ncls, B, ns, tr, tf, iseed in MM.MIL_CASES:
        sd = O.make_mil_weights(wseed, "stress", ncls)
        m = fullModel(data_type='reps', nclasses=ncls, domain='NH_02', rep_dim=384, encoder_type='ViT', modalities='RGB-Flow')
        own = m.state_dict()
        for k, v in sd.items():
            if k in ("frame_pos_table", "clip_pos_table"):
                pre = "frame_pos_embeddings." if k.startswith("frame") else "clip_pos_embeddings."
                for i in range(v.shape[0]):
                    own[pre + str(i)] = v[i:i + 1]
            else:
                own[k] = v
        m.load_state_dict(own, strict=True)
        m = m.to(dev).eval()
        x, f, xp, fp = MM.mil_inputs(B, ns, tr, tf, iseed)
        keep = x.clone()
        x_dev = x.to(dev)
        seq, reps, logits, attn = m(x_dev, f.to(dev), None, None, 'MIL', xp.to(dev), fp.to(dev), None)
        assert torch.equal(x_dev.cpu(), keep), "inputs must not be mutated"
        r_seq, r_reps, r_logits, r_attn = O.mil_forward(sd, x, f, xp, fp, ncls)
        assert seq.shape == (ns, B, 384) and reps.shape == (B, ns, 384) and logits.shape == (B, ncls) and len(attn) == ncls
        for got, ref, gold, tol in ((seq, r_seq, g[f"{name}_seq"], 2e-4), (reps, r_reps, g[f"{name}_reps"], 5e-4),
                                    (logits, r_logits, g[f"{name}_logits"], 5e-4)):
            scale = float(ref.abs().max())
            assert float((got.cpu() - ref).abs().max()) <= tol * max(scale, 1.0), name
            assert float(np.abs(got.cpu().numpy() - gold).max()) <= tol * max(scale, 1.0), name
        for c in range(ncls):
            assert float((attn[c].cpu() - r_attn[c]).abs().max()) <= 1e-4
            assert float(np.abs(attn[c].cpu().numpy() - g[f"{name}_attn"][c]).max()) <= 1e-4
    with pytest.raises(NotImplementedError):
        m(x.to(dev), f.to(dev), None, None, 'ClassificationHead', xp.to(dev), fp.to(dev), None)
