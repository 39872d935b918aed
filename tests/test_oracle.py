"""CPU: pins the oracle (oracle/sais_oracle.py) against the golden vectors produced by the UNMODIFIED reference
classes (oracle/make_golden.py).  fp32 vs fp32, so tolerances are tight (summation-order noise only)."""
import numpy as np
import pytest
import torch

from oracle import make_golden as MG
from oracle import sais_oracle as O


def _load(golden_dir, name):
    return np.load(golden_dir / f"{name}.npz")


@pytest.mark.parametrize("name,style,wseed,n,iseed", MG.VIT_CASES)
def test_vit_oracle_matches_reference(golden_dir, name, style, wseed, n, iseed):
    g = _load(golden_dir, name)
    sd = O.make_vit_weights(wseed, style)
    x = O.normalize_frames(O.make_frames_u8(n, iseed))
    reps = O.vit_forward(sd, x)
    cos, rel = O.embedding_errors(reps, torch.from_numpy(g["reps"]))
    assert cos > 1 - 1e-6 and rel < 2e-5, (cos, rel)
    attn = O.vit_forward(sd, x, return_last_attention=True)
    assert attn.shape == (n, 6, 197, 197)
    np.testing.assert_allclose(attn[:, :, 0, :].numpy(), g["attn_cls"], atol=1e-5)
    np.testing.assert_allclose(attn[:, :, 100, :].numpy(), g["attn_row100"], atol=1e-5)
    np.testing.assert_allclose(attn[0, 3].numpy(), g["attn_frame0_head3"], atol=1e-5)
    toks = O.vit_forward(sd, x, return_tokens=True)
    np.testing.assert_allclose(toks[:, :8].numpy(), g["tokens_first8"], atol=5e-5, rtol=1e-4)


def test_vit_intermediate_layers_oracle_matches_reference(golden_dir):
    """get_intermediate_layers(x, 4) (vision_transformer.py:225-233, eval_linear.py's call) vs the executed reference."""
    g = _load(golden_dir, "vit_inter4")
    outs = O.vit_intermediate_layers(O.make_vit_weights(0, "stress"), O.normalize_frames(O.make_frames_u8(3, 1)), 4)
    assert len(outs) == 4
    for j, t in enumerate(outs):
        np.testing.assert_allclose(t[:, :8].numpy(), g["layers_first8"][j], atol=5e-5, rtol=1e-4)
    assert torch.equal(outs[-1], O.vit_forward(O.make_vit_weights(0, "stress"), O.normalize_frames(O.make_frames_u8(3, 1)),
                                               return_tokens=True))


@pytest.mark.parametrize("name,style,seed,mods,B,t_rgb,t_flow,ragged", MG.HEAD_CASES)
def test_head_oracle_matches_reference(golden_dir, name, style, seed, mods, B, t_rgb, t_flow, ragged):
    g = _load(golden_dir, name)
    sd = O.make_head_weights(seed, style)
    xs, fs, xps, fps = MG.head_inputs(B, t_rgb, t_flow, seed, ragged)
    if len(t_rgb) > 1:
        out, attn = O.full_model_forward(sd, xs, fs, xps, fps, mods)
    else:
        out, attn = O.full_model_forward(sd, xs[0], fs[0], xps[0], fps[0], mods)
        out = [out]
    for v, o in enumerate(out):
        np.testing.assert_allclose(o.numpy(), g[f"out{v}"], atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(attn.numpy(), g["attn"], atol=1e-5)
    # masked keys get exactly zero probability, rows sum to one (SURVEY.md §7 "exact semantics")
    pad = xps[0].reshape(-1, xps[0].shape[-1])
    assert float(attn[pad.unsqueeze(1).expand_as(attn)].abs().max() if pad.any() else 0.0) == 0.0
    np.testing.assert_allclose(attn.sum(-1).numpy(), 1.0, atol=1e-5)


def test_scoring_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, "scoring")
    reps = torch.from_numpy(g["reps"])
    for P in (2, 6):
        probs, sim = O.prototype_probs(reps, O.make_prototypes(P))
        np.testing.assert_allclose(sim.numpy(), g[f"sim_P{P}"], atol=1e-6)
        np.testing.assert_allclose(probs.numpy(), g[f"probs_P{P}"], atol=1e-6)


def test_normalize_matches_torchvision_formula():
    fr = O.make_frames_u8(2, 5)
    x = O.normalize_frames(fr)
    c = 1
    ref = (fr[..., c].float() / 255.0 - O.IMAGENET_MEAN[c]) / O.IMAGENET_STD[c]
    np.testing.assert_allclose(x[:, c].numpy(), ref.numpy(), atol=1e-6)


def test_padding_mask_contract():
    m = O.padding_mask([3, 5], 5)
    assert m.shape == (2, 1, 6)
    assert not m[:, :, 0].any()                      # CLS never padded
    assert m[0, 0].tolist() == [False] * 4 + [True] * 2 and not m[1].any()


def test_mil_oracle_matches_reference_golden(golden_dir):
    """MIL pathway (getClipReps + MIL_Head, prepare_model.py:359-363, 451-488): the restatement equals the executed
    reference (tests/golden/mil.npz, written by oracle/make_golden_mil.py) on both cases."""
    from oracle import make_golden_mil as MM
    g = np.load(golden_dir / "mil.npz")
    for name, wseed, ncls, B, ns, tr, tf, iseed in MM.MIL_CASES:
        sd = O.make_mil_weights(wseed, "stress", ncls)
        x, f, xp, fp = MM.mil_inputs(B, ns, tr, tf, iseed)
        seq, reps, logits, attn = O.mil_forward(sd, x, f, xp, fp, ncls)
        assert seq.shape == (ns, B, 384) and reps.shape == (B, ns, 384) and logits.shape == (B, ncls)
        assert np.abs(seq.numpy() - g[f"{name}_seq"]).max() <= 1e-5
        assert np.abs(reps.numpy() - g[f"{name}_reps"]).max() <= 2e-4
        assert np.abs(logits.numpy() - g[f"{name}_logits"]).max() <= 2e-4
        for c in range(ncls):
            assert np.abs(attn[c].numpy() - g[f"{name}_attn"][c]).max() <= 1e-5
            assert torch.allclose(attn[c].sum(1), torch.ones(B), atol=1e-6)
