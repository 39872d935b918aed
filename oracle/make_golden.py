"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference classes
(imported from /root/reference, build container only) on the seeded weights / inputs of oracle/sais_oracle.py.

    python oracle/make_golden.py            # rewrites tests/golden/

The fixtures hold OUTPUTS only (weights and inputs are regenerated from seeds), so they stay small.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_import  # noqa: E402
from oracle import sais_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"

# (name, weight style, weight seed, n frames, input seed)
VIT_CASES = [("vit_stress", "stress", 0, 3, 1), ("vit_init", "init", 0, 2, 2)]
# (name, style, seed, modalities, B, T_rgb list per view, T_flow list per view, ragged)
HEAD_CASES = [
    ("head_rgbflow", "stress", 0, "RGB-Flow", 3, [10], [4], True),
    ("head_tta", "stress", 1, "RGB-Flow", 2, [15, 12, 9], [2, 2, 1], True),
    ("head_rgb_only", "init", 2, "RGB", 4, [30], [30], False),
    ("head_c1", "stress", 3, "RGB-Flow", 1, [10], [10], False),
]


def head_inputs(B, t_rgb, t_flow, seed, ragged):
    xs, fs, xps, fps = [], [], [], []
    for v, (tr, tf) in enumerate(zip(t_rgb, t_flow)):
        x, xp, _ = O.make_clip_batch(B, tr, seed=10 * seed + v, ragged=ragged)
        f, fp, _ = O.make_clip_batch(B, tf, seed=10 * seed + v + 5, ragged=ragged)
        xs.append(x), fs.append(f), xps.append(xp), fps.append(fp)
    return xs, fs, xps, fps


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    GOLD.mkdir(parents=True, exist_ok=True)
    vits = ref_import.load_vits()
    for name, style, wseed, n, iseed in VIT_CASES:
        sd = O.make_vit_weights(wseed, style)
        model = vits.vit_small(patch_size=16).eval()
        model.load_state_dict(sd, strict=True)
        x = O.normalize_frames(O.make_frames_u8(n, iseed))
        reps = model(x)
        attn = model.get_last_selfattention(x)
        toks = model.get_intermediate_layers(x, 1)[0]
        np.savez_compressed(GOLD / f"{name}.npz", reps=O.np_f32(reps), attn_cls=O.np_f32(attn[:, :, 0, :]),
                            attn_row100=O.np_f32(attn[:, :, 100, :]), attn_frame0_head3=O.np_f32(attn[0, 3]),
                            tokens_first8=O.np_f32(toks[:, :8]))
        print(name, "reps", tuple(reps.shape), "std", float(reps.std()), "attn max", float(attn.max()))

    pm = ref_import.load_prepare_model()
    for name, style, seed, mods, B, t_rgb, t_flow, ragged in HEAD_CASES:
        sd = O.make_head_weights(seed, style)
        model = ref_import.load_head_weights(ref_import.build_full_model(pm, mods), sd)
        xs, fs, xps, fps = head_inputs(B, t_rgb, t_flow, seed, ragged)
        is_list = len(t_rgb) > 1
        clone = lambda ts: [t.clone() for t in ts]  # the reference mutates its inputs in place (:192)
        if is_list:
            nv = len(xs)
            out, attn = model(clone(xs), clone(fs), [None] * nv, [None] * nv, 'Prototypes', xps, fps, None)
        else:
            out, attn = model(xs[0].clone(), fs[0].clone(), None, None, 'Prototypes', xps[0], fps[0], None)
        outs = out if is_list else [out]
        save = {f"out{v}": O.np_f32(o) for v, o in enumerate(outs)}
        save["attn"] = O.np_f32(attn)
        np.savez_compressed(GOLD / f"{name}.npz", **save)
        print(name, "out", tuple(outs[0].shape), "attn", tuple(attn.shape), "row-sum", float(attn.sum(-1).mean()))

    # prototype scoring through the reference's own calcProbs (process_inference_results.py:76-91)
    pir = ref_import.load_process_inference_results()
    g = torch.Generator().manual_seed(77)
    reps = torch.randn(16, 256, generator=g)
    save = {"reps": O.np_f32(reps)}
    for P in (2, 6):
        protos = O.make_prototypes(P)
        pdict = {str(i): protos[i:i + 1] for i in range(P)}
        _, sim, probs = pir.calcProbs({'reps': [[r for r in reps]]}, pdict, 0)
        save[f"sim_P{P}"], save[f"probs_P{P}"] = O.np_f32(sim), O.np_f32(probs)
    np.savez_compressed(GOLD / "scoring.npz", **save)
    print("scoring ok")


if __name__ == "__main__":
    main()
