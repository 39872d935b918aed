"""TEST INFRASTRUCTURE ONLY — tests/golden/mil.npz: outputs of the UNMODIFIED reference ``fullModel.forward(task='MIL')``
(prepare_model.py:246-443 with getClipReps / MIL_Head :451-488), imported from /root/reference in the build container, on
the seeded weights / inputs of oracle/sais_oracle.py.  Separate from make_golden.py so that the other fixtures stay
byte-identical."""
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
# (name, weight seed, nclasses, B, nsnippets, T_rgb, T_flow, input seed)
MIL_CASES = [("mil_a", 0, 3, 2, 5, 6, 3, 41), ("mil_b", 1, 2, 3, 2, 9, 4, 42)]


def mil_inputs(B, ns, t_rgb, t_flow, seed):
    x, xp, _ = O.make_clip_batch(B, t_rgb, seed=seed, nsnip=ns, ragged=True)
    f, fp, _ = O.make_clip_batch(B, t_flow, seed=seed + 5, nsnip=ns, ragged=True)
    return x, f, xp, fp


def load_mil_weights(model, sd):
    own = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k in ("frame_pos_table", "clip_pos_table"):
            pre = "frame_pos_embeddings." if k.startswith("frame") else "clip_pos_embeddings."
            for i in range(v.shape[0]):
                new[pre + str(i)] = v[i:i + 1].clone()
        else:
            new[k] = v.clone()
    missing = [k for k in new if k not in own]
    assert not missing, missing[:5]
    own.update(new)
    model.load_state_dict(own)
    return model


def main():
    torch.set_grad_enabled(False)
    pm = ref_import.load_prepare_model()
    save = {}
    for name, wseed, ncls, B, ns, tr, tf, iseed in MIL_CASES:
        sd = O.make_mil_weights(wseed, "stress", ncls)
        model = load_mil_weights(ref_import.build_full_model(pm, "RGB-Flow", nclasses=ncls), sd)
        x, f, xp, fp = mil_inputs(B, ns, tr, tf, iseed)
        seq, reps, logits, attn = model(x.clone(), f.clone(), None, None, 'MIL', xp, fp, None)
        save[f"{name}_seq"], save[f"{name}_reps"], save[f"{name}_logits"] = O.np_f32(seq), O.np_f32(reps), O.np_f32(logits)
        save[f"{name}_attn"] = np.stack([O.np_f32(attn[c]) for c in range(ncls)])
        print(name, "seq", tuple(seq.shape), "reps", tuple(reps.shape), "logits", logits.tolist())
    np.savez_compressed(GOLD / "mil.npz", **save)
    print("wrote", GOLD / "mil.npz")


if __name__ == "__main__":
    main()
