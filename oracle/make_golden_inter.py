"""TEST INFRASTRUCTURE ONLY — tests/golden/vit_inter4.npz: ``get_intermediate_layers(x, n=4)`` of the UNMODIFIED reference
ViT (dino-main/vision_transformer.py:225-233; the call eval_linear.py makes with n_last_blocks = 4) on the seeded 'stress'
weights / frames of oracle/sais_oracle.py, build container only.  Outputs only: the first 8 tokens of each of the four
final-norm'd token tensors."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_import  # noqa: E402
from oracle import sais_oracle as O  # noqa: E402

CASE = ("stress", 0, 3, 1)  # weight style, weight seed, n frames, input seed (= the vit_stress case of make_golden.py)


def main():
    torch.set_grad_enabled(False)
    style, wseed, n, iseed = CASE
    vits = ref_import.load_vits()
    model = vits.vit_small(patch_size=16).eval()
    model.load_state_dict(O.make_vit_weights(wseed, style), strict=True)
    x = O.normalize_frames(O.make_frames_u8(n, iseed))
    outs = model.get_intermediate_layers(x, 4)
    assert len(outs) == 4 and tuple(outs[0].shape) == (n, 197, 384)
    np.savez_compressed(ROOT / "tests" / "golden" / "vit_inter4.npz",
                        layers_first8=np.stack([O.np_f32(t[:, :8]) for t in outs]))
    print("wrote vit_inter4.npz", [float(t.std()) for t in outs])


if __name__ == "__main__":
    main()
