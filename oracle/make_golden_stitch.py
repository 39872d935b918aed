"""TEST INFRASTRUCTURE ONLY — tests/golden/stitch_windows.npz: the per-view RGB / flow index arithmetic of the
reference's ``VUA_EASE_Stitch`` dataset branch (the step-/skill-assessment inference form of SURVEY.md §8f row 3),
produced by EXECUTING the reference's own statements (``SAIS/scripts/prepare_dataset.py:2280-2396``, the ``__getitem__``
branch from ``df = self.data[self.phase]`` to the third flow view) in the build container.  The module itself cannot be
imported (h5py / moviepy are missing), so the line range is read from the read-only mount at generation time, dedented and
``exec``-ed against stand-ins for ``self`` / the data frame / the HDF5 handles / ``fps_dict``.  Nothing is copied into the
repository; the fixture holds numbers only.

Stand-in embeddings carry their own row index in every component, so the gathered snippets reveal exactly which rows the
reference reads (negative rows wrap like numpy does)."""
from __future__ import annotations

import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
REF = Path("/root/reference/SAIS/scripts/prepare_dataset.py")

N_RGB, N_FLOW = 4000, 400
RACES = ["Needle Withdrawal", "Needle Handling", "Needle Driving"]
# (phase, race id or -1 for the label-free USC form, start frame, end frame, fps of the video)
CASES = []
for phase in ("val", "test", "Gronau_inference", "HMH_inference"):
    for r in range(3):
        for (s, e) in ((50, 260), (1201, 1500), (41, 131), (3000, 3651), (21, 200)):
            CASES.append((phase, r, s, e, 30 if phase == "Gronau_inference" else (24 if s == 1201 else 30)))
# ('USC_inference' is not a case: in this branch the reference defines jump_size for 'Gronau_inference' / 'HMH_inference'
# only (:2362-2367), so its own flow lookup raises NameError for any other '*inference' phase)


class _H5:
    def __init__(self, arr):
        self.arr = arr

    def get(self, name):
        return self.arr


class _Enc:
    def transform(self, x):
        return np.asarray([0])


def sample(phase, race, start, end, fps):
    import pandas as pd
    import torch

    src = REF.read_text().splitlines()
    body = textwrap.dedent("\n".join(src[2280 - 1:2396]))
    assert body.lstrip().startswith("df = self.data[self.phase]") and body.rstrip().endswith("flows3 = flows3.unsqueeze(0)") \
        and "Needle Entry Start Frame" in body, "reference moved"
    video = np.repeat(np.arange(N_RGB, dtype=np.float32)[:, None], 4, axis=1)
    flow = np.repeat(np.arange(N_FLOW, dtype=np.float32)[:, None], 4, axis=1)
    row = {"Video": "vid", "Domain": "VUA", "EASE": "Low"}
    if race < 0:
        row.update({"StartFrame": start, "EndFrame": end})
    else:
        row["RACE"] = RACES[race]
        # the three races read different column pairs (:2299-2307); give every column its own value so a wrong pick shows
        cols = {"Needle Withdrawal": ("Needle Withdrawal Start Frame", "Needle Withdrawal End Frame"),
                "Needle Handling": ("Needle Handling Start Frame", "Needle Entry Start Frame"),
                "Needle Driving": ("Needle Entry Start Frame", "Needle Withdrawal Start Frame")}[RACES[race]]
        for c in ("Needle Withdrawal Start Frame", "Needle Withdrawal End Frame", "Needle Handling Start Frame",
                  "Needle Entry Start Frame"):
            row[c] = -12345
        row[cols[0]], row[cols[1]] = start, end
    df = pd.DataFrame([row])

    class _Self:
        pass

    s = _Self()
    s.phase, s.data, s.hf_rgb, s.hf_of, s.label_encoder, s.domain = phase, {phase: df}, _H5(video), _H5(flow), _Enc(), "VUA"
    ns = {"np": np, "torch": torch, "self": s, "idx": 0, "fps_dict": {"vid": fps, "VUA_HMH": {"vid": fps}}}
    exec(body, ns)
    rgb = [ns[k][0, :, 0].numpy().astype(np.int64) for k in ("snippets", "snippets2", "snippets3")]
    fl = [ns[k][0, :, 0].numpy().astype(np.int64) for k in ("flows", "flows2", "flows3")]
    return rgb, fl


def main():
    sys.dont_write_bytecode = True
    out = {"n_rgb": np.int64(N_RGB), "n_flow": np.int64(N_FLOW)}
    meta = []
    for ci, (phase, race, s, e, fps) in enumerate(CASES):
        rgb, fl = sample(phase, race, s, e, fps)
        for v in range(3):
            out[f"c{ci}_rgb{v}"] = rgb[v]
            out[f"c{ci}_flow{v}"] = fl[v]
        meta.append(f"{phase}|{race}|{s}|{e}|{fps}")
    out["cases"] = np.asarray(meta)
    np.savez_compressed(GOLD / "stitch_windows.npz", **out)
    print("wrote", GOLD / "stitch_windows.npz", len(CASES), "cases")


if __name__ == "__main__":
    main()
