"""TEST INFRASTRUCTURE ONLY — tests/golden/custom_gesture_windows.npz: the window list and the per-view RGB / flow
index arithmetic of the reference's ``Custom_Gestures`` / ``Custom_inference`` dataset, produced by EXECUTING the
reference's own statements (``SAIS/scripts/prepare_dataset.py:1711-1726`` — window list — and ``:2642-2695`` — the
``__getitem__`` branch) in the build container.  The module itself cannot be imported (h5py / moviepy are missing), so
the two line ranges are read from the read-only mount at generation time, dedented and ``exec``-ed against stand-ins for
``self`` / ``countdf`` / the HDF5 handles.  Nothing is copied into the repository; the fixture holds numbers only.

Stand-in embeddings carry their own row index in every component (row r = r everywhere), so the gathered snippets
reveal exactly which rows the reference reads — including the numpy wrap-around of row -1 for the first window and of
flow row -1 (``-1 // 15``)."""
from __future__ import annotations

import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
REF = Path("/root/reference/SAIS/scripts/prepare_dataset.py")

# (total RGB frames of the video, rows in the flow HDF5 dataset)
CASES = [(100, 7), (450, 30), (31, 2), (15, 1), (46, 2), (14, 1), (1800, 117)]


def _lines(lo, hi):
    src = REF.read_text().splitlines()
    return textwrap.dedent("\n".join(src[lo - 1:hi]))


def window_list(total_frames):
    """prepare_dataset.py:1711-1726 executed as written -> DataFrame with StartFrame / EndFrame per window."""
    import pandas as pd
    from tqdm import tqdm

    body = _lines(1711, 1726)
    assert body.lstrip().startswith("duration = 0.5") and "inference_df = pd.concat" in body, "reference moved"
    countdf = pd.DataFrame([["cat", "vid", total_frames]], columns=["category", "label", "count"])
    ns = {"pd": pd, "tqdm": tqdm, "countdf": countdf}
    exec(body, ns)
    return ns["inference_df"].reset_index(drop=True)


class _H5:
    def __init__(self, arr):
        self.arr = arr

    def get(self, name):
        return self.arr


def sample(row, video_reps, flow_reps):
    """prepare_dataset.py:2642-2695 executed as written for one window -> (snippets x3, flows x3) as row ids."""
    import torch

    body = _lines(2642, 2695)
    assert body.lstrip().startswith("startIdx = curr_df['StartFrame']-1") and "flows3 = flows3.unsqueeze(0)" in body, \
        "reference moved"

    class _Self:
        phase = "Custom_inference"
        hf_rgb = _H5(video_reps)
        hf_of = _H5(flow_reps)

    ns = {"np": np, "torch": torch, "self": _Self(), "curr_df": row, "videoname": "vid"}
    exec(body, ns)
    rgb = [ns[k][0, :, 0].numpy().astype(np.int64) for k in ("snippets", "snippets2", "snippets3")]
    flow = [ns[k][0, :, 0].numpy().astype(np.int64) for k in ("flows", "flows2", "flows3")]
    return rgb, flow


def main():
    sys.dont_write_bytecode = True
    out = {}
    for ci, (n_rgb, n_flow) in enumerate(CASES):
        video = np.repeat(np.arange(n_rgb, dtype=np.float32)[:, None], 4, axis=1)
        flow = np.repeat(np.arange(n_flow, dtype=np.float32)[:, None], 4, axis=1)
        df = window_list(n_rgb)
        out[f"c{ci}_start"] = df["StartFrame"].to_numpy().astype(np.int64)
        out[f"c{ci}_end"] = df["EndFrame"].to_numpy().astype(np.int64)
        rgbs, flows = [[], [], []], [[], [], []]
        for w in range(len(df)):
            rgb, fl = sample(df.iloc[w, :], video, flow)
            for v in range(3):
                rgbs[v].append(rgb[v])
                flows[v].append(fl[v])
        for v in range(3):  # RGB views have a fixed length (15 / 12 / 9); flow views are ragged (1-2 rows): flat + lengths
            out[f"c{ci}_rgb{v}"] = np.stack(rgbs[v]) if rgbs[v] else np.zeros((0, 0), dtype=np.int64)
            out[f"c{ci}_flow{v}_len"] = np.asarray([len(f) for f in flows[v]], dtype=np.int64)
            out[f"c{ci}_flow{v}_flat"] = (np.concatenate(flows[v]) if flows[v] else np.zeros(0)).astype(np.int64)
    out["cases"] = np.asarray(CASES, dtype=np.int64)
    np.savez_compressed(GOLD / "custom_gesture_windows.npz", **out)
    print("wrote", GOLD / "custom_gesture_windows.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
