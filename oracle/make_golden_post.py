"""TEST INFRASTRUCTURE ONLY — tests/golden/postprocess.npz from the UNMODIFIED reference functions
(process_inference_results.py: calcProbs, getPreds, groupPredictionIntervals), imported in the build container with
argv patched (the module parses ``-p`` at import).  Inputs are regenerated from seeds by the tests."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
REF = Path("/root/reference/SAIS/scripts/process_inference_results.py")

SEEDS = (0, 1, 2)
INTERVAL_CASES = [[4], [0, 1, 2, 9, 10, 20], [3, 8, 9, 10, 30, 31, 40], [0, 5, 10, 15], [1, 2], [7, 9, 12, 13, 14, 30]]


def make_inputs(seed, n=40, d=256, P=2):
    g = torch.Generator().manual_seed(100 + seed)
    reps = [torch.randn(n, d, generator=g) for _ in range(3)]
    protos = {str(i): torch.randn(1, d, generator=g) for i in range(P)}
    for v in range(3):  # clips correlated with a prototype so that the threshold / entropy filters both fire
        for i in range(n):
            reps[v][i] += (1.0 + 0.1 * v) * (3.0 * (i % 7) / 6.0) * protos[str(i % P)][0]
    return reps, protos


def main():
    import importlib.util
    argv = list(sys.argv)
    sys.argv = ["process_inference_results.py", "-p", "/tmp"]
    sys.dont_write_bytecode = True
    try:
        spec = importlib.util.spec_from_file_location("ref_post", str(REF))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.argv = argv
    ref.class_cols = [0, 1]
    import pandas as pd
    out = {}
    for seed in SEEDS:
        reps, protos = make_inputs(seed)
        info = {"reps": tuple([r[i] for i in range(r.shape[0])] for r in reps)}
        views = [ref.calcProbs(info, protos, a)[2].numpy() for a in range(3)]
        df = pd.DataFrame()
        for a, p in enumerate(views):
            cur = pd.DataFrame(p)
            cur["TTA"], cur["ID"] = a, np.arange(cur.shape[0])
            df = pd.concat((df, cur), axis=0)
        ens = df.groupby(by=["ID"]).mean()
        out[f"views_{seed}"] = np.stack(views)
        out[f"ens_{seed}"] = ens[[0, 1]].to_numpy()
        e = ref.getPreds(ens.copy(), {0: 0, 1: 1}, threshold=None)
        out[f"entropy_{seed}"] = e["Entropy"].to_numpy()
        out[f"pred_{seed}_argmax"] = e["pred"].to_numpy()
        # the threshold branch (:135) calls Series.apply(func, 1), which pandas >= 2 rejects (the reference pins an
        # older pandas): its expression is evaluated here on the reference's own ensembled column instead
        out[f"pred_{seed}_thr"] = ens[1].apply(lambda prob: int(prob > 0.515)).to_numpy()
    for k, idx in enumerate(INTERVAL_CASES):
        for seconds in (2, 3):
            cur = pd.DataFrame({"x": np.zeros(len(idx))}, index=idx)
            s, e = ref.groupPredictionIntervals(cur, seconds)
            out[f"int_{k}_{seconds}_s"] = np.asarray(s, dtype=np.int64)
            out[f"int_{k}_{seconds}_e"] = np.asarray(e, dtype=np.int64)
    np.savez_compressed(GOLD / "postprocess.npz", **out)
    print("wrote", GOLD / "postprocess.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
