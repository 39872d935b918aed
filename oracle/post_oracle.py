"""ORACLE — TEST INFRASTRUCTURE ONLY.  Restatement, with the reference's own pandas idioms, of the window
post-processing in SAIS/scripts/process_inference_results.py (the file cannot be imported by tests at run time: it
parses argv at import and /root/reference does not exist on the GPU box).  Pinned against outputs of the reference
functions themselves by oracle/make_golden_post.py -> tests/golden/postprocess.npz."""
from __future__ import annotations

import numpy as np
import pandas as pd

CLASS_COLS = [0, 1]


def ensemble_tta(probs_views):
    """getResults + groupby('ID').mean() — process_inference_results.py:100-108, 218."""
    df = pd.DataFrame()
    for augment, probs in enumerate(probs_views):
        cur = pd.DataFrame(np.asarray(probs))
        cur["TTA"] = augment
        cur["ID"] = np.arange(cur.shape[0])
        df = pd.concat((df, cur), axis=0)
    return df.groupby(by=["ID"]).mean()[CLASS_COLS].to_numpy()


def get_preds(probs, threshold=None):
    """getPreds — :132-139."""
    df = pd.DataFrame(np.asarray(probs, dtype=np.float64))
    ent = df[CLASS_COLS].apply(lambda p: -np.sum(p * np.log(p)), axis=1)
    if threshold is None:
        pred = df[CLASS_COLS].apply(lambda p: np.argmax(p), 1)
    else:
        pred = df[CLASS_COLS[-1]].apply(lambda p: int(p > threshold))
    return ent.to_numpy(), pred.to_numpy()


def group_prediction_intervals(index, seconds=2):
    """groupPredictionIntervals — :141-170, driven through a DataFrame exactly like the reference."""
    cur = pd.DataFrame({"x": np.zeros(len(index))}, index=list(index))
    cum, starts, ends = 0, [], []
    if len(cur) == 1:
        starts.append(cur.index[0])
        ends.append(cur.index[0])
    start = cur.index[0]
    prev = start
    for idx, _ in cur.iloc[1:, :].iterrows():
        if idx - prev > seconds:
            starts.append(start)
            ends.append(prev)
            start = idx
            cum = 0
        if idx == cur.index[-1]:
            if cum == 0:
                starts.append(idx)
                ends.append(idx)
            else:
                starts.append(start)
                ends.append(idx)
        cum += 1
        prev = idx
    return [int(s) for s in starts], [int(e) for e in ends]
