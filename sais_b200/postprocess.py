"""Host-side steps either side of the temporal head: ragged-clip collation with the
reference's key-padding masks, persistence of the head's outputs in the reference's on-disk format, and the
post-processing of window probabilities into gesture intervals.

Everything here is index / bookkeeping arithmetic on a few hundred numbers per video — it runs on the host in the
reference too; the per-frame and per-clip arithmetic stays in ``libsais_b200.so``.  Reference line numbers are those
of ``SAIS/scripts/``.

* :func:`create_padding_mask`, :func:`pad_collate` — ``prepare_dataset.py:2797-2806, 2808-2871`` (Prototypes task,
  tensor or 3-tuple TTA form);
* :func:`save_reps_and_labels`, :func:`save_attention` — the ``torch.save`` payloads of ``train.py:91,116-119`` /
  ``perform_training.py:168-176,213-214`` that ``process_inference_results.py:96-98`` reads back;
* :func:`tta_mean`, :func:`fold_mean`, :func:`get_preds`, :func:`group_prediction_intervals`,
  :func:`gestures_for_video`, :func:`frames_to_time` — ``process_inference_results.py:96-109, 132-207, 218-249``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------------- collation
def create_padding_mask(lens: Sequence[int], nsnippets: int = 1) -> torch.Tensor:
    """``bool [B, nsnippets, max(lens)+1]``; True = padded key.  Index 0 is the frame_cls token and is never padded:
    ``mask[b, :, len_b+1:] = True`` (prepare_dataset.py:2797-2806)."""
    lens = [int(n) for n in lens]
    mask = torch.zeros((len(lens), nsnippets, (max(lens) if lens else 0) + 1), dtype=torch.bool)
    for row, n in enumerate(lens):
        mask[row, :, n + 1:] = True
    return mask


def _collate_one(clips: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    lens = [int(c.shape[1]) for c in clips]
    nsnip = max(int(c.shape[0]) for c in clips)
    B, T, D = len(clips), max(lens), int(clips[0].shape[2])
    out = torch.zeros((B, nsnip, T, D), dtype=clips[0].dtype, device=clips[0].device)
    for b, c in enumerate(clips):
        out[b, : c.shape[0], : c.shape[1]] = c  # zero padding (pad_sequence(..., padding_value=0), :2857)
    return out, create_padding_mask(lens, nsnip).to(clips[0].device), lens


def pad_collate(clips):
    """Batch of per-clip embedding tensors ``[nsnippets, T_i, D]`` -> ``(padded [B,nsnippets,Tmax,D], mask, lens)``.
    If every sample is a tuple of TTA versions, returns three lists indexed by version (the dict-of-versions of
    prepare_dataset.py:2839-2866), ready for ``fullModel.forward``'s list form."""
    if len(clips) == 0:
        raise ValueError("empty batch")
    if isinstance(clips[0], (tuple, list)):
        nver = len(clips[0])
        per = [_collate_one([c[v] for c in clips]) for v in range(nver)]
        return [p[0] for p in per], [p[1] for p in per], [p[2] for p in per]
    return _collate_one(clips)


# --------------------------------------------------------------------------------------------- persistence
def save_reps_and_labels(savepath: str, phase: str, reps_views, labels=None, videonames=None, logits=None) -> str:
    """``reps_and_labels_<phase>``: ``{'reps': list | (list_v0, list_v1, list_v2), 'labels', 'videonames', 'logits'}`` with
    one ``[256]`` CPU tensor per clip in every list (train.py:91,117 / perform_training.py:168-176,213-214) — the file
    process_inference_results.py:96 loads with ``torch.load``."""
    def as_list(t):
        t = t.detach().to("cpu", torch.float32)
        return [t[i] for i in range(t.shape[0])]

    if isinstance(reps_views, (tuple, list)):
        reps = tuple(as_list(v) for v in reps_views)
    else:
        reps = as_list(reps_views)
    n = len(reps[0]) if isinstance(reps, tuple) else len(reps)
    info = {"reps": reps,
            "labels": list(labels) if labels is not None else [torch.tensor(0)] * n,
            "videonames": list(videonames) if videonames is not None else [""] * n,
            "logits": logits if logits is not None else []}
    os.makedirs(savepath, exist_ok=True)
    path = os.path.join(savepath, f"reps_and_labels_{phase}")
    torch.save(info, path)
    return path


def save_attention(savepath: str, phase: str, attn_batches: Sequence[torch.Tensor]) -> str:
    """``attention_<phase>``: list of per-batch ``[B,S,S]`` head-averaged maps of TTA view 0 (train.py:118-119)."""
    os.makedirs(savepath, exist_ok=True)
    path = os.path.join(savepath, f"attention_{phase}")
    torch.save([a.detach().cpu() for a in attn_batches], path)
    return path


def save_h5(results_dir: str, model_type: str, reps, labels: Sequence[str], kind: str = "rgb") -> str:
    """The HDF5 hand-off between feature extraction and the head, exactly as ``saveH5`` writes it
    (extract_representations.py:389-407): one float32 dataset ``[n_frames,384]`` per video label, rows in frame order,
    in ``<model_type>_RepsAndLabels.h5`` (``_FlowRepsAndLabels.h5`` for optical-flow frames).  ``reps``: ``[n,384]``
    tensor / array, ``labels``: the per-frame video label.  Written with ``h5py`` when it is installed (a dependency of
    the reference scripts) and otherwise with the built-in writer (:mod:`sais_b200.h5lite`: the same contiguous
    version-0-superblock layout libhdf5 produces for this call, so ``h5py.File(path)[label][idx, :]`` reads it back)."""
    reps = reps.detach().cpu().numpy() if isinstance(reps, torch.Tensor) else np.asarray(reps)
    labels = np.asarray(list(labels))
    suffix = {"rgb": "%s_RepsAndLabels.h5", "flow": "%s_FlowRepsAndLabels.h5", "seg": "%s_SegRepsAndLabels.h5"}[kind]
    os.makedirs(results_dir, exist_ok=True)
    path = os.path.join(results_dir, suffix % model_type)
    per_label = {str(label): reps[np.where(labels == label)[0]].astype(np.float32) for label in np.unique(labels)}
    try:
        import h5py
    except ImportError:
        from . import h5lite
        h5lite.write(path, per_label)
        return path
    with h5py.File(path, "w") as hf:  # pragma: no cover - depends on the environment
        for label, rows in per_label.items():
            hf.create_dataset(label, data=rows)
    return path


def load_h5(path: str) -> Dict[str, np.ndarray]:
    """``{video label: [n_frames,384] float32}`` as prepare_dataset.py:317-319, 2658-2667 reads it (``h5py`` when installed,
    :mod:`sais_b200.h5lite` otherwise)."""
    try:
        import h5py
    except ImportError:
        from . import h5lite
        return h5lite.read(path)
    with h5py.File(path, "r") as hf:  # pragma: no cover
        return {k: np.asarray(hf[k]) for k in hf.keys()}


# --------------------------------------------------------------------------------------------- ensembling
def tta_mean(probs_views: Sequence) -> np.ndarray:
    """Mean of the per-view probabilities, ``groupby('ID').mean()`` over the 3 TTA augments
    (process_inference_results.py:100-108, 218)."""
    return np.mean(np.stack([np.asarray(p, dtype=np.float64) for p in probs_views], 0), axis=0)


def fold_mean(probs_folds: Sequence) -> np.ndarray:
    """Mean across folds (process_inference_results.py:226)."""
    return np.mean(np.stack([np.asarray(p, dtype=np.float64) for p in probs_folds], 0), axis=0)


def get_preds(probs, threshold: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """``(entropy [W], pred [W])``: entropy = -sum p log p; pred = argmax, or ``int(p[last] > threshold)`` when a
    threshold is given (0.515 on the shipped path) — process_inference_results.py:132-139, 230."""
    p = np.asarray(probs, dtype=np.float64)
    entropy = -np.sum(p * np.log(p), axis=1)
    pred = np.argmax(p, axis=1) if threshold is None else (p[:, -1] > threshold).astype(np.int64)
    return entropy, pred


def group_prediction_intervals(indices: Sequence[int], seconds: float = 2) -> Tuple[List[int], List[int]]:
    """Merge sorted window indices into ``(starts, ends)`` of maximal runs whose consecutive members are at most ``seconds``
    apart — the behaviour of ``groupPredictionIntervals`` (process_inference_results.py:141-170), written from that
    behaviour rather than from its loop:

    * a run is broken wherever the gap to the previous index exceeds ``seconds``; every run contributes
      ``(first index, last index)``; a run of one index contributes ``(i, i)``;
    * one reference quirk is part of the contract (the golden fixture pins it): when the input has exactly TWO indices
      and they belong to the same run, the reference emits the single interval ``(last, last)`` — its "final entry"
      rule fires before the run has been extended — instead of ``(first, last)``."""
    idx = np.asarray([int(i) for i in indices], dtype=np.int64)
    if idx.size == 0:
        return [], []
    breaks = np.flatnonzero(np.diff(idx) > seconds) + 1          # positions where a new run starts
    first = np.concatenate(([0], breaks))
    last = np.concatenate((breaks - 1, [idx.size - 1]))
    starts, ends = idx[first].tolist(), idx[last].tolist()
    if idx.size == 2 and breaks.size == 0:
        starts = [ends[0]]
    return starts, ends


def frames_to_time(frame: int, fps: int = 30) -> Tuple[int, int, int]:
    """``(hour, min, sec)`` each taken modulo 60 exactly as FramesToTime does (process_inference_results.py:186-199)."""
    sec = int(frame) // fps
    mins = sec // 60
    hours = mins // 60
    return hours % 60, mins % 60, sec % 60


def gestures_for_video(probs, start_frames, end_frames, class_names: Sequence[str], threshold: float = 0.515,
                       entropy_thresh: float = 0.66, seconds: float = 3) -> List[Dict]:
    """Windows of one video (ensembled probabilities ``[W,P]``, in window order) -> gesture intervals, the body of the
    per-video loop of process_inference_results.py:230-249: threshold prediction, entropy filter, interval merge per
    predicted class, mean probability over each interval, argmax label."""
    p = np.asarray(probs, dtype=np.float64)
    entropy, pred = get_preds(p, threshold)
    out: List[Dict] = []
    for cls, name in enumerate(class_names):
        keep = [i for i in range(p.shape[0]) if pred[i] == cls and entropy[i] <= entropy_thresh]
        if not keep:
            continue
        starts, ends = group_prediction_intervals(keep, seconds)
        for s, e in zip(starts, ends):
            rows = [i for i in keep if s <= i <= e]
            mean_p = p[rows].mean(axis=0)
            ent, lab = get_preds(mean_p[None, :], None)
            out.append({"StartFrame": int(start_frames[s]), "EndFrame": int(end_frames[e]), "probs": mean_p,
                        "Entropy": float(ent[0]), "pred": class_names[int(lab[0])], "Gesture": name,
                        "StartTime": frames_to_time(start_frames[s]), "EndTime": frames_to_time(end_frames[e])})
    return out
