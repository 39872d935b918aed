"""Prototype cosine-similarity scoring — drop-in for ``calcNCEMetrics.getProbs``
(``SAIS/scripts/prepare_miscellaneous.py:97-143``) and ``calcProbs`` (``process_inference_results.py:76-91``).

``probs = exp(cos(s, p)) / sum_p exp(cos(s, p))`` with both sides L2-normalised, ``pred = argmax``; computed by
the fused warp-per-clip kernel ``sais_prototype_score``.  TTA ensembling (mean of the per-view probabilities,
``prepare_miscellaneous.py:128-136`` / ``process_inference_results.py:100-108,218``) is a device-side mean.
"""
from __future__ import annotations

import torch

from . import ops


def stack_prototypes(prototypes):
    """dict / ParameterDict ``'0'..'P-1' -> [1,256]`` (prepare_model.py:556-562) -> fp32 ``[P,256]``."""
    if isinstance(prototypes, torch.Tensor):
        return prototypes.detach().float().reshape(-1, prototypes.shape[-1])
    return torch.vstack([p.detach().float().reshape(1, -1) for p in prototypes.values()])


def calcProbs(reps, prototypes):
    """reps fp32 [B,256] (CUDA) -> ``(reps, sim [B,P], probs [B,P])`` like process_inference_results.calcProbs."""
    p = stack_prototypes(prototypes)
    probs, sims, _ = ops.prototype_score(reps, p, want_sims=True)
    return reps, sims, probs


def getProbs(snip_sequence, prototypes):
    """``probs [B,P]`` for one view (prepare_miscellaneous.py:111-126, label bookkeeping omitted)."""
    return ops.prototype_score(snip_sequence, stack_prototypes(prototypes))[0]


def predict(snip_sequence, prototypes):
    """Class prediction; ``snip_sequence`` may be a tensor or the list of TTA views returned by ``fullModel``.
    Returns ``(pred int64 [B], probs [B,P])`` with probs averaged over views."""
    p = stack_prototypes(prototypes)
    views = snip_sequence if isinstance(snip_sequence, (list, tuple)) else [snip_sequence]
    if len(views) == 1:
        probs, _, pred = ops.prototype_score(views[0], p)
        return pred.long(), probs
    probs = torch.stack([ops.prototype_score(v, p)[0] for v in views], 0).mean(0)
    return probs.argmax(1), probs
