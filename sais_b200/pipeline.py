"""Batch loops and sharding around the two models — the callers of the hot path.

* :func:`extract_features` mirrors ``extractFeatures`` (``SAIS/scripts/extract_representations.py:351-378``): walk the
  frames in batches, ``reps = model(inputs)``, collect ``[n_frames,384]``.  Here the frames are raw ``uint8`` HWC
  images in (pinned) host memory; the next batch's host->device copy runs on a side stream while the current batch
  is in the ViT (the reference's DataLoader is ``num_workers=0``, :178).
* :func:`frame_range` / :func:`gather_embeddings`: data-parallel by contiguous frame range, one process per GPU;
  the only exchange on the path is an NCCL all-gather of the per-rank ``[n/R,384]`` embeddings ahead of the
  temporal encoder.  ``gloo`` works too (CPU tensors) for the host-logic tests.
* :func:`sliding_windows` / :func:`gather_windows`: window + TTA index arithmetic of the inference datasets
  (``prepare_dataset.py:1711-1726, 2642-2651``) done with tensor ops so the gather stays on the device.
* :class:`SaisPipeline`: frames -> ViT -> windows -> temporal head -> prototype scores, the whole path.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import scoring


# --------------------------------------------------------------------------------------------- sharding
def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced frame range ``[lo, hi)`` owned by ``rank`` (first ``n % world`` ranks get one more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_items(n_items: int, rank: int, world: int) -> np.ndarray:
    """Round-robin item (clip / window) ownership: ids ``rank, rank+world, ...``."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def gather_embeddings(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """All-gather per-rank embeddings ``[hi-lo, D]`` (ranges from :func:`frame_range`) into ``[n_frames, D]`` on every
    rank.  Uneven ranges are padded to the largest one for the collective and trimmed afterwards."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_frames:
            raise ValueError("single-rank gather expects all frames locally")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [frame_range(n_frames, r, world) for r in range(world)]
    if local.shape[0] != sizes[rank][1] - sizes[rank][0]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank][1] - sizes[rank][0]}")
    width = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == width for lo, hi in sizes):
        out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


# --------------------------------------------------------------------------------------------- windows / TTA
def sliding_windows(n_frames: int, window: int, hop: int, tta_offsets: Sequence[int] = (0,)) -> List[np.ndarray]:
    """Frame-index matrices, one per TTA view: view ``o`` of window ``w`` covers ``[start_w + o, start_w + window)``
    (same end, later start — lengths ``window - o``; prepare_dataset.py:2646-2651).  Returns a list of int64 arrays
    ``[n_windows, window - o]``.  Windows that would run past the last frame are dropped (:1716-1720)."""
    if window <= 0 or hop <= 0:
        raise ValueError("window and hop must be positive")
    starts = np.arange(0, max(n_frames - window, -1) + 1, hop, dtype=np.int64)
    views = []
    for o in tta_offsets:
        if not 0 <= o < window:
            raise ValueError("TTA offset must lie inside the window")
        views.append(starts[:, None] + np.arange(o, window, dtype=np.int64)[None, :])
    return views


def gather_windows(embeddings: torch.Tensor, index: np.ndarray) -> torch.Tensor:
    """``embeddings [n,384]`` + ``index [W,T]`` -> ``[W,1,T,384]`` (nsnippets == 1 on the inference path)."""
    idx = torch.from_numpy(index).to(embeddings.device)
    return embeddings[idx.reshape(-1)].view(index.shape[0], 1, index.shape[1], embeddings.shape[1])


def full_mask(n: int, T: int, device) -> torch.Tensor:
    """``createPaddingMask`` (prepare_dataset.py:2798-2806) for unpadded windows: all False, ``[n,1,T+1]``."""
    return torch.zeros((n, 1, T + 1), dtype=torch.bool, device=device)


# --------------------------------------------------------------------------------------------- feature extraction
def _pin(frames):
    t = torch.from_numpy(frames) if isinstance(frames, np.ndarray) else frames
    if t.device.type == "cpu" and not t.is_pinned() and torch.cuda.is_available():
        t = t.pin_memory()
    return t


@torch.no_grad()
def extract_features(model, frames, batch_size: int = 256, device=None, precision: Optional[str] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``frames``: uint8 ``[n,224,224,3]`` (host, ideally pinned; or already on the device).  Returns fp32
    ``[n,384]`` on the device.  Host batches are double-buffered: copy of batch i+1 overlaps compute of batch i."""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    frames = _pin(frames)
    n = frames.shape[0]
    if out is None:
        out = torch.empty((n, 384), dtype=torch.float32, device=device)
    if n == 0:
        return out
    if frames.device.type == "cuda":
        for lo in range(0, n, batch_size):
            out[lo:lo + batch_size] = model.forward_u8(frames[lo:lo + batch_size], precision=precision)
        return out
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    bufs = [torch.empty((min(batch_size, n), 224, 224, 3), dtype=torch.uint8, device=device) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    starts = list(range(0, n, batch_size))

    def issue(i):
        lo = starts[i]
        hi = min(lo + batch_size, n)
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(freed[i % 2])  # the forward that read this buffer two batches ago
            bufs[i % 2][: hi - lo].copy_(frames[lo:hi], non_blocking=True)
            ready[i % 2].record(copy_stream)

    issue(0)
    for i, lo in enumerate(starts):
        hi = min(lo + batch_size, n)
        if i + 1 < len(starts):
            issue(i + 1)
        main.wait_event(ready[i % 2])
        out[lo:hi] = model.forward_u8(bufs[i % 2][: hi - lo], precision=precision)
        freed[i % 2].record(main)
    return out


# --------------------------------------------------------------------------------------------- whole path
class SaisPipeline:
    """frames (RGB + flow) -> per-frame embeddings -> windows (+TTA) -> temporal head -> prototype probabilities.

    Mirrors the chain ``main.sh:21,24,27,30`` of the reference without the HDF5 / pickle hand-offs.  With
    ``torch.distributed`` initialised, frames are sharded by contiguous frame range, embeddings are all-gathered,
    and windows are sharded round-robin; :meth:`run_video` then returns this rank's windows only."""

    def __init__(self, vit, head, prototypes, window: int = 15, hop: int = 15, tta_offsets: Sequence[int] = (0, 3, 6),
                 flow_stride: int = 1, batch_size: int = 256):
        self.vit, self.head = vit, head
        self.prototypes = scoring.stack_prototypes(prototypes)
        self.window, self.hop, self.tta = window, hop, tuple(tta_offsets)
        self.flow_stride = flow_stride
        self.batch_size = batch_size

    @torch.no_grad()
    def embed(self, frames, rank: int = 0, world: int = 1, precision=None) -> torch.Tensor:
        n = frames.shape[0]
        lo, hi = frame_range(n, rank, world)
        local = extract_features(self.vit, frames[lo:hi], self.batch_size, precision=precision)
        return gather_embeddings(local, n) if world > 1 else local

    @torch.no_grad()
    def score_windows(self, rgb_emb: torch.Tensor, flow_emb: torch.Tensor, window_ids: Optional[np.ndarray] = None):
        """Returns ``(pred [W], probs [W,P], attn [W,S,S], window_ids)`` for the requested windows."""
        views = sliding_windows(rgb_emb.shape[0], self.window, self.hop, self.tta)
        fviews = sliding_windows(flow_emb.shape[0], self.window, self.hop, self.tta)
        nw = min(views[0].shape[0], fviews[0].shape[0])
        ids = np.arange(nw, dtype=np.int64) if window_ids is None else np.asarray(window_ids, dtype=np.int64)
        dev = rgb_emb.device
        xs = [gather_windows(rgb_emb, v[ids]) for v in views]
        fs = [gather_windows(flow_emb, v[ids]) for v in fviews]
        xp = [full_mask(len(ids), x.shape[2], dev) for x in xs]
        fp = [full_mask(len(ids), f.shape[2], dev) for f in fs]
        if len(xs) == 1:
            out, attn = self.head(xs[0], fs[0], None, None, 'Prototypes', xp[0], fp[0], None)
        else:
            none = [None] * len(xs)
            out, attn = self.head(xs, fs, none, none, 'Prototypes', xp, fp, None)
        pred, probs = scoring.predict(out, self.prototypes.to(dev))
        return pred, probs, attn, ids

    @torch.no_grad()
    def run_video(self, rgb_frames, flow_frames, precision=None):
        import torch.distributed as dist

        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        er = self.embed(rgb_frames, rank, world, precision)
        ef = self.embed(flow_frames, rank, world, precision)
        nw = min(sliding_windows(er.shape[0], self.window, self.hop)[0].shape[0],
                 sliding_windows(ef.shape[0], self.window, self.hop)[0].shape[0])
        return self.score_windows(er, ef, shard_items(nw, rank, world))
