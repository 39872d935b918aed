"""Batch loops and sharding around the two models — the callers of the hot path.

* :func:`extract_features` mirrors ``extractFeatures`` (``SAIS/scripts/extract_representations.py:351-378``): walk the
  frames in batches, ``reps = model(inputs)``, collect ``[n_frames,384]``.  Here the frames are raw ``uint8`` HWC
  images in (pinned) host memory; the next batch's host->device copy runs on a side stream while the current batch
  is in the ViT (the reference's DataLoader is ``num_workers=0``, :178).
* :func:`frame_range` / :func:`gather_embeddings` / :class:`EmbeddingGatherer`: data-parallel by contiguous frame
  range, one process per GPU; the only exchange on the path is an NCCL all-gather of the per-rank ``[n/R,384]``
  embeddings ahead of the temporal encoder — in place into a persistent buffer the ViT writes to, and
  asynchronous so that the head / next batch overlap it.  ``gloo`` works too (CPU tensors) for the host-logic tests.
* :class:`PeerGatherer`: the same exchange without a collective — symmetric-memory gather buffers, the ViT's final-LayerNorm
  kernel stores its rows straight into every GPU's buffer (NVSwitch multicast or peer stores over NVLink), one barrier.
* :func:`sliding_windows` / :func:`gather_windows`: dense window + TTA index arithmetic (step-recognition form,
  ``prepare_dataset.py:469-473, 2324``) done with tensor ops so the gather stays on the device.
* :func:`custom_gesture_windows` / :func:`custom_gesture_indices` / :func:`gather_ragged`: the ``Custom_Gestures``
  inference sampling contract of ``main.sh:27`` (``prepare_dataset.py:1711-1726, 2642-2672``), pinned to the reference's
  own statements by the fixture ``tests/golden/custom_gesture_windows.npz``.
* :func:`stitch_indices`: the ``VUA_EASE_Stitch`` form (``prepare_dataset.py:2279-2396``: per stitch sub-phase, stride-10 rows,
  TTA views that shift the whole range), pinned the same way (``tests/golden/stitch_windows.npz``).
* :class:`SaisPipeline`: frames -> ViT -> windows -> temporal head -> prototype scores, the whole path.
* :class:`CapturedStep`: a fixed-geometry step (e.g. one clip: ViT + head + scoring) captured once as a CUDA graph and
  replayed with a single launch — for the latency-bound, one-clip-at-a-time end of the path.
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


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin this process (one process per GPU) to the CPU cores NVML reports as local to ``device_index`` — so that the
    pinned host staging buffers allocated afterwards are first-touched on the GPU's own NUMA node and the host->device
    copies of eight ranks do not cross the socket interconnect.  Returns the CPU list, or ``None`` when NVML is not
    available, reports nothing usable inside this process's cpuset, or the platform has no ``sched_setaffinity``."""
    import os

    if not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


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


def low_footprint_group(ranks=None):
    """A NCCL process group for the embedding exchange that runs every collective on ONE CTA (``ncclConfig.max_ctas = 1``).

    The exchange is latency-bound (393-691 KB per rank); what matters is how many SMs the collective's kernel holds while it
    spins for its peers: this library's persistent kernels fill an SM completely (shared memory, registers, tensor memory),
    so every NCCL CTA takes an SM away from the next compute kernel for as long as the slowest rank is late.  Measured at
    N = 4 (batch 256 + head per step, three lanes): 333 k frames/s with NCCL's default channel count, 345 k with one
    channel — 0.954 -> 0.988 of 4 x the single-GPU figure.  Falls back to the default group when the backend is not NCCL."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    try:
        if dist.get_backend() != "nccl":
            return None
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = 1
        opts.config.min_ctas = 1
        return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
    except Exception:
        return None


class EmbeddingGatherer:
    """The one exchange step of the path, without allocations or copies: a persistent ``[R * width, D]``
    buffer per slot in which rank ``r`` owns rows ``[r * width, r * width + n_r)``.  The ViT's final-LayerNorm kernel writes
    this rank's embeddings straight into :meth:`own_slice` (``forward_u8(..., out=...)``), :meth:`gather_async` issues the
    IN-PLACE all-gather (send buffer = this rank's slice of the receive buffer) asynchronously on the process group's
    own stream, so whatever the caller enqueues next on the compute stream — the temporal head over the rank's own clips,
    the next batch's ViT — overlaps the collective; :meth:`buffer` / :meth:`wait_all` join it.  ``depth`` slots rotate so
    that step ``i + 1`` can write while the gather of step ``i`` is still in flight (a slot is re-used only after its
    pending gather has been joined).  Rows are split evenly (``ceil(n / R)`` per rank) unless ``frame_ranges=True``
    (then exactly as :func:`frame_range` does; ragged splits are compacted by :meth:`buffer`)."""

    def __init__(self, n_rows: int, dim: int, rank: int, world: int, device, frame_ranges: bool = False, depth: int = 2,
                 dtype=torch.float32, group=None):
        self.n, self.dim, self.rank, self.world, self.group = int(n_rows), int(dim), int(rank), int(world), group
        if frame_ranges:
            self.ranges = [frame_range(self.n, r, world) for r in range(world)]
        else:
            w = (self.n + world - 1) // world
            self.ranges = [(min(r * w, self.n), min((r + 1) * w, self.n)) for r in range(world)]
        self.width = max(hi - lo for lo, hi in self.ranges) if world else 0
        self.even = all(hi - lo == self.width for lo, hi in self.ranges)
        self.slots = [torch.zeros((world * self.width, dim), dtype=dtype, device=device) for _ in range(depth)]
        self.pending = [None] * depth

    def _join(self, k):
        if self.pending[k] is not None:
            self.pending[k].wait()  # stream-level: the current stream waits for the collective, the host does not
            self.pending[k] = None

    def own_slice(self, i: int) -> torch.Tensor:
        """This rank's rows of slot ``i % depth`` (joins that slot's previous gather first: WAR)."""
        k = i % len(self.slots)
        self._join(k)
        lo, hi = self.ranges[self.rank]
        return self.slots[k][self.rank * self.width: self.rank * self.width + (hi - lo)]

    def gather_async(self, i: int) -> None:
        if self.world == 1:
            return
        import torch.distributed as dist
        k = i % len(self.slots)
        buf = self.slots[k]
        mine = buf[self.rank * self.width: (self.rank + 1) * self.width]
        self.pending[k] = dist.all_gather_into_tensor(buf, mine, group=self.group, async_op=True)

    def buffer(self, i: int) -> torch.Tensor:
        """All ``n`` rows of slot ``i % depth`` in frame order (joins its gather)."""
        k = i % len(self.slots)
        self._join(k)
        if self.even:
            return self.slots[k][: self.n]
        return torch.cat([self.slots[k][r * self.width: r * self.width + (hi - lo)]
                          for r, (lo, hi) in enumerate(self.ranges)], 0)

    def wait_all(self) -> None:
        for k in range(len(self.slots)):
            self._join(k)


class PeerGatherer:
    """The exchange step WITHOUT a collective: the same job as :class:`EmbeddingGatherer` (every rank ends up with all
    ranks' ``[n/R,384]`` embeddings), but the bytes move inside the ViT's last kernel.

    The gather buffers live in symmetric memory (``torch.distributed._symmetric_memory``: one allocation per rank at the same
    offsets, every peer's copy mapped into this process, plus the NVSwitch multicast mapping where the box has one).
    :meth:`fanout` describes this rank's slice in the other GPUs' mappings; ``forward_u8(frames, out=own_slice(i),
    fanout=fanout(i))`` makes the final-LayerNorm kernel store each embedding row locally AND — one ``multimem.st`` per 16
    bytes to the multicast address, replicated by the switch; or one plain NVLink store per peer — into every other GPU's
    buffer while it computes them.  No NCCL kernel, no extra copy, no SM spinning for its peers during the data movement;
    :meth:`publish` then enqueues one stream-ordered barrier (a one-CTA signal / wait over the symmetric signal pads) after
    which :meth:`buffer` holds every rank's rows.

    Slot protocol (why there are ``2 * depth`` slots): step ``i`` writes slot ``i % (2 * depth)`` on EVERY GPU, so the
    peers must have finished reading that slot's previous content (step ``i - 2 * depth``).  A rank that has passed the
    barrier of step ``i - depth`` on this channel knows every peer has *enqueued past* its own barrier of step
    ``i - depth`` on that channel, i.e. past everything it enqueued for step ``i - 2 * depth`` — provided each rank enqueues
    its reads of ``buffer(j)`` on the stream it called ``publish(j)`` on (or orders them before its ``publish(j + depth)``).
    Steps ``j`` and ``j + depth`` share channel ``j % depth`` — with :class:`Lanes` that is the lane, and one barrier per step
    is the only synchronisation.  Rows are split evenly (``ceil(n / R)`` per rank)."""

    def __init__(self, n_rows: int, dim: int, rank: int, world: int, device, depth: int = 2, group=None,
                 use_multicast: bool = True, timeout_ms: int = 20000):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        if world - 1 > _lib.MAX_PEERS:
            raise ValueError(f"at most {_lib.MAX_PEERS + 1} ranks")
        self.n, self.dim, self.rank, self.world = int(n_rows), int(dim), int(rank), int(world)
        self.depth, self.nslots = int(depth), 2 * int(depth)
        self.timeout_ms = int(timeout_ms)
        self.width = (self.n + world - 1) // world
        self.ranges = [(min(r * self.width, self.n), min((r + 1) * self.width, self.n)) for r in range(world)]
        self.device = torch.device(device)
        slot_elems = world * self.width * dim
        self.slot_bytes = slot_elems * 4
        if self.slot_bytes % 16:
            raise ValueError("slot size must be a multiple of 16 bytes")
        if world == 1:  # nothing to exchange: a plain buffer, no fan-out, publish() only records an event
            self.mem = torch.zeros((self.nslots, self.width, dim), dtype=torch.float32, device=self.device)
            self.handle, self.peer_ptrs, self.multicast_ptr = None, [], 0
        else:
            self.mem = symm.empty((self.nslots, world * self.width, dim), dtype=torch.float32, device=self.device)
            self.mem.zero_()
            torch.cuda.synchronize(self.device)
            self.handle = symm.rendezvous(self.mem, group if group is not None else dist.group.WORLD)
            self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            mc = int(self.handle.multicast_ptr) if use_multicast else 0
            self.multicast_ptr = mc if mc else 0
            self.handle.barrier(channel=0, timeout_ms=self.timeout_ms)  # every rank has zeroed and mapped its buffers
        self._fan = {}
        self.published = [None] * self.nslots
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]

    @property
    def mode(self) -> str:
        return "single-rank" if self.world == 1 else ("multicast" if self.multicast_ptr else "peer-stores")

    def _slot(self, i: int) -> int:
        return i % self.nslots

    def own_slice(self, i: int) -> torch.Tensor:
        """This rank's rows of step ``i``'s slot.  The current stream first waits for this rank's barrier of step
        ``i - depth`` (same channel): once that has passed, no peer still reads what step ``i`` is about to overwrite."""
        prev = self.published[self._slot(i - self.depth)] if i >= self.depth else None
        if prev is not None:
            torch.cuda.current_stream(self.device).wait_event(prev)
        lo, hi = self.ranges[self.rank]
        return self.mem[self._slot(i)][self.rank * self.width: self.rank * self.width + (hi - lo)]

    def fanout(self, i: int):
        """``SaisFanout`` for :meth:`own_slice` ``(i)``: its address in the multicast mapping, or in every peer's mapping
        (``None`` for a single rank: ``forward_u8(..., fanout=None)`` is the plain forward)."""
        from . import _lib

        if self.world == 1:
            return None

        k = self._slot(i)
        f = self._fan.get(k)
        if f is None:
            off = k * self.slot_bytes + self.rank * self.width * self.dim * 4
            f = _lib.SaisFanout()
            if self.multicast_ptr:
                f.multicast, f.n_peers = self.multicast_ptr + off, 0
            else:
                peers = [p for r, p in enumerate(self.peer_ptrs) if r != self.rank]
                f.multicast, f.n_peers = None, len(peers)
                for j, p in enumerate(peers):
                    f.peers[j] = p + off
            self._fan[k] = f
        return f

    def publish(self, i: int) -> None:
        """Call on the stream that ran step ``i``'s forward: one barrier over channel ``i % depth``; once it has passed,
        every rank's rows of slot ``i`` are in this GPU's buffer (and every peer has this rank's)."""
        cur = torch.cuda.current_stream(self.device)
        ch = i % self.depth
        wrote = torch.cuda.Event()
        wrote.record(cur)
        st = self._streams[ch]
        st.wait_event(wrote)
        with torch.cuda.stream(st):  # the barrier spins for the peers on its own stream, not in front of the caller's next kernel
            if self.world > 1:
                self.handle.barrier(channel=1 + ch, timeout_ms=self.timeout_ms)
            done = torch.cuda.Event()
            done.record(st)
        self.published[self._slot(i)] = done

    def buffer(self, i: int) -> torch.Tensor:
        """All ``n`` rows of step ``i`` in frame order; the current stream waits for that step's barrier."""
        ev = self.published[self._slot(i)]
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        slot = self.mem[self._slot(i)]
        if all(hi - lo == self.width for lo, hi in self.ranges):
            return slot[: self.n]
        return torch.cat([slot[r * self.width: r * self.width + (hi - lo)] for r, (lo, hi) in enumerate(self.ranges)], 0)

    def wait_all(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for ev in self.published:
            if ev is not None:
                cur.wait_event(ev)


class SideStream:
    """Software pipelining of the two stages across batches: run the latency-bound stage of batch ``i`` (the temporal head
    + prototype scoring: a chain of ~35 small dependent kernels, ~0.3 ms on an idle GPU) on a second, high-priority CUDA
    stream underneath the throughput-bound stage of batch ``i + 1`` (the ViT, whose persistent kernels leave SMs idle at
    their tails — the fused MLP runs 197 row tiles on 74 CTA pairs, so 50 SMs idle for a third of every launch).

    ``run(fn, *args)`` orders ``fn`` after everything enqueued so far on the current stream and returns its result
    immediately (tensors it returns belong to the side stream); ``guard(slot)`` makes the current stream wait for the
    ``run`` that last used ``slot`` (call it before overwriting what that run reads); ``join()`` makes the current stream
    wait for all side work.  Works because the library's kernels do not trigger their dependents early (csrc/common.cuh,
    "Programmatic dependent launch"): idle SMs stay free for the other stream."""

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.done = [None] * slots

    def run(self, slot: int, fn, *args, **kw):
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            res = fn(*args, **kw)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.done[slot % len(self.done)] = done
        return res

    def mark(self, slot: int) -> None:
        """Re-record ``slot``'s completion point after more work was enqueued on the side stream for it."""
        done = torch.cuda.Event()
        done.record(self.stream)
        self.done[slot % len(self.done)] = done

    def guard(self, slot: int) -> None:
        d = self.done[slot % len(self.done)]
        if d is not None:
            torch.cuda.current_stream(self.device).wait_event(d)

    def join(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class Lanes:
    """``n`` independent pipeline lanes on one GPU: each lane is a CUDA stream, and consecutive batches go to consecutive
    lanes (``with lanes.lane(i): ...``).  Every persistent kernel of the ViT ends in a partially filled last round (the
    fused MLP: 197 row tiles on 74 CTA pairs = 2.66 rounds, the attention: 1,536 items on 148 CTAs = 10.4 rounds), and the
    temporal head is a chain of small kernels; with two lanes one batch's kernel tails and head are filled by the other
    batch's kernels — no SM idles while there is work in either lane.  Because the library's kernels release their
    dependents only at CTA exit (csrc/common.cuh, PDL), an idle SM is really free for the other lane.  Everything a lane
    touches is per stream: the models keep one workspace per CUDA stream, :class:`HostFrameStager` one staging pair per
    stream; give :class:`EmbeddingGatherer` / result buffers ``depth = n`` slots so that slot ``i % n`` belongs to lane
    ``i % n``.  Measured at batch 256 + head per step: 85.5 k -> 90.6 k frames/s on one B200 with two lanes."""

    def __init__(self, device, n: int = 2):
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(int(n), 1))]

    def __len__(self):
        return len(self.streams)

    def lane(self, i: int):
        return torch.cuda.stream(self.streams[i % len(self.streams)])

    def fork(self) -> None:
        """Every lane waits for what has been enqueued on the current stream (call once before the first batch)."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)

    def join(self) -> None:
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)


class CapturedStep:
    """One fixed-geometry step of the path as a CUDA graph: ``fn(*inputs)`` is run a few times, captured once, and every
    later call is a single ``cudaGraphLaunch`` instead of the ~85 (ViT) + ~35 (temporal head) kernel launches it contains.

    What it is for: the latency-bound end of the path — one clip at a time (BASELINE config C1: 10 + 10 frames through the
    ViT, the head and the scoring is ~120 launches of a few microseconds each, so the GPU waits for the host's launch
    calls).  Throughput-bound steps (batch 256) gain nothing: their launches are already hidden behind the kernels.

    The kernels' programmatic-dependent-launch attributes are captured as programmatic graph edges, so the replay keeps the
    same overlap as the stream version; results are the eager path's (same kernels, same order, same workspaces).
    Contract (the usual one for graphs): tensor arguments are copied into the static input tensors captured with the graph
    (same shapes / dtypes as the examples), non-tensor arguments are frozen at capture, and the returned tensors are the
    graph's static outputs — overwritten by the next call, so ``clone()`` what must survive it.  ``fn`` must be free of
    host synchronisation and of host->device copies once warm (true for ``forward_u8``, ``fullModel.forward`` and
    ``scoring.predict`` on a geometry they have seen: offset tables and packed weights are cached)."""

    def __init__(self, fn, *example_inputs, warmup: int = 2):
        tensors = [t for t in example_inputs if torch.is_tensor(t)]
        if not tensors or tensors[0].device.type != "cuda":
            raise ValueError("CapturedStep needs at least one CUDA tensor argument")
        self.device = tensors[0].device
        self.fn = fn
        self.static_in = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        cur = torch.cuda.current_stream(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.device(self.device), torch.no_grad():
            with torch.cuda.stream(self.stream):
                # warm-up on the capture stream itself: per-stream workspaces, geometry tables and per-device function
                # attributes exist before the capture starts (nothing inside it allocates outside the graph's pool)
                for _ in range(max(int(warmup), 1)):
                    fn(*self.static_in)
            cur.wait_stream(self.stream)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.static_out = fn(*self.static_in)
        self.replays = 0

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_in):
            raise ValueError(f"expected {len(self.static_in)} arguments")
        for dst, src in zip(self.static_in, inputs):
            if torch.is_tensor(dst):
                if not torch.is_tensor(src) or src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError("CapturedStep inputs must keep the captured shapes and dtypes")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_out


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


# --------------------------------------------------------------------------------------------- Custom_Gestures sampling
def custom_gesture_windows(total_frames: int, duration_frames: int = 15, hop_frames: int = 15):
    """Window list of the ``Custom_inference`` phase (prepare_dataset.py:1711-1726): 0.5 s windows with a 0.5 s hop at an
    assumed 30 fps -> ``StartFrame = n * hop``, ``EndFrame = StartFrame + duration`` for
    ``n < (total_frames - duration) // hop + 1``.  Returns ``(start_frames, end_frames)`` as int64 arrays."""
    if duration_frames <= 0 or hop_frames <= 0:
        raise ValueError("duration and hop must be positive")
    nsamples = max((int(total_frames) - duration_frames) // hop_frames + 1, 0)
    starts = np.arange(nsamples, dtype=np.int64) * hop_frames
    return starts, starts + duration_frames


def _wrap_rows(idx: np.ndarray, n: int, what: str) -> np.ndarray:
    """numpy fancy-indexing semantics of ``reps[idx, :]``: negative rows count from the end, out-of-range rows raise."""
    if idx.size and (idx.max(initial=0) >= n or idx.min(initial=0) < -n):
        raise IndexError(f"{what} index out of range for {n} rows (as the reference's reps[indices,:] would raise)")
    return np.where(idx < 0, idx + n, idx)


def custom_gesture_indices(start_frames, end_frames, n_rgb: int, n_flow: int, tta_offsets: Sequence[int] = (0, 3, 6),
                           flow_stride: int = 15):
    """Row indices the ``Custom_Gestures`` inference dataset reads for every window and TTA view
    (prepare_dataset.py:2642-2672), quirks included:

    * ``startIdx = StartFrame - 1``, ``endIdx = EndFrame - 1`` — the first window starts at row **-1**, which numpy
      wraps to the video's LAST embedding;
    * RGB rows of view ``o``: ``arange(startIdx + o, endIdx, (endIdx - startIdx) // 10)`` — 15 / 12 / 9 rows for the
      shipped 15-frame windows (same end, later start);
    * flow rows: ``unique(rgb_rows // flow_stride)`` (floor division BEFORE the wrap, so row -1 maps to flow row -1 = the
      last one), restricted to ``< n_flow`` — 0 to 2 rows per view.

    Returns ``(rgb, flow)``: ``rgb[v]`` is an int64 matrix ``[W, L_v]``; ``flow[v]`` a list of W int64 arrays (ragged)."""
    starts = np.asarray(start_frames, dtype=np.int64)
    ends = np.asarray(end_frames, dtype=np.int64)
    if starts.size and bool(np.all(ends - starts == ends[0] - starts[0])):
        # every window has the same duration (what custom_gesture_windows produces): all windows of a view at once — the
        # per-window loop below costs 0.54 s for a 60-minute 30-fps video (7,200 windows), more than the head's kernels
        return _custom_gesture_indices_uniform(starts, ends, n_rgb, n_flow, tta_offsets, flow_stride)
    rgb, flow = [], []
    for o in tta_offsets:
        rows, frows = [], []
        for s, e in zip(starts.tolist(), ends.tolist()):
            s0, e0 = s - 1, e - 1
            jump = (e0 - s0) // 10
            if jump <= 0:
                raise ValueError("windows shorter than 10 frames have no valid stride (reference: arange step 0)")
            raw = np.arange(s0 + o, e0, jump, dtype=np.int64)
            f = np.unique(raw // flow_stride)
            f = f[f < n_flow]
            rows.append(_wrap_rows(raw, n_rgb, "RGB"))
            frows.append(_wrap_rows(f, n_flow, "flow") if f.size else f)
        lens = {len(r) for r in rows}
        if len(lens) > 1:
            raise ValueError("windows of different lengths in one call")
        rgb.append(np.stack(rows) if rows else np.zeros((0, 0), dtype=np.int64))
        flow.append(frows)
    return rgb, flow


# --------------------------------------------------------------------------------------------- VUA_EASE_Stitch sampling
STITCH_RACES = ("Needle Withdrawal", "Needle Handling", "Needle Driving")


def stitch_indices(race: str, start_frame: int, end_frame: int, n_rgb: int, n_flow: int, jump_size: int,
                   phase: str = "inference", tta_offsets: Sequence[int] = (0, 3, 6)):
    """Row indices the ``VUA_EASE_Stitch`` dataset branch reads for one stitch sub-phase and its TTA views
    (prepare_dataset.py:2279-2396; the skill-assessment inference form of ``main.sh``), quirks included:

    * ``race`` selects the sub-phase; the caller passes the two frame numbers the reference takes from the data frame
      (Withdrawal: 'Needle Withdrawal Start / End Frame', Handling: 'Needle Handling Start Frame' / 'Needle Entry Start
      Frame', Driving: 'Needle Entry Start Frame' / 'Needle Withdrawal Start Frame', :2299-2307); both are shifted by -1;
    * ``phase`` 'val' / 'test': Withdrawal ``[s - 40, s + 40)``, Handling ``[s, e - 20)``, Driving ``[s, e - int(0.2 (e - s)))``
      (:2310-2320); any ``'...inference'`` phase: Withdrawal ``[s, s + 60)``, Handling and Driving ``[s, e)`` (:2329-2340);
    * every view is ``arange(start + o, end + o, 10)`` — the END moves with the offset here (unlike ``Custom_Gestures``);
    * flow rows: ``unique(rows // jump_size)`` with ``jump_size`` = 15 for the 30-fps 'Gronau_inference' videos and
      ``int(fps // 2)`` otherwise (:2362-2369), with NO ``< len(flow)`` filter: an out-of-range row raises IndexError like the
      reference's ``flow_reps[flow_indices, :]``; negative rows wrap (numpy semantics), floor division before the wrap.

    Returns ``(rgb, flow)``: lists with one int64 row array per view (ragged between views)."""
    if race not in STITCH_RACES:
        raise ValueError(f"race must be one of {STITCH_RACES}")
    if jump_size <= 0:
        raise ValueError("jump_size must be positive")
    s0, e0 = int(start_frame) - 1, int(end_frame) - 1
    if phase in ("val", "test"):
        if race == "Needle Withdrawal":
            start, end = s0 - 40, s0 + 40
        elif race == "Needle Handling":
            start, end = s0, e0 - 20
        else:
            start, end = s0, e0 - int((e0 - s0) * 0.20)
    elif "inference" in phase:
        start, end = (s0, s0 + 60) if race == "Needle Withdrawal" else (s0, e0)
    else:
        raise ValueError("phase must be 'val', 'test' or an '...inference' phase")
    rgb, flow = [], []
    for o in tta_offsets:
        raw = np.arange(start + o, end + o, 10, dtype=np.int64)
        f = np.unique(raw // jump_size)
        rgb.append(_wrap_rows(raw, n_rgb, "RGB"))
        flow.append(_wrap_rows(f, n_flow, "flow") if f.size else f)
    return rgb, flow


def _custom_gesture_indices_uniform(starts, ends, n_rgb, n_flow, tta_offsets, flow_stride):
    """:func:`custom_gesture_indices` for windows of one common duration, vectorised over the windows (same rows, same
    errors; pinned to the loop form and to the reference fixture by tests/test_pipeline_cpu.py)."""
    s0, e0 = starts - 1, ends - 1
    d = int(e0[0] - s0[0])
    jump = d // 10
    if jump <= 0:
        raise ValueError("windows shorter than 10 frames have no valid stride (reference: arange step 0)")
    rgb, flow = [], []
    for o in tta_offsets:
        raw = (s0[:, None] + o) + np.arange(0, d - o, jump, dtype=np.int64)[None, :]   # arange(s0 + o, e0, jump) per window
        q = raw // flow_stride                                                           # floor division BEFORE the wrap
        first = np.ones(q.shape, dtype=bool)
        first[:, 1:] = q[:, 1:] != q[:, :-1]                                             # unique() of an ascending row
        keep = first & (q < n_flow)
        vals = _wrap_rows(q[keep], n_flow, "flow")                                       # row-major = window order
        rgb.append(_wrap_rows(raw, n_rgb, "RGB"))
        flow.append(np.split(vals, np.cumsum(keep.sum(1))[:-1]))
    return rgb, flow


def gather_ragged(embeddings: torch.Tensor, rows: Sequence[np.ndarray]):
    """Ragged row lists -> zero-padded ``[W,1,Lmax,384]`` + key-padding mask ``bool [W,1,Lmax+1]`` built like
    ``createPaddingMask`` (prepare_dataset.py:2798-2806: ``mask[b,:,len_b+1:] = True``) + the lengths."""
    W = len(rows)
    lens = np.asarray([len(r) for r in rows], dtype=np.int64)
    L = int(lens.max()) if W else 0
    D = embeddings.shape[1]
    dev = embeddings.device
    if not (W and L):
        return (torch.zeros((W, 1, L, D), dtype=embeddings.dtype, device=dev),
                torch.zeros((W, 1, L + 1), dtype=torch.bool, device=dev), lens)
    # vectorised on the host (no per-clip Python work: a 1,000-clip sweep spent more time here than in the head's kernels),
    # one index + one mask transfer per call
    valid = np.arange(L, dtype=np.int64)[None, :] < lens[:, None]
    idx = np.zeros((W, L), dtype=np.int64)
    idx[valid] = np.concatenate([np.asarray(r, dtype=np.int64).reshape(-1) for r in rows])  # row-major fill = clip order
    keypad = np.concatenate([np.zeros((W, 1), dtype=bool), ~valid], axis=1)                   # token 0 (CLS) is never padded
    valid_d = torch.from_numpy(valid).to(dev, non_blocking=True)
    g = embeddings[torch.from_numpy(idx.reshape(-1)).to(dev, non_blocking=True)].view(W, L, D)
    out = (g * valid_d.unsqueeze(-1)).view(W, 1, L, D)
    mask = torch.from_numpy(keypad).to(dev, non_blocking=True).view(W, 1, L + 1)
    return out, mask, lens


# --------------------------------------------------------------------------------------------- feature extraction
def _pin(frames):
    """numpy -> tensor view.  Pageable memory is NOT pinned wholesale here: :class:`HostFrameStager` bounces it through
    two persistent pinned buffers batch by batch, overlapping the host copies with the GPU work."""
    return torch.from_numpy(frames) if isinstance(frames, np.ndarray) else frames


class HostFrameStager:
    """Double-buffered host->device staging of uint8 frame batches, owned by the model object so that it outlives
    a single :func:`extract_features` call.

    Two device buffers ``[batch,224,224,3]``, a private copy stream and one (ready, freed) event pair per buffer.
    Every copy into buffer ``k`` first waits for ``freed[k]`` — recorded on the compute stream after the forward that
    last READ buffer ``k``, whichever call issued it — so back-to-back calls (the single-batch-per-call pattern of
    ``bench.py``) can never overwrite frames a still-running forward is reading.  (Round 1 allocated the buffers per
    call and handed them back to the caching allocator while the forward was in flight: a cross-call WAR race.)"""

    def __init__(self, device, batch_size: int):
        self.device = torch.device(device)
        self.batch_size = int(batch_size)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [torch.empty((self.batch_size, 224, 224, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        self.used = [False, False]  # freed[k] has been recorded at least once
        self.turn = 0               # buffer the next batch goes to (persists across calls)
        # pageable sources go through two persistent pinned bounce buffers, one batch at a time (allocated on first use)
        self.pinned = [None, None]
        self.h2d_done = [None, None]

    def _bounce(self, k: int, src: torch.Tensor) -> torch.Tensor:
        """Pageable ``src`` -> pinned bounce buffer ``k`` (host memcpy of ONE batch; the previous device copy out of that
        buffer must have finished).  Pinning the caller's whole array per call instead — what round 1 did — allocates and
        fills a pinned copy of every frame before the first kernel can start: 92.8 ms per 1,024 frames against 12.0 ms from
        pinned memory (tools/e2e_host_bench.py)."""
        if self.pinned[k] is None:
            self.pinned[k] = torch.empty((self.batch_size, 224, 224, 3), dtype=torch.uint8).pin_memory()
            self.h2d_done[k] = torch.cuda.Event()
        else:
            self.h2d_done[k].synchronize()
        dst = self.pinned[k][: src.shape[0]]
        dst.copy_(src)
        return dst

    def stage(self, src: torch.Tensor) -> int:
        """Enqueue the copy of ``src`` (host, <= batch_size frames) into the next buffer; returns the buffer index."""
        k = self.turn
        self.turn ^= 1
        bounced = not src.is_pinned()
        if bounced:
            src = self._bounce(k, src)
        with torch.cuda.stream(self.copy_stream):
            if self.used[k]:
                self.copy_stream.wait_event(self.freed[k])
            self.bufs[k][: src.shape[0]].copy_(src, non_blocking=True)
            self.ready[k].record(self.copy_stream)
            if bounced:
                self.h2d_done[k].record(self.copy_stream)
        return k

    def acquire(self, k: int, n: int, stream) -> torch.Tensor:
        stream.wait_event(self.ready[k])
        return self.bufs[k][:n]

    def release(self, k: int, stream) -> None:
        self.freed[k].record(stream)
        self.used[k] = True


def _stager_for(model, device, batch_size: int) -> HostFrameStager:
    """The model's stager for the CURRENT stream (one per compute stream: lanes must not share staging buffers)."""
    stagers = getattr(model, "_host_stagers", None)
    if stagers is None:
        stagers = {}
        try:
            object.__setattr__(model, "_host_stagers", stagers)  # plain attribute, not a module / parameter
        except Exception:
            pass
    key = (str(torch.device(device)), torch.cuda.current_stream(device).cuda_stream)
    st = stagers.get(key)
    if st is None or st.batch_size < batch_size:
        st = HostFrameStager(device, batch_size)
        stagers[key] = st
    try:
        object.__setattr__(model, "_host_stager", st)  # (the most recently used one, for introspection / tests)
    except Exception:
        pass
    return st


@torch.no_grad()
def extract_features(model, frames, batch_size: int = 256, device=None, precision: Optional[str] = None,
                     out: Optional[torch.Tensor] = None, fanout=None) -> torch.Tensor:
    """``extractFeatures`` (extract_representations.py:351-378): walk the frames in batches, ``reps = model(inputs)``.
    ``frames``: uint8 ``[n,224,224,3]`` (host — pinned memory is copied from directly, pageable memory / numpy arrays go through
    the stager's pinned bounce buffers one batch at a time — or already on the device).  Returns fp32 ``[n,384]`` on
    the device.  Host batches are double-buffered through the model's :class:`HostFrameStager`: the copy of batch
    i+1 overlaps the ViT on batch i, within a call and across calls.  ``fanout`` (with ``out``): see
    :meth:`PeerGatherer.fanout` — every batch's embeddings also go to the other GPUs' copies of ``out``."""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    if fanout is not None and out is None:
        raise ValueError("fanout needs out= (this rank's slice of the symmetric gather buffer)")

    def fan(lo):
        if fanout is None:
            return {}
        from . import _lib
        return {"fanout": _lib.fanout_shifted(fanout, lo * 384 * 4)}

    frames = _pin(frames)
    n = frames.shape[0]
    if out is None:
        out = torch.empty((n, 384), dtype=torch.float32, device=device)
    if n == 0:
        return out
    if frames.device.type == "cuda":
        for lo in range(0, n, batch_size):
            model.forward_u8(frames[lo:lo + batch_size], precision=precision, out=out[lo:lo + batch_size], **fan(lo))
        return out
    with torch.cuda.device(device):
        st = _stager_for(model, device, min(batch_size, n))
        main = torch.cuda.current_stream(device)
        starts = list(range(0, n, batch_size))
        pending = st.stage(frames[0:min(batch_size, n)])
        for i, lo in enumerate(starts):
            hi = min(lo + batch_size, n)
            k = pending
            if i + 1 < len(starts):
                nlo = starts[i + 1]
                pending = st.stage(frames[nlo:min(nlo + batch_size, n)])
            buf = st.acquire(k, hi - lo, main)
            model.forward_u8(buf, precision=precision, out=out[lo:hi], **fan(lo))
            st.release(k, main)
    return out


# --------------------------------------------------------------------------------------------- whole path
class SaisPipeline:
    """frames (RGB + flow) -> per-frame embeddings -> windows (+TTA) -> temporal head -> prototype probabilities.

    Mirrors the chain ``main.sh:21,24,27,30`` of the reference without the HDF5 / pickle hand-offs.  Two window
    samplers:

    * ``sampling='custom_gestures'`` (what ``main.sh:27`` runs): the ``Custom_inference`` window list and the
      ``Custom_Gestures`` index contract, :func:`custom_gesture_windows` / :func:`custom_gesture_indices` — window rows
      start at ``StartFrame - 1`` (row -1 wraps), TTA views of 15 / 12 / 9 rows, flow rows ``unique(rows // flow_stride)``
      below ``len(flow)``, ragged flow views padded with key-padding masks;
    * ``sampling='plain'``: dense sliding windows over both streams with ``flow_stride == 1`` (the step-recognition form,
      prepare_dataset.py:469-473, 2324; BASELINE config C4).

    With ``torch.distributed`` initialised, frames are sharded by contiguous frame range, embeddings are all-gathered,
    and windows are sharded round-robin; :meth:`run_video` then returns this rank's windows only."""

    def __init__(self, vit, head, prototypes, window: int = 15, hop: int = 15, tta_offsets: Sequence[int] = (0, 3, 6),
                 flow_stride: Optional[int] = None, batch_size: int = 256, sampling: str = "plain"):
        if sampling not in ("plain", "custom_gestures"):
            raise ValueError("sampling must be 'plain' or 'custom_gestures'")
        self.vit, self.head = vit, head
        self.prototypes = scoring.stack_prototypes(prototypes)
        self.window, self.hop, self.tta = window, hop, tuple(tta_offsets)
        self.sampling = sampling
        self.flow_stride = int(flow_stride) if flow_stride is not None else (15 if sampling == "custom_gestures" else 1)
        if sampling == "plain" and self.flow_stride != 1:
            raise ValueError("sampling='plain' cuts RGB and flow windows alike (flow_stride must be 1); use "
                             "sampling='custom_gestures' for the rows // flow_stride mapping")
        self.batch_size = batch_size

    @torch.no_grad()
    def embed(self, frames, rank: int = 0, world: int = 1, precision=None) -> torch.Tensor:
        n = frames.shape[0]
        lo, hi = frame_range(n, rank, world)
        local = extract_features(self.vit, frames[lo:hi], self.batch_size, precision=precision)
        return gather_embeddings(local, n) if world > 1 else local

    def num_windows(self, n_rgb: int, n_flow: int) -> int:
        if self.sampling == "custom_gestures":
            return int(custom_gesture_windows(n_rgb, self.window, self.hop)[0].shape[0])
        return min(sliding_windows(n_rgb, self.window, self.hop)[0].shape[0],
                   sliding_windows(n_flow, self.window, self.hop)[0].shape[0])

    @torch.no_grad()
    def score_windows(self, rgb_emb: torch.Tensor, flow_emb: torch.Tensor, window_ids: Optional[np.ndarray] = None):
        """Returns ``(pred [W], probs [W,P], attn [W,S,S], window_ids)`` for the requested windows."""
        nw = self.num_windows(rgb_emb.shape[0], flow_emb.shape[0])
        ids = np.arange(nw, dtype=np.int64) if window_ids is None else np.asarray(window_ids, dtype=np.int64)
        dev = rgb_emb.device
        if self.sampling == "custom_gestures":
            starts, ends = custom_gesture_windows(rgb_emb.shape[0], self.window, self.hop)
            rgb_idx, flow_rows = custom_gesture_indices(starts[ids], ends[ids], rgb_emb.shape[0], flow_emb.shape[0],
                                                        self.tta, self.flow_stride)
            xs = [gather_windows(rgb_emb, r) for r in rgb_idx]
            xp = [full_mask(len(ids), x.shape[2], dev) for x in xs]
            fs, fp = [], []
            for rows in flow_rows:
                f, m, _ = gather_ragged(flow_emb, rows)
                fs.append(f)
                fp.append(m)
        else:
            views = sliding_windows(rgb_emb.shape[0], self.window, self.hop, self.tta)
            fviews = sliding_windows(flow_emb.shape[0], self.window, self.hop, self.tta)
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
    def score_rows(self, rgb_emb: torch.Tensor, flow_emb: torch.Tensor, rgb_rows, flow_rows):
        """The general ragged form behind every sampler: ``rgb_rows[v][w]`` / ``flow_rows[v][w]`` are the embedding rows of
        TTA view ``v`` of sample ``w`` (e.g. from :func:`stitch_indices`, one call per stitch sub-phase).  Views are padded to
        the batch maximum with key-padding masks built like ``createPaddingMask`` (``pad_collate``,
        prepare_dataset.py:2798-2871) and go through the head's TTA list form.  Returns ``(pred [W], probs [W,P] averaged
        over the views, attn [W,S,S] of RGB view 0)``."""
        if len(rgb_rows) != len(flow_rows) or not rgb_rows:
            raise ValueError("rgb_rows and flow_rows need the same, non-zero number of views")
        xs, xp, fs, fp = [], [], [], []
        for rows, frows in zip(rgb_rows, flow_rows):
            if len(rows) != len(frows):
                raise ValueError("every view needs one RGB and one flow row list per sample")
            x, m, _ = gather_ragged(rgb_emb, rows)
            f, fm, _ = gather_ragged(flow_emb, frows)
            xs.append(x), xp.append(m), fs.append(f), fp.append(fm)
        if len(xs) == 1:
            out, attn = self.head(xs[0], fs[0], None, None, 'Prototypes', xp[0], fp[0], None)
        else:
            none = [None] * len(xs)
            out, attn = self.head(xs, fs, none, none, 'Prototypes', xp, fp, None)
        pred, probs = scoring.predict(out, self.prototypes.to(rgb_emb.device))
        return pred, probs, attn

    @torch.no_grad()
    def clip_vectors(self, emb: torch.Tensor, rgb_offsets, flow_offsets, batch: int = 128) -> torch.Tensor:
        """Clip vectors ``[n,256]`` of ``n`` ragged clips (the skill-assessment sweep, BASELINE config C5): clip ``i`` owns the
        RGB rows ``[rgb_offsets[i], rgb_offsets[i+1])`` and the flow rows ``[flow_offsets[i], flow_offsets[i+1])`` of ``emb``.
        Clips are **bucketed by length** before they are padded into batches of ``batch`` (``pad_collate`` /
        ``createPaddingMask`` semantics, prepare_dataset.py:2798-2871): a clip's vector does not depend on what it is batched
        with (padded frames are masked keys and only the CLS row is consumed — pinned by the padding-invariance tests), so
        the order is free, and sorted batches carry a fraction of the padding of arrival-order batches of 8..64-frame clips
        while 4x larger batches need a quarter of the launches.  Results are returned in the original clip order."""
        r_off = np.asarray(rgb_offsets, dtype=np.int64)
        f_off = np.asarray(flow_offsets, dtype=np.int64)
        n = len(r_off) - 1
        out = torch.empty((n, 256), dtype=torch.float32, device=emb.device)
        if n == 0:
            return out
        order = np.argsort(r_off[1:] - r_off[:-1], kind="stable")
        for b0 in range(0, n, batch):
            ids = order[b0:b0 + batch]
            xs, xp, _ = gather_ragged(emb, [np.arange(r_off[i], r_off[i + 1]) for i in ids])
            fs, fp, _ = gather_ragged(emb, [np.arange(f_off[i], f_off[i + 1]) for i in ids])
            o, _ = self.head(xs, fs, None, None, 'Prototypes', xp, fp, None)
            out[torch.from_numpy(ids).to(emb.device)] = o
        return out

    @torch.no_grad()
    def run_video(self, rgb_frames, flow_frames, precision=None):
        import torch.distributed as dist

        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        er = self.embed(rgb_frames, rank, world, precision)
        ef = self.embed(flow_frames, rank, world, precision)
        nw = self.num_windows(er.shape[0], ef.shape[0])
        return self.score_windows(er, ef, shard_items(nw, rank, world))
