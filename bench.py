#!/usr/bin/env python
"""Benchmark of the SAIS inference hot path on B200 (contract: see the task description / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input per rank:
  256 uint8 frames (128 RGB + 128 optical-flow, 224x224) -> frame normalisation + DINO ViT-S/16 (bf16 operands,
  fp32 accumulate) -> [256,384] embeddings, written straight into this rank's slice of a persistent gather buffer ->
  (N>1: in-place NCCL all-gather of the frame-range shards, asynchronous, overlapped with the head) -> SAIS temporal
  head on this rank's 8 clips x 16 frames (RGB + flow) -> prototype scores (P=2).  Consecutive steps alternate between two
  pipeline lanes (CUDA streams, pipeline.Lanes; --lanes 1 for a single stream): one batch's kernel tails and head are
  filled by the other batch's kernels.  Every step's work completes inside the timed region.
This is BASELINE.json configs[1] (ViT-S/16, batch 256, 1xB200) with the SAIS head of the metric on top.

value : frames/s with the u8 frames already resident in HBM (max over ranks, CUDA events, K steps).
e2e   : same metric through the public API (sais_b200.pipeline.extract_features + fullModel + scoring) with the
        frames in PINNED HOST memory: H2D of every batch and D2H of embeddings + clip scores inside the timed region.
        Before timing, one e2e step is checked bit for bit against the resident step on the same frames.
roofline : the DOMINANT KERNEL of the step (mlp_fused_kernel), CUDA events around every one of its launches (separate
        single-lane pass, so that a launch's time is its own).
extra : the other BASELINE configs measured in the same run — C1 (10+10-frame clip), C3 (head, 512 clips x 30 frames),
        C4 (60-min video, 3,600+3,600 frames, STRONG scaling over the ranks), C5 (1,000 ragged clips sharded by clip).
--impl reference : the reference's algorithm on the host CPU (oracle port, all cores) on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAMES_PER_STEP = 256
CLIPS_PER_STEP, CLIP_T = 8, 16
FLOP_PER_FRAME = 9_196_996_608  # every token through all 12 blocks, as the reference computes it
# The last block runs its attention / proj / MLP on the CLS row only (the only row VisionTransformer.forward returns,
# vision_transformer.py:213-214): 196/197 of (QK^T + PV + proj + fc1 + fc2) of one block are never needed.
FLOP_PER_FRAME_EXECUTED = FLOP_PER_FRAME - (29_805_312 * 2 + 58_097_664 + 2 * 232_390_656) * 196 // 197
METRIC = "frames/sec (ViT-S/16 224^2 RGB+flow + SAIS head)"
REF_FRAMES = 16  # frames per step of the CPU reference arm (8 RGB + 8 flow + head on 1 clip)
# kernel classes of the library's CUDA-event profiler (include/sais_b200.h, sais_profile_end)
CLASSES = ["gemm_other", "vit_attention", "layernorm_rowstats", "patchify", "temporal_attention", "misc",
           "gemm_split3_head", "mlp_fused", "gemm_qkv", "gemm_proj"]
TENSOR_CLASSES = {"gemm_other", "vit_attention", "gemm_split3_head", "mlp_fused", "gemm_qkv", "gemm_proj"}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.thread, self.nv = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            for prefix in ("nvmlClocksEventReason", "nvmlClocksThrottleReason"):
                if hasattr(nv, prefix + n):
                    names[getattr(nv, prefix + n)] = n
                    break
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.nv is None:  # one-shot fallback
            try:
                import subprocess
                out = subprocess.run(["nvidia-smi", f"--id={self.index}",
                                      "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.samples, self.max_mhz = [int(out[0])], int(out[1])
            except Exception:
                pass
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference arm
def make_reference_step(frames=REF_FRAMES):
    """One bounded sample of the workload through the oracle port of the reference forward on the host CPU (fp32, all
    cores): ViT on `frames` frames (half RGB, half flow) + temporal head on one clip + prototype scoring."""
    import torch
    from oracle import sais_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vsd, hsd = O.make_vit_weights(0, "init"), O.make_head_weights(0, "init")
    fr = O.make_frames_u8(frames, 5)
    protos = O.make_prototypes(2)
    T = frames // 2

    def step():
        emb = O.vit_forward(vsd, O.normalize_frames(fr))
        x, f = emb[:T].view(1, 1, T, 384), emb[T:].view(1, 1, T, 384)
        pad = O.padding_mask([T], T)
        out, _ = O.full_model_forward(hsd, x, f, pad, pad)
        return O.prototype_probs(out, protos)[0]

    sample = (f"{frames} frames ({T} RGB + {T} flow) through the ViT + SAIS head on 1 clip + scoring per step, "
              f"oracle port of the reference forward, fp32 torch CPU, {cores} threads")
    return step, cores, sample


def cpu_reference_rate(seconds_budget=12.0):
    """(frames/s, cores, sample description) — best pass within the budget (reported baseline of the GPU arm's line)."""
    step, cores, sample = make_reference_step()
    step()  # warm-up
    times = []
    t_end = time.perf_counter() + seconds_budget
    while len(times) < 2 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if len(times) >= 50:
            break
    return REF_FRAMES / min(times), cores, sample + f" (best of {len(times)} passes)"


def run_reference(args):
    """--impl reference: EXACTLY --steps timed steps after --warmup warm-up steps, each a bounded sample (16 frames + 1
    clip) of the GPU arm's workload; rank 0 alone runs and prints."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    step, cores, sample = make_reference_step()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = REF_FRAMES * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def kernel_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same batch-256 step
    (profiles/mlp_traffic.json, written by tools/ncu_traffic.py); null when no capture is committed."""
    p = ROOT / "profiles" / "mlp_traffic.json"
    if not p.exists():
        return {"traffic": None}
    d = json.loads(p.read_text())
    return {"traffic": d.get("dram_bytes_per_launch"), "traffic_unit": "B/launch", "traffic_source": d.get("source")}


def workload_config(n_gpus):
    return {
        "workload": "C2: DINO ViT-S/16 feature extraction, batch 256 u8 frames 224x224 (128 RGB + 128 flow) per GPU, "
                    "+ SAIS temporal head on 8 clips x 16 frames (RGB+flow) + 2 prototypes",
        "frames_per_gpu_per_step": FRAMES_PER_STEP, "clips_per_gpu_per_step": CLIPS_PER_STEP, "clip_frames": CLIP_T,
        "sharding": "frame range per rank; when n_gpus > 1 every rank's embeddings reach all ranks each step: stored into the "
                    "peers' symmetric gather buffers by the ViT's last kernel (NVSwitch multicast) + one barrier, or "
                    "(--exchange nccl) an in-place asynchronous NCCL all-gather; overlapped with the head either way",
        "l2": "inputs rotate over 4 distinct 38.5 MB frame batches and the 426 MB per-step working set exceeds "
              "the 126 MB L2",
        "parallelism": f"dp{n_gpus}",
    }


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="sais_b200", choices=["sais_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C3 / C4 / C5 legs")
    ap.add_argument("--chunk", type=int, default=256, help="ViT frames per workspace chunk")
    ap.add_argument("--vit-sms", type=int, default=0,
                    help="with --head-stream 1: SMs the ViT's persistent kernels may occupy (even; 0 = all), the rest are left "
                         "to the concurrent head (sais_set_sm_limit).  140 of 148 costs the ViT almost nothing at batch 256 — "
                         "1,536 attention items = 11 rounds on 140 CTAs as on 148, 197 MLP row tiles = 3 rounds on 70 CTA "
                         "pairs as on 74 — and measured +3.3 %% frames/s")
    ap.add_argument("--lanes", type=int, default=3,
                    help="pipeline lanes (pipeline.Lanes): consecutive steps go to consecutive CUDA streams, so one batch's kernel "
                         "tails and head are filled by the other batches' kernels (default 3)")
    ap.add_argument("--exchange", default="peer", choices=["nccl", "peer"],
                    help="N > 1: how the per-step embeddings reach the other ranks.  nccl: in-place asynchronous all-gather "
                         "(pipeline.EmbeddingGatherer); peer: no collective — the ViT's final-LayerNorm kernel stores its rows "
                         "into every GPU's symmetric gather buffer (NVSwitch multicast / NVLink peer stores) and one barrier "
                         "publishes them (pipeline.PeerGatherer)")
    ap.add_argument("--numa-bind", type=int, default=1,
                    help="N > 1: pin every rank to the CPU cores local to its GPU before the pinned host buffers are allocated "
                         "(pipeline.bind_to_gpu_numa); matters for the end-to-end number only")
    ap.add_argument("--head-stream", type=int, default=0,
                    help="1: run the temporal head + scoring of a step on a separate high-priority stream "
                         "(pipeline.SideStream) instead of on the step's lane; with --lanes 1 --vit-sms 140 this is the "
                         "single-lane way of hiding the head (88.6 k frames/s); with two lanes it is not needed")
    args = ap.parse_args()
    if args.impl == "reference":
        args.warmup = max(args.warmup, 1)
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", device_id=dev)

    import sais_b200.vision_transformer as vits
    from sais_b200 import _lib, pipeline, postprocess, scoring
    from sais_b200.prepare_model import fullModel

    numa_cpus = pipeline.bind_to_gpu_numa(local_rank) if (world > 1 and args.numa_bind) else None
    lib = _lib.lib()
    # random-init weights of the named architectures (the modules' own init = the reference's distributions)
    torch.manual_seed(0)
    vit = vits.vit_small(patch_size=16, chunk_frames=args.chunk).to(dev).eval()
    head = fullModel(data_type='reps', nclasses=2, domain='NH_02', rep_dim=384, encoder_type='ViT',
                     modalities='RGB-Flow').to(dev).eval()
    protos = torch.randn(2, 256, device=dev)

    # synthetic frames: 4 distinct batches per rank, rotated so no step re-reads the previous step's input from L2
    nbuf = 4
    g = torch.Generator().manual_seed(100 + rank)
    host_batches = [torch.randint(0, 256, (FRAMES_PER_STEP, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory()
                    for _ in range(nbuf)]
    dev_batches = [hb.to(dev) for hb in host_batches]
    pad = pipeline.full_mask(CLIPS_PER_STEP, CLIP_T, dev)
    n_global = FRAMES_PER_STEP * world
    # pipeline lanes: consecutive steps go to consecutive CUDA streams (pipeline.Lanes); result slot i % depth belongs to
    # lane i % n_lanes, so nothing is shared between lanes but the (read-only) weights
    lanes = pipeline.Lanes(dev, args.lanes)
    depth = len(lanes) if len(lanes) > 1 else 2
    xgroup = pipeline.low_footprint_group() if world > 1 else None   # one-CTA NCCL collectives for the exchange
    peer_x = args.exchange == "peer" and world > 1
    exchange_note = None
    if peer_x:
        # measured at N = 8 (profiles/r02_exchange_ab.md): same device-resident rate as the NCCL all-gather, +2 % end to end
        try:
            gatherer = pipeline.PeerGatherer(n_global, 384, rank, world, dev, depth=depth)
        except Exception as ex:  # no symmetric memory on this box: every rank fails alike -> the NCCL exchange
            peer_x, exchange_note = False, f"peer exchange unavailable ({type(ex).__name__}: {str(ex)[:120]}); NCCL all-gather used"
    if not peer_x:
        gatherer = pipeline.EmbeddingGatherer(n_global, 384, rank, world, dev, depth=depth, group=xgroup)

    def fan(i):
        return {"fanout": gatherer.fanout(i)} if peer_x else {}

    def exchange(i):
        if peer_x:
            gatherer.publish(i)       # one barrier on the gatherer's own stream: the rows are already on their way
        else:
            gatherer.gather_async(i)  # in place, on NCCL's stream; the head below reads own rows only

    def head_and_score(own):
        # this rank's clips: 8 RGB clips from the first half of its frame range, 8 flow clips from the second half
        x = own[:CLIPS_PER_STEP * CLIP_T].view(CLIPS_PER_STEP, 1, CLIP_T, 384)
        f = own[128:128 + CLIPS_PER_STEP * CLIP_T].view(CLIPS_PER_STEP, 1, CLIP_T, 384)
        out, attn = head(x, f, None, None, 'Prototypes', pad, pad, None)
        pred, probs = scoring.predict(out, protos)
        return out, probs, pred

    side = pipeline.SideStream(dev, slots=depth) if args.head_stream else None

    def run_head(i, own):
        """head + scoring of step i: on the step's lane, or (--head-stream 1) on a high-priority side stream ordered after
        the lane's ViT and joined before the slot it reads is overwritten"""
        if side is None:
            return head_and_score(own)
        return side.run(i, head_and_score, own)

    def claim_slot(i):
        own = gatherer.own_slice(i)
        if side is not None:
            side.guard(i)  # the head that read this slot `depth` steps ago
        return own

    vit_limit = _lib.sm_limit(args.vit_sms if (args.head_stream and args.vit_sms) else 0)

    def step_resident(i):
        with lanes.lane(i):
            own = claim_slot(i)                   # the final-LN kernel writes at rank * count of the gather buffer
            with vit_limit:
                vit.forward_u8(dev_batches[i % nbuf], out=own, **fan(i))
            exchange(i)
            return own, run_head(i, own)

    emb_host = [torch.empty((FRAMES_PER_STEP, 384), dtype=torch.float32).pin_memory() for _ in range(depth)]
    probs_host = [torch.empty((CLIPS_PER_STEP, 2), dtype=torch.float32).pin_memory() for _ in range(depth)]

    def step_e2e(i):
        with lanes.lane(i):
            own = claim_slot(i)
            with vit_limit:
                pipeline.extract_features(vit, host_batches[i % nbuf], batch_size=FRAMES_PER_STEP, device=dev, out=own, **fan(i))
            exchange(i)
            out, probs, pred = run_head(i, own)
            # device->host reads of the step's results: embeddings + clip probabilities, behind the head on its stream
            with torch.cuda.stream(side.stream if side is not None else torch.cuda.current_stream(dev)):
                emb_host[i % depth].copy_(own, non_blocking=True)
                probs_host[i % depth].copy_(probs, non_blocking=True)
                if side is not None:
                    side.mark(i)
            return own, (out, probs, pred)

    def join_side():
        lanes.join()
        if side is not None:
            side.join()

    def barrier():
        join_side()
        gatherer.wait_all()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate of the e2e path: host-frame step == resident step on the same frames, bit for bit
    for i in range(2):
        own_r, (out_r, probs_r, _) = step_resident(i)
        barrier()  # (the head may run on the side stream: join before reading its results)
        keep = (own_r.clone(), out_r.clone(), probs_r.clone())
        own_e, (out_e, probs_e, _) = step_e2e(i)
        barrier()
        if not (torch.equal(own_e, keep[0]) and torch.equal(emb_host[i % depth], keep[0].cpu())):
            raise SystemExit("bench: e2e embeddings differ from the device-resident step")
        # (the small-batch head accumulates its split-K partials in L2 in arrival order: last-bit run-to-run differences)
        if not (torch.allclose(out_e, keep[1], rtol=1e-3, atol=1e-4) and torch.allclose(probs_host[i % depth], keep[2].cpu(), atol=1e-4)):
            raise SystemExit("bench: e2e clip vectors / probabilities differ from the device-resident step")
        if world > 1:  # the gathered buffer holds every rank's rows: compare a checksum of rank r's slice with rank r's own
            full = gatherer.buffer(i)
            sums = full.view(world, -1).double().sum(1)
            mine = sums[rank].clone()
            allsums = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allsums, mine)  # (default group: also checks the two communicators agree on the ranks)
            if not torch.equal(torch.stack(allsums), sums):
                raise SystemExit("bench: gathered embeddings differ from the owners' rows")

    def timed(fn, steps, warmup, sampler=None):
        for i in range(warmup):
            fn(i)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.sais_launch_count()
        e0.record()
        lanes.fork()         # the lanes start behind e0
        for i in range(steps):
            fn(warmup + i)
        join_side()          # every lane / head of the timed steps ...
        gatherer.wait_all()  # ... and every gather completes inside the timed region
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), lib.sais_launch_count() - l0, clocks

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)

    # roofline leg: same steps with every launch bracketed by CUDA events, per kernel class
    import ctypes as C
    ncls = len(CLASSES)
    ms_c, work_c, n_c = (C.c_double * ncls)(), (C.c_double * ncls)(), (C.c_int64 * ncls)()
    prof_steps = min(args.steps, 10)
    barrier()
    lib.sais_profile_begin()
    for i in range(prof_steps):
        step_resident(i * len(lanes))  # one lane only: per-kernel times must not include the other lane's kernels
    _lib.check(lib.sais_profile_end(ms_c, work_c, n_c, ncls), "sais_profile_end")
    barrier()

    # effective SM clock inside the loop: a one-thread probe kernel between steps (cycle counter vs nanosecond timer);
    # separate pass so that the timed region above stays untouched
    probe_steps = min(args.steps, 20)
    probe = torch.zeros((probe_steps, 4), dtype=torch.int64, device=dev)
    for i in range(3):
        step_resident(i)
    for i in range(probe_steps):
        step_resident(i)
        with lanes.lane(i):
            _lib.check(lib.sais_clock_probe(probe[i].data_ptr(), 4000, torch.cuda.current_stream(dev).cuda_stream),
                       "clock_probe")
    barrier()
    pr = probe.cpu().double()
    ghz = ((pr[:, 3] - pr[:, 1]) / (pr[:, 2] - pr[:, 0]).clamp(min=1.0)).tolist()
    ghz_sorted = sorted(ghz)

    extra = {}
    if not args.no_extra:
        extra = run_extra_configs(torch, dist, dev, rank, world, vit, head, pipeline, postprocess, scoring, barrier, lanes, xgroup)

    if rank == 0:
        burst, sustained, hbm, src = load_peaks()
        value = n_global * args.steps / (ms * 1e-3)
        e2e = n_global * args.steps / (ms_e2e * 1e-3)
        total_prof = sum(ms_c) or 1.0
        shares = {}
        for i, c in enumerate(CLASSES):
            if n_c[i] == 0:
                continue
            d = {"ms_per_step": ms_c[i] / prof_steps, "launches_per_step": n_c[i] / prof_steps,
                 "share": ms_c[i] / total_prof, "avg_launch_us": 1e3 * ms_c[i] / n_c[i]}
            if c in TENSOR_CLASSES:
                d["tflops"] = work_c[i] / 1e12 / (ms_c[i] * 1e-3)
            else:
                d["gbs"] = work_c[i] / 1e9 / (ms_c[i] * 1e-3)
            shares[c] = d
        # the regime decides the denominator: the timed region runs at (near) max clock -> burst peak, else sustained
        ghz_med = ghz_sorted[len(ghz_sorted) // 2]
        max_ghz = ((clocks or {}).get("sm_max_mhz") or 1965) / 1e3
        use_burst = ghz_med >= 0.95 * max_ghz
        peak, peak_kind = (burst, "bf16_tflops (burst)") if use_burst else (sustained, "bf16_tflops_sustained")
        k = CLASSES.index("mlp_fused")
        mlp_ms = ms_c[k] / max(n_c[k], 1)
        mlp_flop = work_c[k] / max(n_c[k], 1)
        mlp_tflops = mlp_flop / 1e12 / (mlp_ms * 1e-3) if mlp_ms > 0 else 0.0
        tok = FRAMES_PER_STEP * 197
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(world), lanes=len(lanes), head_stream=bool(args.head_stream),
                           exchange=(("peer:" + gatherer.mode) if peer_x else ("nccl" if world > 1 else "none")),
                           **({"exchange_note": exchange_note} if exchange_note else {}),
                           **({"numa_bound_cpus": len(numa_cpus)} if numa_cpus else {}),
                           vit_sms=args.vit_sms or "all"),
            "flop_per_frame": {"reference_forward": FLOP_PER_FRAME, "executed": FLOP_PER_FRAME_EXECUTED,
                               "note": "last block evaluated on the CLS rows only (dead rows of the reference "
                                       "forward are not computed); fractions below use the EXECUTED flops"},
            "step_tc_frac": {"of_burst": value / world * FLOP_PER_FRAME_EXECUTED / (burst * 1e12),
                             "of_sustained": value / world * FLOP_PER_FRAME_EXECUTED / (sustained * 1e12),
                             "target": 0.60},
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": FRAMES_PER_STEP * 224 * 224 * 3,
                    "d2h_bytes_per_step": FRAMES_PER_STEP * 384 * 4 + CLIPS_PER_STEP * 2 * 4,
                    "verified": "one e2e step == resident step bit for bit before timing"},
            "gpu_launches": int(launches),
            "clocks": dict(clocks or {}, sm_ghz_in_loop_median=ghz_med, sm_ghz_in_loop_min=ghz_sorted[0],
                           sm_ghz_in_loop_max=ghz_sorted[-1],
                           in_loop_note="clock64 / globaltimer of a probe kernel enqueued after every step of a separate "
                                        "20-step pass"),
            "roofline": {"bound": "tensor",
                         "kernel": "mlp_fused_kernel (fc1 + GELU + fc2 + residual, %d launches per step, %.1f %% of the "
                                   "step's kernel time)" % (round(n_c[k] / prof_steps), 100 * ms_c[k] / total_prof),
                         "achieved": mlp_tflops, "peak": peak, "unit": "TFLOP/s", "frac": mlp_tflops / peak,
                         "peak_kind": peak_kind, "peak_source": src, "frac_of_burst": mlp_tflops / burst,
                         "frac_of_sustained": mlp_tflops / sustained,
                         "avg_launch_ms": mlp_ms, "flop_per_launch": mlp_flop,
                         "algorithmic_bytes_per_launch": tok * (768 + 32 + 1536 * 2 + 768 + 32) + 2 * 2 * 1536 * 384,
                         "algorithmic_bytes_note": "per launch: bf16 rows + statistics in, fp32 residual read-modify-write, "
                                                   "bf16 copy + statistics out (cast warps), weights once",
                         **kernel_traffic()},
            "kernel_classes": shares,
            "extra_configs": extra,
        }
        if not args.no_cpu_baseline and world == 1:  # the CPU leg is timed at N = 1 only (bench contract)
            v, cores, sample = cpu_reference_rate()
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ other BASELINE configs
def run_extra_configs(torch, dist, dev, rank, world, vit, head, pipeline, postprocess, scoring, barrier, lanes, xgroup=None):
    """C1 / C3 on every rank (rank 0 reports), C4 strong scaling and C5 sharded by clip over all ranks.  Device-timed with
    CUDA events, max over ranks; each leg: 1-2 warm-up passes + a few timed ones (bounded, ~2 s in total at N = 1)."""
    out = {}

    def time_passes(fn, warm, reps):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    protos2 = torch.randn(2, 256, device=dev, generator=g)
    protos3 = torch.randn(3, 256, device=dev, generator=g)

    # ---- C1: one 10-frame RGB + flow clip (20 ViT frames + head, S = 11), single GPU semantics on every rank
    c1_frames = torch.randint(0, 256, (20, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
    pad10 = pipeline.full_mask(1, 10, dev)

    def c1():
        e = vit.forward_u8(c1_frames)
        o, _ = head(e[:10].view(1, 1, 10, 384), e[10:].view(1, 1, 10, 384), None, None, 'Prototypes', pad10, pad10, None)
        return scoring.predict(o, protos2)

    ms = time_passes(c1, 3, 20)
    out["C1_clip_10+10_frames"] = {"ms_per_clip": ms, "frames_per_s": 20 / (ms * 1e-3), "n_gpus": 1,
                                   "note": "latency-bound: 20 frames is 0.08 of one ViT row-tile round"}
    # the same clip step as ONE CUDA-graph launch (pipeline.CapturedStep): ~120 kernel launches -> 1
    try:
        def c1_step(frames):
            e = vit.forward_u8(frames)
            o, _ = head(e[:10].view(1, 1, 10, 384), e[10:].view(1, 1, 10, 384), None, None, 'Prototypes', pad10, pad10, None)
            return scoring.predict(o, protos2)

        graphed = pipeline.CapturedStep(c1_step, c1_frames)
        ms_g = time_passes(lambda: graphed(c1_frames), 3, 20)
        out["C1_clip_10+10_frames"].update({"ms_per_clip_cuda_graph": ms_g, "frames_per_s_cuda_graph": 20 / (ms_g * 1e-3)})
        # ... and with the small-batch MLP policy (fc1 / fc2 GEMM pair below 48 frames: lower latency, but a frame's
        # embedding then depends on the batch size in the last bits, so it is opt-in: _lib.mlp_policy / sais_set_mlp_policy)
        from sais_b200 import _lib as _l
        with _l.mlp_policy(1):
            graphed_lat = pipeline.CapturedStep(c1_step, c1_frames)
        ms_l = time_passes(lambda: graphed_lat(c1_frames), 3, 20)
        out["C1_clip_10+10_frames"].update({"ms_per_clip_cuda_graph_latency_policy": ms_l,
                                            "frames_per_s_cuda_graph_latency_policy": 20 / (ms_l * 1e-3)})
    except Exception as ex:  # reported, never fatal for the headline
        out["C1_clip_10+10_frames"]["cuda_graph_error"] = f"{type(ex).__name__}: {str(ex)[:160]}"

    # ---- C3: temporal head alone, 512 clips x 30 frames (RGB + flow), attention maps out
    x = torch.randn(512, 1, 30, 384, device=dev, generator=g)
    f = torch.randn(512, 1, 30, 384, device=dev, generator=g)
    pad30 = pipeline.full_mask(512, 30, dev)

    def c3():
        o, a = head(x, f, None, None, 'Prototypes', pad30, pad30, None)
        return scoring.predict(o, protos2)

    ms = time_passes(c3, 2, 10)
    flop = 512 * 1.0847e9
    out["C3_temporal_512x30"] = {"ms_per_batch": ms, "clips_per_s": 512 / (ms * 1e-3), "n_gpus": 1,
                                 "tflops_fp32_equivalent": flop / 1e12 / (ms * 1e-3),
                                 "note": "split-precision (3 bf16 MMAs per product): executed tensor work is 3x this"}

    # ---- C4: 60-minute 1 fps video, 3,600 RGB + 3,600 flow frames, STRONG scaling: frame ranges over the ranks,
    # all-gather of both [3600,384] streams, windows 20 / hop 10 x TTA {0,3,6}, P = 3, windows sharded round-robin
    n = 3600
    lo, hi = pipeline.frame_range(n, rank, world)
    pool = torch.randint(0, 256, (512, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)  # frames of the shard

    def shard_frames(k):  # contiguous 256-frame chunks cut from the pool: no copies inside the timed region
        return pool[(k * 256) % 512: (k * 256) % 512 + 256]

    pipe = pipeline.SaisPipeline(vit, head, protos3, window=20, hop=10, tta_offsets=(0, 3, 6), batch_size=256)
    # (default communicator here: one 5.5 MB exchange per stream and video, nothing to overlap it with — bandwidth matters,
    # the one-CTA group cost 1 ms per video at N = 8)
    g_rgb = pipeline.EmbeddingGatherer(n, 384, rank, world, dev, frame_ranges=True)
    g_flow = pipeline.EmbeddingGatherer(n, 384, rank, world, dev, frame_ranges=True)

    def c4():
        lanes.fork()
        for li, gat in enumerate((g_rgb, g_flow)):  # the RGB and the flow frames of the shard: one pipeline lane each
            with lanes.lane(li):
                own = gat.own_slice(0)
                k = 0
                for b0 in range(0, hi - lo, 256):
                    nb = min(256, hi - lo - b0)
                    vit.forward_u8(shard_frames(k)[:nb], out=own[b0:b0 + nb])
                    k += 1
                gat.gather_async(0)
        lanes.join()
        g_rgb.wait_all()
        g_flow.wait_all()
        nw = pipe.num_windows(n, n)
        return pipe.score_windows(g_rgb.buffer(0), g_flow.buffer(0), pipeline.shard_items(nw, rank, world))

    ms = time_passes(c4, 1, 3)
    out["C4_60min_video"] = {"ms_per_video": ms, "frames_per_s": 2 * n / (ms * 1e-3), "n_gpus": world,
                             "scaling": "strong", "frames": 2 * n, "windows": 359, "tta_views": 3}

    # ---- C5: 1,000 clips of 8-64 RGB frames (+ ceil(len/2) flow frames), sharded by clip id mod world: ViT over the
    # shard's frames, length-bucketed padded batches of 128 clips through the head, all-gather of the [n_own,256] clip vectors
    import numpy as np
    rng = np.random.default_rng(3)
    lens = rng.integers(8, 65, 1000)
    own_ids = np.arange(rank, 1000, world)
    rl = lens[own_ids]
    fl = (rl + 1) // 2
    nr, nf = int(rl.sum()), int(fl.sum())
    emb = torch.empty((nr + nf, 384), dtype=torch.float32, device=dev)
    r_off = np.concatenate([[0], np.cumsum(rl)])
    f_off = nr + np.concatenate([[0], np.cumsum(fl)])
    width = (1000 + world - 1) // world
    vec_own = torch.zeros((width, 256), dtype=torch.float32, device=dev)
    vec_all = torch.empty((world * width, 256), dtype=torch.float32, device=dev)

    pipe5 = pipeline.SaisPipeline(vit, head, protos2)

    def c5():
        lanes.fork()
        k = 0
        for b0 in range(0, nr + nf, 256):  # 256-frame chunks of the shard alternate between the lanes
            nb = min(256, nr + nf - b0)
            with lanes.lane(k):
                vit.forward_u8(shard_frames(k)[:nb], out=emb[b0:b0 + nb])
            k += 1
        lanes.join()
        # length-bucketed padded batches of 128 clips through the head (pipeline.SaisPipeline.clip_vectors)
        vec_own[:len(own_ids)] = pipe5.clip_vectors(emb, r_off, f_off, batch=128)
        if world > 1:
            dist.all_gather_into_tensor(vec_all, vec_own)
            return scoring.predict(vec_all, protos2)
        return scoring.predict(vec_own, protos2)

    ms = time_passes(c5, 1, 2)
    tot = int((lens + (lens + 1) // 2).sum())
    out["C5_1000_clips"] = {"ms_per_sweep": ms, "clips_per_s": 1000 / (ms * 1e-3), "frames_per_s": tot / (ms * 1e-3),
                            "n_gpus": world, "scaling": "strong", "frames": tot}
    return out


if __name__ == "__main__":
    main()
