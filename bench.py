#!/usr/bin/env python
"""Benchmark of the SAIS inference hot path on B200 (contract: see the task description / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input per rank:
  256 uint8 frames (128 RGB + 128 optical-flow, 224x224) -> frame normalisation + DINO ViT-S/16 (bf16 operands,
  fp32 accumulate) -> [256,384] embeddings -> (N>1: NCCL all-gather of the frame-range shards) -> SAIS temporal
  head on this rank's 8 clips x 16 frames (RGB + flow) -> prototype scores (P=2).
This is BASELINE.json configs[1] (ViT-S/16, batch 256, 1xB200) with the SAIS head of the metric on top.

value : frames/s with the u8 frames already resident in HBM (max over ranks, CUDA events, K steps).
e2e   : same metric through the public API (sais_b200.pipeline.extract_features + fullModel + scoring) with the
        frames in PINNED HOST memory: H2D of every batch and D2H of embeddings + clip scores inside the timed region.
--impl reference : the reference's algorithm on the host CPU (oracle port, all cores) on a bounded sample.
"""
from __future__ import annotations

import argparse
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
def cpu_reference_rate(seconds_budget=12.0, frames=16):
    """Oracle port of the reference forward on the host CPU (fp32, all cores): ViT on `frames` frames (half RGB,
    half flow) + temporal head + scoring.  Returns (frames/s, cores, sample description, seconds per pass)."""
    import torch
    from oracle import sais_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vsd, hsd = O.make_vit_weights(0, "init"), O.make_head_weights(0, "init")
    fr = O.make_frames_u8(frames, 5)
    protos = O.make_prototypes(2)
    T = frames // 2

    def one_pass():
        emb = O.vit_forward(vsd, O.normalize_frames(fr))
        x, f = emb[:T].view(1, 1, T, 384), emb[T:].view(1, 1, T, 384)
        pad = O.padding_mask([T], T)
        out, _ = O.full_model_forward(hsd, x, f, pad, pad)
        return O.prototype_probs(out, protos)[0]

    one_pass()  # warm-up
    times = []
    t_end = time.perf_counter() + seconds_budget
    while len(times) < 2 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
        if len(times) >= 50:
            break
    best = min(times)
    sample = f"{frames} frames ({T} RGB + {T} flow) + head on 1 clip, best of {len(times)} passes, fp32 torch CPU"
    return frames / best, cores, sample, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401
    frames = 16
    from oracle import sais_oracle as O
    cores = os.cpu_count() or 1
    import torch
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

    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = frames * steps / dt
    sample = f"{frames} frames ({T} RGB + {T} flow) + SAIS head on 1 clip per step, fp32 torch CPU, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def gemm_traffic():
    """DRAM bytes per bf16 GEMM launch from the committed ncu --set full capture (profiles/gemm_traffic.json, written
    by tools/ncu_traffic.py from the same batch-256 step); null when no capture is committed."""
    p = ROOT / "profiles" / "gemm_traffic.json"
    if not p.exists():
        return {"traffic": None}
    d = json.loads(p.read_text())
    return {"traffic": d.get("dram_bytes_per_launch"), "traffic_unit": "B/launch", "traffic_source": d.get("source")}


def workload_config(n_gpus):
    return {
        "workload": "C2: DINO ViT-S/16 feature extraction, batch 256 u8 frames 224x224 (128 RGB + 128 flow) per GPU, "
                    "+ SAIS temporal head on 8 clips x 16 frames (RGB+flow) + 2 prototypes",
        "frames_per_gpu_per_step": FRAMES_PER_STEP, "clips_per_gpu_per_step": CLIPS_PER_STEP, "clip_frames": CLIP_T,
        "sharding": "frame range per rank; all-gather of embeddings when n_gpus > 1",
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
    ap.add_argument("--chunk", type=int, default=256, help="ViT frames per workspace chunk")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

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
    from sais_b200 import _lib, pipeline, scoring
    from sais_b200.prepare_model import fullModel

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

    def head_and_score(emb_all):
        # this rank's clips: 8 RGB clips from the first half of its frame range, 8 flow clips from the second half
        lo = rank * FRAMES_PER_STEP
        x = emb_all[lo:lo + CLIPS_PER_STEP * CLIP_T].view(CLIPS_PER_STEP, 1, CLIP_T, 384)
        f = emb_all[lo + 128:lo + 128 + CLIPS_PER_STEP * CLIP_T].view(CLIPS_PER_STEP, 1, CLIP_T, 384)
        out, attn = head(x, f, None, None, 'Prototypes', pad, pad, None)
        pred, probs = scoring.predict(out, protos)
        return out, probs, pred

    def step_resident(i):
        emb = vit.forward_u8(dev_batches[i % nbuf])
        emb_all = pipeline.gather_embeddings(emb, n_global) if world > 1 else emb
        return head_and_score(emb_all)

    emb_host = torch.empty((FRAMES_PER_STEP, 384), dtype=torch.float32).pin_memory()
    probs_host = torch.empty((CLIPS_PER_STEP, 2), dtype=torch.float32).pin_memory()
    emb_dev = torch.empty((FRAMES_PER_STEP, 384), dtype=torch.float32, device=dev)

    def step_e2e(i):
        emb = pipeline.extract_features(vit, host_batches[i % nbuf], batch_size=FRAMES_PER_STEP, device=dev,
                                        out=emb_dev)
        emb_all = pipeline.gather_embeddings(emb, n_global) if world > 1 else emb
        out, probs, pred = head_and_score(emb_all)
        emb_host.copy_(emb, non_blocking=True)
        probs_host.copy_(probs, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for i in range(warmup):
            fn(i)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.sais_launch_count()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
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
    ncls = 7
    ms_c, work_c, n_c = (C.c_double * ncls)(), (C.c_double * ncls)(), (C.c_int64 * ncls)()
    prof_steps = min(args.steps, 10)
    barrier()
    lib.sais_profile_begin()
    for i in range(prof_steps):
        step_resident(i)
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
        _lib.check(lib.sais_clock_probe(probe[i].data_ptr(), 4000, torch.cuda.current_stream().cuda_stream), "clock_probe")
    barrier()
    pr = probe.cpu().double()
    ghz = ((pr[:, 3] - pr[:, 1]) / (pr[:, 2] - pr[:, 0]).clamp(min=1.0)).tolist()
    ghz_sorted = sorted(ghz)

    if rank == 0:
        burst, sustained, hbm, src = load_peaks()
        value = n_global * args.steps / (ms * 1e-3)
        e2e = n_global * args.steps / (ms_e2e * 1e-3)
        gemm_ms = ms_c[0] / max(n_c[0], 1)
        gemm_tflops = (work_c[0] / 1e12) / (ms_c[0] * 1e-3) if ms_c[0] > 0 else 0.0
        classes = ["gemm", "vit_attention", "layernorm", "patchify", "temporal_attention", "misc", "gemm_split3_head"]
        total_prof = sum(ms_c) or 1.0
        shares = {c: {"ms_per_step": ms_c[i] / prof_steps, "launches_per_step": n_c[i] / prof_steps,
                      "share": ms_c[i] / total_prof} for i, c in enumerate(classes)}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "flop_per_frame": {"reference_forward": FLOP_PER_FRAME, "executed": FLOP_PER_FRAME_EXECUTED,
                               "note": "last block evaluated on the CLS rows only (dead rows of the reference "
                                       "forward are not computed); fractions below use the EXECUTED flops"},
            "tc_frac_of_measured_sustained": value / world * FLOP_PER_FRAME_EXECUTED / (sustained * 1e12),
            "tc_frac_of_measured_burst": value / world * FLOP_PER_FRAME_EXECUTED / (burst * 1e12),
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": FRAMES_PER_STEP * 224 * 224 * 3,
                    "d2h_bytes_per_step": FRAMES_PER_STEP * 384 * 4 + CLIPS_PER_STEP * 2 * 4},
            "gpu_launches": int(launches),
            "clocks": dict(clocks or {}, sm_ghz_in_loop_median=ghz_sorted[len(ghz_sorted) // 2], sm_ghz_in_loop_min=ghz_sorted[0],
                           sm_ghz_in_loop_max=ghz_sorted[-1],
                           in_loop_note="clock64 / globaltimer of a probe kernel enqueued after every step of a separate "
                                        "20-step pass"),
            "roofline": {"bound": "tensor",
                         "kernel": "GEMM class, bf16: gemm_tcgen05_kernel + mlp_fused_kernel (the %d tensor-core GEMM "
                                   "launches of a step: patch, qkv, proj, fused fc1+GELU+fc2, CLS-row GEMMs of the "
                                   "last block)" % round(n_c[0] / prof_steps),
                         "achieved": gemm_tflops, "peak": sustained, "unit": "TFLOP/s",
                         "frac": gemm_tflops / sustained, "frac_of_burst": gemm_tflops / burst, "peak_source": src,
                         "avg_launch_ms": gemm_ms, "flop_per_launch": work_c[0] / max(n_c[0], 1),
                         **gemm_traffic()},
            "kernel_classes": shares,
        }
        if not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_reference_rate()
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
