"""BASELINE.json configurations C2-C5 at their FULL sizes on the GPU, checked through size-independent properties plus
oracle comparisons on sampled items (the oracle finishes a handful of frames / clips in seconds, not thousands):

* C2  ViT-S/16 feature extraction, batch 256: sampled frames vs the oracle, batch-composition invariance;
* C3  temporal encoder + head, 512 clips x 30 frames (RGB + flow): sampled clips vs the oracle, clip-permutation
      equivariance, attention rows sum to one;
* C4  60-minute 1 fps video (3,600 RGB + 3,600 flow frames) sharded by frame range over 8 ranks, windows 20 / hop 10 /
      TTA {0,3,6}, P = 3: the union of the rank shards equals the single-rank run bit for bit, windows equal clips run alone;
* C5  1,000 clips of 8-64 frames (flow = ceil(len/2)) sharded by clip id mod 8 with padded batches: padding invariance
      against clips run alone, sampled clips vs the oracle, predicted class identical above the margin.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sais_oracle as O  # noqa: E402
from test_gpu_models import ATTN_ABS, COS_MIN, MARGIN, REL_MAX, REL_MAX_FP32, _head, _vit  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def vit(dev):
    return _vit(O.make_vit_weights(0, "stress"), dev)


@pytest.fixture(scope="module")
def head_sd():
    return O.make_head_weights(0, "stress")


def _frames(n, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)


# ------------------------------------------------------------------------------------------------ C2
def test_c2_vit_batch_256(dev, vit):
    frames = _frames(256, 11, dev)
    emb = vit.forward_u8(frames)
    assert emb.shape == (256, 384) and bool(torch.isfinite(emb).all())
    pick = [0, 37, 128, 255]
    ref = O.vit_forward(O.make_vit_weights(0, "stress"), O.normalize_frames(frames[pick].cpu()))
    cos, rel = O.embedding_errors(emb[pick].cpu(), ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    # a frame's embedding does not depend on what else is in the batch, where it sits, or how the batch is chunked
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(5)).to(dev)
    assert torch.equal(vit.forward_u8(frames[perm].contiguous()), emb[perm])
    assert torch.equal(vit.forward_u8(frames[100:133].contiguous()), emb[100:133])


# ------------------------------------------------------------------------------------------------ C3
def test_c3_temporal_head_512x30(dev, head_sd):
    model = _head(head_sd, dev, "RGB-Flow")
    g = torch.Generator().manual_seed(33)
    x = torch.randn(512, 1, 30, 384, generator=g)
    f = torch.randn(512, 1, 30, 384, generator=g)
    pad = O.padding_mask([30] * 512, 30)
    out, attn = model(x.to(dev), f.to(dev), None, None, 'Prototypes', pad.to(dev), pad.to(dev), None)
    assert out.shape == (512, 256) and attn.shape == (512, 31, 31)
    assert float((attn.sum(-1) - 1).abs().max()) <= 1e-5
    pick = [0, 99, 300, 511]
    r_out, r_attn = O.full_model_forward(head_sd, x[pick], f[pick], pad[pick], pad[pick])
    cos, rel = O.embedding_errors(out[pick].cpu(), r_out)
    assert cos >= 1 - 1e-6 and rel <= REL_MAX_FP32, (cos, rel)
    assert float((attn[pick].cpu() - r_attn).abs().max()) <= 1e-4
    perm = torch.randperm(512, generator=g)
    out_p, attn_p = model(x[perm].to(dev), f[perm].to(dev), None, None, 'Prototypes', pad.to(dev), pad.to(dev), None)
    assert torch.equal(out_p, out[perm.to(dev)]) and torch.equal(attn_p, attn[perm.to(dev)])


# ------------------------------------------------------------------------------------------------ C4
def test_c4_hour_video_sharded_by_frame_range(dev, vit, head_sd):
    from sais_b200 import pipeline, scoring
    n, world = 3600, 8
    head = _head(head_sd, dev, "RGB-Flow")
    protos = O.make_prototypes(3, seed=2).to(dev)
    pipe = pipeline.SaisPipeline(vit, head, protos, window=20, hop=10, tta_offsets=(0, 3, 6), batch_size=256)
    emb = {}
    for name, seed in (("rgb", 41), ("flow", 42)):
        whole = torch.empty((n, 384), device=dev)
        shards = []
        for r in range(world):  # every rank's frame range through the ViT, exactly as rank r would run it
            lo, hi = pipeline.frame_range(n, r, world)
            g = torch.Generator(device=dev).manual_seed(seed * 100 + r)
            fr = torch.randint(0, 256, (hi - lo, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
            e = pipeline.extract_features(vit, fr, batch_size=256)
            whole[lo:hi] = e
            shards.append((lo, hi, fr[:2].clone(), e[:2].clone()))
        emb[name] = whole
        lo, hi, fr2, e2 = shards[5]
        ref = O.vit_forward(O.make_vit_weights(0, "stress"), O.normalize_frames(fr2.cpu()))
        cos, rel = O.embedding_errors(e2.cpu(), ref)
        assert cos >= COS_MIN and rel <= REL_MAX, (name, cos, rel)
    pred, probs, attn, ids = pipe.score_windows(emb["rgb"], emb["flow"])
    nw = (n - 20) // 10 + 1
    assert ids.shape[0] == nw == 359 and probs.shape == (nw, 3) and attn.shape == (nw, 21, 21)
    assert float((probs.sum(1) - 1).abs().max()) <= 1e-5
    # window shards (round-robin over ranks) reassemble to the single-rank result bit for bit
    got_pred, got_probs = torch.empty_like(pred), torch.empty_like(probs)
    for r in range(world):
        own = pipeline.shard_items(nw, r, world)
        p_r, pr_r, _, ids_r = pipe.score_windows(emb["rgb"], emb["flow"], own)
        assert np.array_equal(ids_r, own)
        got_pred[torch.from_numpy(own).to(dev)] = p_r
        got_probs[torch.from_numpy(own).to(dev)] = pr_r
    assert torch.equal(got_pred, pred) and torch.equal(got_probs, probs)
    # one window against the oracle head on the same embeddings (TTA list form: three views ending at the same frame)
    w = 123
    xs = [emb["rgb"][w * 10 + o: w * 10 + 20].view(1, 1, -1, 384).cpu() for o in (0, 3, 6)]
    fs = [emb["flow"][w * 10 + o: w * 10 + 20].view(1, 1, -1, 384).cpu() for o in (0, 3, 6)]
    pads = [O.padding_mask([t.shape[2]], t.shape[2]) for t in xs]
    r_out, _ = O.full_model_forward(head_sd, xs, fs, pads, pads)
    r_probs = torch.stack([O.prototype_probs(o, protos.cpu())[0] for o in r_out], 0).mean(0)
    assert float((probs[w].cpu() - r_probs[0]).abs().max()) <= 1e-4


# ------------------------------------------------------------------------------------------------ C5
def test_c5_thousand_ragged_clips_sharded_by_clip(dev, head_sd):
    from sais_b200 import postprocess, scoring
    head = _head(head_sd, dev, "RGB-Flow")
    protos = O.make_prototypes(2, seed=2)
    rng = np.random.default_rng(3)
    lens = rng.integers(8, 65, 1000)
    g = torch.Generator().manual_seed(55)
    rgb = [torch.randn(1, int(n), 384, generator=g) for n in lens]
    flow = [torch.randn(1, int((n + 1) // 2), 384, generator=g) for n in lens]
    world, batch = 8, 32
    outs = torch.empty((1000, 256))
    preds = torch.empty(1000, dtype=torch.long)
    for r in range(world):
        own = np.arange(r, 1000, world)  # clip id mod 8
        for b0 in range(0, len(own), batch):
            ids = own[b0:b0 + batch]
            x, xpad, _ = postprocess.pad_collate([rgb[i] for i in ids])
            f, fpad, _ = postprocess.pad_collate([flow[i] for i in ids])
            out, attn = head(x.to(dev), f.to(dev), None, None, 'Prototypes', xpad.to(dev), fpad.to(dev), None)
            assert attn.shape == (len(ids), x.shape[2] + 1, x.shape[2] + 1)
            # padded keys carry exactly zero probability, real rows sum to one
            m = xpad.reshape(len(ids), -1).to(dev)
            assert float(attn[m.unsqueeze(1).expand_as(attn)].abs().max()) == 0.0 if bool(m.any()) else True
            pred, _ = scoring.predict(out, protos.to(dev))
            outs[ids] = out.cpu()
            preds[ids] = pred.cpu()
    assert bool(torch.isfinite(outs).all())
    # padding / batch-composition invariance: a clip run alone (no padding at all) gives the same vector
    for i in (0, 123, 500, 777, 999):
        x, f = rgb[i].unsqueeze(0), flow[i].unsqueeze(0)
        xp, fp = O.padding_mask([x.shape[2]], x.shape[2]), O.padding_mask([f.shape[2]], f.shape[2])
        alone, _ = head(x.to(dev), f.to(dev), None, None, 'Prototypes', xp.to(dev), fp.to(dev), None)
        assert float((alone.cpu()[0] - outs[i]).abs().max()) <= 2e-5 * float(outs[i].abs().max()) + 1e-6
        r_out, _ = O.full_model_forward(head_sd, x, f, xp, fp)
        cos, rel = O.embedding_errors(outs[i:i + 1], r_out)
        assert cos >= 1 - 1e-6 and rel <= REL_MAX_FP32, (i, cos, rel)
    # predicted class identical to the oracle's wherever the top-2 margin exceeds the tolerance (sampled clips)
    pick = list(range(0, 1000, 50))
    xs, xps, _ = postprocess.pad_collate([rgb[i] for i in pick])
    fs, fps, _ = postprocess.pad_collate([flow[i] for i in pick])
    r_out, _ = O.full_model_forward(head_sd, xs, fs, xps, fps)
    r_probs, r_sim = O.prototype_probs(r_out, protos)
    safe = O.top2_margin(r_sim) > MARGIN
    assert safe.any() and torch.equal(preds[pick][safe], r_probs.argmax(1)[safe])


# ------------------------------------------------------------------------------------------------ main.sh:27 sampling
def test_custom_gestures_pipeline_against_oracle(dev, golden_dir, head_sd):
    """`SaisPipeline(sampling='custom_gestures')` on the embeddings of one video: the windows / TTA rows / flow rows are the
    reference's (golden fixture written from its own statements), ragged flow views are padded with key-padding masks,
    and every sampled window equals the oracle head run on that window ALONE with the fixture's rows."""
    from sais_b200 import pipeline
    g = np.load(golden_dir / "custom_gesture_windows.npz")
    ci = 1
    n_rgb, n_flow = g["cases"][ci].tolist()  # 450 RGB rows, 30 flow rows
    gen = torch.Generator().manual_seed(77)
    rgb_emb, flow_emb = torch.randn(n_rgb, 384, generator=gen), torch.randn(n_flow, 384, generator=gen)
    head = _head(head_sd, dev, "RGB-Flow")
    protos = O.make_prototypes(2, seed=2)
    pipe = pipeline.SaisPipeline(None, head, protos, window=15, hop=15, tta_offsets=(0, 3, 6), sampling="custom_gestures")
    pred, probs, attn, ids = pipe.score_windows(rgb_emb.to(dev), flow_emb.to(dev))
    nw = len(g[f"c{ci}_start"])
    assert ids.shape[0] == nw == 30 and probs.shape == (nw, 2) and attn.shape == (nw, 16, 16)
    flat = [g[f"c{ci}_flow{v}_flat"] for v in range(3)]
    lens = [g[f"c{ci}_flow{v}_len"] for v in range(3)]
    offs = [np.concatenate([[0], np.cumsum(l)]) for l in lens]
    for w in (0, 1, 13, nw - 1):
        xs = [rgb_emb[torch.from_numpy(g[f"c{ci}_rgb{v}"][w])].view(1, 1, -1, 384) for v in range(3)]
        fs = [flow_emb[torch.from_numpy(flat[v][offs[v][w]:offs[v][w + 1]])].view(1, 1, -1, 384) for v in range(3)]
        xp = [O.padding_mask([t.shape[2]], t.shape[2]) for t in xs]
        fp = [O.padding_mask([t.shape[2]], t.shape[2]) for t in fs]
        r_out, r_attn = O.full_model_forward(head_sd, xs, fs, xp, fp)
        r_probs = torch.stack([O.prototype_probs(o, protos)[0] for o in r_out], 0).mean(0)
        assert float((probs[w].cpu() - r_probs[0]).abs().max()) <= 1e-4, w
        assert float((attn[w].cpu() - r_attn[0]).abs().max()) <= 1e-4, w
    # sharding the windows over ranks reassembles to the same result
    got = torch.empty_like(probs)
    for r in range(4):
        own = pipeline.shard_items(nw, r, 4)
        _, pr, _, _ = pipe.score_windows(rgb_emb.to(dev), flow_emb.to(dev), own)
        got[torch.from_numpy(own).to(dev)] = pr
    assert float((got - probs).abs().max()) <= 1e-6


def test_stitch_sampling_through_the_head_against_oracle(dev, golden_dir, head_sd):
    """VUA_EASE_Stitch form end to end: rows from pipeline.stitch_indices (pinned to the reference's statements by
    tests/golden/stitch_windows.npz) -> SaisPipeline.score_rows (ragged views of different lengths in one padded batch,
    S up to 67) -> every sample equals the oracle head run on that stitch ALONE with the fixture's rows."""
    from sais_b200 import pipeline
    g = np.load(golden_dir / "stitch_windows.npz")
    n_rgb, n_flow = int(g["n_rgb"]), int(g["n_flow"])
    gen = torch.Generator().manual_seed(78)
    rgb_emb, flow_emb = torch.randn(n_rgb, 384, generator=gen), torch.randn(n_flow, 384, generator=gen)
    head = _head(head_sd, dev, "RGB-Flow")
    protos = O.make_prototypes(2, seed=2)
    pipe = pipeline.SaisPipeline(None, head, protos, sampling="plain")
    cases = [ci for ci, m in enumerate(g["cases"]) if str(m).split("|")[0] in ("Gronau_inference", "HMH_inference")]
    rgb_rows, flow_rows = [[], [], []], [[], [], []]
    for ci in cases:
        phase, race, s, e, fps = str(g["cases"][ci]).split("|")
        jump = 15 if phase == "Gronau_inference" else int(int(fps) // 2)
        r, f = pipeline.stitch_indices(pipeline.STITCH_RACES[int(race)], int(s), int(e), n_rgb, n_flow, jump, phase)
        for v in range(3):
            rgb_rows[v].append(r[v])
            flow_rows[v].append(f[v])
    pred, probs, attn = pipe.score_rows(rgb_emb.to(dev), flow_emb.to(dev), rgb_rows, flow_rows)
    assert probs.shape == (len(cases), 2) and pred.shape == (len(cases),)
    for k in (0, 5, 14, len(cases) - 1, len(cases) - 3):
        ci = cases[k]
        xs = [rgb_emb[torch.from_numpy(g[f"c{ci}_rgb{v}"])].view(1, 1, -1, 384) for v in range(3)]
        fs = [flow_emb[torch.from_numpy(g[f"c{ci}_flow{v}"])].view(1, 1, -1, 384) for v in range(3)]
        xp = [O.padding_mask([t.shape[2]], t.shape[2]) for t in xs]
        fp = [O.padding_mask([t.shape[2]], t.shape[2]) for t in fs]
        r_out, r_attn = O.full_model_forward(head_sd, xs, fs, xp, fp)
        r_probs = torch.stack([O.prototype_probs(o, protos)[0] for o in r_out], 0).mean(0)
        assert float((probs[k].cpu() - r_probs[0]).abs().max()) <= 1e-4, ci
        S = xs[0].shape[2] + 1
        assert float((attn[k, :S, :S].cpu() - r_attn[0]).abs().max()) <= 1e-4, ci


def test_clip_vectors_bucketed_batches_equal_arrival_order_batches(dev, head_sd):
    """SaisPipeline.clip_vectors (length-bucketed batches of 128) vs arrival-order pad_collate batches of 32 on 300 ragged clips:
    a clip's vector does not depend on its batch (same bound as the padding-invariance check of the C5 test), results come
    back in the original order."""
    from sais_b200 import pipeline, postprocess
    head = _head(head_sd, dev, "RGB-Flow")
    rng = np.random.default_rng(4)
    lens = rng.integers(8, 65, 300)
    fl = (lens + 1) // 2
    r_off = np.concatenate([[0], np.cumsum(lens)])
    f_off = int(lens.sum()) + np.concatenate([[0], np.cumsum(fl)])
    emb = torch.randn(int(lens.sum() + fl.sum()), 384, generator=torch.Generator().manual_seed(56)).to(dev)
    pipe = pipeline.SaisPipeline(None, head, O.make_prototypes(2, seed=2))
    got = pipe.clip_vectors(emb, r_off, f_off, batch=128)
    want = torch.empty_like(got)
    for b0 in range(0, 300, 32):
        ids = range(b0, min(b0 + 32, 300))
        xs, xp, _ = postprocess.pad_collate([emb[r_off[i]:r_off[i + 1]].unsqueeze(0) for i in ids])
        fs, fp, _ = postprocess.pad_collate([emb[f_off[i]:f_off[i + 1]].unsqueeze(0) for i in ids])
        o, _ = head(xs, fs, None, None, 'Prototypes', xp.to(dev), fp.to(dev), None)
        want[b0:b0 + len(ids)] = o
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-6
    assert pipe.clip_vectors(emb[:0], [0], [0]).shape == (0, 256)
