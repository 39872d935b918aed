"""CPU: the sharding / window / TTA arithmetic around the hot path and the N>1 exchange step (world_size-2 ``gloo``).
No kernels are launched: the all-gather is exercised on CPU tensors, the ViT is replaced by a stand-in."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sais_b200 import pipeline


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (3600, 8), (3601, 8), (125, 3)])
def test_frame_range_partitions_every_frame_once(n, world):
    ranges = [pipeline.frame_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (lo, hi), (lo2, _) in zip(ranges, ranges[1:]):
        assert hi == lo2 and lo <= hi
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1  # balanced
    with pytest.raises(ValueError):
        pipeline.frame_range(n, world, world)


def test_shard_items_round_robin():
    parts = [pipeline.shard_items(1000, r, 8) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(1000))
    assert parts[3][:3].tolist() == [3, 11, 19]
    assert pipeline.shard_items(2, 5, 8).size == 0  # more ranks than items: empty shard, not an error


def test_sliding_windows_tta_views_share_the_window_end():
    """Custom_Gestures inference windows: 15 frames, hop 15, TTA offsets 0/3/6 -> lengths 15/12/9, same end
    (reference prepare_dataset.py:1711-1726, 2646-2651)."""
    views = pipeline.sliding_windows(100, 15, 15, (0, 3, 6))
    assert [v.shape for v in views] == [(6, 15), (6, 12), (6, 9)]
    for v in views:
        assert np.array_equal(v[:, -1], views[0][:, -1])
    assert views[1][2, 0] == 2 * 15 + 3 and views[2][0, 0] == 6
    # step-recognition form: 20 frames, hop 10 (prepare_dataset.py:469-473,2324)
    v = pipeline.sliding_windows(3600, 20, 10)[0]
    assert v.shape == (359, 20) and v[-1, -1] == 3599
    assert pipeline.sliding_windows(10, 15, 15)[0].shape == (0, 15)  # video shorter than one window
    with pytest.raises(ValueError):
        pipeline.sliding_windows(10, 15, 15, (15,))


def test_gather_windows_and_mask_layout():
    emb = torch.arange(50 * 384, dtype=torch.float32).view(50, 384)
    idx = pipeline.sliding_windows(50, 15, 15, (0, 3))[1]
    w = pipeline.gather_windows(emb, idx)
    assert w.shape == (3, 1, 12, 384)
    assert torch.equal(w[1, 0, 0], emb[15 + 3]) and torch.equal(w[2, 0, -1], emb[44])
    m = pipeline.full_mask(3, 12, "cpu")
    assert m.shape == (3, 1, 13) and m.dtype == torch.bool and not m.any()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_frames * 384, dtype=torch.float32).view(n_frames, 384)
        lo, hi = pipeline.frame_range(n_frames, rank, world)
        got = pipeline.gather_embeddings(full[lo:hi].clone(), n_frames)
        ok = torch.equal(got, full)
        # window ownership after the gather: round-robin, disjoint, complete
        nw = pipeline.sliding_windows(n_frames, 20, 10)[0].shape[0]
        mine = pipeline.shard_items(nw, rank, world)
        counts = torch.zeros(nw, dtype=torch.int64)
        counts[torch.from_numpy(mine)] += 1
        dist.all_reduce(counts)
        ok = ok and bool((counts == 1).all())
        # a rank holding the wrong number of rows must be rejected, not silently padded
        try:
            pipeline.gather_embeddings(full[lo:hi + 1 if hi < n_frames else hi - 1].clone(), n_frames)
            bad = False
        except ValueError:
            bad = True
        q.put((rank, ok, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [64, 61])  # even split (single collective) and ragged split (pad + trim)
def test_gather_embeddings_world2_gloo(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get() for _ in range(2))
    assert res == [(0, True, True), (1, True, True)]


class _FakeVit(torch.nn.Module):
    """Stand-in with the product ViT's ``forward_u8`` contract (embedding = per-frame mean colour, repeated)."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward_u8(self, frames, precision=None):
        return frames.float().mean(dim=(1, 2)).repeat(1, 128)


def test_extract_features_empty_video():
    """batch loop of extractFeatures (extract_representations.py:365-371) on a video with no frames."""
    frames = torch.zeros((0, 224, 224, 3), dtype=torch.uint8)
    assert pipeline.extract_features(_FakeVit(), frames, 8, device="cpu").shape == (0, 384)


# ------------------------------------------------------------------------------------- Custom_Gestures sampling contract
def test_custom_gesture_sampling_matches_reference_statements(golden_dir):
    """Window list + per-view RGB / flow rows equal what the reference's own statements read
    (prepare_dataset.py:1711-1726, 2642-2695 executed by oracle/make_golden_windows.py): StartFrame-1 with the row -1
    wrap-around, TTA views of 15 / 12 / 9 rows, flow rows unique(rows // 15) below len(flow)."""
    g = np.load(golden_dir / "custom_gesture_windows.npz")
    for ci, (n_rgb, n_flow) in enumerate(g["cases"].tolist()):
        starts, ends = pipeline.custom_gesture_windows(n_rgb)
        assert np.array_equal(starts, g[f"c{ci}_start"]) and np.array_equal(ends, g[f"c{ci}_end"]), (n_rgb, n_flow)
        rgb, flow = pipeline.custom_gesture_indices(starts, ends, n_rgb, n_flow)
        for v in range(3):
            want = g[f"c{ci}_rgb{v}"]
            if len(starts) == 0:
                assert rgb[v].shape[0] == 0
                continue
            assert np.array_equal(rgb[v], want), (ci, v)
            lens = g[f"c{ci}_flow{v}_len"]
            assert [len(f) for f in flow[v]] == lens.tolist(), (ci, v)
            flat = np.concatenate(flow[v]) if len(flow[v]) else np.zeros(0, dtype=np.int64)
            assert np.array_equal(flat, g[f"c{ci}_flow{v}_flat"]), (ci, v)
    # spot checks of the quirks themselves
    starts, ends = pipeline.custom_gesture_windows(100)
    rgb, flow = pipeline.custom_gesture_indices(starts, ends, 100, 7)
    assert rgb[0][0, 0] == 99 and rgb[0].shape == (6, 15) and rgb[1].shape == (6, 12) and rgb[2].shape == (6, 9)
    assert flow[0][0].tolist() == [6, 0] and flow[1][0].tolist() == [0]
    assert pipeline.custom_gesture_windows(14)[0].size == 0
    with pytest.raises(IndexError):  # a window past the end of the embeddings raises like reps[indices,:] does
        pipeline.custom_gesture_indices([90], [105], 100, 7)


def test_gather_ragged_masks_like_create_padding_mask():
    emb = torch.arange(10 * 384, dtype=torch.float32).view(10, 384)
    rows = [np.array([9, 0]), np.array([3]), np.zeros(0, dtype=np.int64)]
    out, mask, lens = pipeline.gather_ragged(emb, rows)
    assert out.shape == (3, 1, 2, 384) and mask.shape == (3, 1, 3) and lens.tolist() == [2, 1, 0]
    assert torch.equal(out[0, 0, 0], emb[9]) and torch.equal(out[0, 0, 1], emb[0]) and torch.equal(out[1, 0, 0], emb[3])
    assert not out[1, 0, 1].any() and not out[2].any()  # zero padding
    assert mask.tolist() == [[[False, False, False]], [[False, False, True]], [[False, True, True]]]
    from sais_b200 import postprocess
    assert torch.equal(mask, postprocess.create_padding_mask([2, 1, 0]))


def test_pipeline_sampling_modes_are_validated():
    with pytest.raises(ValueError):
        pipeline.SaisPipeline(None, None, torch.zeros(2, 256), sampling="plain", flow_stride=15)
    p = pipeline.SaisPipeline(None, None, torch.zeros(2, 256), sampling="custom_gestures")
    assert p.flow_stride == 15 and p.num_windows(100, 7) == 6
    assert pipeline.SaisPipeline(None, None, torch.zeros(2, 256), window=20, hop=10).num_windows(3600, 3600) == 359


def _gatherer_worker(rank, world, port, n_rows, frame_ranges, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gat = pipeline.EmbeddingGatherer(n_rows, 384, rank, world, "cpu", frame_ranges=frame_ranges)
        ok = True
        for step in range(5):  # slots rotate; every step's rows differ
            full = torch.arange(n_rows * 384, dtype=torch.float32).view(n_rows, 384) + 1000.0 * step
            lo, hi = gat.ranges[rank]
            own = gat.own_slice(step)
            assert own.shape == (hi - lo, 384)
            own.copy_(full[lo:hi])          # stands in for forward_u8(..., out=own)
            gat.gather_async(step)
            got = gat.buffer(step)
            ok = ok and torch.equal(got, full)
        gat.wait_all()
        covered = sorted(gat.ranges)
        ok = ok and covered[0][0] == 0 and covered[-1][1] == n_rows and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,frame_ranges", [(64, False), (61, True), (61, False)])
def test_embedding_gatherer_in_place_world2_gloo(n_rows, frame_ranges):
    """The in-place, asynchronous all-gather of bench.py / C4 (persistent slots, the owner writes its slice, gather joins on
    demand) on CPU tensors over gloo, world size 2: even split, frame_range split and a ragged ceil split."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, n_rows, frame_ranges, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [(0, True), (1, True)]


def test_embedding_gatherer_single_rank_is_a_plain_buffer():
    gat = pipeline.EmbeddingGatherer(10, 384, 0, 1, "cpu")
    own = gat.own_slice(0)
    own.fill_(3.0)
    gat.gather_async(0)
    assert gat.buffer(0).data_ptr() == own.data_ptr() and gat.buffer(0).shape == (10, 384)


def test_stitch_sampling_matches_reference_statements(golden_dir):
    """VUA_EASE_Stitch index arithmetic (prepare_dataset.py:2279-2396) == the fixture written by executing the reference's
    own statements (oracle/make_golden_stitch.py): 4 phases x 3 stitch sub-phases x 5 frame ranges, three TTA views each,
    including 24-fps videos (jump 12), rows that wrap below 0 and a 66-row view."""
    from sais_b200 import pipeline

    g = np.load(golden_dir / "stitch_windows.npz")
    n_rgb, n_flow = int(g["n_rgb"]), int(g["n_flow"])
    assert len(g["cases"]) == 60
    for ci, meta in enumerate(g["cases"]):
        phase, race, s, e, fps = str(meta).split("|")
        jump = 15 if phase == "Gronau_inference" else int(int(fps) // 2)
        rgb, flow = pipeline.stitch_indices(pipeline.STITCH_RACES[int(race)], int(s), int(e), n_rgb, n_flow, jump, phase)
        for v in range(3):
            assert np.array_equal(rgb[v], g[f"c{ci}_rgb{v}"]), (meta, v)
            assert np.array_equal(flow[v], g[f"c{ci}_flow{v}"]), (meta, v)
    with pytest.raises(IndexError):  # no '< len(flow)' filter in this branch: the reference's fancy index raises
        pipeline.stitch_indices("Needle Handling", 3000, 3651, 4000, 100, 15, "Gronau_inference")
    with pytest.raises(ValueError):
        pipeline.stitch_indices("Needle Poking", 1, 2, 10, 10, 15)


def _loop_form(pipeline, s, e, n_rgb, n_flow, stride):
    """The reference arithmetic for ONE window, statement by statement (prepare_dataset.py:2642-2672)."""
    s0, e0 = int(s) - 1, int(e) - 1
    jump = (e0 - s0) // 10
    rgb, flow = [], []
    for o in (0, 3, 6):
        raw = np.arange(s0 + o, e0, jump, dtype=np.int64)
        q = np.unique(raw // stride)
        q = q[q < n_flow]
        rgb.append([np.where(raw < 0, raw + n_rgb, raw)])
        flow.append([np.where(q < 0, q + n_flow, q)])
    return rgb, flow


def test_custom_gesture_indices_vectorised_form_equals_the_loop_form():
    """Windows of one duration take the vectorised path; windows of mixed durations the per-window loop: same rows, same
    errors (the reference fixture above pins the loop's arithmetic)."""
    from sais_b200 import pipeline

    rng = np.random.default_rng(7)
    for n_rgb, n_flow, dur, stride in [(450, 30, 15, 15), (2000, 40, 30, 12), (100, 3, 15, 15), (64, 64, 20, 1)]:
        starts = np.sort(rng.integers(0, n_rgb - dur, 37)).astype(np.int64)
        starts[0] = 0  # the row -1 wrap
        ends = starts + dur
        fast = pipeline.custom_gesture_indices(starts, ends, n_rgb, n_flow, (0, 3, 6), stride)
        slow_rgb, slow_flow = [], []
        for s, e in zip(starts, ends):  # the per-window reference form: the loop body of custom_gesture_indices, one window at a time
            r, f = _loop_form(pipeline, s, e, n_rgb, n_flow, stride)
            slow_rgb.append([v[0] for v in r]), slow_flow.append([v[0] for v in f])
        for v in range(3):
            assert np.array_equal(fast[0][v], np.stack([w[v] for w in slow_rgb]))
            for w in range(len(starts)):
                assert np.array_equal(fast[1][v][w], slow_flow[w][v]), (n_rgb, v, w)
                assert fast[1][v][w].dtype == np.int64
    # windows of two durations with equal row counts (20 and 30 frames -> 10 rows each) go through the per-window loop
    # (only view 0: the later views of the two durations differ in length, which one call refuses like the reference's collate)
    mixed = pipeline.custom_gesture_indices([40, 100], [60, 130], 450, 30, (0,), 15)
    for w, (s, e) in enumerate([(40, 60), (100, 130)]):
        r, f = _loop_form(pipeline, s, e, 450, 30, 15)
        assert np.array_equal(mixed[0][0][w], r[0][0]) and np.array_equal(mixed[1][0][w], f[0][0])
    with pytest.raises(IndexError):
        pipeline.custom_gesture_indices([90, 95], [105, 110], 100, 7)
    with pytest.raises(ValueError):
        pipeline.custom_gesture_indices([0, 20], [9, 29], 100, 7)
