"""GPU parity of the HOST-frame path — the call a user makes and the path behind ``bench.py``'s ``e2e`` number:
``pipeline.extract_features`` with uint8 frames in (pinned) host memory, double-buffered H2D copies on a side stream
(extract_representations.py:365-371 is the loop it replaces).  Round 1 had no test here and a cross-call buffer race."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sais_oracle as O  # noqa: E402
from test_gpu_models import COS_MIN, REL_MAX, _vit  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def vit(dev):
    return _vit(O.make_vit_weights(0, "stress"), dev)


def _host_frames(n, seed, pinned=True):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, generator=g)
    return t.pin_memory() if pinned else t


@pytest.mark.parametrize("pinned", [True, False])
def test_host_frames_three_batches_and_ragged_tail(dev, vit, pinned):
    """3 full batches + a ragged tail from host memory == forward_u8 on the same frames, bit for bit; sampled frames
    equal the oracle."""
    from sais_b200 import pipeline

    n, bs = 3 * 48 + 17, 48
    frames = _host_frames(n, 101, pinned)
    got = pipeline.extract_features(vit, frames, batch_size=bs, device=dev)
    torch.cuda.synchronize()
    assert got.shape == (n, 384) and got.dtype == torch.float32 and got.device.type == "cuda"
    fd = frames.to(dev)
    want = torch.cat([vit.forward_u8(fd[lo:lo + bs].contiguous()) for lo in range(0, n, bs)], 0)
    assert torch.equal(got, want)
    # numpy input takes the same route
    got_np = pipeline.extract_features(vit, frames[:50].numpy(), batch_size=bs, device=dev)
    assert torch.equal(got_np[:48], want[:48])
    pick = [0, 47, 48, n - 1]
    ref = O.vit_forward(O.make_vit_weights(0, "stress"), O.normalize_frames(frames[pick]))
    cos, rel = O.embedding_errors(got[pick].cpu(), ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_back_to_back_single_batch_calls_do_not_race(dev, vit):
    """The bench pattern: one call per batch, different frames every call, no host synchronisation in between.  Each
    call's embeddings must equal the device-resident result of ITS OWN frames (the staging buffers are shared between
    calls; a missing cross-call ordering shows up as embeddings of a later batch)."""
    from sais_b200 import pipeline

    bs, ncalls = 64, 10
    batches = [_host_frames(bs, 200 + i) for i in range(ncalls)]
    want = [vit.forward_u8(b.to(dev)) for b in batches]
    torch.cuda.synchronize()
    outs = [torch.empty((bs, 384), dtype=torch.float32, device=dev) for _ in range(ncalls)]
    for rep in range(3):
        for i in range(ncalls):
            pipeline.extract_features(vit, batches[i], batch_size=bs, device=dev, out=outs[i])
        torch.cuda.synchronize()
        for i in range(ncalls):
            assert torch.equal(outs[i], want[i]), (rep, i)
    # the stager persists on the model (no per-call allocation of staging buffers)
    st = vit._host_stager
    pipeline.extract_features(vit, batches[0], batch_size=bs, device=dev)
    assert vit._host_stager is st


def test_alternating_batch_sizes_share_the_stager(dev, vit):
    from sais_b200 import pipeline

    a, b = _host_frames(40, 301), _host_frames(9, 302)
    wa, wb = vit.forward_u8(a.to(dev)), vit.forward_u8(b.to(dev))
    for _ in range(3):
        ga = pipeline.extract_features(vit, a, batch_size=16, device=dev)
        gb = pipeline.extract_features(vit, b, batch_size=16, device=dev)
        ga2 = pipeline.extract_features(vit, a, batch_size=64, device=dev)  # grows the stager
    torch.cuda.synchronize()
    want_a16 = torch.cat([vit.forward_u8(a.to(dev)[lo:lo + 16].contiguous()) for lo in range(0, 40, 16)], 0)
    assert torch.equal(ga, want_a16) and torch.equal(gb, wb) and torch.equal(ga2, wa)


def test_forward_u8_out_argument(dev, vit):
    from sais_b200 import SaisError

    fr = _host_frames(5, 401).to(dev)
    buf = torch.zeros((8, 384), dtype=torch.float32, device=dev)
    r = vit.forward_u8(fr, out=buf[2:7])
    assert r.data_ptr() == buf[2:7].data_ptr()
    assert torch.equal(buf[2:7], vit.forward_u8(fr)) and not buf[:2].any() and not buf[7:].any()
    with pytest.raises(SaisError):
        vit.forward_u8(fr, out=buf[:4])


def test_captured_step_replays_the_eager_clip_path(dev, vit):
    """pipeline.CapturedStep: one C1-shaped clip (10 RGB + 10 flow frames -> ViT -> temporal head -> prototype scores)
    captured as a CUDA graph; replays on new frames equal the eager path (ViT embeddings bit for bit; the small-batch head
    adds its split-K partials in L2 in arrival order, so the clip vector carries the same 1e-3 bound as bench.py's e2e gate)."""
    from sais_b200 import pipeline, scoring
    from test_gpu_models import _head

    head = _head(O.make_head_weights(0), dev, 'RGB-Flow')
    protos = torch.randn(2, 256, generator=torch.Generator().manual_seed(9)).to(dev)
    pad = pipeline.full_mask(1, 10, dev)

    def clip_step(frames):
        e = vit.forward_u8(frames)
        out, attn = head(e[:10].view(1, 1, 10, 384), e[10:].view(1, 1, 10, 384), None, None, 'Prototypes', pad, pad, None)
        pred, probs = scoring.predict(out, protos)
        return e, out, attn, pred, probs

    step = pipeline.CapturedStep(clip_step, _host_frames(20, 500).to(dev))
    for seed in (501, 502, 503):
        fr = _host_frames(20, seed).to(dev)
        want = [t.clone() for t in clip_step(fr)]
        got = [t.clone() for t in step(fr)]
        assert torch.equal(got[0], want[0])
        assert torch.allclose(got[1], want[1], atol=1e-3, rtol=1e-3)
        assert torch.allclose(got[2], want[2], atol=1e-4)
        assert torch.equal(got[3], want[3]) and torch.allclose(got[4], want[4], atol=1e-4)
    assert step.replays == 3
    torch.cuda.synchronize()


def test_mlp_policy_small_batch_path_within_tolerance(dev, vit):
    """sais_set_mlp_policy(1): batches below 48 frames take the fc1 / fc2 GEMM pair (single-clip latency).  Same tolerances
    against the oracle; the default policy keeps a frame's embedding independent of the batch size bit for bit."""
    from sais_b200 import _lib

    fr = _host_frames(20, 600).to(dev)
    base = vit.forward_u8(fr)
    with _lib.mlp_policy(1):
        lat = vit.forward_u8(fr)
        big = vit.forward_u8(_host_frames(64, 601).to(dev))
    assert torch.equal(vit.forward_u8(fr), base)                       # policy restored
    assert torch.equal(big, vit.forward_u8(_host_frames(64, 601).to(dev)))  # >= 48 frames: fused kernel either way
    ref = O.vit_forward(O.make_vit_weights(0, "stress"), O.normalize_frames(fr[:4].cpu()))
    cos, rel = O.embedding_errors(lat[:4].cpu(), ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    assert float((lat - base).abs().max()) <= 2e-2 * float(base.abs().max())
    with pytest.raises(_lib.SaisError):
        with _lib.mlp_policy(7):
            pass


def test_peer_gatherer_single_rank_is_a_plain_buffer(dev, vit):
    """world == 1: no symmetric memory, no fan-out, same interface (the N = 1 case of code written for N ranks)."""
    from sais_b200 import pipeline

    gat = pipeline.PeerGatherer(10, 384, 0, 1, dev, depth=2)
    assert gat.mode == "single-rank" and gat.fanout(0) is None
    for i in range(5):
        fr = _host_frames(10, 700 + i).to(dev)
        own = gat.own_slice(i)
        vit.forward_u8(fr, out=own, fanout=gat.fanout(i))
        gat.publish(i)
        assert torch.equal(gat.buffer(i), vit.forward_u8(fr))
    gat.wait_all()
