"""GPU parity of the exchange step WITHOUT a collective (needs 2 GPUs; skipped on a 1-GPU box): pipeline.PeerGatherer +
``forward_u8(..., fanout=...)`` — the ViT's final-LayerNorm kernel stores this rank's embedding rows into every GPU's
symmetric gather buffer (NVSwitch multicast, and plain NVLink peer stores), one barrier publishes them.  Every rank must
end up with every rank's rows bit for bit (each rank recomputes the other's frames itself: same weights, same kernels)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import sais_oracle as O
        from sais_b200 import pipeline
        from test_gpu_models import _vit

        vit = _vit(O.make_vit_weights(0, "stress"), dev)
        n_per, steps = 24, 7
        ok, modes = True, []
        for use_mc in (True, False):
            gat = pipeline.PeerGatherer(n_per * world, 384, rank, world, dev, depth=2, use_multicast=use_mc)
            modes.append(gat.mode)
            lanes = pipeline.Lanes(dev, 2)
            lanes.fork()
            frames, fulls = {}, []
            for i in range(steps):
                for r in range(world):  # every rank can generate every rank's frames of every step
                    g = torch.Generator().manual_seed(1000 * i + r)
                    frames[(i, r)] = torch.randint(0, 256, (n_per, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev)
            torch.cuda.synchronize()
            for i in range(steps):
                with lanes.lane(i):
                    own = gat.own_slice(i)
                    if i % 2 == 0:
                        vit.forward_u8(frames[(i, rank)], out=own, fanout=gat.fanout(i))
                    else:  # the host-frame path, two batches per step: the fan-out addresses follow the batch offsets
                        pipeline.extract_features(vit, frames[(i, rank)].cpu(), batch_size=16, device=dev, out=own,
                                                  fanout=gat.fanout(i))
                    gat.publish(i)
                    fulls.append(gat.buffer(i).clone())  # read on the lane that will run step i + depth (the slot protocol)
            lanes.join()
            gat.wait_all()
            torch.cuda.synchronize()
            for i in range(steps):
                for r in range(world):
                    want = vit.forward_u8(frames[(i, r)])
                    ok = ok and bool(torch.equal(fulls[i][r * n_per:(r + 1) * n_per], want))
            torch.cuda.synchronize()
            dist.barrier()
        q.put((rank, ok, modes))
    finally:
        dist.destroy_process_group()


def test_peer_gatherer_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = sorted(q.get() for _ in range(2))
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
    assert res[0][2][1] == "peer-stores"  # (the first mode is "multicast" where the box has NVLS multicast)
