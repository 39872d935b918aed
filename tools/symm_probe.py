"""Feasibility probe (dev tool, N >= 2 GPUs under torchrun): torch symmetric memory on this box — peer-mapped buffer
pointers, NVLS multicast support, a stream-ordered barrier, and a P2P store round trip."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
buf = symm.empty((4, 1024), dtype=torch.float32, device=dev)
buf.fill_(float(rank))
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok; world", hdl.world_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs],
      "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0, "signal pad bytes", hdl.signal_pad_size, flush=True)
hdl.barrier(channel=0, timeout_ms=5000)
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (4, 1024), torch.float32)
remote[rank].fill_(100.0 + rank)          # P2P store into the peer's row `rank`
hdl.barrier(channel=1, timeout_ms=5000)
torch.cuda.synchronize()
src = (rank - 1) % world
print(rank, "row written by", src, "=", float(buf[src, 0]), "(expect", 100.0 + src, ")", flush=True)
dist.barrier()
dist.destroy_process_group()
