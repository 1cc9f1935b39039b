"""Dev tool (GPUs, torchrun): the peer-memory all-reduce against NCCL on random buffers, many epochs, and its latency."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pde_read_b200.parallel import PeerAllReduce   # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
peer = PeerAllReduce(None, 1 << 17)
if rank == 0:
    print("peer ok:", peer.ok, flush=True)
if not peer.ok:
    sys.exit(0)
g = torch.Generator(device=dev).manual_seed(100 + rank)
bad = 0
for it in range(200):
    n = [4, 21000, 84004, 131072][it % 4]
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    # the exact sum in rank order, for the bit-identity claim
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    want = parts[0].clone()
    for p in parts[1:]:
        want += p
    peer.allreduce_(x)
    if not torch.equal(x, want):
        bad += 1
    if (x - ref).abs().max().item() > 1e-5 * max(1.0, ref.abs().max().item()):
        bad += 1
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
x = torch.randn(84004, device=dev)
for name, fn in (("peer", lambda: peer.allreduce_(x)), ("nccl", lambda: dist.all_reduce(x))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("%s all-reduce of 84004 floats on %d GPUs: %.1f us" % (name, world, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
if rank == 0:
    print("mismatches over 200 epochs (all ranks):", int(t.item()), flush=True)
dist.destroy_process_group()
