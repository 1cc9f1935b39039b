"""Dev tool (torchrun, N GPUs): torch.profiler table of a few sharded C2 steps on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench                                                     # noqa: E402
from pde_read_b200 import Loss_Functions, functional as F         # noqa: E402
from pde_read_b200.Loss_Functions import Collocation_Loss         # noqa: E402
from pde_read_b200.parallel import ShardReducer                   # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
c = dict(bench.CONFIGS["C2"])
U, N, _, _ = bench.build_nets(c, dev)
red = ShardReducer()
Loss_Functions.set_reducer(red)
P = F.generate_points(np.asarray(c["box"], dtype=np.float64), c["M"], dev, seed=1000 + rank)
params = list(U.parameters()) + list(N.parameters())


def step():
    for p in params:
        p.grad = None
    loss = Collocation_Loss(U, N, c["m"], c["n"], P, torch.float32, dev)
    loss.backward()
    return loss


for _ in range(10):
    step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("ms/step (50 steps, no profiler): %.4f" % (e0.elapsed_time(e1) / 50))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
dist.destroy_process_group()
