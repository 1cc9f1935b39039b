"""Point sharding across the GPUs of one box (one process per GPU, torch.distributed).

Residuals at different collocation points are independent; the only cross-point operations on the
hot path are the mean in Collocation_Loss (reference Loss_Functions.py:65), the parameter-gradient
sums of its backward, and the Gram sums of extraction.  So every rank takes a contiguous slice of
the points with replicated weights and the exchange is ONE small all-reduce:

  * discovery : float32 [dU | dN | sum(R^2) as a (hi, lo) float pair]   (84 KB .. 330 KB)
  * extraction: float64 [G | A^T b | b^T b]                               (3 KB .. 508 KB)

These payloads are latency-bound on NVLink/NVSwitch, so NCCL's all-reduce is used as is (there is
no compute step to fuse the transfer with tile by tile).  The backend is whatever the process
group was created with: "nccl" on GPUs, "gloo" in the CPU tests of this logic.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_points: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`; the first (num_points % world_size) ranks get one
    extra point, so slices differ in length by at most one and cover every point exactly once."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    base, extra = divmod(num_points, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_points(coords: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(coords.shape[0], rank, world_size)
    return coords[lo:hi].contiguous()


class ShardReducer:
    """All-reduce helper handed to Collocation_Loss (via Loss_Functions.set_reducer) and to
    Extraction.Extract_PDE."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._totals = {}

    def total_points(self, local_points: int) -> int:
        """Sum of the ranks' local point counts (cached per local count: a training run keeps its
        shard size, so this costs one tiny all-reduce once)."""
        key = int(local_points)
        if key not in self._totals:
            t = torch.tensor([key], dtype=torch.int64)
            dev = self._device()
            t = t.to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._totals[key] = int(t.item())
        return self._totals[key]

    def _device(self) -> torch.device:
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def reduce(self, flat: Optional[torch.Tensor], sum_sq: torch.Tensor) -> torch.Tensor:
        """Sum `flat` (float32 gradients with >= 2 spare floats at the end) and the float64 scalar
        `sum_sq` over the ranks in one all-reduce; returns the global sum_sq (float64, shape [1]).
        The scalar rides in the tail as a (hi, lo) float pair, which keeps ~48 bits of it."""
        if flat is None:
            buf = sum_sq.to(torch.float64).reshape(1).clone()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            return buf
        ss = sum_sq.to(torch.float64).reshape(1)
        hi = ss.to(torch.float32)
        lo = (ss - hi.to(torch.float64)).to(torch.float32)
        flat[-2:-1] = hi
        flat[-1:] = lo
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat[-2:-1].to(torch.float64) + flat[-1:].to(torch.float64)

    def reduce_gram(self, G: torch.Tensor, Atb: torch.Tensor, btb: torch.Tensor):
        C = G.shape[0]
        buf = torch.cat([G.reshape(-1), Atb.reshape(-1), btb.reshape(-1)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf[:C * C].view(C, C), buf[C * C:C * C + C], buf[C * C + C:]


def sharded_loss_and_grads(local_fn: Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                           coords: torch.Tensor, reducer: ShardReducer) -> Tuple[float, torch.Tensor]:
    """Reference semantics of a sharded collocation step, independent of where the arithmetic runs.

    `local_fn(shard, M_total)` returns (sum of R^2 over the shard as float64[1], flat float32
    gradient of sum(R^2)/M_total with >= 2 spare floats at the end).  On GPUs `local_fn` is the fused
    CUDA call; the CPU tests inject the oracle.  Returns (global mean(R^2), reduced flat gradient).
    """
    shard = shard_points(coords, reducer.rank, reducer.world_size)
    M_total = coords.shape[0]
    ss, flat = local_fn(shard, M_total)
    ss = reducer.reduce(flat, ss)
    return float(ss.item()) / float(M_total), flat
