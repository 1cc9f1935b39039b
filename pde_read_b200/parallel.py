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
    Extraction.Extract_PDE.

    No rank ever decides locally whether a collective happens: every call below issues exactly the same
    sequence of all-reduces on every rank, and the global point count is never cached -- it travels with the
    data.  A rank computes the gradient of its UNSCALED sum(R^2) (``M_total = 0`` in the C ABI); the point count
    comes back in the tail of the same all-reduce and the 1 / count of the mean is applied on the device."""

    TAIL = 4      # floats at the end of the packed gradient buffer: sum(R^2) hi, lo | count >> 12, count & 4095

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def total_points(self, local_points: int) -> int:
        """Sum of the ranks' local point counts.  Always a collective (every rank must call it); returns a
        host integer, so it synchronises -- the training path uses reduce() instead."""
        t = torch.tensor([int(local_points)], dtype=torch.int64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def _device(self) -> torch.device:
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def reduce(self, flat: Optional[torch.Tensor], sum_sq: torch.Tensor, local_points: int) -> torch.Tensor:
        """Sum `flat` (float32 gradients of the unscaled sum(R^2), with TAIL spare floats at the end) and the
        float64 scalar `sum_sq` over the ranks in ONE all-reduce.  Returns a float64 tensor [3] on the device of
        `sum_sq`: (global sum(R^2), global point count, 1 / global count).  With ``flat is None``
        (forward-only) a float64 pair is reduced instead."""
        M_as = 1
        if flat is None:
            buf = torch.cat([sum_sq.to(torch.float64).reshape(1),
                             torch.tensor([float(local_points)], dtype=torch.float64, device=sum_sq.device)])
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            return torch.cat([buf, (float(M_as) / buf[1:2])])
        tail = flat[-self.TAIL:]
        out = torch.empty(3, dtype=torch.float64, device=flat.device)
        if flat.is_cuda:
            from . import _lib
            lib = _lib.load()
            stream = torch.cuda.current_stream(flat.device).cuda_stream
            ss = sum_sq if (sum_sq.dtype == torch.float64 and sum_sq.is_cuda) else sum_sq.to(flat.device, torch.float64)
            _lib.check(lib.pderead_loss_tail_pack(ss.data_ptr(), int(local_points), tail.data_ptr(), stream),
                       "pderead_loss_tail_pack")
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(lib.pderead_loss_tail_unpack(tail.data_ptr(), M_as, out.data_ptr(), stream),
                       "pderead_loss_tail_unpack")
            return out
        # CPU tensors (gloo tests of this logic): the same packing with torch ops
        ss = sum_sq.to(torch.float64).reshape(1)
        hi = ss.to(torch.float32)
        tail[0:1] = hi
        tail[1:2] = (ss - hi.to(torch.float64)).to(torch.float32)
        tail[2] = float(int(local_points) >> 12)
        tail[3] = float(int(local_points) & 4095)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        out[0] = tail[0].double() + tail[1].double()
        out[1] = tail[2].double() * 4096.0 + tail[3].double()
        out[2] = float(M_as) / out[1]
        return out

    def reduce_gram(self, G: torch.Tensor, Atb: torch.Tensor, btb: torch.Tensor):
        C = G.shape[0]
        buf = torch.cat([G.reshape(-1), Atb.reshape(-1), btb.reshape(-1)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf[:C * C].view(C, C), buf[C * C:C * C + C], buf[C * C + C:]

    def reduce_moments(self, moments: torch.Tensor) -> torch.Tensor:
        """Sum a small float64 vector (count, sums and sums of squares of the batch statistics of a
        Batch_Norm=True PDE network, Network.py:100-104) over the ranks."""
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=self.group)
        return moments


def all_reduce_grads(params, reducer: "ShardReducer") -> None:
    """Sum the .grad of `params` over the ranks in one all-reduce (general autograd path, e.g. a Batch_Norm=True PDE
    network; the fused collocation path reduces its flat buffer itself)."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=reducer.group)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def sharded_loss_and_grads(local_fn: Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                           coords: torch.Tensor, reducer: ShardReducer) -> Tuple[float, torch.Tensor]:
    """Reference semantics of a sharded collocation step, independent of where the arithmetic runs.

    `local_fn(shard)` returns (sum of R^2 over the shard as float64[1], flat float32 gradient of that
    unscaled sum with ShardReducer.TAIL spare floats at the end).  On GPUs `local_fn` is the fused CUDA call;
    the CPU tests inject the oracle.  Returns (global mean(R^2), reduced flat gradient of the global mean).
    """
    shard = shard_points(coords, reducer.rank, reducer.world_size)
    ss, flat = local_fn(shard)
    out = reducer.reduce(flat, ss, shard.shape[0])
    flat[:-ShardReducer.TAIL] *= out[2].to(torch.float32)
    return float(out[0].item()) / float(out[1].item()), flat
