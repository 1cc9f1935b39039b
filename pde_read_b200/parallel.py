"""Point sharding across the GPUs of one box (one process per GPU, torch.distributed).

Residuals at different collocation points are independent; the only cross-point operations on the
hot path are the mean in Collocation_Loss (reference Loss_Functions.py:65), the parameter-gradient
sums of its backward, and the Gram sums of extraction.  So every rank takes a contiguous slice of
the points with replicated weights and the exchange is ONE small all-reduce:

  * discovery : float32 [dU | dN | sum(R^2) as a (hi, lo) float pair]   (84 KB .. 330 KB)
  * extraction: float64 [G | A^T b | b^T b]                               (3 KB .. 508 KB)

These payloads are latency-bound on NVLink/NVSwitch.  The discovery exchange -- once per training step -- runs on
the library's own one-shot all-reduce over peer memory when every rank sits on one box (PeerAllReduce below:
each rank writes its buffer into every peer's receive area and sums its own area in rank order; one kernel, no
NCCL on the step), and on NCCL's all-reduce otherwise; the extraction sums, once per run, stay on NCCL.  The
backend is whatever the process group was created with: "nccl" on GPUs, "gloo" in the CPU tests of this logic.
"""
from __future__ import annotations

import ctypes
import os
import socket
import zlib
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


class PeerAllReduce:
    """In-place sum of a float32 CUDA buffer over the ranks of one box through peer memory
    (csrc/peer.cu, pderead_peer_allreduce).  Construction is a collective: every rank allocates its receive area,
    the CUDA IPC handles travel through the process group once, and all ranks agree on success; `ok` is False
    (on every rank alike) when the ranks span several hosts, peer mapping fails, or PDEREAD_PEER_ALLREDUCE=0."""

    def __init__(self, group, cap_floats: int):
        from . import _lib
        self.ok = False
        self.group, self.cap = group, int(cap_floats + 3) // 4 * 4
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._lib, self._areas, self._local = _lib.load(), [], None
        dev = torch.device("cuda", torch.cuda.current_device())
        want = os.environ.get("PDEREAD_PEER_ALLREDUCE", "1") != "0" and 1 < self.world <= 32
        handle = (ctypes.c_ubyte * 64)()
        local = ctypes.c_void_p()
        good = want
        if good:
            good = self._lib.pderead_peer_alloc(self.world, self.cap, ctypes.byref(local), handle) == 0
        # one exchange carries the handle, the host name (hashed) and this rank's verdict so far
        host = zlib.crc32(socket.gethostname().encode()) & 0xFFFFFF
        # (an all-reduce of a one-hot [world, 68] table, not an all-gather: NCCL_ALGO=Tree has no all-gather)
        table = torch.zeros(self.world, 68, dtype=torch.int32, device=dev)
        table[self.rank] = torch.tensor(list(bytes(handle)) + [host >> 16 & 255, host >> 8 & 255, host & 255, 1 if good else 0],
                                        dtype=torch.int32, device=dev)
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
        rows = [bytes(r) for r in table.cpu().tolist()]
        good = good and all(r[67] == 1 for r in rows) and all(r[64:67] == rows[0][64:67] for r in rows)
        if good:
            for r, row in enumerate(rows):
                if r == self.rank:
                    self._areas.append(local.value)
                    continue
                p = ctypes.c_void_p()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(row[:64])
                if self._lib.pderead_peer_open(hb, ctypes.byref(p)) != 0:
                    good = False
                    break
                self._areas.append(p.value)
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(int(flag.item()))
        self._local = local.value
        if self.ok:
            self._table = torch.tensor(self._areas, dtype=torch.int64, device=dev)
        else:
            # not usable (on every rank alike): give the mappings and the area back, the caller stays on NCCL
            for r, p in enumerate(self._areas):
                if r != self.rank and p:
                    self._lib.pderead_peer_close(ctypes.c_void_p(p))
            if local.value:
                self._lib.pderead_peer_free(local)
            self._areas, self._local = [], None

    def allreduce_(self, flat: torch.Tensor) -> None:
        from . import _lib
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        _lib.check(self._lib.pderead_peer_allreduce(flat.data_ptr(), flat.numel(), self._table.data_ptr(), self._local,
                                                    self.rank, self.world, self.cap, stream), "pderead_peer_allreduce")


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
        self._peer = None           # PeerAllReduce, built at the first gradient exchange (a collective, like that exchange)

    def _sum_flat(self, flat: torch.Tensor) -> None:
        """In-place sum of the packed float32 buffer over the ranks: peer memory on one box, else NCCL / gloo.  Which one
        is decided collectively (PeerAllReduce.ok is the same on every rank; buffer sizes are the same on every rank)."""
        if flat.is_cuda and dist.get_backend(self.group) == "nccl":
            n = flat.numel()
            if self._peer is None or (self._peer.ok and n > self._peer.cap):
                self._peer = PeerAllReduce(self.group, max(n, 1 << 17))
            if self._peer.ok and n % 4 == 0 and flat.data_ptr() % 16 == 0:
                self._peer.allreduce_(flat)
                return
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

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
            self._sum_flat(flat)
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
