"""Tensor-level host glue between PyTorch and the C ABI of libpderead_b200.so.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspaces), provides the
stream, and carries the autograd edges so that ``Loss.backward()`` in the reference's training
closure (reference Test_Train.py:73-102) reaches the hand-written reverse kernels.  All arithmetic
of the hot path happens inside the library; nothing in this file computes on the CPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_CODES, PdereadError, pderead_net, pderead_net_grad


# ----------------------------------------------------------------------------------------------
# network description
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class NetSpec:
    num_hidden: int
    width: int
    input_dim: int
    activation: str          # "Tanh" | "Sin" | "Rational"

    @property
    def num_tensors(self) -> int:
        return 2 * (self.num_hidden + 1) + (2 * self.num_hidden if self.activation == "Rational" else 0)


def net_spec(net, allow_input_norm: bool = False) -> NetSpec:
    """Spec of a pde_read_b200.Network.Neural_Network (or anything with the same attributes).

    The fused kernels evaluate the Linear / activation stack.  A network with ``Batch_Norm=True`` (the
    reference's "PDE Network - Normalize Inputs", Network.py:100-104,194-195) normalises its INPUTS with a
    torch BatchNorm1d first; callers that apply that layer themselves pass ``allow_input_norm=True``."""
    if getattr(net, "Batch_Norm", False) and not allow_input_norm:
        raise NotImplementedError("this entry point fuses U and N and cannot insert the input BatchNorm of a "
                                  "Batch_Norm=True network; use the unfused path (Network.forward)")
    if net.Output_Dim != 1:
        raise NotImplementedError("the CUDA path supports Output_Dim == 1 only (got %d)" % net.Output_Dim)
    return NetSpec(net.Num_Hidden_Layers, net.Layers[0].out_features, net.Input_Dim, net.Activation_Name)


def net_tensors(net) -> List[torch.Tensor]:
    """Parameter tensors in the fixed order the autograd functions use:
    w0, b0, ..., wL, bL, then (Rational only) a0, b0, ..., a_{L-1}, b_{L-1}."""
    out: List[torch.Tensor] = []
    for lin in net.Layers:
        out += [lin.weight, lin.bias]
    if net.Activation_Name == "Rational":
        for act in net.Activation_Functions:
            out += [act.a, act.b]
    return out


def _check_tensor(t: torch.Tensor, what: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise PdereadError("%s must live on a CUDA device (there is no CPU path); got %s" % (what, t.device))
    if t.dtype != dtype:
        raise PdereadError("%s must be %s, got %s" % (what, dtype, t.dtype))
    return t.contiguous()


def _check_param(t: torch.Tensor, what: str) -> torch.Tensor:
    """Network parameters are passed by pointer and the pointer table is cached (_fill_net): a non-contiguous
    parameter would be copied once and the kernels would keep reading the stale copy after an in-place
    optimiser step, so it is refused instead."""
    if not t.is_contiguous():
        raise PdereadError("%s must be contiguous (parameters are passed to the kernels by pointer)" % what)
    return _check_tensor(t, what)


_NET_CACHE: dict = {}


def _fill_net(spec: NetSpec, tensors: Sequence[torch.Tensor]) -> Tuple[pderead_net, list]:
    """ctypes description of a network.  Building it touches every parameter tensor, so the result is cached by
    (spec, device pointers): optimisers update parameters in place and a training loop hits the cache every step."""
    key = (spec, tuple([t.data_ptr() for t in tensors]))
    hit = _NET_CACHE.get(key)
    if hit is not None:
        return hit
    out = _fill_net_uncached(spec, tensors)
    if len(_NET_CACHE) > 64:
        _NET_CACHE.clear()
    _NET_CACHE[key] = out
    return out


def _fill_net_uncached(spec: NetSpec, tensors: Sequence[torch.Tensor]) -> Tuple[pderead_net, list]:
    if len(tensors) != spec.num_tensors:
        raise PdereadError("expected %d parameter tensors, got %d" % (spec.num_tensors, len(tensors)))
    if spec.num_hidden < 1 or spec.num_hidden > _lib.MAX_HIDDEN_LAYERS:
        raise PdereadError("Num_Hidden_Layers must be in 1..%d" % _lib.MAX_HIDDEN_LAYERS)
    s = pderead_net()
    s.num_hidden_layers = spec.num_hidden
    s.width = spec.width
    s.input_dim = spec.input_dim
    s.activation = ACT_CODES[spec.activation]
    keep = []
    L = spec.num_hidden
    for i in range(L + 1):
        w = _check_param(tensors[2 * i].detach(), "Layers.%d.weight" % i)
        b = _check_param(tensors[2 * i + 1].detach(), "Layers.%d.bias" % i)
        exp_w = (spec.width if i < L else 1, spec.input_dim if i == 0 else spec.width)
        if tuple(w.shape) != exp_w or b.numel() != exp_w[0]:
            raise PdereadError("Layers.%d has shape %s / %s, expected %s" % (i, tuple(w.shape), tuple(b.shape), exp_w))
        keep += [w, b]
        s.weight[i] = w.data_ptr()
        s.bias[i] = b.data_ptr()
    if spec.activation == "Rational":
        base = 2 * (L + 1)
        for i in range(L):
            a = _check_param(tensors[base + 2 * i].detach(), "Activation_Functions.%d.a" % i)
            b = _check_param(tensors[base + 2 * i + 1].detach(), "Activation_Functions.%d.b" % i)
            if a.numel() != 4 or b.numel() != 3:
                raise PdereadError("Rational coefficients must have 4 and 3 entries")
            keep += [a, b]
            s.rat_a[i] = a.data_ptr()
            s.rat_b[i] = b.data_ptr()
    return s, keep


class _LazyViews:
    """The per-parameter gradient views of a flat buffer, created on first use (a training step only needs
    the flat buffer; creating 32 view tensors per call costs more host time than launching the kernels)."""

    def __init__(self, flat: torch.Tensor, layout):
        self._flat, self._layout, self._views = flat, layout, None

    def _make(self):
        if self._views is None:
            self._views = [self._flat.as_strided(shape, stride, off) for (off, shape, stride) in self._layout]
        return self._views

    def __iter__(self):
        return iter(self._make())

    def __len__(self):
        return len(self._layout)

    def __getitem__(self, i):
        return self._make()[i]

    def __add__(self, other):
        return list(self) + list(other)

    def __radd__(self, other):
        return list(other) + list(self)


def _contig_strides(shape):
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    return tuple(reversed(st))


def _grad_views(spec: NetSpec, tensors: Sequence[torch.Tensor], flat: torch.Tensor, offset: int = 0
                ) -> Tuple[pderead_net_grad, "_LazyViews", int]:
    """One gradient slot per parameter tensor inside `flat` (float32, device) starting at `offset` (every slot
    16-byte aligned); a pderead_net_grad pointing at the slots and lazily created views of them."""
    g = pderead_net_grad()
    L = spec.num_hidden
    base_ptr = flat.data_ptr()
    off = offset
    ptrs, layout = [], []
    for t in tensors:
        shape = tuple(t.shape)
        n = t.numel()
        ptrs.append(base_ptr + 4 * off)
        layout.append((off, shape, _contig_strides(shape)))
        off += (n + 3) // 4 * 4
    for i in range(L + 1):
        g.weight[i] = ptrs[2 * i]
        g.bias[i] = ptrs[2 * i + 1]
    if spec.activation == "Rational":
        b0 = 2 * (L + 1)
        for i in range(L):
            g.rat_a[i] = ptrs[b0 + 2 * i]
            g.rat_b[i] = ptrs[b0 + 2 * i + 1]
    return g, _LazyViews(flat, layout), off


def flat_grad_numel(tensors: Sequence[torch.Tensor]) -> int:
    return sum((t.numel() + 3) // 4 * 4 for t in tensors)


# ----------------------------------------------------------------------------------------------
# workspace / stream
# ----------------------------------------------------------------------------------------------
_workspaces = {}

# Test hook: when set, workspaces are sized for at most this many points per chunk, which forces
# the library's multi-chunk path on small inputs.
WORKSPACE_CAP_POINTS: Optional[int] = None


def _ws_points(M: int) -> int:
    return M if WORKSPACE_CAP_POINTS is None else min(M, int(WORKSPACE_CAP_POINTS))


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """A per-device scratch buffer owned by PyTorch's allocator, grown on demand."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _workspaces.get(key)
    if WORKSPACE_CAP_POINTS is not None and buf is not None and buf.numel() != max(int(nbytes), 256):
        buf = None
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)
        floor = 0 if WORKSPACE_CAP_POINTS is not None else (1 << 20)
        buf = torch.empty(max(int(nbytes), floor, 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------
# raw calls (no autograd)
# ----------------------------------------------------------------------------------------------
def sol_derivatives_forward(spec: NetSpec, tensors, m: int, n: int, coords: torch.Tensor
                            ) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    lib = _lib.load()
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    with torch.cuda.device(coords.device):
        s, keep = _fill_net(spec, tensors)
        dxn = torch.empty((M, n + 1), dtype=torch.float32, device=coords.device)
        dtm = torch.empty((M,), dtype=torch.float32, device=coords.device) if m > 0 else None
        if M == 0:
            return dtm, dxn
        nb = lib.pderead_sol_derivatives_workspace_bytes(ctypes.byref(s), m, n, M, 0)
        ws = workspace(coords.device, nb)
        _lib.check(lib.pderead_sol_derivatives_forward(ctypes.byref(s), m, n, coords.data_ptr(), M, _ptr(dtm),
                                                       dxn.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "pderead_sol_derivatives_forward")
    return dtm, dxn


def sol_derivatives_backward(spec: NetSpec, tensors, m: int, n: int, coords: torch.Tensor,
                             grad_dtm: Optional[torch.Tensor], grad_dxn: Optional[torch.Tensor]) -> List[torch.Tensor]:
    lib = _lib.load()
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    with torch.cuda.device(coords.device):
        s, keep = _fill_net(spec, tensors)
        flat = torch.zeros(flat_grad_numel(tensors), dtype=torch.float32, device=coords.device)
        g, views, _ = _grad_views(spec, tensors, flat)
        if M == 0:
            return views
        gd = _check_tensor(grad_dtm, "grad Dtm_U") if grad_dtm is not None else None
        gx = _check_tensor(grad_dxn, "grad Dxn_U") if grad_dxn is not None else None
        nb = lib.pderead_sol_derivatives_workspace_bytes(ctypes.byref(s), m, n, _ws_points(M), 1)
        ws = workspace(coords.device, nb)
        _lib.check(lib.pderead_sol_derivatives_backward(ctypes.byref(s), m, n, coords.data_ptr(), M, _ptr(gd), _ptr(gx),
                                                        ctypes.byref(g), 0.0, ws.data_ptr(), ws.numel(), _stream()),
                   "pderead_sol_derivatives_backward")
    return views


def _pick_ws(device, nb: int, private_ws: Optional[list]) -> torch.Tensor:
    """The shared growable workspace, or (CUDA-graph capture: the graph keeps raw pointers) a buffer owned by the
    caller's list that is never reallocated once the graph exists."""
    if private_ws is None:
        return workspace(device, nb)
    if not private_ws or private_ws[0].numel() < nb:
        private_ws[:] = [torch.empty(max(int(nb), 256), dtype=torch.uint8, device=device)]
    return private_ws[0]


def mlp_forward(spec: NetSpec, tensors, X: torch.Tensor, private_ws: Optional[list] = None) -> torch.Tensor:
    lib = _lib.load()
    X = _check_tensor(X.detach(), "network input")
    if X.dim() != 2 or X.shape[1] != spec.input_dim:
        raise PdereadError("network input must be [B, %d], got %s" % (spec.input_dim, tuple(X.shape)))
    M = X.shape[0]
    with torch.cuda.device(X.device):
        s, keep = _fill_net(spec, tensors)
        out = torch.empty((M,), dtype=torch.float32, device=X.device)
        if M == 0:
            return out
        nb = lib.pderead_mlp_workspace_bytes(ctypes.byref(s), M, 0)
        ws = _pick_ws(X.device, nb, private_ws)
        _lib.check(lib.pderead_mlp_forward(ctypes.byref(s), X.data_ptr(), M, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _stream()), "pderead_mlp_forward")
    return out


def mlp_backward(spec: NetSpec, tensors, X: torch.Tensor, grad_out: torch.Tensor, need_grad_in: bool
                 ) -> Tuple[Optional[torch.Tensor], List[torch.Tensor]]:
    lib = _lib.load()
    X = _check_tensor(X.detach(), "network input")
    M = X.shape[0]
    with torch.cuda.device(X.device):
        s, keep = _fill_net(spec, tensors)
        flat = torch.zeros(flat_grad_numel(tensors), dtype=torch.float32, device=X.device)
        g, views, _ = _grad_views(spec, tensors, flat)
        gin = torch.zeros_like(X) if need_grad_in else None
        if M == 0:
            return gin, views
        go = _check_tensor(grad_out.reshape(-1), "grad of network output")
        nb = lib.pderead_mlp_workspace_bytes(ctypes.byref(s), _ws_points(M), 1)
        ws = workspace(X.device, nb)
        _lib.check(lib.pderead_mlp_backward(ctypes.byref(s), X.data_ptr(), M, go.data_ptr(), _ptr(gin), ctypes.byref(g),
                                            0.0, ws.data_ptr(), ws.numel(), _stream()), "pderead_mlp_backward")
    return gin, views


def mlp_backward_accumulate(spec: NetSpec, tensors, X: torch.Tensor, grad_out: torch.Tensor, flat: torch.Tensor,
                            offset: int = 0, private_ws: Optional[list] = None) -> None:
    """Reverse pass of the plain MLP whose parameter gradients are ADDED (beta = 1) to the slots of this network
    inside an existing flat gradient buffer (layout of _grad_views at `offset`)."""
    lib = _lib.load()
    X = _check_tensor(X.detach(), "network input")
    M = X.shape[0]
    if M == 0:
        return
    with torch.cuda.device(X.device):
        s, keep = _fill_net(spec, tensors)
        g, _, _ = _grad_views(spec, tensors, flat, offset)
        go = _check_tensor(grad_out.reshape(-1), "grad of network output")
        nb = lib.pderead_mlp_workspace_bytes(ctypes.byref(s), _ws_points(M), 1)
        ws = _pick_ws(X.device, nb, private_ws)
        _lib.check(lib.pderead_mlp_backward(ctypes.byref(s), X.data_ptr(), M, go.data_ptr(), None, ctypes.byref(g),
                                            1.0, ws.data_ptr(), ws.numel(), _stream()), "pderead_mlp_backward")


def residual_forward(sol_spec, sol_tensors, pde_spec, pde_tensors, m: int, n: int, coords: torch.Tensor,
                     want_resid: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """-> (R [M] or None, sum(R^2) as a float64 device scalar)."""
    lib = _lib.load()
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    with torch.cuda.device(coords.device):
        su, k1 = _fill_net(sol_spec, sol_tensors)
        sn, k2 = _fill_net(pde_spec, pde_tensors)
        R = torch.empty((M,), dtype=torch.float32, device=coords.device) if want_resid else None
        ss = torch.zeros((1,), dtype=torch.float64, device=coords.device)
        if M == 0:
            return R, ss
        nb = lib.pderead_residual_workspace_bytes(ctypes.byref(su), ctypes.byref(sn), m, n, _ws_points(M), 0)
        ws = workspace(coords.device, nb)
        _lib.check(lib.pderead_residual_forward(ctypes.byref(su), ctypes.byref(sn), m, n, coords.data_ptr(), M, _ptr(R),
                                                ss.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "pderead_residual_forward")
    return R, ss


def collocation_loss_grad(sol_spec, sol_tensors, pde_spec, pde_tensors, m: int, n: int, coords: torch.Tensor,
                          M_total: Optional[int] = None, grad_resid: Optional[torch.Tensor] = None,
                          want_resid: bool = False, extra_tail: int = 0, private_ws: Optional[list] = None):
    """One fused call: sum(R^2) over these points and d(sum(R^2)/M_total)/d(parameters).

    Returns (sum_sq float64[1], flat float32 gradient buffer, sol grad views, pde grad views,
    R or None).  `flat` holds the U gradients, then the N gradients, then `extra_tail` spare
    floats (used by the multi-GPU path to ride the loss on the same all-reduce).
    """
    lib = _lib.load()
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    if M_total is None:
        M_total = M                     # (0 = gradients of the unscaled sum(R^2): the sharded path)
    with torch.cuda.device(coords.device):
        su, k1 = _fill_net(sol_spec, sol_tensors)
        sn, k2 = _fill_net(pde_spec, pde_tensors)
        nU, nN = flat_grad_numel(sol_tensors), flat_grad_numel(pde_tensors)
        flat = torch.zeros(nU + nN + extra_tail, dtype=torch.float32, device=coords.device)
        gu, vu, off = _grad_views(sol_spec, sol_tensors, flat, 0)
        gn, vn, off = _grad_views(pde_spec, pde_tensors, flat, off)
        ss = torch.zeros((1,), dtype=torch.float64, device=coords.device)
        R = torch.empty((M,), dtype=torch.float32, device=coords.device) if want_resid else None
        if M == 0:
            return ss, flat, vu, vn, R
        gr = _check_tensor(grad_resid.reshape(-1), "grad of residual") if grad_resid is not None else None
        nb = lib.pderead_residual_workspace_bytes(ctypes.byref(su), ctypes.byref(sn), m, n, _ws_points(M), 1)
        ws = _pick_ws(coords.device, nb, private_ws)
        _lib.check(lib.pderead_collocation_loss_grad(ctypes.byref(su), ctypes.byref(sn), m, n, coords.data_ptr(), M,
                                                     int(M_total), _ptr(gr), _ptr(R), ss.data_ptr(), ctypes.byref(gu),
                                                     ctypes.byref(gn), 0.0, ws.data_ptr(), ws.numel(), _stream()),
                   "pderead_collocation_loss_grad")
    return ss, flat, vu, vn, R


# ----------------------------------------------------------------------------------------------
# CUDA-graph replay of the fused step (launch-bound when the point set is small, e.g. the reference's
# default 5000 collocation points: ~10 launches of ~30 us of work each)
# ----------------------------------------------------------------------------------------------
import os as _os

USE_CUDA_GRAPHS = _os.environ.get("PDEREAD_CUDA_GRAPHS", "0") not in ("0", "", "off")
_GRAPH_CACHE = {}
_GRAPH_CACHE_MAX = 8


def collocation_loss_grad_graphed(sol_spec, sol_tensors, pde_spec, pde_tensors, m: int, n: int, coords: torch.Tensor,
                                  M_total: Optional[int] = None, extra_tail: int = 0):
    """Same result as collocation_loss_grad(...)[:4] + (None,), but the kernel sequence is captured once
    into a CUDA graph per (networks, orders, point count) and replayed: one launch per step instead of ten.
    The parameter tensors are read in place at replay time, so optimiser updates are seen; a graph is rebuilt
    if any parameter tensor is replaced by a new one."""
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    if M_total is None:
        M_total = M
    dev = coords.device
    key = (dev.index, sol_spec, pde_spec, m, n, M, int(M_total), extra_tail,
           tuple(t.data_ptr() for t in sol_tensors), tuple(t.data_ptr() for t in pde_tensors))
    ent = _GRAPH_CACHE.get(key)
    if ent is None:
        if len(_GRAPH_CACHE) >= _GRAPH_CACHE_MAX:
            _GRAPH_CACHE.pop(next(iter(_GRAPH_CACHE)))
        static_coords = coords.clone()
        own_ws: list = []                                  # workspace owned by this graph (never reallocated)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up: allocates the workspace outside the capture
            for _ in range(2):
                collocation_loss_grad(sol_spec, sol_tensors, pde_spec, pde_tensors, m, n, static_coords,
                                      M_total=M_total, extra_tail=extra_tail, private_ws=own_ws)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ss, flat, vu, vn, _ = collocation_loss_grad(sol_spec, sol_tensors, pde_spec, pde_tensors, m, n,
                                                        static_coords, M_total=M_total, extra_tail=extra_tail,
                                                        private_ws=own_ws)
        ent = (graph, static_coords, ss, flat, [tuple(v.shape) for v in vu], [tuple(v.shape) for v in vn],
               list(sol_tensors) + list(pde_tensors) + own_ws)   # keep parameters and workspace alive with the graph
        _GRAPH_CACHE[key] = ent
    graph, static_coords, ss, flat, shu, shn, _keep = ent
    static_coords.copy_(coords)
    graph.replay()
    out_flat = flat.clone()                                  # the static buffers are overwritten by the next replay
    views, off = [], 0
    for shape in shu + shn:
        numel = 1
        for d in shape:
            numel *= d
        views.append(out_flat[off:off + numel].view(shape))
        off += (numel + 3) // 4 * 4
    return ss.clone(), out_flat, views[:len(shu)], views[len(shu):], None


def release_graphs() -> None:
    _GRAPH_CACHE.clear()


# ----------------------------------------------------------------------------------------------
# autograd functions
# ----------------------------------------------------------------------------------------------
class SolDerivativesFn(torch.autograd.Function):
    """(Dtm_U, Dxn_U) = Evaluate_Derivatives(...) with gradients w.r.t. the U parameters."""

    @staticmethod
    def forward(ctx, coords, spec, m, n, *tensors):
        dtm, dxn = sol_derivatives_forward(spec, tensors, m, n, coords)
        ctx.spec, ctx.m, ctx.n = spec, m, n
        ctx.save_for_backward(coords, *tensors)
        return dtm, dxn

    @staticmethod
    def backward(ctx, g_dtm, g_dxn):
        coords, *tensors = ctx.saved_tensors
        views = sol_derivatives_backward(ctx.spec, tensors, ctx.m, ctx.n, coords, g_dtm, g_dxn)
        grads = [v if t.requires_grad else None for v, t in zip(views, tensors)]
        return (None, None, None, None, *grads)


class MLPFn(torch.autograd.Function):
    """out[M] = Neural_Network(X) for scalar-output nets, differentiable in X and the parameters."""

    @staticmethod
    def forward(ctx, X, spec, *tensors):
        out = mlp_forward(spec, tensors, X)
        ctx.spec = spec
        ctx.save_for_backward(X, *tensors)
        return out

    @staticmethod
    def backward(ctx, g_out):
        X, *tensors = ctx.saved_tensors
        gin, views = mlp_backward(ctx.spec, tensors, X, g_out, ctx.needs_input_grad[0])
        grads = [v if t.requires_grad else None for v, t in zip(views, tensors)]
        return (gin, None, *grads)


class ResidualFn(torch.autograd.Function):
    """R[M] = PDE_Residual(...); backward runs the fused reverse pass with the incoming dL/dR."""

    @staticmethod
    def forward(ctx, coords, sol_spec, pde_spec, m, n, n_sol, *tensors):
        sol_t, pde_t = tensors[:n_sol], tensors[n_sol:]
        R, _ = residual_forward(sol_spec, sol_t, pde_spec, pde_t, m, n, coords)
        ctx.meta = (sol_spec, pde_spec, m, n, n_sol)
        ctx.save_for_backward(coords, *tensors)
        return R

    @staticmethod
    def backward(ctx, g_R):
        sol_spec, pde_spec, m, n, n_sol = ctx.meta
        coords, *tensors = ctx.saved_tensors
        sol_t, pde_t = tensors[:n_sol], tensors[n_sol:]
        _, _, vu, vn, _ = collocation_loss_grad(sol_spec, sol_t, pde_spec, pde_t, m, n, coords,
                                                grad_resid=g_R.contiguous())
        grads = [v if t.requires_grad else None for v, t in zip(list(vu) + list(vn), tensors)]
        return (None, None, None, None, None, None, *grads)


class CollocationLossFn(torch.autograd.Function):
    """mean(R^2) with the parameter gradients computed in the same fused pass (the loss is a
    scalar, so the reverse pass can run eagerly and backward() only scales by the incoming grad).

    With a reducer (points sharded over ranks) every rank computes the gradient of its unscaled sum(R^2), the
    global point count comes back with the all-reduce, and backward() applies 1 / count on the device: no
    rank-local decision, no host synchronisation."""

    @staticmethod
    def forward(ctx, coords, sol_spec, pde_spec, m, n, n_sol, reducer, *tensors):
        sol_t, pde_t = tensors[:n_sol], tensors[n_sol:]
        M = coords.shape[0]
        # (grad mode is always off inside Function.forward; needs_input_grad is the real signal)
        need_grad = any(ctx.needs_input_grad[7:])
        M_total = 0 if reducer is not None else M          # 0: unscaled sums (C ABI)
        factor = None
        if need_grad:
            step = collocation_loss_grad_graphed if (USE_CUDA_GRAPHS and M > 0) else collocation_loss_grad
            ss, flat, vu, vn, _ = step(sol_spec, sol_t, pde_spec, pde_t, m, n, coords,
                                       M_total=M_total, extra_tail=reducer.TAIL if reducer is not None else 0)
            if reducer is not None:
                red = reducer.reduce(flat, ss, M)
                loss = (red[0] / red[1]).to(torch.float32).reshape(())
                factor = red[2].to(torch.float32)
                ctx.save_for_backward(flat, factor)
            else:
                loss = (ss / float(M_total)).to(torch.float32).reshape(())
                ctx.save_for_backward(flat)
            ctx.shapes = [tuple(t.shape) for t in tensors]
        else:
            _, ss = residual_forward(sol_spec, sol_t, pde_spec, pde_t, m, n, coords, want_resid=False)
            if reducer is not None:
                red = reducer.reduce(None, ss, M)
                loss = (red[0] / red[1]).to(torch.float32).reshape(())
            else:
                loss = (ss / float(max(M_total, 1))).to(torch.float32).reshape(())
        ctx.need_grad = need_grad
        ctx.has_factor = factor is not None
        ctx.flags = list(ctx.needs_input_grad[7:])
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.need_grad:
            return (None,) * (7 + len(ctx.flags))
        if ctx.has_factor:
            flat, factor = ctx.saved_tensors
            scaled = flat * (g * factor)
        else:
            (flat,) = ctx.saved_tensors
            scaled = flat * g          # one launch for all parameter tensors
        grads, off = [], 0
        for shape, f in zip(ctx.shapes, ctx.flags):
            numel = 1
            for d in shape:
                numel *= d
            grads.append(scaled.as_strided(shape, _contig_strides(shape), off) if f else None)
            off += (numel + 3) // 4 * 4
        return (None, None, None, None, None, None, None, *grads)


# ----------------------------------------------------------------------------------------------
# input normalisation of the PDE network (Batch_Norm=True; reference Network.py:100-104,194-195), two-phase
# ----------------------------------------------------------------------------------------------
def input_moments(X: torch.Tensor, reducer=None) -> torch.Tensor:
    """float64 [1 + 2 din]: point count, per-column sums and sums of squares of X, summed over the ranks when a
    reducer (pde_read_b200.parallel.ShardReducer) is given."""
    lib = _lib.load()
    X = _check_tensor(X.detach(), "network input")
    M, din = X.shape
    with torch.cuda.device(X.device):
        mom = torch.zeros(1 + 2 * din, dtype=torch.float64, device=X.device)
        _lib.check(lib.pderead_input_moments(X.data_ptr(), M, din, mom.data_ptr(), _stream()), "pderead_input_moments")
    if reducer is not None:
        mom = reducer.reduce_moments(mom)
    return mom


class FoldedNet:
    """A Batch_Norm=True network with its input normalisation folded into the first Linear layer for one call:
    `tensors` is the parameter list the MLP kernels run on (folded copies in slots 0 and 1)."""

    def __init__(self, net, moments: Optional[torch.Tensor], training: bool):
        lib = _lib.load()
        self.spec = net_spec(net, allow_input_norm=True)
        self.orig = net_tensors(net)
        bn = net.Norm_Layer
        dev = self.orig[0].device
        W, din = self.spec.width, self.spec.input_dim
        self.training = bool(training)
        self.moments = moments
        self.gamma, self.beta = bn.weight, bn.bias
        with torch.cuda.device(dev):
            s, keep = _fill_net_uncached(self.spec, self.orig)
            self.w0f = torch.empty((W, din), dtype=torch.float32, device=dev)
            self.b0f = torch.empty((W,), dtype=torch.float32, device=dev)
            self.stats = torch.empty((2 * din,), dtype=torch.float32, device=dev)
            momentum = 0.1 if bn.momentum is None else float(bn.momentum)
            track = bn.track_running_stats and bn.running_mean is not None
            if not self.training and not track:
                raise PdereadError("eval-mode input normalisation needs running statistics")
            _lib.check(lib.pderead_bn_fold(ctypes.byref(s), _ptr(self.gamma.detach() if self.gamma is not None else None),
                                           _ptr(self.beta.detach() if self.beta is not None else None),
                                           _ptr(moments), 1 if self.training else 0, momentum, float(bn.eps),
                                           _ptr(bn.running_mean) if track else None,
                                           _ptr(bn.running_var) if track else None,
                                           _ptr(bn.num_batches_tracked) if (track and self.training) else None,
                                           self.w0f.data_ptr(), self.b0f.data_ptr(), self.stats.data_ptr(), _stream()),
                       "pderead_bn_fold")
        self.tensors = [self.w0f, self.b0f] + list(self.orig[2:])

    def backward_fixup(self, flat_views, X: torch.Tensor, gin: Optional[torch.Tensor], reducer=None):
        """After the reverse pass of the folded network wrote its gradients into `flat_views` (slot 0 = gW0', slot 1 =
        gb0'): rewrite slot 0 to gW0 in place, return (d gamma, d beta) and add the batch-statistics term to `gin`."""
        lib = _lib.load()
        dev = X.device
        din = self.spec.input_dim
        with torch.cuda.device(dev):
            s, keep = _fill_net_uncached(self.spec, self.orig)
            gw0, gb0 = flat_views[0], flat_views[1]
            sums = torch.empty(2 * din, dtype=torch.float64, device=dev)
            _lib.check(lib.pderead_bn_grad_sums(ctypes.byref(s), gw0.data_ptr(), gb0.data_ptr(), sums.data_ptr(), _stream()),
                       "pderead_bn_grad_sums")
            gsums = sums
            if reducer is not None and self.training:
                gsums = reducer.reduce_moments(sums.clone())
            gg = torch.empty(din, dtype=torch.float32, device=dev)
            gb = torch.empty(din, dtype=torch.float32, device=dev)
            seed = torch.empty(2 * din, dtype=torch.float32, device=dev)
            _lib.check(lib.pderead_bn_backward(ctypes.byref(s), _ptr(self.gamma.detach() if self.gamma is not None else None),
                                               _ptr(self.beta.detach() if self.beta is not None else None),
                                               self.stats.data_ptr(), _ptr(self.moments), sums.data_ptr(), gsums.data_ptr(),
                                               1 if self.training else 0, gw0.data_ptr(), gb0.data_ptr(), gg.data_ptr(),
                                               gb.data_ptr(), seed.data_ptr(), _stream()), "pderead_bn_backward")
            if gin is not None and self.training:
                Xc = _check_tensor(X.detach(), "network input")
                _lib.check(lib.pderead_bn_input_grad(gin.data_ptr(), Xc.data_ptr(), Xc.shape[0], din, seed.data_ptr(),
                                                     _stream()), "pderead_bn_input_grad")
        return gg, gb


class ShardedMeanSquareFn(torch.autograd.Function):
    """mean(R^2) over the points of ALL ranks (R = this rank's residuals).  The value is all-reduced; the gradient
    w.r.t. R is 2 R / M_total, so after backward() every rank holds its share of the parameter gradients and
    ``parallel.all_reduce_grads`` sums them (the fused path without input normalisation does both in one packed
    all-reduce; this one is the general-autograd path)."""

    @staticmethod
    def forward(ctx, R, reducer):
        ss = (R.double() ** 2).sum().reshape(1)
        red = reducer.reduce(None, ss, R.shape[0])
        ctx.save_for_backward(R, red)
        return (red[0] / red[1]).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        R, red = ctx.saved_tensors
        return (R * (2.0 * g * red[2].to(torch.float32))), None


class BNInputMLPFn(torch.autograd.Function):
    """out[M] = Neural_Network(X) for a Batch_Norm=True network (train or eval mode), differentiable in X, the
    Linear / activation parameters and the normalisation layer's weight and bias."""

    @staticmethod
    def forward(ctx, X, net, reducer, gamma, beta, *tensors):
        training = net.training
        X = X.contiguous()
        mom = input_moments(X, reducer) if training else None
        if training and float(X.shape[0]) <= 1 and reducer is None:
            raise ValueError("Expected more than 1 value per channel when training (input normalisation)")
        fold = FoldedNet(net, mom, training)
        out = mlp_forward(fold.spec, fold.tensors, X)
        ctx.fold, ctx.reducer = fold, reducer
        ctx.save_for_backward(X)
        ctx.tensors = tensors
        return out

    @staticmethod
    def backward(ctx, g_out):
        (X,) = ctx.saved_tensors
        fold = ctx.fold
        gin, views = mlp_backward(fold.spec, fold.tensors, X, g_out, True)
        gg, gb = fold.backward_fixup(views, X, gin, ctx.reducer)
        grads = [v if t.requires_grad else None for v, t in zip(views, ctx.tensors)]
        return (gin if ctx.needs_input_grad[0] else None, None, None,
                gg if ctx.needs_input_grad[3] else None, gb if ctx.needs_input_grad[4] else None, *grads)


# ----------------------------------------------------------------------------------------------
# extraction helpers (no autograd)
# ----------------------------------------------------------------------------------------------
def library_num_columns(n: int, degree: int) -> int:
    c = _lib.load().pderead_library_num_columns(n, degree)
    if c < 1:
        raise PdereadError("library with n=%d, degree=%d is not supported (at most %d columns)" %
                           (n, degree, _lib.MAX_LIBRARY_COLUMNS))
    return c


def library_multi_indices(n: int, degree: int):
    """[C, degree] int32 table (-1 padded); row 0 is the constant column."""
    import numpy as np
    C = library_num_columns(n, degree)
    tab = np.full((C, max(degree, 1)), -1, dtype=np.int32)
    if degree > 0:
        _lib.check(_lib.load().pderead_library_multi_indices(n, degree, tab.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))),
                   "pderead_library_multi_indices")
    return tab


def library_dense(dxn: torch.Tensor, n: int, degree: int) -> torch.Tensor:
    lib = _lib.load()
    dxn = _check_tensor(dxn, "Dxn_U")
    M = dxn.shape[0]
    C = library_num_columns(n, degree)
    with torch.cuda.device(dxn.device):
        out = torch.empty((M, C), dtype=torch.float32, device=dxn.device)
        _lib.check(lib.pderead_library_dense(dxn.data_ptr(), M, n, degree, out.data_ptr(), _stream()),
                   "pderead_library_dense")
    return out


def _gram_outputs(C: int, device):
    G = torch.zeros((C, C), dtype=torch.float64, device=device)
    Atb = torch.zeros((C,), dtype=torch.float64, device=device)
    btb = torch.zeros((1,), dtype=torch.float64, device=device)
    return G, Atb, btb


def library_gram(dxn: torch.Tensor, b: torch.Tensor, n: int, degree: int, out=None):
    lib = _lib.load()
    dxn = _check_tensor(dxn, "Dxn_U")
    b = _check_tensor(b.reshape(-1), "b")
    C = library_num_columns(n, degree)
    with torch.cuda.device(dxn.device):
        G, Atb, btb = out if out is not None else _gram_outputs(C, dxn.device)
        _lib.check(lib.pderead_library_gram(dxn.data_ptr(), b.data_ptr(), dxn.shape[0], n, degree, G.data_ptr(),
                                            Atb.data_ptr(), btb.data_ptr(), None, 0, _stream()), "pderead_library_gram")
    return G, Atb, btb


def dense_gram(A: torch.Tensor, b: torch.Tensor, out=None):
    lib = _lib.load()
    A = _check_tensor(A, "A")
    b = _check_tensor(b.reshape(-1), "b")
    M, C = A.shape
    with torch.cuda.device(A.device):
        G, Atb, btb = out if out is not None else _gram_outputs(C, A.device)
        _lib.check(lib.pderead_dense_gram(A.data_ptr(), b.data_ptr(), M, C, G.data_ptr(), Atb.data_ptr(), btb.data_ptr(),
                                          None, 0, _stream()), "pderead_dense_gram")
    return G, Atb, btb


def extraction_gram(sol_spec, sol_tensors, pde_spec, pde_tensors, n: int, degree: int, coords: torch.Tensor, out=None):
    lib = _lib.load()
    coords = _check_tensor(coords.detach(), "Coords")
    M = coords.shape[0]
    C = library_num_columns(n, degree)
    with torch.cuda.device(coords.device):
        su, k1 = _fill_net(sol_spec, sol_tensors)
        sn, k2 = _fill_net(pde_spec, pde_tensors)
        G, Atb, btb = out if out is not None else _gram_outputs(C, coords.device)
        if M == 0:
            return G, Atb, btb
        nb = lib.pderead_extraction_gram_workspace_bytes(ctypes.byref(su), ctypes.byref(sn), n, degree, _ws_points(M))
        ws = workspace(coords.device, nb)
        _lib.check(lib.pderead_extraction_gram(ctypes.byref(su), ctypes.byref(sn), n, degree, coords.data_ptr(), M,
                                               G.data_ptr(), Atb.data_ptr(), btb.data_ptr(), ws.data_ptr(), ws.numel(),
                                               _stream()), "pderead_extraction_gram")
    return G, Atb, btb


def rfe_gram(G: torch.Tensor, Atb: torch.Tensor, btb: torch.Tensor):
    """-> X [C,C] f32, Residual [C+1] f32, order [C] i32, rank [C] i32, change [C] f32 (device)."""
    lib = _lib.load()
    G = _check_tensor(G, "G", torch.float64)
    Atb = _check_tensor(Atb, "Atb", torch.float64)
    btb = _check_tensor(btb.reshape(-1), "btb", torch.float64)
    C = G.shape[0]
    dev = G.device
    with torch.cuda.device(dev):
        X = torch.empty((C, C), dtype=torch.float32, device=dev)
        Res = torch.empty((C + 1,), dtype=torch.float32, device=dev)
        order = torch.empty((C,), dtype=torch.int32, device=dev)
        rank = torch.empty((C,), dtype=torch.int32, device=dev)
        change = torch.empty((C,), dtype=torch.float32, device=dev)
        nb = lib.pderead_rfe_workspace_bytes(C)
        ws = workspace(dev, nb)
        _lib.check(lib.pderead_rfe_gram(G.data_ptr(), Atb.data_ptr(), btb.data_ptr(), C, X.data_ptr(), Res.data_ptr(),
                                        order.data_ptr(), rank.data_ptr(), change.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _stream()), "pderead_rfe_gram")
    return X, Res, order, rank, change


def generate_points(bounds, num_points: int, device, seed: int, offset: int = 0) -> torch.Tensor:
    import numpy as np
    lib = _lib.load()
    bounds = np.ascontiguousarray(np.asarray(bounds, dtype=np.float64))
    dim = bounds.shape[0]
    device = torch.device(device)
    if device.type != "cuda":
        raise PdereadError("points are generated on a CUDA device (there is no CPU path)")
    with torch.cuda.device(device):
        out = torch.empty((num_points, dim), dtype=torch.float32, device=device)
        _lib.check(lib.pderead_generate_points(bounds.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), dim, num_points,
                                               seed & (2 ** 64 - 1), offset, out.data_ptr(), _stream()),
                   "pderead_generate_points")
    return out


def fma_peak_probe(device, variant: int, iters: int = 4096) -> float:
    """TFLOP/s of a register-resident FMA loop (roofline denominator for this compute-bound path)."""
    lib = _lib.load()
    device = torch.device(device)
    with torch.cuda.device(device):
        sink = torch.zeros(4, dtype=torch.float32, device=device)
        flops = ctypes.c_double(0.0)
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.pderead_fma_peak_probe(variant, iters, sink.data_ptr(), ctypes.byref(flops), _stream()),
                       "pderead_fma_peak_probe")
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = max(best, flops.value / (ms * 1e-3) / 1e12)
    return best
