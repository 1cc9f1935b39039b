"""Multi-tensor optimisers for the training closure.

The reference builds ``torch.optim.Adam`` / ``torch.optim.SGD(momentum=0.9, nesterov=True)`` over
``list(Sol_NN.parameters()) + list(PDE_NN.parameters())`` (reference main.py:63-71) and steps them once per
epoch with the training closure (Test_Train.py:105).  ``Fused_Adam`` and ``Fused_SGD`` are drop-in subclasses of
those two classes: same constructor arguments, same ``step(closure)``, and ``state_dict()`` /
``load_state_dict()`` in torch's own layout, so the reference's ``Optimizer_State`` checkpoints (main.py:93-98,
278-283) round-trip in both directions.  What changes is where the numbers live:

* on construction the parameters are re-pointed at slices of ONE flat float32 buffer (every tensor 16-byte aligned --
  the layout the fused collocation call uses for its flat gradient buffer), and the optimiser state is one flat
  buffer per moment;
* ``step`` is one launch of ``pderead_adam_step`` / ``pderead_sgd_step`` over the flat buffers (the step count lives on
  the device and is advanced by the kernel, so the launch can be captured in a CUDA graph);
* when the closure is the graph-free training closure of ``Test_Train.Discovery_Training`` the whole epoch --
  collocation forward + reverse, data term, optimiser step -- is captured once per point-set shape and replayed
  (``PDEREAD_TRAIN_GRAPH=0`` disables the capture).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from . import functional as F


class _FlatMixin:
    """Flat parameter / gradient / state buffers shared by the fused optimisers."""

    def _flatten(self) -> None:
        params: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        if len(self.param_groups) != 1:
            raise ValueError("the fused optimisers take one parameter group (as reference main.py:63-71 builds them)")
        if not params:
            raise ValueError("no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.PdereadError("the fused optimisers run on a CUDA device (there is no CPU path)")
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise _lib.PdereadError("all parameters must be float32 tensors on one CUDA device")
        self._params = params
        self._offsets, off = [], 0
        for p in params:
            self._offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self._numel = off
        self._flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, self._offsets):
                view = self._flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                      # parameters now alias the flat buffer
        self._flat_g = torch.zeros(off, dtype=torch.float32, device=dev)     # used when gradients arrive per tensor
        self._dev_state = torch.zeros(2, dtype=torch.float32, device=dev)    # [step count, ticket]
        self._external_flat_grad: Optional[torch.Tensor] = None
        F._NET_CACHE.clear()                       # cached network descriptions hold the old pointers

    # -- gradient hand-over -------------------------------------------------------------------------------
    def accept_flat_grad(self, flat: torch.Tensor) -> None:
        """Called by the graph-free training closure: `flat` holds d(loss)/d(parameters) in this optimiser's own
        layout (U tensors, then N tensors, each 16-byte aligned); step() then reads it directly."""
        if flat.numel() < self._numel:
            raise _lib.PdereadError("flat gradient buffer is too short for this optimiser's layout")
        self._external_flat_grad = flat

    def flat_layout_matches(self, tensors) -> bool:
        """True if `tensors` (U tensors then N tensors, in the closure's order) are this optimiser's parameters in
        its own order, i.e. the closure's flat gradient layout is this optimiser's layout."""
        return len(tensors) == len(self._params) and all(a is b for a, b in zip(tensors, self._params))

    def _gather_grads(self) -> torch.Tensor:
        flat = self._external_flat_grad
        self._external_flat_grad = None
        if flat is not None:
            return flat
        g = self._flat_g
        g.zero_()
        for p, o in zip(self._params, self._offsets):
            if p.grad is not None:
                g[o:o + p.numel()].view(p.shape).copy_(p.grad)
        return g

    def _views(self, flat: torch.Tensor):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self._params, self._offsets)]

    def device_step_count(self) -> int:
        return int(self._dev_state[0].item())


class Fused_Adam(_FlatMixin, torch.optim.Adam):
    """torch.optim.Adam (same arguments; amsgrad / maximize are not supported) with a single-launch step."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, **kw):
        if amsgrad or kw.get("maximize"):
            raise NotImplementedError("Fused_Adam implements the reference's plain Adam (main.py:66)")
        torch.optim.Adam.__init__(self, params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._flatten()
        self._m = torch.zeros_like(self._flat_p)
        self._v = torch.zeros_like(self._flat_p)

    def launch_step(self, flat_grad: torch.Tensor) -> None:
        """The kernel launch alone (what a captured training graph contains)."""
        g = self.param_groups[0]
        lib = _lib.load()
        with torch.cuda.device(self._flat_p.device):
            _lib.check(lib.pderead_adam_step(self._flat_p.data_ptr(), flat_grad.data_ptr(), self._m.data_ptr(),
                                             self._v.data_ptr(), self._numel, float(g["lr"]), float(g["betas"][0]),
                                             float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), 1.0,
                                             self._dev_state.data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "pderead_adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.launch_step(self._gather_grads())
        return loss

    # -- checkpoint layout of torch.optim.Adam ------------------------------------------------------------
    def _publish_state(self) -> None:
        step = torch.tensor(float(self.device_step_count()), dtype=torch.float32)
        for p, m, v in zip(self._params, self._views(self._m), self._views(self._v)):
            self.state[p] = {"step": step.clone(), "exp_avg": m, "exp_avg_sq": v}

    def state_dict(self):
        if self.device_step_count() > 0:
            self._publish_state()
        return torch.optim.Adam.state_dict(self)

    def load_state_dict(self, state_dict) -> None:
        torch.optim.Adam.load_state_dict(self, state_dict)
        steps = []
        with torch.no_grad():
            for p, m, v in zip(self._params, self._views(self._m), self._views(self._v)):
                st = self.state.get(p)
                if st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    steps.append(float(st["step"]))
            self._dev_state.zero_()
            if steps:
                self._dev_state[0] = max(steps)
        self._publish_state()


class Fused_SGD(_FlatMixin, torch.optim.SGD):
    """torch.optim.SGD (momentum / nesterov / weight_decay; dampening 0) with a single-launch step."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0, weight_decay: float = 0.0,
                 nesterov: bool = False, **kw):
        if dampening != 0.0 or kw.get("maximize"):
            raise NotImplementedError("Fused_SGD implements the reference's SGD(momentum=0.9, nesterov=True) (main.py:71)")
        torch.optim.SGD.__init__(self, params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                 nesterov=nesterov)
        self._flatten()
        self._buf = torch.zeros_like(self._flat_p)

    def launch_step(self, flat_grad: torch.Tensor) -> None:
        g = self.param_groups[0]
        lib = _lib.load()
        with torch.cuda.device(self._flat_p.device):
            _lib.check(lib.pderead_sgd_step(self._flat_p.data_ptr(), flat_grad.data_ptr(), self._buf.data_ptr(),
                                            self._numel, float(g["lr"]), float(g["momentum"]), int(bool(g["nesterov"])),
                                            float(g["weight_decay"]), 1.0, self._dev_state.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "pderead_sgd_step")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.launch_step(self._gather_grads())
        return loss

    def _publish_state(self) -> None:
        for p, b in zip(self._params, self._views(self._buf)):
            self.state[p] = {"momentum_buffer": b}

    def state_dict(self):
        if self.device_step_count() > 0 and self.param_groups[0]["momentum"] != 0:
            self._publish_state()
        return torch.optim.SGD.state_dict(self)

    def load_state_dict(self, state_dict) -> None:
        torch.optim.SGD.load_state_dict(self, state_dict)
        loaded = False
        with torch.no_grad():
            for p, b in zip(self._params, self._views(self._buf)):
                st = self.state.get(p)
                if st and st.get("momentum_buffer") is not None:
                    b.copy_(st["momentum_buffer"])
                    loaded = True
            self._dev_state.zero_()
            if loaded:
                self._dev_state[0] = 1.0           # any non-zero count: the buffers are live
        if loaded:
            self._publish_state()
