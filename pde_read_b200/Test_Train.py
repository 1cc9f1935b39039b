"""Discovery_Training / Discovery_Testing with the reference's signatures (reference
Code/Test_Train.py:10-20 and :109-118).  The closure is where Loss.backward() triggers the reverse
kernels; it works unchanged under Adam, SGD and LBFGS (which calls it several times per step)."""
from __future__ import annotations

import os
from typing import Tuple

import torch

from . import functional as F
from .Loss_Functions import Collocation_Loss, Data_Loss
from .Network import Neural_Network


def Discovery_Training(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Collocation_Coords: torch.Tensor,
        Data_Coords: torch.Tensor,
        Data_Values: torch.Tensor,
        Optimizer: torch.optim.Optimizer,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> None:
    """One epoch: a single Optimizer.step over loss = Collocation_Loss + Data_Loss."""
    Sol_NN.train()
    PDE_NN.train()
    fast = _fast_closure_ok(Sol_NN, PDE_NN, Collocation_Coords, Data_Coords, Data_Values, Data_Type)
    if fast and _graph_ok(Sol_NN, PDE_NN, Collocation_Coords, Optimizer):
        # whole epoch (closure + optimiser step) as one CUDA-graph replay
        _graphed_epoch(Sol_NN, PDE_NN, max(int(Time_Derivative_Order), 1), int(Spatial_Derivative_Order),
                       Collocation_Coords, Data_Coords, Data_Values, Optimizer)
        return

    def Discovery_Closure():
        if torch.is_grad_enabled():
            Optimizer.zero_grad()
        if fast and torch.is_grad_enabled():
            return _fused_closure(Sol_NN, PDE_NN, max(int(Time_Derivative_Order), 1), int(Spatial_Derivative_Order),
                                  Collocation_Coords, Data_Coords, Data_Values, Optimizer)
        Loss = (Collocation_Loss(Sol_NN, PDE_NN, Time_Derivative_Order, Spatial_Derivative_Order, Collocation_Coords,
                                 Data_Type, Device)
                + Data_Loss(Sol_NN, Data_Coords, Data_Values, Data_Type, Device))
        if Loss.requires_grad:
            Loss.backward()
        return Loss

    Optimizer.step(Discovery_Closure)


# The closure above is what the reference runs (Test_Train.py:73-102).  With ~50 small parameter tensors the
# autograd bookkeeping around the two fused kernels costs more host time than the kernels take at the
# reference's default 5000 points, so when nothing needs the graph (plain training: every parameter trainable,
# no input BatchNorm, no point sharding) the same numbers are produced without it: the collocation call returns
# d(mean R^2)/d(theta) for both networks in one flat buffer, the data term's reverse kernel ADDS into U's slots
# of that buffer, and the parameters' .grad become views of it.  PDEREAD_FAST_CLOSURE=0 keeps the autograd path.
_FAST_CLOSURE = os.environ.get("PDEREAD_FAST_CLOSURE", "1") not in ("0", "", "off")


def _fast_closure_ok(Sol_NN, PDE_NN, Collocation_Coords, Data_Coords, Data_Values, Data_Type) -> bool:
    from . import Loss_Functions
    if not _FAST_CLOSURE or Loss_Functions._reducer is not None or Data_Type != torch.float32:
        return False
    if getattr(Sol_NN, "Batch_Norm", False) or getattr(PDE_NN, "Batch_Norm", False):
        return False
    if not (Collocation_Coords.is_cuda and Data_Coords.is_cuda) or Collocation_Coords.shape[0] == 0 or Data_Coords.shape[0] < 2:
        return False
    # targets as the reference's loader delivers them: one value per data point (anything else takes the reference's
    # own broadcasting semantics through the autograd path)
    if not (Data_Values.is_cuda and Data_Values.dim() == 1 and Data_Values.shape[0] == Data_Coords.shape[0]
            and Data_Values.dtype in (torch.float32, torch.float64)):
        return False
    return all(p.requires_grad for p in list(Sol_NN.parameters()) + list(PDE_NN.parameters()))


def _loss_and_flat_grad(su, tu, sn, tn, m, n, Collocation_Coords, Data_Coords, Data_Values, private_ws=None):
    """(loss, flat gradient, U views, N views) of Collocation_Loss + Data_Loss without an autograd graph."""
    ws1 = private_ws[0] if private_ws is not None else None
    ws2 = private_ws[1] if private_ws is not None else None
    ss, flat, vu, vn, _ = F.collocation_loss_grad(su, tu, sn, tn, m, n, Collocation_Coords, private_ws=ws1)
    Coll = ss.reshape(()) / float(Collocation_Coords.shape[0])
    # data term: mean((U(x) - u)^2) (reference Loss_Functions.py:115-122; float64 targets promote the loss)
    u = F.mlp_forward(su, tu, Data_Coords, private_ws=ws2)
    diff = u - Data_Values.reshape(-1)
    Data = (diff ** 2).mean()
    F.mlp_backward_accumulate(su, tu, Data_Coords, (diff * (2.0 / diff.numel())).to(torch.float32), flat, 0,
                              private_ws=ws2)
    return (Coll.to(Data.dtype) + Data).detach(), flat, vu, vn


def _fused_closure(Sol_NN, PDE_NN, m, n, Collocation_Coords, Data_Coords, Data_Values, Optimizer=None) -> torch.Tensor:
    su, sn = F.net_spec(Sol_NN), F.net_spec(PDE_NN)
    tu, tn = F.net_tensors(Sol_NN), F.net_tensors(PDE_NN)
    Loss, flat, vu, vn = _loss_and_flat_grad(su, tu, sn, tn, m, n, Collocation_Coords, Data_Coords, Data_Values)
    if Optimizer is not None and hasattr(Optimizer, "accept_flat_grad") and Optimizer.flat_layout_matches(tu + tn):
        # multi-tensor optimiser (Optimizers.py): it reads the flat buffer directly, no per-tensor .grad traffic
        Optimizer.accept_flat_grad(flat)
        return Loss
    for p, g in zip(tu + tn, list(vu) + list(vn)):
        if p.grad is None:
            p.grad = g
        else:
            p.grad.add_(g)
    return Loss


# ---- the whole epoch as one CUDA graph ---------------------------------------------------------------------------
# At the reference's default sizes (5000 collocation + 4000 data points) an epoch is ~12 launches of 10-60 us of
# work each: launch-bound.  With a multi-tensor optimiser (Optimizers.Fused_Adam / Fused_SGD) nothing in the epoch
# needs the host -- the step count lives on the device -- so closure + optimiser step are captured once per
# (networks, optimiser settings, point counts) and replayed: the epoch's points are copied into the graph's static
# input buffers and one launch runs everything.  PDEREAD_TRAIN_GRAPH=0 switches the capture off.
_TRAIN_GRAPH = os.environ.get("PDEREAD_TRAIN_GRAPH", "1") not in ("0", "", "off")
_TRAIN_GRAPH_MAX_POINTS = int(os.environ.get("PDEREAD_TRAIN_GRAPH_MAX_POINTS", "65536"))
_EPOCH_GRAPHS: dict = {}
LAST_EPOCH_LOSS = None        # device scalar of the most recent graphed epoch (the closure's return value)


def _graph_ok(Sol_NN, PDE_NN, Collocation_Coords, Optimizer) -> bool:
    if not _TRAIN_GRAPH or not hasattr(Optimizer, "launch_step"):
        return False
    if Collocation_Coords.shape[0] > _TRAIN_GRAPH_MAX_POINTS:
        return False
    return Optimizer.flat_layout_matches(F.net_tensors(Sol_NN) + F.net_tensors(PDE_NN))


def _graphed_epoch(Sol_NN, PDE_NN, m, n, Collocation_Coords, Data_Coords, Data_Values, Optimizer) -> torch.Tensor:
    global LAST_EPOCH_LOSS
    su, sn = F.net_spec(Sol_NN), F.net_spec(PDE_NN)
    tu, tn = F.net_tensors(Sol_NN), F.net_tensors(PDE_NN)
    dev = Collocation_Coords.device
    g = Optimizer.param_groups[0]
    hyper = tuple(sorted((k, str(v)) for k, v in g.items() if k != "params"))
    key = (id(Optimizer), su, sn, m, n, Collocation_Coords.shape[0], Data_Coords.shape[0], Data_Values.dtype, dev.index,
           hyper, tuple(t.data_ptr() for t in tu + tn))
    ent = _EPOCH_GRAPHS.get(key)
    if ent is None:
        if len(_EPOCH_GRAPHS) >= 4:
            _EPOCH_GRAPHS.pop(next(iter(_EPOCH_GRAPHS)))
        s_coll, s_dc, s_dv = Collocation_Coords.clone(), Data_Coords.clone(), Data_Values.clone()
        own_ws = [[], []]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):           # warm-up of the closure only (no optimiser step: parameters untouched)
            for _ in range(2):
                _loss_and_flat_grad(su, tu, sn, tn, m, n, s_coll, s_dc, s_dv, private_ws=own_ws)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            Loss, flat, vu, vn = _loss_and_flat_grad(su, tu, sn, tn, m, n, s_coll, s_dc, s_dv, private_ws=own_ws)
            Optimizer.launch_step(flat)
        for p, gv in zip(tu + tn, list(vu) + list(vn)):
            p.grad = gv                          # .grad = views of the graph's static gradient buffer
        ent = (graph, s_coll, s_dc, s_dv, Loss, flat, own_ws, [Data_Coords.data_ptr(), Data_Coords._version,
                                                               Data_Values.data_ptr(), Data_Values._version])
        _EPOCH_GRAPHS[key] = ent
    graph, s_coll, s_dc, s_dv, Loss, flat, _ws, seen = ent
    s_coll.copy_(Collocation_Coords)
    now = [Data_Coords.data_ptr(), Data_Coords._version, Data_Values.data_ptr(), Data_Values._version]
    if now != seen or s_dc.data_ptr() == 0:
        s_dc.copy_(Data_Coords)
        s_dv.copy_(Data_Values)
        seen[:] = now
    graph.replay()
    LAST_EPOCH_LOSS = Loss
    return Loss


def release_epoch_graphs() -> None:
    _EPOCH_GRAPHS.clear()


def Discovery_Testing(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Collocation_Coords: torch.Tensor,
        Data_Coords: torch.Tensor,
        Data_Values: torch.Tensor,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> Tuple[float, float]:
    """(collocation loss, data loss) as floats.  Forward-only: the jet kernels need no autograd graph
    to obtain the derivatives, so the losses are evaluated under no_grad (the reference has to build
    the graph here, Test_Train.py:119-122)."""
    Sol_NN.eval()
    PDE_NN.eval()
    with torch.no_grad():
        Coll_Loss = Collocation_Loss(Sol_NN, PDE_NN, Time_Derivative_Order, Spatial_Derivative_Order,
                                     Collocation_Coords, Data_Type, Device).item()
        Data_loss = Data_Loss(Sol_NN, Data_Coords, Data_Values, Data_Type, Device).item()
    return (Coll_Loss, Data_loss)
