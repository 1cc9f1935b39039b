"""Collocation_Loss / Data_Loss with the reference's signatures (reference
Code/Loss_Functions.py:11-18 and :70-75)."""
from __future__ import annotations

import torch

from . import functional as F
from .Network import Neural_Network

# Optional cross-GPU reducer (pde_read_b200.parallel.ShardReducer).  When set, Collocation_Loss
# treats its points as this rank's shard and all-reduces loss and gradients.
_reducer = None


def set_reducer(reducer) -> None:
    global _reducer
    _reducer = reducer


def get_reducer():
    return _reducer


def Collocation_Loss(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Collocation_Coords: torch.Tensor,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """mean(R^2) over the collocation points (reference Loss_Functions.py:55-65).  The returned
    scalar carries a grad_fn; its backward yields the same parameter gradients as the reference's
    nested-autograd graph, computed by the fused forward+reverse CUDA pass."""
    if Data_Type != torch.float32:
        raise NotImplementedError("the CUDA path computes in float32")
    m = max(int(Time_Derivative_Order), 1)
    n = int(Spatial_Derivative_Order)
    if getattr(PDE_NN, "Batch_Norm", False):
        # input normalisation couples the points: U-jets -> moments (all-reduced over ranks) -> N with the
        # normalisation folded into its first layer; the reverse pass adds the batch-statistics terms to d/dDxn_U
        from .PDE_Residual import PDE_Residual
        R = PDE_Residual(Sol_NN, PDE_NN, m, n, Collocation_Coords, Data_Type, Device)
        if _reducer is None:
            return (R ** 2).mean()
        return F.ShardedMeanSquareFn.apply(R, _reducer)
    su, sn = F.net_spec(Sol_NN), F.net_spec(PDE_NN)
    tu, tn = F.net_tensors(Sol_NN), F.net_tensors(PDE_NN)
    return F.CollocationLossFn.apply(Collocation_Coords, su, sn, m, n, len(tu), _reducer, *tu, *tn)


def Data_Loss(
        Sol_NN: Neural_Network,
        Data_Coords: torch.Tensor,
        Data_Values: torch.Tensor,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """mean((U(Data_Coords) - Data_Values)^2) (reference Loss_Functions.py:115-122).  The network
    evaluation and its reverse pass are the CUDA MLP kernels; the subtraction, square and mean are
    the same torch ops as in the reference (so float64 targets promote the loss as they do there)."""
    u_approx_batch = Sol_NN(Data_Coords).squeeze()
    return ((u_approx_batch - Data_Values) ** 2).mean()
