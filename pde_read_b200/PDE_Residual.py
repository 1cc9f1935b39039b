"""Evaluate_Derivatives / PDE_Residual with the reference's signatures (reference
Code/PDE_Residual.py:9-15 and :151-158), computed by Taylor-jet CUDA kernels instead of nested
torch.autograd.grad."""
from __future__ import annotations

from typing import Tuple

import torch

from . import functional as F
from .Network import Neural_Network


def Evaluate_Derivatives(
        Sol_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Coords: torch.Tensor,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> Tuple[torch.Tensor, torch.Tensor]:
    """(D_t^m U [M], [U, D_x U, ..., D_x^n U] [M, n+1]) at Coords [M, 2] (column 0 = t, 1 = x).

    Both results carry a grad_fn w.r.t. the parameters of Sol_NN, as in the reference.  Unlike the
    reference this does not set ``Coords.requires_grad`` and does not produce ``Coords.grad``.
    """
    if Data_Type != torch.float32:
        raise NotImplementedError("the CUDA path computes in float32 (the reference's main.py hard-codes it)")
    m = max(int(Time_Derivative_Order), 1)      # the reference always returns the first t-derivative for m <= 1
    n = int(Spatial_Derivative_Order)
    if n < 1:
        raise ValueError("Spatial_Derivative_Order must be >= 1 (the reference indexes column 1 unconditionally)")
    spec = F.net_spec(Sol_NN)
    Dtm_U, Dxn_U = F.SolDerivativesFn.apply(Coords, spec, m, n, *F.net_tensors(Sol_NN))
    return Dtm_U, Dxn_U


def PDE_Residual(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Coords: torch.Tensor,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """R [M] = D_t^m U - PDE_NN(U, D_x U, ..., D_x^n U) (reference PDE_Residual.py:199-215)."""
    if Data_Type != torch.float32:
        raise NotImplementedError("the CUDA path computes in float32")
    m = max(int(Time_Derivative_Order), 1)
    n = int(Spatial_Derivative_Order)
    if getattr(PDE_NN, "Batch_Norm", False):
        # "PDE Network - Normalize Inputs": batch statistics couple all points, so the residual is composed from the
        # U-jet kernel, the moments kernel and the MLP kernel on the network with the normalisation folded into its
        # first layer (two-phase path, functional.BNInputMLPFn) instead of the single fused call
        Dtm_U, Dxn_U = Evaluate_Derivatives(Sol_NN, m, n, Coords, Data_Type, Device)
        return Dtm_U - PDE_NN(Dxn_U).view(-1)
    su, sn = F.net_spec(Sol_NN), F.net_spec(PDE_NN)
    tu, tn = F.net_tensors(Sol_NN), F.net_tensors(PDE_NN)
    return F.ResidualFn.apply(Coords, su, sn, m, n, len(tu), *tu, *tn)
