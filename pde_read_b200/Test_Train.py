"""Discovery_Training / Discovery_Testing with the reference's signatures (reference
Code/Test_Train.py:10-20 and :109-118).  The closure is where Loss.backward() triggers the reverse
kernels; it works unchanged under Adam, SGD and LBFGS (which calls it several times per step)."""
from __future__ import annotations

from typing import Tuple

import torch

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

    def Discovery_Closure():
        if torch.is_grad_enabled():
            Optimizer.zero_grad()
        Loss = (Collocation_Loss(Sol_NN, PDE_NN, Time_Derivative_Order, Spatial_Derivative_Order, Collocation_Coords,
                                 Data_Type, Device)
                + Data_Loss(Sol_NN, Data_Coords, Data_Values, Data_Type, Device))
        if Loss.requires_grad:
            Loss.backward()
        return Loss

    Optimizer.step(Discovery_Closure)


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
