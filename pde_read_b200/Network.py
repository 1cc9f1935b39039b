"""Solution / PDE network containers with the reference's constructor and state_dict layout.

Mirrors reference Code/Network.py: ``Rational`` (:7-57), ``Sin`` (:60-67) and ``Neural_Network``
(:72-202).  The classes only *hold* parameters (so checkpoints written by the reference load
unchanged, main.py:78-98); evaluating a network dispatches to the CUDA library.
"""
from __future__ import annotations

import math
import sys

import torch

from . import functional as F


class Rational(torch.nn.Module):
    """Trainable (3,2) rational activation R(x) = (a0+a1 x+a2 x^2+a3 x^3)/(b0+b1 x+b2 x^2),
    initialised to the best rational approximation of ReLU (reference Network.py:20-30)."""

    def __init__(self, Data_Type: torch.dtype = torch.float32, Device: torch.device = torch.device("cpu")):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor((0.0218, 0.5, 1.5957, 1.1915), dtype=Data_Type, device=Device))
        self.b = torch.nn.Parameter(torch.tensor((1.0, 0.0, 2.3830), dtype=Data_Type, device=Device))

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        # Stand-alone elementwise use (the reference's test_Rational).  Inside Neural_Network the
        # activation is fused into the CUDA layer kernels and this method is not called.
        a, b = self.a, self.b
        return (a[0] + X * (a[1] + X * (a[2] + a[3] * X))) / (b[0] + X * (b[1] + b[2] * X))


class Sin(torch.nn.Module):
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return torch.sin(input)


class Neural_Network(torch.nn.Module):
    """Fully connected network: Num_Hidden_Layers hidden layers of Neurons_Per_Layer neurons,
    an activation after each, then a linear output layer (reference Network.py:72-170)."""

    def __init__(self,
                 Num_Hidden_Layers: int = 3,
                 Neurons_Per_Layer: int = 20,
                 Input_Dim: int = 1,
                 Output_Dim: int = 1,
                 Data_Type: torch.dtype = torch.float32,
                 Device: torch.device = torch.device("cpu"),
                 Activation_Function: str = "Tanh",
                 Batch_Norm: bool = False):
        assert Num_Hidden_Layers > 0, "Num_Hidden_Layers must be positive. Got %d" % Num_Hidden_Layers
        assert Neurons_Per_Layer > 0, "Neurons_Per_Layer must be positive. Got %d" % Neurons_Per_Layer
        assert Input_Dim > 0, "Input_Dim must be positive. Got %d" % Input_Dim
        assert Output_Dim > 0, "Output_Dim must be positive. Got %d" % Output_Dim
        super().__init__()
        self.Input_Dim = Input_Dim
        self.Output_Dim = Output_Dim
        self.Num_Hidden_Layers = Num_Hidden_Layers
        self.Batch_Norm = Batch_Norm
        self.Activation_Name = Activation_Function

        self.Layers = torch.nn.ModuleList()
        if Batch_Norm:
            self.Norm_Layer = torch.nn.BatchNorm1d(num_features=Input_Dim, dtype=Data_Type, device=Device)
        widths = [Input_Dim] + [Neurons_Per_Layer] * Num_Hidden_Layers + [Output_Dim]
        for i in range(Num_Hidden_Layers + 1):
            self.Layers.append(torch.nn.Linear(widths[i], widths[i + 1], bias=True).to(dtype=Data_Type, device=Device))

        # Initial weights: Xavier-normal with gain 5/3 (Tanh) or 1.41 (Rational), or the SIREN-style
        # uniform(-3/sqrt(width), 3/sqrt(width)) for Sin; biases start at zero (Network.py:133-154).
        if Activation_Function in ("Tanh", "Rational"):
            gain = 5.0 / 3.0 if Activation_Function == "Tanh" else 1.41
            for lin in self.Layers:
                torch.nn.init.xavier_normal_(lin.weight, gain=gain)
                torch.nn.init.zeros_(lin.bias)
        elif Activation_Function == "Sin":
            bound = 3.0 / math.sqrt(Neurons_Per_Layer)
            for lin in self.Layers:
                torch.nn.init.uniform_(lin.weight, -bound, bound)
                torch.nn.init.zeros_(lin.bias)

        self.Activation_Functions = torch.nn.ModuleList()
        for _ in range(Num_Hidden_Layers):
            if Activation_Function == "Tanh":
                self.Activation_Functions.append(torch.nn.Tanh())
            elif Activation_Function == "Sin":
                self.Activation_Functions.append(Sin())
            elif Activation_Function == "Rational":
                self.Activation_Functions.append(Rational(Data_Type=Data_Type, Device=Device))
            else:
                print("Unknown Activation Function. Got %s" % Activation_Function)
                print("Thrown by Neural_Network.__init__. Aborting.")
                sys.exit()

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        """[B, Input_Dim] -> [B, Output_Dim] (reference Network.py:174-202), evaluated by the fused
        CUDA MLP kernel; differentiable w.r.t. X and the parameters."""
        if self.Batch_Norm:
            # input normalisation (batch statistics in train mode, running statistics in eval mode, running
            # statistics updated as a side effect).  Norm_Layer only HOLDS the parameters and buffers of the
            # reference's BatchNorm1d (checkpoint layout); the arithmetic is the two-phase CUDA path: moments of the
            # inputs (all-reduced when the points are sharded), normalisation folded into the first Linear layer
            from . import Loss_Functions
            bn = self.Norm_Layer
            out = F.BNInputMLPFn.apply(X, self, Loss_Functions._reducer, bn.weight, bn.bias, *F.net_tensors(self))
            return out.view(-1, 1)
        spec = F.net_spec(self, allow_input_norm=True)
        out = F.MLPFn.apply(X.contiguous(), spec, *F.net_tensors(self))
        return out.view(-1, 1)
