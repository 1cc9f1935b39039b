"""Generate_Points with the reference's signature (reference Code/Points.py:7-62).

The reference draws every coordinate with Python's ``random.uniform`` and writes it into the tensor
element by element (2*M interpreter iterations, and on a CUDA device one host-to-device write per
coordinate).  Here the points are produced by a Philox-4x32-10 counter kernel on the device: the
same distribution (independent uniforms in the box), not the same stream.  The stream is
reproducible: it is a function of (seed, running offset), and ``Set_Seed`` resets both.
"""
from __future__ import annotations

import numpy
import torch

from . import functional as F

_state = {"seed": 0x5DE2EAD, "offset": 0}


def Set_Seed(seed: int) -> None:
    """Seed the device point stream (the counterpart of ``random.seed`` for the reference)."""
    _state["seed"] = int(seed)
    _state["offset"] = 0


def Generate_Points(
        Bounds: numpy.ndarray,
        Num_Points: int,
        Data_Type: torch.dtype = torch.float32,
        Device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """Num_Points uniformly distributed points in the box Bounds = [[a_1, b_1], ..., [a_d, b_d]]
    as a [Num_Points, d] tensor on Device (which must be a CUDA device)."""
    Bounds = numpy.asarray(Bounds)
    Num_Dim = Bounds.shape[0]
    for j in range(Num_Dim):
        assert Bounds[j, 0] <= Bounds[j, 1]
    if Data_Type != torch.float32:
        raise NotImplementedError("points are generated in float32 (what the reference's main.py asks for)")
    pts = F.generate_points(Bounds.astype(numpy.float64), int(Num_Points), Device, _state["seed"], _state["offset"])
    _state["offset"] += int(Num_Points)
    return pts
