"""CPU oracle for the PDE-READ hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement of the algorithm the reference (punkduckable/PDE-READ,
mounted at /root/reference in the build container) runs on the path this repo accelerates:

    Collocation_Loss -> PDE_Residual -> Evaluate_Derivatives (+ Loss.backward())
    Generate_Library -> Recursive_Feature_Elimination -> Rank_Candidate_Solutions

The arithmetic of that path lives in two un-vendored dependencies of the reference, torch
(ATen + autograd, nested ``autograd.grad(create_graph=True)``) and numpy (``linalg.lstsq``,
LAPACK dgelsd).  The reference pins no versions (README.md:43-47); this image carries
torch 2.11.0 and numpy 2.3.  The oracle calls the *same* two libraries at the same call
sites, so it follows the reference's numerics operation for operation while being written
independently of its sources.

Who may import this file: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker / CPU baseline only.  Nothing under
``pde_read_b200/`` imports it; the product path raises if its CUDA library is missing.

Parity pin: ``tests/golden/make_golden.py`` runs the unmodified reference from
/root/reference/Code in the build container and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those vectors and against
the closed-form known-answer tests of the reference's own ``Test/Testing.py``.

Networks are passed as ``(state, activation)`` where ``state`` is a mapping with the reference's
``state_dict`` keys (``Layers.i.weight`` [out,in], ``Layers.i.bias``, and for Rational nets
``Activation_Functions.i.a`` [4] / ``.b`` [3]; reference Network.py:97-166) and ``activation``
is "Tanh", "Sin" or "Rational".
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np
import torch

State = Mapping[str, torch.Tensor]

RATIONAL_A_INIT = (0.0218, 0.5, 1.5957, 1.1915)   # reference Network.py:20-24
RATIONAL_B_INIT = (1.0, 0.0, 2.3830)              # reference Network.py:27-30


# ----------------------------------------------------------------------------------------------
# Network (reference Network.py)
# ----------------------------------------------------------------------------------------------

def num_hidden_layers(state: State) -> int:
    """Number of hidden layers = number of Linear modules minus the output layer
    (reference Network.py:110-130)."""
    n_linear = sum(1 for k in state if k.startswith("Layers.") and k.endswith(".weight"))
    return n_linear - 1


def init_network(num_hidden: int, width: int, input_dim: int, activation: str,
                 dtype: torch.dtype = torch.float32, seed: int | None = None) -> Dict[str, torch.Tensor]:
    """Build a parameter dict with the reference's shapes and initial distribution.

    Follows reference Network.py:133-154: Xavier-normal weights with gain 5/3 (Tanh) or 1.41
    (Rational) and zero biases; for Sin, uniform in +-3/sqrt(width) and zero biases.  Rational
    coefficients start at the ReLU best-approximation values (Network.py:20-30).  The random
    stream is NOT the reference's (it draws through nn.Linear's constructor first); only the
    distribution matches, which is all the synthetic-weights benchmark needs.
    """
    gen = torch.Generator()
    if seed is not None:
        gen.manual_seed(seed)
    dims = [input_dim] + [width] * num_hidden + [1]
    state: Dict[str, torch.Tensor] = {}
    for i in range(num_hidden + 1):
        fan_in, fan_out = dims[i], dims[i + 1]
        if activation in ("Tanh", "Rational"):
            gain = 5.0 / 3.0 if activation == "Tanh" else 1.41
            std = gain * math.sqrt(2.0 / (fan_in + fan_out))
            w = torch.randn((fan_out, fan_in), generator=gen, dtype=torch.float64) * std
        elif activation == "Sin":
            a = 3.0 / math.sqrt(width)
            w = (torch.rand((fan_out, fan_in), generator=gen, dtype=torch.float64) * 2 - 1) * a
        else:
            raise ValueError("unknown activation %r" % activation)
        state["Layers.%d.weight" % i] = w.to(dtype)
        state["Layers.%d.bias" % i] = torch.zeros(fan_out, dtype=dtype)
    if activation == "Rational":
        for i in range(num_hidden):
            state["Activation_Functions.%d.a" % i] = torch.tensor(RATIONAL_A_INIT, dtype=dtype)
            state["Activation_Functions.%d.b" % i] = torch.tensor(RATIONAL_B_INIT, dtype=dtype)
    return state


def _activation(state: State, activation: str, i: int, X: torch.Tensor) -> torch.Tensor:
    if activation == "Tanh":
        return torch.tanh(X)                      # reference Network.py:158-160 (torch.nn.Tanh)
    if activation == "Sin":
        return torch.sin(X)                       # reference Network.py:66-67
    if activation == "Rational":
        # reference Network.py:53-57: Horner numerator (cubic) over Horner denominator (quadratic),
        # elementwise, no pole guard.
        a = state["Activation_Functions.%d.a" % i]
        b = state["Activation_Functions.%d.b" % i]
        num = a[0] + X * (a[1] + X * (a[2] + a[3] * X))
        den = b[0] + X * (b[1] + b[2] * X)
        return num / den
    raise ValueError("unknown activation %r" % activation)


def network_forward(state: State, activation: str, X: torch.Tensor) -> torch.Tensor:
    """MLP forward, reference Network.py:198-202 (Batch_Norm=False branch): activation after
    every hidden Linear, none after the last."""
    L = num_hidden_layers(state)
    for i in range(L):
        X = torch.nn.functional.linear(X, state["Layers.%d.weight" % i], state["Layers.%d.bias" % i])
        X = _activation(state, activation, i, X)
    return torch.nn.functional.linear(X, state["Layers.%d.weight" % L], state["Layers.%d.bias" % L])


# ----------------------------------------------------------------------------------------------
# Derivatives / residual / losses (reference PDE_Residual.py, Loss_Functions.py)
# ----------------------------------------------------------------------------------------------

def evaluate_derivatives(state: State, activation: str, m: int, n: int,
                         coords: torch.Tensor, create_graph: bool = True
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(D_t^m U [M], [U, D_x U, ..., D_x^n U] [M, n+1]) by nested autograd.

    Follows reference PDE_Residual.py:51-147: one ``autograd.grad`` of U w.r.t. the [M,2] coords
    with ones as grad_outputs gives (U_t, U_x) because the batch Jacobian is diagonal (:92-104);
    each further order differentiates the previous column again and keeps column 0 for t
    (:109-125) or column 1 for x (:129-145).  Column 0 of coords is t, column 1 is x (:33-34).
    Unlike the reference this does not flip ``requires_grad`` on the caller's tensor.
    """
    X = coords.detach().clone().requires_grad_(True)
    U = network_forward(state, activation, X).reshape(-1)

    def d(f: torch.Tensor) -> torch.Tensor:
        return torch.autograd.grad(f, X, torch.ones_like(f), retain_graph=True,
                                   create_graph=True)[0]

    g = d(U)
    Dtm = g[:, 0]
    cols = [U, g[:, 1]]
    for _ in range(2, m + 1):
        Dtm = d(Dtm)[:, 0]
    for _ in range(2, n + 1):
        cols.append(d(cols[-1])[:, 1])
    Dxn = torch.stack(cols[: n + 1], dim=1)
    if not create_graph:
        Dtm, Dxn = Dtm.detach(), Dxn.detach()
    return Dtm, Dxn


def pde_residual(sol_state: State, sol_act: str, pde_state: State, pde_act: str,
                 m: int, n: int, coords: torch.Tensor) -> torch.Tensor:
    """R = D_t^m U - N(U, D_x U, ..., D_x^n U); reference PDE_Residual.py:199-215."""
    Dtm, Dxn = evaluate_derivatives(sol_state, sol_act, m, n, coords)
    return Dtm - network_forward(pde_state, pde_act, Dxn).reshape(-1)


def collocation_loss(sol_state: State, sol_act: str, pde_state: State, pde_act: str,
                     m: int, n: int, coords: torch.Tensor) -> torch.Tensor:
    """mean(R^2); reference Loss_Functions.py:55-65."""
    return (pde_residual(sol_state, sol_act, pde_state, pde_act, m, n, coords) ** 2).mean()


def data_loss(sol_state: State, sol_act: str, data_coords: torch.Tensor,
              data_values: torch.Tensor) -> torch.Tensor:
    """mean((U(coords) - values)^2); reference Loss_Functions.py:115-122."""
    return ((network_forward(sol_state, sol_act, data_coords).squeeze() - data_values) ** 2).mean()


def _leaf_copy(state: State) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}


def collocation_loss_and_grads(sol_state: State, sol_act: str, pde_state: State, pde_act: str,
                               m: int, n: int, coords: torch.Tensor, chunk: int = 2000,
                               total_points: int | None = None
                               ) -> Tuple[float, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Loss and d(loss)/d(parameters) of both nets, i.e. what the closure at reference
    Test_Train.py:73-102 leaves in ``.grad`` for the collocation term.

    The reference builds one autograd graph over all M points (0.25-3.5 MB per point), which
    cannot hold large M, so this evaluates sum(R^2)/M chunk by chunk and accumulates the
    gradients -- mean over M is linear in the per-chunk sums, so the result is the same
    quantity.  ``total_points`` overrides the divisor (used to emulate one rank of a sharded run).
    """
    M = coords.shape[0]
    denom = float(total_points if total_points is not None else M)
    S = _leaf_copy(sol_state)
    P = _leaf_copy(pde_state)
    total = 0.0
    for lo in range(0, M, chunk):
        R = pde_residual(S, sol_act, P, pde_act, m, n, coords[lo:lo + chunk])
        part = (R ** 2).sum() / denom
        part.backward()
        total += float(part.detach())
    gs = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in S.items()}
    gp = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    return total, gs, gp


# ----------------------------------------------------------------------------------------------
# Points (reference Points.py)
# ----------------------------------------------------------------------------------------------

def generate_points(bounds: np.ndarray, num_points: int) -> torch.Tensor:
    """Uniform points in a box with Python's ``random`` stream in the reference's draw order:
    every coordinate of dimension 0 first, then dimension 1, ... (reference Points.py:53-60)."""
    bounds = np.asarray(bounds, dtype=np.float64)
    out = torch.empty((num_points, bounds.shape[0]), dtype=torch.float32)
    for j in range(bounds.shape[0]):
        lo, hi = float(bounds[j, 0]), float(bounds[j, 1])
        col = [random.uniform(lo, hi) for _ in range(num_points)]
        out[:, j] = torch.tensor(col, dtype=torch.float64).to(torch.float32)
    return out


# ----------------------------------------------------------------------------------------------
# Extraction (reference Extraction.py)
# ----------------------------------------------------------------------------------------------

def count_multi_indices(num_values: int, order: int) -> int:
    """Number of non-decreasing ``order``-tuples over {0..num_values-1}; what reference
    Extraction.py:11-77 counts by recursion, in closed form C(num_values+order-1, order)."""
    return math.comb(num_values + order - 1, order)


def multi_indices(num_values: int, order: int) -> np.ndarray:
    """All non-decreasing ``order``-tuples over {0..num_values-1} in lexicographic order, the
    order reference Extraction.py:81-174 fills its table in."""
    rows: List[Tuple[int, ...]] = []

    def rec(prefix: Tuple[int, ...], lo: int) -> None:
        if len(prefix) == order:
            rows.append(prefix)
            return
        for v in range(lo, num_values):
            rec(prefix + (v,), v)

    rec((), 0)
    return np.asarray(rows, dtype=np.int64).reshape(len(rows), order)


def library_columns(n: int, degree: int) -> List[Tuple[int, ...]]:
    """Column -> multi-index map of the library: () for the constant, then degree 1..K blocks
    (reference Extraction.py:349-372)."""
    cols: List[Tuple[int, ...]] = [()]
    for k in range(1, degree + 1):
        cols.extend(tuple(int(v) for v in row) for row in multi_indices(n + 1, k))
    return cols


def generate_library(sol_state: State, sol_act: str, pde_state: State, pde_act: str,
                     m: int, n: int, coords: torch.Tensor, degree: int, batch: int = 1000
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, list]:
    """(b, Library, num_multi_indices, multi_indices_list); reference Extraction.py:178-378.

    Derivatives are evaluated in ``batch``-point graphs and detached (:305-341); b is the PDE
    network at those derivatives (:344), *not* D_t^m U; each library column starts at 1 and is
    multiplied by its factors left to right in float32 (:349-369).
    """
    M = coords.shape[0]
    parts = []
    # Reference batching (:305-341): full batches while more than one batch remains, then the rest.
    lo = 0
    if M > batch:
        for lo in range(0, M - batch, batch):
            parts.append(evaluate_derivatives(sol_state, sol_act, m, n, coords[lo:lo + batch],
                                              create_graph=False)[1])
        parts.append(evaluate_derivatives(sol_state, sol_act, m, n, coords[lo + batch:],
                                          create_graph=False)[1])
    else:
        parts.append(evaluate_derivatives(sol_state, sol_act, m, n, coords, create_graph=False)[1])
    Dxn = torch.cat(parts, dim=0).to(torch.float32)
    with torch.no_grad():
        b = network_forward(pde_state, pde_act, Dxn).squeeze().numpy().astype(np.float32)
    D = Dxn.numpy().astype(np.float32)

    counts = np.empty(degree + 1, dtype=np.int64)
    counts[0] = 1
    mi_list: list = [np.array(1, dtype=np.int64)]
    for k in range(1, degree + 1):
        counts[k] = count_multi_indices(n + 1, k)
        mi_list.append(multi_indices(n + 1, k))
    C = int(counts.sum())
    Library = np.ones((M, C), dtype=np.float32)
    pos = 1
    for k in range(1, degree + 1):
        for row in mi_list[k]:
            for idx in row:
                Library[:, pos] = Library[:, pos] * D[:, idx]
            pos += 1
    return b, Library, counts, mi_list


def _seq_sum_f32(v: np.ndarray) -> np.float32:
    """Left-to-right float32 running sum -- bit-for-bit what the reference's Python loops
    ``Sum += x*x`` produce on float32 numpy scalars (Extraction.py:447-452, 480-483, 486-489)."""
    if v.size == 0:
        return np.float32(0)
    return np.cumsum(v.astype(np.float32), dtype=np.float32)[-1]


def recursive_feature_elimination(A: np.ndarray, b: np.ndarray
                                  ) -> Tuple[np.ndarray, np.ndarray, List[int]]:
    """(X, Residual, elimination_order); reference Extraction.py:382-501.

    Columns are scaled to unit L2 norm (:443-453); step j solves least squares on the remaining
    columns with numpy.linalg.lstsq(rcond=None) (:464), removes the remaining feature of smallest
    magnitude -- strict '<', first index wins ties (:467-476) -- and records ||b - A_s x||_2
    (:479-483); Residual[C] = ||b||_2 (:486-489); X rows are finally un-scaled (:497-499).
    The elimination order is returned as well (the reference does not expose it).
    """
    M, C = A.shape
    A = A.astype(np.float32, copy=False)
    b = b.astype(np.float32, copy=False)
    norms = np.empty(C, dtype=np.float32)
    A_s = np.empty((M, C), dtype=np.float32)
    for j in range(C):
        norms[j] = math.sqrt(_seq_sum_f32(A[:, j] * A[:, j]))
        A_s[:, j] = A[:, j] / norms[j]
    X = np.zeros((C, C), dtype=np.float32)
    remaining = np.ones(C, dtype=bool)
    Residual = np.empty(C + 1, dtype=np.float32)
    order: List[int] = []
    for j in range(C):
        X[remaining, j] = np.linalg.lstsq(A_s[:, remaining], b, rcond=None)[0]
        idx = np.flatnonzero(remaining)
        mags = np.abs(X[idx, j])
        drop = int(idx[int(np.argmin(mags))])       # argmin returns the first minimum
        remaining[drop] = False
        order.append(drop)
        diff = b - A_s @ X[:, j]
        Residual[j] = math.sqrt(_seq_sum_f32(diff * diff))
    Residual[C] = math.sqrt(_seq_sum_f32(b * b))
    X = X / norms[:, None]
    return X.astype(np.float32), Residual, order


def rank_candidate_solutions(X: np.ndarray, Residual: np.ndarray
                             ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(X_Ranked, Residual_Ranked, Residual_Change, index_rankings); reference
    Extraction.py:505-572.  Relative residual jump (Res[i+1]-Res[i])/Res[i] (:532-535), sorted
    ascending by the reference's selection sort (swap with the first strict minimum of the tail,
    :544-560), so the last column is the most likely PDE."""
    C = Residual.size - 1
    change = np.empty(C, dtype=np.float32)
    for i in range(C):
        change[i] = (Residual[i + 1] - Residual[i]) / Residual[i]
    rank = np.arange(C, dtype=np.int64)
    for i in range(C):
        k = i + int(np.argmin(change[i:]))          # first strict minimum of the tail
        change[i], change[k] = change[k], change[i]
        rank[i], rank[k] = rank[k], rank[i]
    Xr = np.ascontiguousarray(X[:, rank]).astype(np.float32)
    Rr = Residual[rank].astype(np.float32)
    return Xr, Rr, change, rank


def format_pde(coeffs: Sequence[float], m: int, n: int, degree: int) -> str:
    """Human-readable PDE string in the layout reference Extraction.py:576-635 prints."""
    cols = library_columns(n, degree)
    out = "D_t U = " if m == 1 else "D_t^%u U = " % m
    if coeffs[0] != 0:
        out += "%f " % coeffs[0]
    for c, mi in zip(coeffs[1:], cols[1:]):
        if c != 0:
            out += " + %f" % c
            for idx in mi:
                out += "(U)" if idx == 0 else ("(D_x U)" if idx == 1 else "(D_x^%u U)" % idx)
    return out
