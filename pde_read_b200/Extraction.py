"""Extraction mode with the reference's function names and return shapes (reference
Code/Extraction.py).  The heavy steps -- derivatives of U, the PDE network, the monomial library,
the Gram system and Recursive Feature Elimination -- run on the GPU; index bookkeeping and the
pretty-printer stay on the host.
"""
from __future__ import annotations

import itertools
import math
from typing import List, Tuple

import numpy as np
import torch

from . import functional as F
from .Network import Neural_Network


# ----------------------------------------------------------------------------------------------
# multi-index bookkeeping (host integers; reference Extraction.py:11-174)
# ----------------------------------------------------------------------------------------------
def Recursive_Counter(
        num_sub_index_values: int,
        num_sub_indices: int,
        sub_index: int = 0,
        sub_index_value: int = 0,
        counter: int = 0) -> int:
    """Number of non-decreasing multi-indices with ``num_sub_indices`` entries drawn from
    {0, ..., num_sub_index_values-1}.  (The recursion arguments of the reference are accepted:
    entries from position ``sub_index`` on are free and must be >= ``sub_index_value``.)"""
    assert num_sub_indices > 0, "num_sub_indices must be a POSITIVE integer. Got %d." % num_sub_indices
    assert num_sub_index_values > 0, "num_sub_index_values must be a POSITIVE integer. Got %d." % num_sub_index_values
    free = num_sub_indices - sub_index
    values = num_sub_index_values - sub_index_value
    return counter + math.comb(values + free - 1, free)


def Recursive_Multi_Indices(
        multi_indices: np.ndarray,
        num_sub_index_values: int,
        num_sub_indices: int,
        sub_index: int = 0,
        sub_index_value: int = 0,
        position: int = 0) -> int:
    """Fill the rows of ``multi_indices`` (from ``position``) with every non-decreasing tuple in
    lexicographic order -- the column order of the library -- and return the next free row."""
    assert num_sub_indices > 0, "num_sub_indices must be a POSITIVE integer. Got %d." % num_sub_indices
    assert num_sub_index_values > 0, "num_sub_index_values must be a POSITIVE integer. Got %d." % num_sub_index_values
    free = num_sub_indices - sub_index
    for tup in itertools.combinations_with_replacement(range(sub_index_value, num_sub_index_values), free):
        multi_indices[position, sub_index:num_sub_indices] = tup
        position += 1
    return position


def _index_tables(n: int, degree: int):
    counts = np.empty(degree + 1, dtype=np.int64)
    counts[0] = 1
    mi_list: List[np.ndarray] = [np.array(1, dtype=np.int64)]
    for k in range(1, degree + 1):
        counts[k] = Recursive_Counter(n + 1, k)
        mi = np.empty((counts[k], k), dtype=np.int64)
        Recursive_Multi_Indices(mi, n + 1, k)
        mi_list.append(mi)
    return counts, mi_list


# ----------------------------------------------------------------------------------------------
# library (reference Extraction.py:178-378)
# ----------------------------------------------------------------------------------------------
def Generate_Library(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Coords: torch.Tensor,
        Poly_Degree: int,
        Device: torch.device = torch.device("cpu")) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List]:
    """(PDE_NN at the points [M] f32, Library [M, C] f32, num_multi_indices [K+1] i64,
    multi_indices_list).  Dense library for inspection / small M; large extractions should use
    Extract_PDE, which never materialises it."""
    n = int(Spatial_Derivative_Order)
    su, tu = F.net_spec(Sol_NN), F.net_tensors(Sol_NN)
    # b = N(U, D_x U, ...): the reference computes D_t^m U here too and discards it (:305-344)
    _, Dxn_U = F.sol_derivatives_forward(su, tu, 0, n, Coords)
    with torch.no_grad():
        b = PDE_NN(Dxn_U).view(-1)          # Network.forward: input BatchNorm (if any) + fused MLP kernel
    Library = F.library_dense(Dxn_U, n, int(Poly_Degree))
    counts, mi_list = _index_tables(n, int(Poly_Degree))
    return (b.cpu().numpy().astype(np.float32), Library.cpu().numpy().astype(np.float32), counts, mi_list)


# ----------------------------------------------------------------------------------------------
# RFE + ranking (reference Extraction.py:382-572)
# ----------------------------------------------------------------------------------------------
def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise F.PdereadError("Recursive_Feature_Elimination runs on a CUDA device; none is available")
    return torch.device("cuda", torch.cuda.current_device())


def Recursive_Feature_Elimination(A: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(X [C,C] f32, Residual [C+1] f32): column j of X is the least-squares fit on the features
    that survive j eliminations.  The Gram system A^T A, A^T b, b^T b is accumulated in float64 on
    the GPU and the C shrinking solves run there as well."""
    dev = _default_device()
    At = torch.as_tensor(np.ascontiguousarray(A, dtype=np.float32)).to(dev)
    bt = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float32).reshape(-1)).to(dev)
    G, Atb, btb = F.dense_gram(At, bt)
    X, Res, _, _, _ = F.rfe_gram(G, Atb, btb)
    return X.cpu().numpy(), Res.cpu().numpy()


def Rank_Candidate_Solutions(X: np.ndarray, Residual: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(X_Ranked, Residual_Ranked, Residual_Change): candidates ordered by the relative jump of the
    residual when their least important term is removed, ascending -- the last column is the most
    likely PDE.  Ties keep the reference's selection-sort order."""
    Residual = np.asarray(Residual, dtype=np.float32)
    C = Residual.size - 1
    change = ((Residual[1:] - Residual[:-1]) / Residual[:-1]).astype(np.float32)
    rank = np.arange(C, dtype=np.int64)
    for i in range(C):
        k = i + int(np.argmin(change[i:]))
        change[[i, k]] = change[[k, i]]
        rank[[i, k]] = rank[[k, i]]
    X_Ranked = np.ascontiguousarray(np.asarray(X, dtype=np.float32)[:, rank])
    Residual_Ranked = Residual[rank].astype(np.float32)
    return X_Ranked, Residual_Ranked, change


def Extract_PDE(
        Sol_NN: Neural_Network,
        PDE_NN: Neural_Network,
        Time_Derivative_Order: int,
        Spatial_Derivative_Order: int,
        Coords: torch.Tensor,
        Poly_Degree: int,
        reducer=None):
    """Fused extraction: coords -> (X, Residual, X_Ranked, Residual_Ranked, Residual_Change,
    num_multi_indices, multi_indices_list, order) without building the [M, C] library.
    With a ``reducer`` (pde_read_b200.parallel.ShardReducer) Coords is this rank's shard and the
    float64 Gram system is all-reduced before the solves."""
    n, K = int(Spatial_Derivative_Order), int(Poly_Degree)
    if getattr(PDE_NN, "Batch_Norm", False):
        # two-phase: U-jets -> moments over all extraction points (summed over ranks when sharded) -> N with the
        # input normalisation folded into its first layer
        su, tu = F.net_spec(Sol_NN), F.net_tensors(Sol_NN)
        _, Dxn_U = F.sol_derivatives_forward(su, tu, 0, n, Coords)
        mom = F.input_moments(Dxn_U, reducer) if PDE_NN.training else None
        fold = F.FoldedNet(PDE_NN, mom, PDE_NN.training)
        b = F.mlp_forward(fold.spec, fold.tensors, Dxn_U)
        G, Atb, btb = F.library_gram(Dxn_U, b, n, K)
    else:
        su, sn = F.net_spec(Sol_NN), F.net_spec(PDE_NN)
        tu, tn = F.net_tensors(Sol_NN), F.net_tensors(PDE_NN)
        G, Atb, btb = F.extraction_gram(su, tu, sn, tn, n, K, Coords)
    if reducer is not None:
        G, Atb, btb = reducer.reduce_gram(G, Atb, btb)
    X, Res, order, rank, change = F.rfe_gram(G, Atb, btb)
    counts, mi_list = _index_tables(n, K)
    Xh, Resh = X.cpu().numpy(), Res.cpu().numpy()
    rk = rank.cpu().numpy().astype(np.int64)
    return (Xh, Resh, np.ascontiguousarray(Xh[:, rk]), Resh[rk], change.cpu().numpy(), counts, mi_list,
            order.cpu().numpy())


# ----------------------------------------------------------------------------------------------
# printing (reference Extraction.py:576-635)
# ----------------------------------------------------------------------------------------------
def Print_Extracted_PDE(
        Extracted_PDE: np.ndarray,
        Time_Derivative_Order: int,
        num_multi_indices: np.ndarray,
        multi_indices_list) -> None:
    """Print one candidate as a human-readable PDE, in the reference's format."""
    head = "D_t U = " if Time_Derivative_Order == 1 else "D_t^%u U = " % Time_Derivative_Order
    parts = [head]
    if Extracted_PDE[0] != 0:
        parts.append("%f " % Extracted_PDE[0])
    position = 1
    for k in range(1, num_multi_indices.shape[0]):
        for i in range(num_multi_indices[k]):
            coeff = Extracted_PDE[position]
            if coeff != 0:
                term = " + %f" % coeff
                for idx in multi_indices_list[k][i]:
                    term += "(U)" if idx == 0 else ("(D_x U)" if idx == 1 else "(D_x^%u U)" % idx)
                parts.append(term)
            position += 1
    print("".join(parts))
