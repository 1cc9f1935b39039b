"""The device jet recurrences (pde_read_b200/csrc/jet_math.cuh), compiled for the host in double
precision, against torch.autograd: Taylor coefficients of act(z(eps)) and their adjoints."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "jet_harness", "jet_harness.cpp")
OUT = os.path.join(HERE, "jet_harness", "_build", "libjet_harness.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "pde_read_b200", "csrc", "jet_math.cuh")
    if (not os.path.exists(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++17", "-o", OUT, SRC])
    lib = ctypes.CDLL(OUT)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.jet_harness.argtypes = [ctypes.c_int] * 4 + [dp] * 5
    lib.jet_harness.restype = ctypes.c_int
    lib.jet_harness_tanh_row.argtypes = [ctypes.c_int] * 3 + [dp] * 3
    lib.jet_harness_tanh_row.restype = ctypes.c_int
    lib.jet_harness_rational_vjp_y.argtypes = [ctypes.c_int] * 2 + [dp] * 5
    lib.jet_harness_rational_vjp_y.restype = ctypes.c_int
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _act(kind, z, ab):
    if kind == 0:
        return torch.tanh(z)
    if kind == 1:
        return torch.sin(z)
    a, b = ab[:4], ab[4:]
    return (a[0] + z * (a[1] + z * (a[2] + a[3] * z))) / (b[0] + z * (b[1] + b[2] * z))


def _series_coeffs(kind, z0, zs, ab):
    """normalised Taylor coefficients 1..N of act(z0 + sum_k zs[k-1] eps^k) at eps=0, by nested autograd"""
    N = zs.shape[0]
    eps = torch.zeros((), dtype=torch.float64, requires_grad=True)
    zz = z0 + sum(zs[k - 1] * eps ** k for k in range(1, N + 1)) if N > 0 else z0 + 0 * eps
    f = _act(kind, zz, ab)
    out = []
    cur = f
    for k in range(1, N + 1):
        cur = torch.autograd.grad(cur, eps, create_graph=True)[0]
        out.append(cur / math.factorial(k))
    return out


def _torch_jet(kind, z, ab, nx, nt):
    z0 = z[0]
    y0 = _act(kind, z0, ab)
    yx = _series_coeffs(kind, z0, z[1:1 + nx], ab)
    yt = _series_coeffs(kind, z0, z[1 + nx:1 + nx + nt], ab)
    return torch.stack([y0] + yx + yt)


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("nx,nt", [(0, 0), (1, 0), (2, 0), (4, 0), (1, 1), (2, 1), (3, 1), (4, 1), (2, 2),
                                    (4, 2), (5, 3), (0, 2)])
def test_jet_forward_and_vjp(harness, kind, nx, nt):
    rng = np.random.default_rng(1000 * kind + 10 * nx + nt)
    J = 1 + nx + nt
    for trial in range(4):
        z = rng.normal(size=J) * 0.7
        ab = np.array([0.0218, 0.5, 1.5957, 1.1915, 1.0, 0.0, 2.3830]) + rng.normal(size=7) * 0.05
        yb = rng.normal(size=J)
        y = np.zeros(J)
        assert harness.jet_harness(kind, nx, nt, 0, _ptr(z), _ptr(ab), _ptr(yb), _ptr(y), _ptr(np.zeros(7))) == 0
        zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
        abt = torch.tensor(ab, dtype=torch.float64, requires_grad=True)
        yt = _torch_jet(kind, zt, abt, nx, nt)
        np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-11, atol=1e-12)
        L = (yt * torch.tensor(yb)).sum()
        gz, gab = torch.autograd.grad(L, [zt, abt], allow_unused=True)
        zb = np.zeros(J)
        abb = np.zeros(7)
        assert harness.jet_harness(kind, nx, nt, 1, _ptr(z), _ptr(ab), _ptr(yb), _ptr(zb), _ptr(abb)) == 0
        np.testing.assert_allclose(zb, gz.numpy(), rtol=1e-10, atol=1e-11)
        if kind == 2:
            np.testing.assert_allclose(abb, gab.numpy(), rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("nx,nt", [(0, 0), (1, 1), (2, 1), (3, 1), (4, 1), (2, 2), (4, 2), (5, 3)])
def test_tanh_row_forms_match_the_plain_ones(harness, nx, nt):
    """The reverse kernels' tanh forms (y_0 carried in the row, forward jet supplied to the VJP, z_0 never read)
    give the same numbers as the plain forward / VJP."""
    rng = np.random.default_rng(77 + 10 * nx + nt)
    J = 1 + nx + nt
    for trial in range(4):
        z = rng.normal(size=J) * 0.8
        yb = rng.normal(size=J)
        ab = np.zeros(7)
        y, y_row, zb, zb_row = np.zeros(J), np.zeros(J), np.zeros(J), np.zeros(J)
        assert harness.jet_harness(0, nx, nt, 0, _ptr(z), _ptr(ab), _ptr(yb), _ptr(y), _ptr(np.zeros(7))) == 0
        assert harness.jet_harness_tanh_row(nx, nt, 0, _ptr(z), _ptr(yb), _ptr(y_row)) == 0
        np.testing.assert_allclose(y_row, y, rtol=1e-14, atol=1e-15)
        assert harness.jet_harness(0, nx, nt, 1, _ptr(z), _ptr(ab), _ptr(yb), _ptr(zb), _ptr(np.zeros(7))) == 0
        assert harness.jet_harness_tanh_row(nx, nt, 1, _ptr(z), _ptr(yb), _ptr(zb_row)) == 0
        assert np.isfinite(zb_row).all()
        np.testing.assert_allclose(zb_row, zb, rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("nx,nt", [(0, 0), (1, 1), (2, 1), (3, 1), (4, 1), (2, 2), (4, 2), (5, 3)])
def test_rational_vjp_with_supplied_forward_jet(harness, nx, nt):
    """rational_vjp_y (forward jet y supplied, numerator and series division skipped) == rational_vjp."""
    rng = np.random.default_rng(177 + 10 * nx + nt)
    J = 1 + nx + nt
    for trial in range(4):
        z = rng.normal(size=J) * 0.8
        ab = np.array([0.0218, 0.5, 1.5957, 1.1915, 1.0, 0.0, 2.3830]) + rng.normal(size=7) * 0.05
        yb = rng.normal(size=J)
        zb, abb, zb_y, abb_y = np.zeros(J), np.zeros(7), np.zeros(J), np.zeros(7)
        assert harness.jet_harness(2, nx, nt, 1, _ptr(z), _ptr(ab), _ptr(yb), _ptr(zb), _ptr(abb)) == 0
        assert harness.jet_harness_rational_vjp_y(nx, nt, _ptr(z), _ptr(ab), _ptr(yb), _ptr(zb_y), _ptr(abb_y)) == 0
        np.testing.assert_allclose(zb_y, zb, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(abb_y, abb, rtol=1e-13, atol=1e-14)
