"""Pin oracle/pde_read_oracle.py against (a) the reference's own closed-form known-answer tests
(Test/Testing.py) and (b) golden vectors the unmodified reference produced (tests/golden)."""
import glob
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import pde_read_oracle as orc
from tests.helpers import GOLDEN, load_golden, split_state, rel_err


def ones_net(width):
    """Reference Testing.py:29-51: 1 hidden tanh layer, all weights/biases 1 =>
    u(t,x) = width*tanh(t+x+1) + 1."""
    return {"Layers.0.weight": torch.ones(width, 2), "Layers.0.bias": torch.ones(width),
            "Layers.1.weight": torch.ones(1, width), "Layers.1.bias": torch.ones(1)}


def test_network_closed_form():
    # Testing.py:56-80, tolerance Epsilon = 5e-6
    st = ones_net(17)
    x = torch.rand(300, 2, generator=torch.Generator().manual_seed(1))
    got = orc.network_forward(st, "Tanh", x).reshape(-1)
    want = 17 * torch.tanh(x[:, 0] + x[:, 1] + 1.0) + 1.0
    assert (got - want).abs().max().item() < 5e-6


def test_rational_closed_form():
    # Testing.py:82-131: R(x) and R'(x) against explicit formulas, tol 2.5e-5
    st = {"Activation_Functions.0.a": torch.tensor(orc.RATIONAL_A_INIT),
          "Activation_Functions.0.b": torch.tensor(orc.RATIONAL_B_INIT)}
    T = torch.rand(40, 30, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    R = orc._activation(st, "Rational", 0, T)
    a, b = orc.RATIONAL_A_INIT, orc.RATIONAL_B_INIT
    x = T.detach()
    Nx = a[0] + a[1] * x + a[2] * x ** 2 + a[3] * x ** 3
    Dx = b[0] + b[1] * x + b[2] * x ** 2
    assert (R.detach() - Nx / Dx).abs().max().item() < 2.5e-5
    R.sum().backward()
    dN = a[1] + 2 * a[2] * x + 3 * a[3] * x ** 2
    dD = b[1] + 2 * b[2] * x
    assert (T.grad - (dN / Dx - dD * Nx / (Dx * Dx))).abs().max().item() < 2.5e-5


def test_evaluate_derivatives_closed_form():
    # Testing.py:253-300 (stale name there; formulas still valid): u, u_t, u_x, u_xx of the ones-net
    n_h = 9
    st = ones_net(n_h)
    P = torch.rand(200, 2, generator=torch.Generator().manual_seed(3))
    Dtm, Dxn = orc.evaluate_derivatives(st, "Tanh", 1, 2, P)
    th = torch.tanh(P[:, 0] + P[:, 1] + 1.0)
    eps = 5e-6
    assert (Dxn[:, 0] - (n_h * th + 1)).abs().max().item() < 5 * eps
    assert (Dtm - n_h * (1 - th ** 2)).abs().max().item() < 10 * eps
    assert (Dxn[:, 1] - n_h * (1 - th ** 2)).abs().max().item() < 10 * eps
    assert (Dxn[:, 2] - (-2 * n_h * th * (1 - th ** 2))).abs().max().item() < 15 * eps


def test_multi_index_counts_and_tuples():
    # Testing.py:314-432: counts n, n(n+1)/2, sum k(k+1)/2; every sorted tuple present, in order
    for nv in range(1, 7):
        assert orc.count_multi_indices(nv, 1) == nv
        assert orc.count_multi_indices(nv, 2) == nv * (nv + 1) // 2
        assert orc.count_multi_indices(nv, 3) == sum(k * (k + 1) // 2 for k in range(1, nv + 1))
        for order in (1, 2, 3, 4):
            mi = orc.multi_indices(nv, order)
            assert mi.shape == (orc.count_multi_indices(nv, order), order)
            rows = [tuple(r) for r in mi]
            assert rows == sorted(set(rows))
            assert all(all(r[i] <= r[i + 1] for i in range(order - 1)) for r in rows)
    assert [tuple(r) for r in orc.multi_indices(4, 2)] == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2),
                                                          (1, 3), (2, 2), (2, 3), (3, 3)]


def test_library_closed_form_column_order():
    # Testing.py:436-489: n=1, K=2 library = [1, u, u_x, u^2, u*u_x, u_x^2]
    st = ones_net(5)
    pde = orc.init_network(2, 6, 2, "Tanh", seed=4)
    P = torch.rand(50, 2, generator=torch.Generator().manual_seed(5))
    b, Lib, counts, mil = orc.generate_library(st, "Tanh", pde, "Tanh", 1, 1, P, 2)
    th = np.tanh((P[:, 0] + P[:, 1] + 1.0).numpy().astype(np.float64))
    u, ux = 5 * th + 1, 5 * (1 - th ** 2)
    want = np.stack([np.ones_like(u), u, ux, u * u, u * ux, ux * ux], axis=1)
    assert list(counts) == [1, 2, 3]
    assert np.max(np.abs(Lib - want)) < 5e-5 * max(1.0, np.max(np.abs(want)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "random_*.npz"))))
def test_random_nets_match_reference(path):
    z = np.load(path)
    LU, WU, LN, WN, m, n = [int(v) for v in z["meta"]]
    aU, aN = [str(v) for v in z["acts"]]
    sol, pde = split_state(z, "sol."), split_state(z, "pde.")
    P = torch.from_numpy(z["coords"])
    Dtm, Dxn = orc.evaluate_derivatives(sol, aU, m, n, P, create_graph=False)
    # same libraries, same op order => agreement to a few ulp
    assert rel_err(Dtm.numpy(), z["dtm"]) < 2e-6
    for k in range(n + 1):
        assert rel_err(Dxn[:, k].numpy(), z["dxn"][:, k]) < 2e-6
    R = orc.pde_residual(sol, aU, pde, aN, m, n, P).detach()
    assert rel_err(R.numpy(), z["resid"]) < 5e-6
    loss, gs, gp = orc.collocation_loss_and_grads(sol, aU, pde, aN, m, n, P, chunk=1 << 20)
    assert abs(loss - float(z["loss"])) <= 2e-6 * abs(float(z["loss"]))
    for k, g in gs.items():
        assert rel_err(g.numpy(), z["gsol." + k]) < 2e-5, k
    for k, g in gp.items():
        assert rel_err(g.numpy(), z["gpde." + k]) < 2e-5, k
    # chunked accumulation gives the same gradient (linearity of the mean)
    loss_c, gs_c, gp_c = orc.collocation_loss_and_grads(sol, aU, pde, aN, m, n, P, chunk=17)
    assert abs(loss_c - loss) <= 1e-5 * abs(loss)
    for k in gs:
        assert rel_err(gs_c[k].numpy(), gs[k].numpy()) < 5e-5, k


def test_checkpoint_collocation_matches_reference():
    ck = load_golden("burgers_ckpt.npz")
    z = load_golden("burgers_colloc.npz")
    sol, pde = split_state(ck, "sol."), split_state(ck, "pde.")
    P = torch.from_numpy(z["coords"])
    # coords are reproducible from the seed the golden script used
    g = torch.Generator().manual_seed(0)
    P2 = torch.rand((5000, 2), generator=g) * torch.tensor([10.0, 15.9375]) + torch.tensor([0.0, -8.0])
    assert torch.equal(P, P2)
    H = z["dtm"].shape[0]
    Dtm, Dxn = orc.evaluate_derivatives(sol, "Rational", 1, 2, P[:H], create_graph=False)
    assert rel_err(Dtm.numpy(), z["dtm"]) < 2e-6
    assert rel_err(Dxn.numpy(), z["dxn"]) < 2e-6
    # SURVEY 8c item 2 spot values
    assert abs(float(Dtm[0]) - 0.0514453) < 1e-6
    np.testing.assert_allclose(Dxn[0].numpy(), [-0.4827163, 0.1250211, 0.0197168], atol=1e-6)
    loss, gs, gp = orc.collocation_loss_and_grads(sol, "Rational", pde, "Rational", 1, 2, P, chunk=1 << 20)
    assert abs(loss - 3.229976e-4) < 2e-10
    assert abs(loss - float(z["loss"])) < 1e-6 * float(z["loss"])
    for k, gr in gs.items():
        assert rel_err(gr.numpy(), z["gsol." + k]) < 2e-5, k
    for k, gr in gp.items():
        assert rel_err(gr.numpy(), z["gpde." + k]) < 2e-5, k


def test_data_loss_matches_reference():
    ck = load_golden("burgers_ckpt.npz")
    z = load_golden("burgers_data_loss.npz")
    sol = {k: v.clone().requires_grad_(True) for k, v in split_state(ck, "sol.").items()}
    loss = orc.data_loss(sol, "Rational", torch.from_numpy(z["inputs"]), torch.from_numpy(z["targets"]))
    assert loss.dtype == torch.float64          # float64 targets promote the loss
    assert abs(loss.item() - float(z["loss"])) < 1e-7 * float(z["loss"])
    loss.backward()
    for k, p in sol.items():
        assert rel_err(p.grad.numpy(), z["gsol." + k]) < 2e-5, k


@pytest.mark.parametrize("name,n,K", [("library_n1_k2.npz", 1, 2), ("library_n2_k3.npz", 2, 3),
                                       ("library_n3_k2.npz", 3, 2)])
def test_library_matches_reference(name, n, K):
    z = load_golden(name)
    sol, pde = split_state(z, "sol."), split_state(z, "pde.")
    b, Lib, counts, mil = orc.generate_library(sol, "Tanh", pde, "Tanh", 1, n, torch.from_numpy(z["coords"]), K)
    assert list(counts) == list(z["counts"])
    for k in range(1, K + 1):
        assert np.array_equal(mil[k], z["mi_%d" % k])
    assert rel_err(b, z["b"]) < 2e-6
    assert rel_err(Lib, z["Library"]) < 2e-6


def test_extraction_matches_reference():
    """Shipped checkpoint, random.seed(0), 20000 points, K=3."""
    ck = load_golden("burgers_ckpt.npz")
    z = load_golden("burgers_extract_k3.npz")
    bounds = load_golden("burgers_data_loss.npz")["bounds"]
    sol, pde = split_state(ck, "sol."), split_state(ck, "pde.")
    random.seed(0)
    E = orc.generate_points(bounds, 20000)
    assert np.array_equal(E[:512].numpy(), z["coords_head"])
    b, Lib, counts, mil = orc.generate_library(sol, "Rational", pde, "Rational", 1, 2, E, 3)
    assert rel_err(b[:512], z["b_head"]) < 2e-6
    assert rel_err(Lib[:512], z["lib_head"]) < 2e-6
    X, Res, order = orc.recursive_feature_elimination(Lib, b)
    assert order == [int(v) for v in z["order"]]
    assert order == [2, 0, 11, 7, 10, 16, 12, 4, 14, 18, 9, 1, 6, 17, 19, 15, 8, 13, 3, 5]
    np.testing.assert_allclose(Res, z["Residual"], rtol=2e-6)
    assert rel_err(X, z["X"]) < 1e-5
    Xr, Rr, chg, rank = orc.rank_candidate_solutions(X, Res)
    np.testing.assert_allclose(Rr, z["Residual_Ranked"], rtol=2e-6)
    np.testing.assert_allclose(chg, z["Residual_Change"], rtol=1e-4, atol=1e-7)
    assert rel_err(Xr, z["X_Ranked"]) < 1e-5
    # printed #1 PDE: D_t U = 0.076364 (D_x^2 U) - 0.747831 (U)(D_x U)
    best = Xr[:, -1]
    assert set(np.flatnonzero(best).tolist()) == {3, 5}
    assert abs(best[3] - 0.076364) < 2e-6 and abs(best[5] + 0.747831) < 2e-6
    s = orc.format_pde(best, 1, 2, 3)
    assert "(D_x^2 U)" in s and "(U)(D_x U)" in s
