"""GPU tests of the multi-tensor optimisers (reference main.py:63-71, Test_Train.py:73-105):
single-launch Adam / SGD-Nesterov against torch's own implementations, checkpoint round trips in torch's
Optimizer_State layout, and a 20-epoch Adam trajectory of Discovery_Training (whole epoch replayed from one CUDA
graph) against the float64 oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import pde_read_oracle as orc
from tests.helpers import load_golden, rel_err, split_state
from tests.test_gpu_parity import DEV, make_net, to64

pytestmark = pytest.mark.gpu


def _param_set(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(50, 2), (50,), (50, 50), (50,), (1, 50), (1,), (4,), (3,), (7, 13)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]


def _set_grads(params, step, scale=1.0):
    g = torch.Generator().manual_seed(1000 + step)
    for p in params:
        p.grad = (torch.randn(*p.shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_fused_step_matches_torch(kind):
    from pde_read_b200.Optimizers import Fused_Adam, Fused_SGD
    pa, pb = _param_set(0), _param_set(0)
    if kind == "adam":
        ref, ours = torch.optim.Adam(pa, lr=3e-3), Fused_Adam(pb, lr=3e-3)
    else:
        ref, ours = (torch.optim.SGD(pa, lr=3e-3, momentum=0.9, nesterov=True),
                     Fused_SGD(pb, lr=3e-3, momentum=0.9, nesterov=True))
    for step in range(7):
        _set_grads(pa, step)
        _set_grads(pb, step)
        ref.step()
        ours.step()
    for a, b in zip(pa, pb):
        assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6
    assert ours.device_step_count() == 7
    # checkpoint layout: torch's own keys, both directions
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert sd_ref["state"].keys() == sd_ours["state"].keys()
    for k in sd_ref["state"]:
        assert sd_ref["state"][k].keys() == sd_ours["state"][k].keys()
        for name, v in sd_ref["state"][k].items():
            assert rel_err(sd_ours["state"][k][name].cpu().numpy(), v.cpu().numpy()) < 2e-5, (k, name)
    pc, pd = _param_set(0), _param_set(0)
    with torch.no_grad():
        for c, d, a in zip(pc, pd, pa):
            c.copy_(a)
            d.copy_(a)
    if kind == "adam":
        ref2, ours2 = torch.optim.Adam(pc, lr=3e-3), Fused_Adam(pd, lr=3e-3)
    else:
        ref2, ours2 = (torch.optim.SGD(pc, lr=3e-3, momentum=0.9, nesterov=True),
                       Fused_SGD(pd, lr=3e-3, momentum=0.9, nesterov=True))
    ref2.load_state_dict(copy.deepcopy(sd_ours))        # fused -> torch
    ours2.load_state_dict(copy.deepcopy(sd_ref))        # torch -> fused
    for step in range(7, 10):
        _set_grads(pc, step)
        _set_grads(pd, step)
        _set_grads(pa, step)
        ref.step()
        ref2.step()
        ours2.step()
    for a, c, d in zip(pa, pc, pd):
        assert rel_err(c.detach().cpu().numpy(), a.detach().cpu().numpy()) < 5e-6
        assert rel_err(d.detach().cpu().numpy(), a.detach().cpu().numpy()) < 5e-6


def _oracle_epoch(sol, pde, acts, m, n, P, Xd, yd, opt, params):
    """One reference epoch in float64 on the CPU: closure of Test_Train.py:73-102 + Optimizer.step."""
    loss_c, gs, gp = orc.collocation_loss_and_grads(sol, acts[0], pde, acts[1], m, n, P)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sol.items()}
    dl = orc.data_loss(leaf, acts[0], Xd, yd)
    dl.backward()
    for k in sol:
        sol[k].grad = gs[k] + leaf[k].grad
    for k in pde:
        pde[k].grad = gp[k].clone()
    opt.step()
    return float(loss_c) + float(dl.item())


def test_adam_trajectory_matches_oracle():
    """C1 shape (shipped Burgers checkpoint: U 5x50 Rational, N 2x100 Rational, m = 1, n = 2; 5000 collocation points
    redrawn every epoch + 1000 data points): 20 epochs of Discovery_Training with Fused_Adam -- each epoch ONE CUDA-graph
    replay of closure + optimiser step -- against 20 float64 epochs of the oracle with torch.optim.Adam."""
    from pde_read_b200 import Test_Train
    from pde_read_b200.Optimizers import Fused_Adam
    ck = load_golden("burgers_ckpt.npz")
    su, sn = split_state(ck, "sol."), split_state(ck, "pde.")
    dz = load_golden("burgers_data_loss.npz")
    Xd = torch.from_numpy(dz["inputs"]).float()
    yd = torch.from_numpy(dz["targets"]).double()
    bounds = dz["bounds"].astype(np.float64)
    epochs, lr, M = 20, 1e-3, 5000
    pts = []
    for t in range(epochs):
        g = torch.Generator().manual_seed(500 + t)
        u = torch.rand(M, 2, generator=g)
        pts.append((u * torch.tensor(bounds[:, 1] - bounds[:, 0], dtype=torch.float32)
                    + torch.tensor(bounds[:, 0], dtype=torch.float32)))

    # oracle, float64
    sol = {k: v.double().clone() for k, v in su.items()}
    pde = {k: v.double().clone() for k, v in sn.items()}
    opt_o = torch.optim.Adam(list(sol.values()) + list(pde.values()), lr=lr)
    want_losses = [_oracle_epoch(sol, pde, ("Rational", "Rational"), 1, 2, pts[t].double(), Xd.double(), yd, opt_o,
                                 None) for t in range(epochs)]

    # ours
    U, N = make_net(su, "Rational", 2), make_net(sn, "Rational", 3)
    start = {k: p.detach().clone() for k, p in list(U.named_parameters())}
    opt = Fused_Adam(list(U.parameters()) + list(N.parameters()), lr=lr)
    Xd_d, yd_d = Xd.to(DEV), yd.to(DEV)
    Test_Train.release_epoch_graphs()
    got_losses = []
    for t in range(epochs):
        Test_Train.Discovery_Training(U, N, 1, 2, pts[t].to(DEV), Xd_d, yd_d, opt, torch.float32, DEV)
        got_losses.append(float(Test_Train.LAST_EPOCH_LOSS.item()))
    assert len(Test_Train._EPOCH_GRAPHS) == 1, "the epoch must have been captured once and replayed"
    assert opt.device_step_count() == epochs
    np.testing.assert_allclose(got_losses, want_losses, rtol=2e-3)
    # parameters: Adam moves every entry by ~lr per step whatever its gradient, so entries whose gradient is at the
    # float32 noise floor may walk differently; the bulk must agree to a small fraction of the distance travelled
    moved, diff = [], []
    for net, ref in ((U, sol), (N, pde)):
        for k, p in net.named_parameters():
            a, b = p.detach().cpu().double().numpy().ravel(), ref[k].detach().numpy().ravel()
            diff.append(np.abs(a - b))
    diff = np.concatenate(diff)
    travelled = epochs * lr
    assert np.median(diff) < 2e-3 * travelled, np.median(diff)
    assert np.quantile(diff, 0.99) < 0.1 * travelled, np.quantile(diff, 0.99)
    assert diff.max() < 2.0 * travelled
    # and the parameters really moved
    assert max(float((p.detach() - start[k]).abs().max()) for k, p in U.named_parameters()) > 5 * lr


def test_graphed_epoch_equals_eager_fused_closure():
    """Same optimiser, same points: the CUDA-graph replay and the eager fused closure + launch take the same step."""
    from pde_read_b200 import Test_Train
    from pde_read_b200.Optimizers import Fused_Adam
    su = orc.init_network(3, 30, 2, "Tanh", seed=61)
    sn = orc.init_network(2, 20, 3, "Tanh", seed=62)
    gen = torch.Generator().manual_seed(63)
    Ps = [(torch.rand(900, 2, generator=gen) * 2 - 1).to(DEV) for _ in range(4)]
    Xd = (torch.rand(300, 2, generator=gen) * 2 - 1).to(DEV)
    yd = torch.cos(Xd[:, 1]) * Xd[:, 0]
    out = []
    old = Test_Train._TRAIN_GRAPH
    try:
        for graph in (False, True):
            Test_Train._TRAIN_GRAPH = graph
            Test_Train.release_epoch_graphs()
            U, N = make_net(su, "Tanh", 2), make_net(sn, "Tanh", 3)
            opt = Fused_Adam(list(U.parameters()) + list(N.parameters()), lr=2e-3)
            for P in Ps:
                Test_Train.Discovery_Training(U, N, 1, 2, P, Xd, yd, opt, torch.float32, DEV)
            out.append([p.detach().clone() for p in list(U.parameters()) + list(N.parameters())])
    finally:
        Test_Train._TRAIN_GRAPH = old
    for a, b in zip(*out):
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 2e-5
