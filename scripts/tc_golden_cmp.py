"""Dev tool (GPU): both kernel paths against the fp32 / fp64 golden references of the shipped checkpoint,
and against the fp64 oracle for a deep narrow Tanh net."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc
from pde_read_b200 import _lib, functional as F
from pde_read_b200.Loss_Functions import Collocation_Loss
from pde_read_b200.Network import Neural_Network
from tests.helpers import load_golden, rel_err, split_state

dev = torch.device("cuda:0")
lib = _lib.load()


def mk(st, act, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=act)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


ck, z = load_golden("burgers_ckpt.npz"), load_golden("burgers_colloc.npz")
P = torch.from_numpy(z["coords"]).to(dev)
for tc in (0, 1):
    lib.pderead_debug_set_tensor_cores(2 if tc else 0)
    U = mk(split_state(ck, "sol."), "Rational", 2)
    N = mk(split_state(ck, "pde."), "Rational", 3)
    loss = Collocation_Loss(U, N, 1, 2, P, torch.float32, dev)
    loss.backward()
    print("tc", tc, "loss", loss.item())
    for pref, net in (("gsol.", U), ("gpde.", N)):
        for k, p in net.named_parameters():
            g = p.grad.cpu().numpy()
            e32, e64, noise = rel_err(g, z[pref + k]), rel_err(g, z[pref + k + ".f64"]), rel_err(z[pref + k], z[pref + k + ".f64"])
            flag = "  <-- FAIL" if (e32 > 1e-4 and e64 > max(1e-4, 1.5 * noise)) else ""
            print("  %-40s e32 %.2e e64 %.2e refnoise %.2e%s" % (pref + k, e32, e64, noise, flag))

# deep narrow tanh net vs the fp64 oracle
m, n, M = 2, 4, 1000
su = orc.init_network(6, 20, 2, "Tanh", seed=1)
sn = orc.init_network(6, 20, n + 1, "Tanh", seed=2)
Pc = torch.rand(M, 2, generator=torch.Generator().manual_seed(5)) * 2 - 1
want, gs, gp = orc.collocation_loss_and_grads({k: v.double() for k, v in su.items()}, "Tanh",
                                              {k: v.double() for k, v in sn.items()}, "Tanh", m, n, Pc.double())
w32, gs32, gp32 = orc.collocation_loss_and_grads(su, "Tanh", sn, "Tanh", m, n, Pc)
for tc in (0, 1):
    lib.pderead_debug_set_tensor_cores(2 if tc else 0)
    U, N = mk(su, "Tanh", 2), mk(sn, "Tanh", n + 1)
    loss = Collocation_Loss(U, N, m, n, Pc.to(dev), torch.float32, dev)
    loss.backward()
    worst = 0
    for net, ref, r32 in ((U, gs, gs32), (N, gp, gp32)):
        for k, p in net.named_parameters():
            e = rel_err(p.grad.cpu().numpy(), ref[k].numpy())
            worst = max(worst, e)
    noise = max(rel_err(r32[k].numpy(), ref[k].numpy()) for ref, r32 in ((gs, gs32), (gp, gp32)) for k in ref)
    print("tanh20 tc", tc, "loss rel", abs(loss.item() - want) / want, "worst grad rel vs f64", worst, "torch-fp32 noise", noise)
lib.pderead_debug_set_tensor_cores(1)
