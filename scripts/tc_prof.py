"""Dev tool (GPU): a few launches of one path for ncu.  usage: tc_prof.py <cfg> <what> [M]"""
import sys

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import functional as F           # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
CFG = {"C2": ("Tanh", 5, 50, 1, 2), "C3": ("Rational", 5, 50, 1, 3), "C4": ("Rational", 5, 100, 1, 4)}
act, L, W, m, n = CFG[sys.argv[1]]
what = sys.argv[2]
M = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


U = mk(orc.init_network(L, W, 2, act, seed=1), act, 2)
N = mk(orc.init_network(L, W, n + 1, act, seed=2), act, n + 1)
su, sn = F.net_spec(U), F.net_spec(N)
tu, tn = F.net_tensors(U), F.net_tensors(N)
P = torch.rand(M, 2, device=dev) * 2 - 1
for _ in range(3):
    if what == "ufwd":
        F.sol_derivatives_forward(su, tu, m, n, P)
    elif what == "resid":
        F.residual_forward(su, tu, sn, tn, m, n, P, want_resid=False)
    else:
        F.collocation_loss_grad(su, tu, sn, tn, m, n, P)
torch.cuda.synchronize()
print("ok")
