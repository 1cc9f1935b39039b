"""Dev tool (GPU): one tiny fused step + forward + extraction in both kernel families (for compute-sanitizer)."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import _lib, functional as F     # noqa: E402
from pde_read_b200.Extraction import Extract_PDE    # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


for mode in (0, 2):
    lib.pderead_debug_set_tensor_cores(mode)
    for act, L, W, m, n, M in (("Tanh", 3, 50, 1, 2, 333), ("Rational", 3, 100, 1, 3, 97), ("Sin", 2, 24, 2, 1, 65)):
        U = mk(orc.init_network(L, W, 2, act, seed=1), act, 2)
        N = mk(orc.init_network(2, W, n + 1, act, seed=2), act, n + 1)
        su, sn = F.net_spec(U), F.net_spec(N)
        tu, tn = F.net_tensors(U), F.net_tensors(N)
        P = torch.rand(M, 2, device=dev) * 2 - 1
        ss, flat, vu, vn, R = F.collocation_loss_grad(su, tu, sn, tn, m, n, P, want_resid=True)
        dtm, dxn = F.sol_derivatives_forward(su, tu, m, n, P)
        F.residual_forward(su, tu, sn, tn, m, n, P)
        Extract_PDE(U, N, m, n, P, 2)
        torch.cuda.synchronize()
        print("mode", mode, act, W, "ok", float(ss), flush=True)
lib.pderead_debug_set_tensor_cores(1)
