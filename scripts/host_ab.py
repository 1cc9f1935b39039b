"""Dev tool (GPU): host wall time per call of the plain-MLP entry points with the register-tiled forward kernel on / off."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import _lib, functional as F     # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


U = mk(orc.init_network(5, 50, 2, "Rational", seed=1), "Rational", 2)
N = mk(orc.init_network(2, 100, 3, "Rational", seed=2), "Rational", 3)
su, sn, tu, tn = F.net_spec(U), F.net_spec(N), F.net_tensors(U), F.net_tensors(N)
X = torch.rand(5000, 3, device=dev)
P = torch.rand(5000, 2, device=dev)
go = torch.ones(5000, device=dev)
for on in (1, 0, 1, 0):
    lib.pderead_debug_set_plain_mlp(on)
    for name, fn in (("mlp_forward", lambda: F.mlp_forward(sn, tn, X)),
                     ("mlp_backward", lambda: F.mlp_backward(sn, tn, X, go, True)),
                     ("colloc_loss_grad", lambda: F.collocation_loss_grad(su, tu, sn, tn, 1, 2, P))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("plain=%d %-18s host enqueue %.1f us/call, with drain %.1f us/call" % (on, name, (t1 - t0) / 300 * 1e6, (t2 - t0) / 300 * 1e6), flush=True)
