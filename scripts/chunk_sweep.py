"""Dev tool (GPU): whole fused step time vs points per chunk (is an L2-resident stash worth more launches?)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import functional as F           # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
CFG = {"C2": ("Tanh", 5, 50, 5, 50, 1, 2, 100_000), "C3": ("Rational", 5, 50, 5, 50, 1, 3, 250_000),
       "C4": ("Rational", 5, 100, 5, 100, 1, 4, 200_000)}


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in (sys.argv[1:] or ["C2"]):
    act, LU, WU, LN, WN, m, n, M = CFG[name]
    U = mk(orc.init_network(LU, WU, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(LN, WN, n + 1, act, seed=2), act, n + 1)
    su, sn = F.net_spec(U), F.net_spec(N)
    tu, tn = F.net_tensors(U), F.net_tensors(N)
    P = torch.rand(M, 2, device=dev) * 2 - 1
    for cap in (None, 65536, 37888, 18944, 9472, 4736):
        F.WORKSPACE_CAP_POINTS = cap
        F.release_workspaces()
        for _ in range(3):
            F.collocation_loss_grad(su, tu, sn, tn, m, n, P)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            F.collocation_loss_grad(su, tu, sn, tn, m, n, P)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps(dict(cfg=name, M=M, cap=cap, ms_min=min(ts), ms_med=sorted(ts)[3])), flush=True)
F.WORKSPACE_CAP_POINTS = None
