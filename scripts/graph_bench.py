"""Dev tool (GPU): training-closure time per step with and without CUDA-graph replay."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import functional as F           # noqa: E402
from pde_read_b200.Loss_Functions import Collocation_Loss   # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


for name, act, LU, WU, LN, WN, m, n, M in (("C1", "Rational", 5, 50, 2, 100, 1, 2, 5000), ("C2", "Tanh", 5, 50, 5, 50, 1, 2, 100_000)):
    U = mk(orc.init_network(LU, WU, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(LN, WN, n + 1, act, seed=2), act, n + 1)
    params = list(U.parameters()) + list(N.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    P = torch.rand(M, 2, device=dev) * 2 - 1
    for graphs in (False, True):
        F.USE_CUDA_GRAPHS = graphs

        def closure():
            opt.zero_grad()
            loss = Collocation_Loss(U, N, m, n, P, torch.float32, dev)
            loss.backward()
            return loss
        for _ in range(5):
            opt.step(closure)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 50
        for _ in range(K):
            opt.step(closure)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print(json.dumps(dict(cfg=name, M=M, cuda_graphs=graphs, ms_per_adam_step=dt * 1e3, pts_per_s=M / dt)), flush=True)
