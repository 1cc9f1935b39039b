"""Quick on-box probe: FMA peaks and raw kernel timings for the benchmark shapes (dev tool)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import functional as F           # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
out = {}
for v in (0, 1):
    out["fma_peak_variant%d_tflops" % v] = F.fma_peak_probe(dev, v, 8192)
print(json.dumps(out), flush=True)


def mk(st, act, din):
    L = orc.num_hidden_layers(st)
    W = st["Layers.0.weight"].shape[0]
    net = Neural_Network(L, W, din, 1, Device=dev, Activation_Function=act)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


cfgs = [("C2", "Tanh", 50, 1, 2, 100_000, 303000), ("C3", "Rational", 50, 1, 3, 125_000, 363600),
        ("C4", "Rational", 100, 1, 4, 500_000, 1688400), ("C4t", "Tanh", 100, 1, 4, 500_000, 1688400)]
for name, act, W, m, n, M, flop in cfgs:
    U = mk(orc.init_network(5, W, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(5, W, n + 1, act, seed=2), act, n + 1)
    su, sn = F.net_spec(U), F.net_spec(N)
    tu, tn = F.net_tensors(U), F.net_tensors(N)
    P = torch.rand(M, 2, device=dev)
    t_fb = timeit(lambda: F.collocation_loss_grad(su, tu, sn, tn, m, n, P))
    t_f = timeit(lambda: F.residual_forward(su, tu, sn, tn, m, n, P, want_resid=False))
    t_u = timeit(lambda: F.sol_derivatives_forward(su, tu, m, n, P))
    r = dict(cfg=name, M=M, fwdbwd_ms=t_fb, fwd_ms=t_f, ufwd_ms=t_u, pts_per_s=M / (t_fb[0] * 1e-3),
             tflops=flop * M / (t_fb[0] * 1e-3) / 1e12)
    print(json.dumps(r), flush=True)
