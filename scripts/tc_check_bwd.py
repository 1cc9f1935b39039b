"""Dev tool (GPU): tensor-core reverse path vs CUDA-core reverse path on the same inputs, plus timings."""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import _lib, functional as F     # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


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
    return min(ts)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


cfgs = [("tanh50", "Tanh", 5, 50, 1, 2), ("rat50", "Rational", 5, 50, 1, 3), ("rat40", "Rational", 2, 40, 1, 1),
        ("sin24", "Sin", 3, 24, 2, 2), ("tanh64", "Tanh", 4, 64, 1, 2), ("tanh20", "Tanh", 6, 20, 2, 4),
        ("rat100", "Rational", 5, 100, 1, 4), ("tanh100n2", "Tanh", 2, 100, 1, 2)]
only = sys.argv[1:] or None
for name, act, L, W, m, n in cfgs:
    if only and name not in only:
        continue
    U = mk(orc.init_network(L, W, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(L, W, n + 1, act, seed=2), act, n + 1)
    su, sn = F.net_spec(U), F.net_spec(N)
    tu, tn = F.net_tensors(U), F.net_tensors(N)
    for M in (1000, 77, 200_000):
        P = torch.rand(M, 2, device=dev) * 2 - 1
        out = {}
        for tc in (0, 1):
            lib.pderead_debug_set_tensor_cores(2 if tc else 0)
            ss, flat, vu, vn, R = F.collocation_loss_grad(su, tu, sn, tn, m, n, P, want_resid=True)
            torch.cuda.synchronize()
            out[tc] = (ss.clone(), [v.clone() for v in vu], [v.clone() for v in vn], R.clone())
            if M >= 100_000:
                out[("t", tc)] = timeit(lambda: F.collocation_loss_grad(su, tu, sn, tn, m, n, P))
        r = dict(cfg=name, M=M, ss=rel(out[1][0], out[0][0]), R=rel(out[1][3], out[0][3]),
                 gU=max(rel(a, b) for a, b in zip(out[1][1], out[0][1])),
                 gN=max(rel(a, b) for a, b in zip(out[1][2], out[0][2])),
                 gU_each=[round(rel(a, b), 7) for a, b in zip(out[1][1], out[0][1])][:6])
        if M >= 100_000:
            r["fwdbwd_ms_fma_tc"] = [out[("t", 0)], out[("t", 1)]]
        print(json.dumps(r), flush=True)
lib.pderead_debug_set_tensor_cores(1)
