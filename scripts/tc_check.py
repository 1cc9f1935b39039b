"""Dev tool (GPU): tensor-core path vs CUDA-core path on the same inputs, plus timings."""
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


which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
cfgs = [("tanh50", "Tanh", 5, 50, 1, 2), ("rat50", "Rational", 5, 50, 1, 3), ("rat100", "Rational", 5, 100, 1, 4),
        ("tanh100", "Tanh", 5, 100, 1, 4), ("sin24", "Sin", 3, 24, 2, 2), ("rat40", "Rational", 2, 40, 1, 1),
        ("tanh64", "Tanh", 4, 64, 1, 2), ("tanh128", "Tanh", 3, 128, 1, 2), ("tanh20", "Tanh", 6, 20, 2, 4)]
for name, act, L, W, m, n in cfgs:
    U = mk(orc.init_network(L, W, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(L, W, n + 1, act, seed=2), act, n + 1)
    su, sn = F.net_spec(U), F.net_spec(N)
    tu, tn = F.net_tensors(U), F.net_tensors(N)
    for M in (1000, 77, 200_000):
        P = torch.rand(M, 2, device=dev) * 2 - 1
        out = {}
        for tc in (0, 1):
            lib.pderead_debug_set_tensor_cores(2 if tc else 0)
            dtm, dxn = F.sol_derivatives_forward(su, tu, m, n, P)
            R, ss = F.residual_forward(su, tu, sn, tn, m, n, P)
            torch.cuda.synchronize()
            out[tc] = (dtm.clone(), dxn.clone(), R.clone(), ss.clone())
            if M >= 100_000:
                out[("t", tc)] = (timeit(lambda: F.sol_derivatives_forward(su, tu, m, n, P)),
                                  timeit(lambda: F.residual_forward(su, tu, sn, tn, m, n, P, want_resid=False)))
        r = dict(cfg=name, M=M, dtm=rel(out[1][0], out[0][0]),
                 dxn=[rel(out[1][1][:, k], out[0][1][:, k]) for k in range(n + 1)],
                 R=rel(out[1][2], out[0][2]), ss=rel(out[1][3], out[0][3]))
        if M >= 100_000:
            r["ufwd_ms_fma_tc"] = [out[("t", 0)][0], out[("t", 1)][0]]
            r["resid_ms_fma_tc"] = [out[("t", 0)][1], out[("t", 1)][1]]
        print(json.dumps(r), flush=True)
lib.pderead_debug_set_tensor_cores(1)
