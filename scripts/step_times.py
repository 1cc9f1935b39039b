"""Dev tool (GPU): per-kernel times of one fused step, both paths, via the stage-event hook."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import _lib, functional as F     # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
CFG = {"C2": ("Tanh", 5, 50, 5, 50, 1, 2, 100_000), "C3": ("Rational", 5, 50, 5, 50, 1, 3, 125_000),
       "C4": ("Rational", 5, 100, 5, 100, 1, 4, 200_000), "C1": ("Rational", 5, 50, 2, 100, 1, 2, 5_000),
       "C4T": ("Tanh", 5, 100, 5, 100, 1, 4, 200_000)}


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


for name in (sys.argv[1:] or ["C2"]):
    act, LU, WU, LN, WN, m, n, M = CFG[name]
    U = mk(orc.init_network(LU, WU, 2, act, seed=1), act, 2)
    N = mk(orc.init_network(LN, WN, n + 1, act, seed=2), act, n + 1)
    su, sn = F.net_spec(U), F.net_spec(N)
    tu, tn = F.net_tensors(U), F.net_tensors(N)
    P = torch.rand(M, 2, device=dev) * 2 - 1
    for tc in (0, 1):
        lib.pderead_debug_set_tensor_cores(2 if tc else 0)
        for _ in range(3):
            F.collocation_loss_grad(su, tu, sn, tn, m, n, P)
        torch.cuda.synchronize()
        acc = None
        for rep in range(5):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(40)]
            e0 = torch.cuda.Event(enable_timing=True)
            for e in evs:
                e.record()
            torch.cuda.synchronize()
            arr = (ctypes.c_void_p * 40)(*[e.cuda_event for e in evs])
            lib.pderead_debug_stage_events(arr, 40)
            e0.record()
            F.collocation_loss_grad(su, tu, sn, tn, m, n, P)
            nrec = lib.pderead_debug_stage_events_recorded()
            lib.pderead_debug_stage_events(None, 0)
            torch.cuda.synchronize()
            ts, prev = [], e0
            for k in range(nrec):
                ts.append(prev.elapsed_time(evs[k]))
                prev = evs[k]
            acc = np.array(ts) if acc is None else np.minimum(acc, np.array(ts))
        print(json.dumps(dict(cfg=name, tc=tc, M=M, total_ms=float(acc.sum()), stages_ms=[round(float(x), 4) for x in acc])), flush=True)
lib.pderead_debug_set_tensor_cores(1)
