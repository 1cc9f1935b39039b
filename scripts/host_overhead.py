"""Dev tool (GPU): host-side time to enqueue one fused step (no synchronisation inside the timed region)."""
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


U = mk(orc.init_network(5, 50, 2, "Rational", seed=1), "Rational", 2)
N = mk(orc.init_network(2, 100, 3, "Rational", seed=2), "Rational", 3)
su, sn = F.net_spec(U), F.net_spec(N)
tu, tn = F.net_tensors(U), F.net_tensors(N)
P = torch.rand(5000, 2, device=dev)
params = list(U.parameters()) + list(N.parameters())
for name, fn in (("collocation_loss_grad (ctypes call + torch.zeros)", lambda: F.collocation_loss_grad(su, tu, sn, tn, 1, 2, P)),
                 ("Collocation_Loss + backward (public API)", lambda: Collocation_Loss(U, N, 1, 2, P, torch.float32, dev).backward())):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("%-55s host enqueue time: median %.1f us, p10 %.1f us" % (name, ts[100] * 1e6, ts[20] * 1e6))

import cProfile
import pstats

pr = cProfile.Profile()
fn = lambda: Collocation_Loss(U, N, 1, 2, P, torch.float32, dev).backward()
torch.cuda.synchronize()
pr.enable()
for _ in range(200):
    fn()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
