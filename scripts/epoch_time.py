"""Dev tool (GPU): wall time of one Discovery_Training epoch (Adam) at the reference's default sizes, with the
autograd closure, the graph-free closure, the single-launch Fused_Adam, and the whole epoch replayed from one CUDA graph."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import pde_read_oracle as orc          # noqa: E402  (weights init only)
from pde_read_b200 import Test_Train                # noqa: E402
from pde_read_b200.Network import Neural_Network    # noqa: E402
from pde_read_b200.Optimizers import Fused_Adam     # noqa: E402

dev = torch.device("cuda:0")


def mk(st, a, din):
    net = Neural_Network(orc.num_hidden_layers(st), st["Layers.0.weight"].shape[0], din, 1, Device=dev, Activation_Function=a)
    net.load_state_dict({k: v.to(dev) for k, v in st.items()})
    return net


for M, Md in ((5000, 4000), (50000, 20000)):
    for fast in (False, True):
        Test_Train._FAST_CLOSURE = fast
        U = mk(orc.init_network(5, 50, 2, "Rational", seed=1), "Rational", 2)
        N = mk(orc.init_network(2, 100, 3, "Rational", seed=2), "Rational", 3)
        params = list(U.parameters()) + list(N.parameters())
        for foreach, fused in ((None, None), (None, True)):
            opt = torch.optim.Adam(params, lr=1e-3, foreach=foreach, fused=fused)
            P = torch.rand(M, 2, device=dev)
            Xd = torch.rand(Md, 2, device=dev)
            yd = torch.rand(Md, device=dev, dtype=torch.float64)
            for _ in range(30):
                Test_Train.Discovery_Training(U, N, 1, 2, P, Xd, yd, opt, torch.float32, dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            K = 300
            for _ in range(K):
                Test_Train.Discovery_Training(U, N, 1, 2, P, Xd, yd, opt, torch.float32, dev)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / K
            print("M=%d Md=%d fast_closure=%s adam_fused=%s: %.1f us/epoch" % (M, Md, fast, fused, dt * 1e6), flush=True)

    Test_Train._FAST_CLOSURE = True
    for graph in (False, True):
        Test_Train._TRAIN_GRAPH = graph
        Test_Train.release_epoch_graphs()
        U = mk(orc.init_network(5, 50, 2, "Rational", seed=1), "Rational", 2)
        N = mk(orc.init_network(2, 100, 3, "Rational", seed=2), "Rational", 3)
        opt = Fused_Adam(list(U.parameters()) + list(N.parameters()), lr=1e-3)
        Xd = torch.rand(Md, 2, device=dev)
        yd = torch.rand(Md, device=dev, dtype=torch.float64)
        Ps = [torch.rand(M, 2, device=dev) for _ in range(8)]      # new collocation points every epoch, as main.py does
        for i in range(30):
            Test_Train.Discovery_Training(U, N, 1, 2, Ps[i % 8], Xd, yd, opt, torch.float32, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 300
        for i in range(K):
            Test_Train.Discovery_Training(U, N, 1, 2, Ps[i % 8], Xd, yd, opt, torch.float32, dev)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print("M=%d Md=%d Fused_Adam whole-epoch graph=%s: %.1f us/epoch" % (M, Md, graph, dt * 1e6), flush=True)
