"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules from /root/reference/Code (read-only), feeds them seeded
inputs and stores inputs + outputs as small .npz fixtures.  Nothing from the reference's sources
is copied; only numbers it computed.  The fixtures pin ``oracle/pde_read_oracle.py`` (CPU tests)
and are compared directly with the CUDA path (GPU tests).
"""
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "Code"))

import Extraction            # noqa: E402  (reference)
import Loss_Functions        # noqa: E402  (reference)
import Network               # noqa: E402  (reference)
import PDE_Residual          # noqa: E402  (reference)
import Points                # noqa: E402  (reference)

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def state_np(prefix, net):
    # (copies: BatchNorm buffers are updated in place by a later forward pass)
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}


def grads_np(prefix, net):
    out = {}
    for k, p in net.named_parameters():
        out[prefix + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy().copy()
    return out


def load_ckpt_nets(dtype=torch.float32):
    saved = torch.load(os.path.join(REF, "Saves", "Burgers_Sine_N50_P5000_Rat_Adam"), map_location="cpu")
    U = Network.Neural_Network(Num_Hidden_Layers=5, Neurons_Per_Layer=50, Input_Dim=2, Output_Dim=1,
                               Data_Type=dtype, Activation_Function="Rational")
    N = Network.Neural_Network(Num_Hidden_Layers=2, Neurons_Per_Layer=100, Input_Dim=3, Output_Dim=1,
                               Data_Type=dtype, Activation_Function="Rational")
    U.load_state_dict(saved["Sol_Network_State"])
    N.load_state_dict(saved["PDE_Network_State"])
    return U.to(dtype), N.to(dtype)


def run_colloc(U, N, m, n, coords, dtype):
    """Reference forward + backward; returns dict of outputs."""
    for net in (U, N):
        for p in net.parameters():
            p.grad = None
    C = coords.detach().clone().to(dtype)
    Dtm, Dxn = PDE_Residual.Evaluate_Derivatives(U, m, n, C, Data_Type=dtype)
    R = Dtm - N(Dxn).view(-1)
    C2 = coords.detach().clone().to(dtype)
    loss = Loss_Functions.Collocation_Loss(U, N, m, n, C2, Data_Type=dtype)
    loss.backward()
    return dict(dtm=Dtm.detach().numpy().copy(), dxn=Dxn.detach().numpy().copy(),
                resid=R.detach().numpy().copy(), loss=np.float64(loss.item()))


def golden_checkpoint():
    U, N = load_ckpt_nets()
    out = {}
    out.update(state_np("sol.", U))
    out.update(state_np("pde.", N))
    np.savez_compressed(os.path.join(HERE, "burgers_ckpt.npz"), **out)

    # --- collocation: SURVEY 8c item 2 --------------------------------------------------
    g = torch.Generator().manual_seed(0)
    P = torch.rand((5000, 2), generator=g) * torch.tensor([10.0, 15.9375]) + torch.tensor([0.0, -8.0])
    r32 = run_colloc(U, N, 1, 2, P, torch.float32)
    g32 = {}
    g32.update(grads_np("gsol.", U))
    g32.update(grads_np("gpde.", N))
    U64, N64 = load_ckpt_nets(torch.float64)
    r64 = run_colloc(U64, N64, 1, 2, P, torch.float64)
    g64 = {}
    # float64-reference gradients are stored rounded to float32 (tie-breaker only)
    g64.update({k + ".f64": v.astype(np.float32) for k, v in grads_np("gsol.", U64).items()})
    g64.update({k + ".f64": v.astype(np.float32) for k, v in grads_np("gpde.", N64).items()})
    H = 512   # per-point outputs kept for the first H points; loss and gradients cover all 5000
    np.savez_compressed(os.path.join(HERE, "burgers_colloc.npz"),
                        coords=P.numpy(), dtm=r32["dtm"][:H], dxn=r32["dxn"][:H], resid=r32["resid"][:H],
                        loss=r32["loss"], loss_f64=r64["loss"],
                        dtm_f64=r64["dtm"][:H], dxn_f64=r64["dxn"][:H], resid_f64=r64["resid"][:H],
                        **g32, **g64)
    print("colloc: first point", P[0].tolist(), "dtm", r32["dtm"][0], "dxn", r32["dxn"][0],
          "R", r32["resid"][0], "loss", r32["loss"], "loss64", r64["loss"])

    # --- data loss on a slice of the shipped data set -------------------------------------
    d = np.load(os.path.join(REF, "Data", "DataSets", "Burgers_Sine_N50_P5000.npz"))
    Xd = torch.from_numpy(d["Train_Inputs"][:256])
    Yd = torch.from_numpy(d["Train_Targets"][:256])
    for p in U.parameters():
        p.grad = None
    dl = Loss_Functions.Data_Loss(U, Xd, Yd)
    dl.backward()
    np.savez_compressed(os.path.join(HERE, "burgers_data_loss.npz"),
                        inputs=Xd.numpy(), targets=Yd.numpy(), loss=np.float64(dl.item()),
                        bounds=d["Input_Bounds"], **grads_np("gsol.", U))
    print("data loss", dl.item(), dl.dtype)

    # --- extraction: SURVEY 8c item 3 -------------------------------------------------------
    random.seed(0)
    bounds = d["Input_Bounds"].astype(np.float64)
    E = Points.Generate_Points(bounds, 20000, torch.float32, torch.device('cpu'))
    b, Lib, counts, mil = Extraction.Generate_Library(U, N, 1, 2, E, 3)
    X, Res = Extraction.Recursive_Feature_Elimination(Lib, b)
    Xr, Rr, chg = Extraction.Rank_Candidate_Solutions(X, Res)
    # elimination order = feature that leaves the support between consecutive candidates
    C = Lib.shape[1]
    order = []
    for j in range(C):
        sup = set(np.flatnonzero(X[:, j]).tolist())
        nxt = set(np.flatnonzero(X[:, j + 1]).tolist()) if j + 1 < C else set()
        gone = sorted(sup - nxt)
        order.append(gone[0] if len(gone) == 1 else -1)
    np.savez_compressed(os.path.join(HERE, "burgers_extract_k3.npz"),
                        coords_head=E[:512].detach().numpy(), b_head=b[:512], lib_head=Lib[:512],
                        b_norm=np.float64(np.linalg.norm(b.astype(np.float64))),
                        counts=counts, X=X, Residual=Res, order=np.asarray(order, dtype=np.int64),
                        X_Ranked=Xr, Residual_Ranked=Rr, Residual_Change=chg)
    print("extract: first point", E[0].tolist(), "order", order)
    print("Residual", Res)


def golden_random():
    """Random-init nets through the reference constructor (Network.py:133-166) for every
    activation and derivative order the kernels are instantiated for."""
    cases = [
        # name,           actU,       LU, WU, actN,       LN, WN, m, n, M
        ("tanh_m1n2",     "Tanh",      3, 24, "Tanh",      2, 16, 1, 2, 96),
        ("tanh_m2n3",     "Tanh",      2, 20, "Sin",       2, 12, 2, 3, 70),
        ("sin_m1n1",      "Sin",       2, 18, "Tanh",      3, 10, 1, 1, 64),
        ("sin_m2n4",      "Sin",       3, 16, "Rational",  2, 14, 2, 4, 50),
        ("rat_m1n3",      "Rational",  4, 22, "Rational",  2, 20, 1, 3, 80),
        ("rat_m1n4",      "Rational",  3, 30, "Tanh",      2, 24, 1, 4, 64),
        ("rat_m2n2",      "Rational",  2, 26, "Rational",  3, 18, 2, 2, 33),
        ("tanh_w50_n2",   "Tanh",      5, 50, "Tanh",      5, 50, 1, 2, 128),
        # the BASELINE shapes C3 (KdV: Rational 5x50 both, n = 3) and C4 (KS: Rational 5x100 both, n = 4) at a size the
        # reference's nested autograd holds (C2's shape is the entry above, C1 / C5 the shipped checkpoint)
        ("c3_rat_w50_n3",  "Rational",  5, 50, "Rational",  5, 50, 1, 3, 96),
        ("c4_rat_w100_n4", "Rational",  5, 100, "Rational", 5, 100, 1, 4, 64),
    ]
    only = os.environ.get("GOLDEN_ONLY")
    for idx, (name, aU, LU, WU, aN, LN, WN, m, n, M) in enumerate(cases):
        if only and name not in only.split(","):
            continue
        torch.manual_seed(100 + idx)
        U = Network.Neural_Network(LU, WU, 2, 1, Activation_Function=aU)
        N = Network.Neural_Network(LN, WN, n + 1, 1, Activation_Function=aN)
        # non-zero biases and perturbed rational coefficients so every gradient path is exercised
        with torch.no_grad():
            for net in (U, N):
                for k, p in net.named_parameters():
                    if k.endswith(".bias"):
                        p.uniform_(-0.3, 0.3)
                    if k.endswith(".a") or k.endswith(".b"):
                        p.add_(torch.randn_like(p) * 0.05)
        P = torch.rand((M, 2)) * torch.tensor([2.0, 3.0]) + torch.tensor([0.0, -1.5])
        r32 = run_colloc(U, N, m, n, P, torch.float32)
        out = dict(coords=P.numpy(), dtm=r32["dtm"], dxn=r32["dxn"], resid=r32["resid"], loss=r32["loss"],
                   meta=np.array([LU, WU, LN, WN, m, n], dtype=np.int64),
                   acts=np.array([aU, aN]))
        out.update(state_np("sol.", U))
        out.update(state_np("pde.", N))
        out.update(grads_np("gsol.", U))
        out.update(grads_np("gpde.", N))
        U64 = Network.Neural_Network(LU, WU, 2, 1, Data_Type=torch.float64, Activation_Function=aU)
        N64 = Network.Neural_Network(LN, WN, n + 1, 1, Data_Type=torch.float64, Activation_Function=aN)
        U64.load_state_dict(U.state_dict())
        N64.load_state_dict(N.state_dict())
        U64, N64 = U64.double(), N64.double()
        r64 = run_colloc(U64, N64, m, n, P, torch.float64)
        out.update(dtm_f64=r64["dtm"], dxn_f64=r64["dxn"], resid_f64=r64["resid"], loss_f64=r64["loss"])
        out.update({k + ".f64": v.astype(np.float32) for k, v in grads_np("gsol.", U64).items()})
        out.update({k + ".f64": v.astype(np.float32) for k, v in grads_np("gpde.", N64).items()})
        np.savez_compressed(os.path.join(HERE, "random_%s.npz" % name), **out)
        print(name, "loss", r32["loss"], "loss64", r64["loss"])


def golden_library_small():
    """Library column-order KAT of the reference's test_Generate_Library (Testing.py:436-489),
    run against the current API: n=1, K=2 and n=2, K=3 on a small tanh net."""
    torch.manual_seed(7)
    U = Network.Neural_Network(2, 12, 2, 1, Activation_Function="Tanh")
    for n, K in ((1, 2), (2, 3), (3, 2)):
        N = Network.Neural_Network(2, 8, n + 1, 1, Activation_Function="Tanh")
        P = torch.rand((40, 2))
        b, Lib, counts, mil = Extraction.Generate_Library(U, N, 1, n, P, K)
        out = dict(coords=P.detach().numpy(), b=b, Library=Lib, counts=counts)
        for k in range(1, K + 1):
            out["mi_%d" % k] = mil[k]
        out.update(state_np("sol.", U))
        out.update(state_np("pde.", N))
        np.savez_compressed(os.path.join(HERE, "library_n%d_k%d.npz" % (n, K)), **out)


def golden_extract_k5():
    """BASELINE config 5 shape at a size the reference finishes in seconds: shipped checkpoint, random.seed(0),
    20 000 extraction points, K = 5 (C = 56) through the unmodified Extraction.py:178-572."""
    U, N = load_ckpt_nets()
    d = np.load(os.path.join(REF, "Data", "DataSets", "Burgers_Sine_N50_P5000.npz"))
    random.seed(0)
    bounds = d["Input_Bounds"].astype(np.float64)
    E = Points.Generate_Points(bounds, 20000, torch.float32, torch.device('cpu'))
    b, Lib, counts, mil = Extraction.Generate_Library(U, N, 1, 2, E, 5)
    X, Res = Extraction.Recursive_Feature_Elimination(Lib, b)
    Xr, Rr, chg = Extraction.Rank_Candidate_Solutions(X, Res)
    C = Lib.shape[1]
    order = []
    for j in range(C):
        sup = set(np.flatnonzero(X[:, j]).tolist())
        nxt = set(np.flatnonzero(X[:, j + 1]).tolist()) if j + 1 < C else set()
        gone = sorted(sup - nxt)
        order.append(gone[0] if len(gone) == 1 else -1)
    # margin of every elimination: (second smallest - smallest) / smallest |x| among the remaining features,
    # in the unit-norm scaling RFE compares in (Extraction.py:467-476) -- lets the tests skip near-ties
    norms = np.sqrt((Lib.astype(np.float64) ** 2).sum(axis=0))
    margin = np.zeros(C)
    for j in range(C):
        xs = np.sort(np.abs(X[:, j].astype(np.float64) * norms)[np.flatnonzero(X[:, j])])
        margin[j] = (xs[1] - xs[0]) / xs[0] if xs.size > 1 and xs[0] > 0 else np.inf
    np.savez_compressed(os.path.join(HERE, "burgers_extract_k5.npz"),
                        coords_head=E[:256].detach().numpy(), b_head=b[:256], lib_head=Lib[:256],
                        counts=counts, X=X, Residual=Res, order=np.asarray(order, dtype=np.int64),
                        margin=margin, X_Ranked=Xr, Residual_Ranked=Rr, Residual_Change=chg)
    print("extract k5: order", order)
    print("Residual", Res)
    print("margins", np.round(margin, 4))


def golden_batchnorm():
    """PDE network with "Normalize Inputs" (Network.py:100-104,194-195: BatchNorm1d on the inputs of N) through the
    unmodified reference: train-mode collocation loss + every gradient (incl. Norm_Layer.weight / .bias) + the running
    statistics after the call, then the eval-mode loss."""
    torch.manual_seed(321)
    m, n, M = 1, 2, 80
    U = Network.Neural_Network(3, 20, 2, 1, Activation_Function="Tanh")
    N = Network.Neural_Network(2, 16, n + 1, 1, Activation_Function="Rational", Batch_Norm=True)
    with torch.no_grad():
        for net in (U, N):
            for k, p in net.named_parameters():
                if k.endswith(".bias") and not k.startswith("Norm_Layer"):
                    p.uniform_(-0.3, 0.3)
        N.Norm_Layer.weight.copy_(torch.tensor([1.3, 0.7, 1.1]))
        N.Norm_Layer.bias.copy_(torch.tensor([0.1, -0.2, 0.05]))
    P = torch.rand((M, 2)) * torch.tensor([2.0, 3.0]) + torch.tensor([0.0, -1.5])
    out = dict(coords=P.numpy())
    out.update(state_np("sol.", U))
    out.update(state_np("pde.", N))          # includes Norm_Layer.* and the initial running statistics
    U.train(); N.train()
    loss = Loss_Functions.Collocation_Loss(U, N, m, n, P.clone())
    loss.backward()
    out["loss_train"] = np.float64(loss.item())
    out.update(grads_np("gsol.", U))
    out.update(grads_np("gpde.", N))
    out["running_mean_after"] = N.Norm_Layer.running_mean.numpy().copy()
    out["running_var_after"] = N.Norm_Layer.running_var.numpy().copy()
    out["num_batches_after"] = np.int64(N.Norm_Layer.num_batches_tracked.item())
    U.eval(); N.eval()
    out["loss_eval"] = np.float64(Loss_Functions.Collocation_Loss(U, N, m, n, P.clone()).item())
    np.savez_compressed(os.path.join(HERE, "batchnorm_rat_m1n2.npz"), **out)
    print("batchnorm: train loss", out["loss_train"], "eval loss", out["loss_eval"], "running mean", out["running_mean_after"])


def golden_stock_artifacts():
    """The reference's own artifacts for the end-to-end drop-in test, as DATA: the text of its stock
    Settings.txt with only the four lines a user must touch to extract from the shipped checkpoint on a GPU changed
    (load both networks, Extraction mode, GPU), the Adam state of the shipped checkpoint for two parameter tensors
    (layout check), and a slice of the shipped data set (the driver needs Input_Bounds and the four arrays)."""
    with open(os.path.join(REF, "Settings.txt")) as fh:
        text = fh.read()
    edits = (("Load Sol Network State [bool] :                  False", "Load Sol Network State [bool] :                  True"),
             ("Load PDE Network State [bool] :                  False", "Load PDE Network State [bool] :                  True"),
             ("Discovery, or Extraction mode [str] :            Discovery", "Discovery, or Extraction mode [str] :            Extraction"),
             ("Train on CPU or GPU [GPU, CPU] :                 CPU", "Train on CPU or GPU [GPU, CPU] :                 GPU"))
    for a, b in edits:
        assert text.count(a) == 1, a
        text = text.replace(a, b)
    d = np.load(os.path.join(REF, "Data", "DataSets", "Burgers_Sine_N50_P5000.npz"))
    saved = torch.load(os.path.join(REF, "Saves", "Burgers_Sine_N50_P5000_Rat_Adam"), map_location="cpu")
    opt = saved["Optimizer_State"]
    np.savez_compressed(os.path.join(HERE, "stock_artifacts.npz"),
                        settings_text=np.frombuffer(text.encode(), dtype=np.uint8),
                        Train_Inputs=d["Train_Inputs"][:512], Train_Targets=d["Train_Targets"][:512],
                        Test_Inputs=d["Test_Inputs"][:128], Test_Targets=d["Test_Targets"][:128],
                        Input_Bounds=d["Input_Bounds"],
                        opt_step=np.float64(float(opt["state"][0]["step"])),
                        opt_exp_avg_0=opt["state"][0]["exp_avg"].numpy(), opt_exp_avg_sq_0=opt["state"][0]["exp_avg_sq"].numpy(),
                        opt_num_params=np.int64(len(opt["param_groups"][0]["params"])),
                        opt_lr=np.float64(opt["param_groups"][0]["lr"]))
    print("stock artifacts: settings %d bytes, optimizer step %s, %d params" % (len(text), opt["state"][0]["step"],
                                                                                 len(opt["param_groups"][0]["params"])))


if __name__ == "__main__":
    golden_checkpoint()
    golden_random()
    golden_library_small()
    golden_extract_k5()
    golden_batchnorm()
    golden_stock_artifacts()
    print("done")
