"""End-to-end GPU test of the drop-in: the Settings.txt-driven driver of the reference
(`Code/main.py:14-287`) on a synthetic data set -- Discovery (Adam, collocation + data loss, checkpoint with
the reference's keys), then a second Discovery run that resumes from the checkpoint (LBFGS: the closure is
re-entered several times per step), then Extraction from the saved networks.

The data set is the exact solution u = exp(-0.5 t) sin(x) of the heat equation u_t = 0.5 u_xx in the npz layout
of the reference (`Data/Create_Data_Set.py`: Train/Test_Inputs, Train/Test_Targets (float64), Input_Bounds).
"""
import os
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SETTINGS = """# PDE-READ settings (reference format: fixed order, "<name> [<type>] : <value>")
Load Sol Network State [bool] :                  %(load)s
Load PDE Network State [bool] :                  %(load)s
Load Optimizer State [bool] :                    False
  Load File Name [str] :                         Heat_Test

Save State [Bool] :                              %(save)s
  Save File Name [str] :                         Heat_Test

Discovery, or Extraction mode [str] :            %(mode)s
DataSet Name [str] :                             Heat_Synth
Number of Training Collocation Points [int] :    3000
Number of Testing Collocation Points [int] :     500
Extracted PDE maximum term degree [int] :        2
Number of Extraction Points [int] :              5000
Train on CPU or GPU [GPU, CPU] :                 GPU
PDE - Time Derivative Order [int] :              1
PDE - Spatial Derivative Order [int] :           2
Sol Network - Number of Hidden Layers [int] :    3
Sol Network - Neurons per Hidden Layer [int] :   30
Sol Network - Activation Function [str] :        %(act)s
PDE Network - Normalize Inputs [bool] :          False
PDE Network - Number of Hidden Layers [int] :    2
PDE Network - Neurons per Hidden Layer [int] :   20
PDE Network - Activation Function [str] :        %(act)s
Optimizer [Adam, LBFGS, SGD] :                   %(opt)s
Number of Epochs [int] :                         %(epochs)d
Learning Rate [float] :                          %(lr)s
Epochs between testing [int] :                   %(every)d
"""


def _make_root(tmp_path):
    root = tmp_path / "checkout"
    (root / "Data" / "DataSets").mkdir(parents=True)
    (root / "Saves").mkdir()
    rng = np.random.default_rng(0)

    def sample(n):
        X = np.stack([rng.uniform(0.0, 2.0, n), rng.uniform(-3.0, 3.0, n)], axis=1).astype(np.float32)
        u = np.exp(-0.5 * X[:, 0].astype(np.float64)) * np.sin(X[:, 1].astype(np.float64))
        return X, u

    Xtr, utr = sample(2000)
    Xte, ute = sample(400)
    np.savez(root / "Data" / "DataSets" / "Heat_Synth.npz", Train_Inputs=Xtr, Train_Targets=utr, Test_Inputs=Xte,
             Test_Targets=ute, Input_Bounds=np.array([[0.0, 2.0], [-3.0, 3.0]], dtype=np.float32))
    return root


def _run(root, capsys, **kw):
    from pde_read_b200.main import main
    (root / "Settings.txt").write_text(SETTINGS % kw)
    main(str(root))
    return capsys.readouterr().out


def _totals(out, which):
    return [float(v) for v in re.findall(r"%s:\s+Collocation = [0-9.eE+-]+\s+Data = [0-9.eE+-]+\s+Total = ([0-9.eE+-]+)" % which, out)]


@pytest.mark.parametrize("act", ["Tanh", "Rational"])
def test_discovery_checkpoint_resume_extraction(tmp_path, capsys, act):
    torch.manual_seed(0)
    root = _make_root(tmp_path)
    out = _run(root, capsys, mode="Discovery", load="False", save="True", act=act, opt="Adam", epochs=300, lr="0.002", every=50)
    test_tot = _totals(out, "Test")
    assert len(test_tot) >= 6
    assert test_tot[-1] < 0.2 * test_tot[0], (test_tot[0], test_tot[-1])           # training works through the CUDA path
    ck = torch.load(root / "Saves" / "Heat_Test", map_location="cpu")
    assert set(ck) == {"Sol_Network_State", "PDE_Network_State", "Optimizer_State"}     # reference main.py:280-282
    assert "Layers.0.weight" in ck["Sol_Network_State"]
    if act == "Rational":
        assert "Activation_Functions.0.a" in ck["Sol_Network_State"]

    # resume with LBFGS (closure called several times per step, with and without grad)
    out2 = _run(root, capsys, mode="Discovery", load="True", save="True", act=act, opt="LBFGS", epochs=4, lr="0.5", every=1)
    tot2 = _totals(out2, "Train")
    assert tot2 and np.isfinite(tot2).all()
    assert tot2[0] < 2.0 * test_tot[-1] + 1e-3          # the resumed run starts where the first one stopped

    # extraction from the checkpoint: prints five candidate PDEs in the reference's format
    out3 = _run(root, capsys, mode="Extraction", load="True", save="False", act=act, opt="Adam", epochs=0, lr="0.001", every=1)
    assert out3.count("most likely PDE gives a residual of") == 5
    assert "D_t U =" in out3


# ---- the reference's own artifacts -----------------------------------------------------------------
def _stock_root(tmp_path):
    """A checkout laid out like the reference's: its stock Settings.txt (mode / load / device lines changed only,
    see tests/golden/make_golden.golden_stock_artifacts), its shipped checkpoint in Saves/ (re-saved with torch.save
    from the golden weights, keys of reference main.py:280-282) and its data set's bounds in Data/DataSets/."""
    from tests.helpers import load_golden, split_state
    z, ck = load_golden("stock_artifacts.npz"), load_golden("burgers_ckpt.npz")
    root = tmp_path / "stock"
    (root / "Data" / "DataSets").mkdir(parents=True)
    (root / "Saves").mkdir()
    (root / "Settings.txt").write_text(bytes(z["settings_text"]).decode())
    np.savez(root / "Data" / "DataSets" / "Burgers_Sine_N50_P5000.npz",
             **{k: z[k] for k in ("Train_Inputs", "Train_Targets", "Test_Inputs", "Test_Targets", "Input_Bounds")})
    torch.save({"Sol_Network_State": split_state(ck, "sol."), "PDE_Network_State": split_state(ck, "pde."),
                "Optimizer_State": {}}, root / "Saves" / "Burgers_Sine_N50_P5000_Rat_Adam")
    return root


def _first_pde(out):
    """coefficients {term: value} of the PDE printed as '#1 most likely'"""
    lines = out.splitlines()
    idx = [i for i, ln in enumerate(lines) if ln.startswith("The #1 most likely PDE")][0]
    terms = re.findall(r"\+ (-?[0-9.]+)((?:\([^)]*\))+)", lines[idx + 1])
    return {t: float(v) for v, t in terms}, lines[idx]


def test_stock_settings_and_shipped_checkpoint_extract_the_golden_pde(tmp_path, capsys, monkeypatch):
    """Shipped checkpoint + the reference's Settings.txt (K = 3, 20 000 points) through `python -m pde_read_b200.main`:
    (a) with the reference's own extraction points (random.seed(0), Points.py:53-60) the printed #1 PDE is the golden
    D_t U = 0.076364 (D_x^2 U) - 0.747831 (U)(D_x U), residual 3.8586, within 1e-4;
    (b) with the device sampler (distribution-equal points) the same support and coefficients within 2 %."""
    import random
    from oracle import pde_read_oracle as orc
    from pde_read_b200 import main as main_mod
    root = _stock_root(tmp_path)

    def reference_points(Bounds, Num_Points, Data_Type, Device):
        random.seed(0)
        return orc.generate_points(np.asarray(Bounds, dtype=np.float64), Num_Points).to(Device)

    monkeypatch.setattr(main_mod, "Generate_Points", reference_points)
    main_mod.main(str(root))
    out = capsys.readouterr().out
    assert out.count("most likely PDE gives a residual of") == 5
    terms, head = _first_pde(out)
    assert set(terms) == {"(D_x^2 U)", "(U)(D_x U)"}, terms
    assert abs(terms["(D_x^2 U)"] - 0.076364) <= 1e-4 * 0.076364 + 1.5e-6
    assert abs(terms["(U)(D_x U)"] + 0.747831) <= 1e-4 * 0.747831
    assert "residual of 3.8586" in head and "121.98%" in head
    monkeypatch.undo()

    main_mod.main(str(root))
    out = capsys.readouterr().out
    terms, head = _first_pde(out)
    assert set(terms) == {"(D_x^2 U)", "(U)(D_x U)"}, terms
    assert abs(terms["(D_x^2 U)"] - 0.0764) < 0.02 * 0.0764 + 2e-3 and abs(terms["(U)(D_x U)"] + 0.748) < 0.02 * 0.748


def test_plot_style_batched_residual_caller():
    """reference Plot/Plot_Solution.py:61-106 (Evaluate_Residual): PDE_Residual called on 1024-point batches of a grid
    and a ragged last batch, results detached and concatenated -- shipped checkpoint, against the oracle."""
    from oracle import pde_read_oracle as orc
    from pde_read_b200.PDE_Residual import PDE_Residual
    from tests.helpers import TOL_RESIDUAL, assert_close_to_reference, load_golden, split_state
    from tests.test_gpu_parity import DEV, make_net, to64
    ck = load_golden("burgers_ckpt.npz")
    su, sn = split_state(ck, "sol."), split_state(ck, "pde.")
    U, N = make_net(su, "Rational", 2), make_net(sn, "Rational", 3)
    t, x = torch.meshgrid(torch.linspace(0.0, 10.0, 61), torch.linspace(-8.0, 7.9375, 57), indexing="ij")
    Coords = torch.stack([t.reshape(-1), x.reshape(-1)], dim=1)          # 3477 points: 3 full batches + 405
    Batch_Size, Num = 1024, Coords.shape[0]
    Residual = np.empty(Num, dtype=np.float32)
    Cd = Coords.to(DEV)
    i = 0
    for i in range(0, Num - Batch_Size, Batch_Size):
        Residual[i:i + Batch_Size] = PDE_Residual(U, N, 1, 2, Cd[i:i + Batch_Size, :]).view(-1).detach().cpu().numpy()
    Residual[i + Batch_Size:] = PDE_Residual(U, N, 1, 2, Cd[i + Batch_Size:, :]).view(-1).detach().cpu().numpy()
    want = orc.pde_residual(su, "Rational", sn, "Rational", 1, 2, Coords).detach().numpy()
    want64 = orc.pde_residual(to64(su), "Rational", to64(sn), "Rational", 1, 2, Coords.double()).detach().numpy()
    assert_close_to_reference(Residual, want, want64, TOL_RESIDUAL, "batched residual")


def test_heat_discovery_then_extraction_recovers_the_equation(tmp_path, capsys):
    """Known answer: the data are u = exp(-0.5 t) sin x, a solution of u_t = 0.5 u_xx.  On this solution u_xx = -u,
    so the sparse PDEs consistent with the data are D_t U = a (D_x^2 U) + b (U) with a - b = 0.5 (a = 0.5, b = 0 and
    a = 0, b = -0.5 are the two one-term forms).  After Discovery the extracted #1 PDE must be of that family:
    only those two terms carry weight and a - b is 0.5 to 15 %."""
    torch.manual_seed(0)
    root = _make_root(tmp_path)
    _run(root, capsys, mode="Discovery", load="False", save="True", act="Tanh", opt="Adam", epochs=1500, lr="0.002", every=500)
    out = _run(root, capsys, mode="Extraction", load="True", save="False", act="Tanh", opt="Adam", epochs=0, lr="0.001", every=1)
    terms, head = _first_pde(out)
    a, b = terms.get("(D_x^2 U)", 0.0), terms.get("(U)", 0.0)
    assert abs((a - b) - 0.5) < 0.075, terms
    others = {k: v for k, v in terms.items() if k not in ("(D_x^2 U)", "(U)")}
    assert all(abs(v) < 0.05 for v in others.values()), terms
