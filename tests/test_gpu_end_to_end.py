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
