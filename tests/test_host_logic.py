"""CPU tests of the host-side mirror of the reference API: settings parsing, multi-index
bookkeeping, ranking, printing, network containers, C-ABI symbol table, failure without CUDA."""
import io
import os
import re

import numpy as np
import pytest
import torch

from oracle import pde_read_oracle as orc
from tests.helpers import load_golden, split_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETTINGS_TEXT = """# comment line
Load Sol Network State [bool] :                  True
Load PDE Network State [bool] :                  true
Load Optimizer State [bool] :                    False
  Load File Name [str] :                         Burgers_Sine_N50_P5000_Rat_Adam

Save State [Bool] :                              True
  Save File Name [str] :                         Out_File   # trailing comment

# The "Discovery, or Extraction mode" setting ...
Discovery, or Extraction mode [str] :            %(mode)s
DataSet Name [str] :                             Burgers_Sine_N50_P5000
Number of Training Collocation Points [int] :    5000
Number of Testing Collocation Points [int] :     1000
Extracted PDE maximum term degree [int] :        3
Number of Extraction Points [int] :              20000
Train on CPU or GPU [GPU, CPU] :                 CPU
PDE - Time Derivative Order [int] :              1
PDE - Spatial Derivative Order [int] :           2
Sol Network - Number of Hidden Layers [int] :    5
Sol Network - Neurons per Hidden Layer [int] :   50
Sol Network - Activation Function [str] :        Rat
PDE Network - Normalize Inputs [bool] :          False
PDE Network - Number of Hidden Layers [int] :    2
PDE Network - Neurons per Hidden Layer [int] :   100
PDE Network - Activation Function [str] :        tanh
Optimizer [Adam, LBFGS, SGD] :                   lbfgs
Number of Epochs [int] :                         1000
Learning Rate [float] :                          .001
Epochs between testing [int] :                   10
"""


def _write(tmp_path, mode):
    p = tmp_path / "Settings.txt"
    p.write_text(SETTINGS_TEXT % {"mode": mode})
    return str(p)


def test_settings_reader_discovery(tmp_path):
    from pde_read_b200.Settings_Reader import Settings_Reader
    S = Settings_Reader(_write(tmp_path, "Discovery"))
    assert S.Load_Sol_Network_State is True and S.Load_PDE_Network_State is True and S.Load_Optimizer_State is False
    assert S.Load_File_Name == "Burgers_Sine_N50_P5000_Rat_Adam"
    assert S.Save_To_File is True and S.Save_File_Name == "Out_File"
    assert S.Mode == "Discovery" and S.DataSet_Name == "Burgers_Sine_N50_P5000"
    assert (S.Num_Train_Colloc_Points, S.Num_Test_Colloc_Points) == (5000, 1000)
    assert not hasattr(S, "Num_Extraction_Points")          # only read in Extraction mode
    assert S.Device == torch.device("cpu")
    assert (S.PDE_Time_Derivative_Order, S.PDE_Spatial_Derivative_Order) == (1, 2)
    assert (S.Sol_Num_Hidden_Layers, S.Sol_Neurons_Per_Layer, S.Sol_Activation_Function) == (5, 50, "Rational")
    assert (S.PDE_Num_Hidden_Layers, S.PDE_Neurons_Per_Layer, S.PDE_Activation_Function) == (2, 100, "Tanh")
    assert S.PDE_Normalize_Inputs is False and S.Optimizer == "LBFGS"
    assert (S.Epochs, S.Learning_Rate, S.Epochs_Between_Prints) == (1000, 0.001, 10)


def test_settings_reader_extraction_and_errors(tmp_path):
    from pde_read_b200.Settings_Reader import Read_Error, Settings_Reader
    S = Settings_Reader(_write(tmp_path, "extraction"))
    assert S.Mode == "Extraction" and (S.Extracted_Term_Degree, S.Num_Extraction_Points) == (3, 20000)
    assert not hasattr(S, "Num_Train_Colloc_Points")
    bad = tmp_path / "bad.txt"
    bad.write_text((SETTINGS_TEXT % {"mode": "Discovery"}).replace("lbfgs", "Newton"))
    with pytest.raises(Read_Error):
        Settings_Reader(str(bad))
    # order matters: the reader scans forward once, a setting placed too early is never found
    lines = (SETTINGS_TEXT % {"mode": "Discovery"}).splitlines(keepends=True)
    moved = [l for l in lines if not l.startswith("Save State")] + ["Save State [Bool] : True\n"]
    (tmp_path / "moved.txt").write_text("".join(moved))
    with pytest.raises(Read_Error):
        Settings_Reader(str(tmp_path / "moved.txt"))


@pytest.mark.skipif(not os.path.exists("/root/reference/Settings.txt"), reason="reference tree not mounted")
def test_settings_reader_reads_the_reference_file():
    from pde_read_b200.Settings_Reader import Settings_Reader
    S = Settings_Reader("/root/reference/Settings.txt")
    assert S.Mode == "Discovery" and S.Sol_Activation_Function == "Rational" and S.PDE_Neurons_Per_Layer == 100
    assert S.Optimizer == "Adam" and S.Epochs == 1000 and S.Save_File_Name == "Burgers_Sine_N50_P5000_Rat_Adam"


def test_multi_index_helpers_match_oracle():
    from pde_read_b200.Extraction import Recursive_Counter, Recursive_Multi_Indices
    for nv in range(1, 7):
        for k in range(1, 6):
            cnt = Recursive_Counter(num_sub_index_values=nv, num_sub_indices=k)
            assert cnt == orc.count_multi_indices(nv, k)
            mi = np.empty((cnt, k), dtype=np.int64)
            assert Recursive_Multi_Indices(multi_indices=mi, num_sub_index_values=nv, num_sub_indices=k) == cnt
            assert np.array_equal(mi, orc.multi_indices(nv, k))
    # identities of the reference's test_Recursive_Counter (Testing.py:320-364)
    for n in range(1, 8):
        assert Recursive_Counter(n, 1) == n and Recursive_Counter(n, 2) == n * (n + 1) // 2
        assert Recursive_Counter(n, 3) == sum(k * (k + 1) // 2 for k in range(1, n + 1))


def test_rank_and_print_match_reference_golden(capsys):
    from pde_read_b200.Extraction import Print_Extracted_PDE, Rank_Candidate_Solutions, _index_tables
    z = load_golden("burgers_extract_k3.npz")
    Xr, Rr, chg = Rank_Candidate_Solutions(z["X"], z["Residual"])
    assert np.array_equal(Xr, z["X_Ranked"])
    assert np.array_equal(Rr, z["Residual_Ranked"])
    np.testing.assert_allclose(chg, z["Residual_Change"], rtol=1e-6)
    counts, mil = _index_tables(2, 3)
    assert list(counts) == list(z["counts"])
    Print_Extracted_PDE(Xr[:, -1], 1, counts, mil)
    out = capsys.readouterr().out.strip()
    assert out == "D_t U =  + 0.076364(D_x^2 U) + -0.747831(U)(D_x U)"
    assert out == orc.format_pde(Xr[:, -1], 1, 2, 3)
    Print_Extracted_PDE(np.array([1.5] + [0.0] * 19, dtype=np.float32), 2, counts, mil)
    assert capsys.readouterr().out.strip() == "D_t^2 U = 1.500000"


def test_network_container_matches_reference_layout():
    from pde_read_b200.Network import Neural_Network, Rational
    ck = load_golden("burgers_ckpt.npz")
    U = Neural_Network(5, 50, 2, 1, Activation_Function="Rational")
    N = Neural_Network(2, 100, 3, 1, Activation_Function="Rational")
    sol, pde = split_state(ck, "sol."), split_state(ck, "pde.")
    assert set(U.state_dict().keys()) == set(sol.keys()) and set(N.state_dict().keys()) == set(pde.keys())
    U.load_state_dict(sol)
    N.load_state_dict(pde)
    # parameters() order = Linear weights/biases, then the activations' a, b (optimizer state compatibility)
    names = [k for k, _ in U.named_parameters()]
    assert names[:12] == [("Layers.%d." % (i // 2)) + ("weight" if i % 2 == 0 else "bias") for i in range(12)]
    assert names[12:] == [("Activation_Functions.%d." % (i // 2)) + ("a" if i % 2 == 0 else "b") for i in range(10)]
    assert sum(p.numel() for p in U.parameters()) == 10436 and sum(p.numel() for p in N.parameters()) == 10615
    # initial values (reference Network.py:20-30, 133-154)
    r = Rational()
    assert torch.allclose(r.a, torch.tensor(orc.RATIONAL_A_INIT)) and torch.allclose(r.b, torch.tensor(orc.RATIONAL_B_INIT))
    torch.manual_seed(0)
    T = Neural_Network(3, 400, 2, 1, Activation_Function="Tanh")
    w = T.Layers[1].weight
    assert abs(w.std().item() - (5.0 / 3.0) * (2.0 / 800) ** 0.5) < 0.003 and T.Layers[1].bias.abs().max() == 0
    S = Neural_Network(3, 400, 2, 1, Activation_Function="Sin")
    assert S.Layers[1].weight.abs().max().item() <= 3.0 / 20.0 + 1e-6
    # the standalone Rational module evaluates the reference formula (Testing.py:82-131)
    x = torch.rand(50, 7)
    a, b = orc.RATIONAL_A_INIT, orc.RATIONAL_B_INIT
    want = (a[0] + a[1] * x + a[2] * x ** 2 + a[3] * x ** 3) / (b[0] + b[1] * x + b[2] * x ** 2)
    assert (r(x) - want).abs().max().item() < 2.5e-5


def test_product_path_has_no_cpu_fallback():
    from pde_read_b200 import functional as F
    from pde_read_b200.Loss_Functions import Collocation_Loss
    from pde_read_b200.Network import Neural_Network
    from pde_read_b200.Points import Generate_Points
    U = Neural_Network(2, 8, 2, 1, Activation_Function="Tanh")
    N = Neural_Network(2, 8, 3, 1, Activation_Function="Tanh")
    with pytest.raises(F.PdereadError):
        U(torch.rand(4, 2))
    with pytest.raises(F.PdereadError):
        Collocation_Loss(U, N, 1, 2, torch.rand(4, 2))
    with pytest.raises(F.PdereadError):
        Generate_Points(np.array([[0.0, 1.0], [0.0, 1.0]]), 10, torch.float32, torch.device("cpu"))
    bn = Neural_Network(2, 8, 3, 1, Activation_Function="Tanh", Batch_Norm=True)
    with pytest.raises(NotImplementedError):
        F.net_spec(bn)                      # fused U+N entry points refuse a network with input BatchNorm
    assert F.net_spec(bn, allow_input_norm=True).input_dim == 3


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/pderead_b200.h declares (no compute)."""
    from pde_read_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, "include", "pderead_b200.h")).read()
    declared = set(re.findall(r"\b(pderead_[a-z0-9_]+)\s*\(", header))
    declared -= {"pderead_status", "pderead_activation"}
    assert len(declared) >= 25
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for " + name
    assert set(_lib.PROTOTYPES) == declared
    assert lib.pderead_library_num_columns(2, 3) == 20 and lib.pderead_library_num_columns(2, 5) == 56
    assert lib.pderead_library_num_columns(3, 5) == 126 and lib.pderead_library_num_columns(4, 5) == 252
    assert lib.pderead_status_string(-4) == b"workspace too small"
    # struct layout agrees with the header (17+17+16+16 pointers after four int32)
    import ctypes
    assert ctypes.sizeof(_lib.pderead_net) == 16 + 8 * 66 and ctypes.sizeof(_lib.pderead_net_grad) == 8 * 66


def test_timer():
    from pde_read_b200.Timing import Timer, Timer_Exception
    t = Timer()
    with pytest.raises(Timer_Exception):
        t.Stop()
    t.Start()
    assert t.Stop() >= 0.0


def test_hot_kernels_keep_their_register_budget():
    """Occupancy of the bench kernels rests on register caps (96 registers -> four CTAs of the C2 U kernels per SM,
    80 -> three CTAs of the narrow plain-MLP kernels, 128 for the long-jet kernels) with no or next to no spilling;
    a change that makes ptxas spill shows up here (cuobjdump on the built library, no GPU needed) instead of as an
    unexplained slowdown -- it happened once this round: reverse-U 0.66 -> 0.94 ms with 88 bytes of stack."""
    import re
    import shutil
    import subprocess
    from pde_read_b200 import build as B
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(B.LIB):
        pytest.skip("library not built")
    out = subprocess.run([tool, "--dump-resource-usage", B.LIB], capture_output=True, text=True).stdout
    use = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        use[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    want = {   # mangled kernel: (max registers, max stack bytes)
        "_ZN3pdr18mlp_jet_fwd_kernelILi0ELi2ELi1ELb1ELi1ELi0EEEvNS_7FwdArgsE": (96, 16),    # C2 U forward (stash)
        "_ZN3pdr18mlp_jet_bwd_kernelILi0ELi2ELi1ELb1ELb0EEEvNS_7BwdArgsE": (96, 16),        # C2 U reverse (fused)
        "_ZN3pdr18mlp_jet_fwd_kernelILi2ELi2ELi0ELb0ELi1ELi0EEEvNS_7FwdArgsE": (96, 16),    # C5 U forward (x-series)
        "_ZN3pdr18mlp_jet_fwd_kernelILi2ELi3ELi1ELb1ELi1ELi0EEEvNS_7FwdArgsE": (128, 16),   # C3 U forward
        "_ZN3pdr18mlp_jet_bwd_kernelILi2ELi3ELi1ELb1ELb0EEEvNS_7BwdArgsE": (128, 16),       # C3 U reverse
        "_ZN3pdr18mlp_jet_fwd_kernelILi2ELi4ELi1ELb1ELi1ELi1EEEvNS_7FwdArgsE": (128, 16),   # C4 U forward (staged weights)
        "_ZN3pdr18mlp_jet_bwd_kernelILi2ELi4ELi1ELb0ELb0EEEvNS_7BwdArgsE": (128, 16),       # C4 U reverse (wide block)
        "_ZN3pdr20mlp_plain_fwd_kernelILi0ELb1ELi4ELb1EEEvNS_7FwdArgsE": (80, 32),          # C2 N forward (narrow)
        "_ZN3pdr20mlp_plain_bwd_kernelILi0ELb1EEEvNS_7BwdArgsE": (80, 48),                  # C2 N reverse (narrow)
    }
    for name, (regs, stack) in want.items():
        assert name in use, "kernel not found in the library: %s" % name
        assert use[name][0] <= regs and use[name][1] <= stack, (name, use[name], (regs, stack))
