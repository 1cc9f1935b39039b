"""Settings.txt-driven driver with the reference's behaviour (reference Code/main.py:14-287).

    cd <reference checkout>/Code && python -m pde_read_b200.main        # same relative paths
    python -m pde_read_b200.main --root /path/to/checkout               # or point at the checkout

Discovery mode trains both networks (Adam / LBFGS / SGD-Nesterov) on collocation + data loss and
saves a checkpoint with the reference's keys; Extraction mode loads a checkpoint and prints the five
most likely PDEs.  Checkpoints are interchangeable with the reference (same state_dict keys).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy
import torch

from .Data_Setup import Data_Loader
from .Extraction import Extract_PDE, Print_Extracted_PDE
from .Network import Neural_Network
from .Optimizers import Fused_Adam, Fused_SGD
from .Points import Generate_Points
from .Settings_Reader import Settings_Reader
from .Test_Train import Discovery_Testing, Discovery_Training
from .Timing import Timer


def main(root: str = "..", settings_path: str | None = None) -> None:
    Settings = Settings_Reader(settings_path if settings_path is not None else os.path.join(root, "Settings.txt"))
    print("Loaded the following settings:")
    for (setting, value) in Settings.__dict__.items():
        print(("%-30s = " % setting) + str(value))
    if Settings.Device.type != "cuda":
        raise SystemExit("pde_read_b200 needs a CUDA device: set \"Train on CPU or GPU\" to GPU on a B200 box "
                         "(there is no CPU path)")

    Setup_Timer = Timer()
    Setup_Timer.Start()
    Epochs: int = Settings.Epochs
    m, n = Settings.PDE_Time_Derivative_Order, Settings.PDE_Spatial_Derivative_Order

    Sol_NN = Neural_Network(Num_Hidden_Layers=Settings.Sol_Num_Hidden_Layers,
                            Neurons_Per_Layer=Settings.Sol_Neurons_Per_Layer,
                            Input_Dim=2, Output_Dim=1, Data_Type=torch.float32, Device=Settings.Device,
                            Activation_Function=Settings.Sol_Activation_Function, Batch_Norm=False)
    PDE_NN = Neural_Network(Num_Hidden_Layers=Settings.PDE_Num_Hidden_Layers,
                            Neurons_Per_Layer=Settings.PDE_Neurons_Per_Layer,
                            Input_Dim=n + 1, Output_Dim=1, Data_Type=torch.float32, Device=Settings.Device,
                            Activation_Function=Settings.PDE_Activation_Function,
                            Batch_Norm=Settings.PDE_Normalize_Inputs)

    Optimizer = None
    if Settings.Mode == "Discovery":
        Params = list(Sol_NN.parameters()) + list(PDE_NN.parameters())
        # (reference main.py:63-71; Adam and SGD are the single-launch subclasses of the same torch classes, so the
        #  Optimizer_State of a checkpoint keeps torch's layout)
        if Settings.Optimizer == "Adam":
            Optimizer = Fused_Adam(Params, lr=Settings.Learning_Rate)
        elif Settings.Optimizer == "LBFGS":
            Optimizer = torch.optim.LBFGS(Params, lr=Settings.Learning_Rate)
        elif Settings.Optimizer == "SGD":
            Optimizer = Fused_SGD(Params, lr=Settings.Learning_Rate, momentum=0.9, nesterov=True)
        else:
            print("Optimizer is %s when it should be \"Adam\", \"LBFGS\", or \"SGD\"" % Settings.Optimizer)
            print("Aborting.")
            sys.exit()

    if Settings.Load_Sol_Network_State or Settings.Load_PDE_Network_State or Settings.Load_Optimizer_State:
        Saved_State = torch.load(os.path.join(root, "Saves", Settings.Load_File_Name), map_location=Settings.Device)
        if Settings.Load_Sol_Network_State:
            Sol_NN.load_state_dict(Saved_State["Sol_Network_State"])
        if Settings.Load_PDE_Network_State:
            PDE_NN.load_state_dict(Saved_State["PDE_Network_State"])
        if Settings.Load_Optimizer_State and Settings.Mode != "Extraction":
            Optimizer.load_state_dict(Saved_State["Optimizer_State"])
            for param_group in Optimizer.param_groups:       # the new learning rate wins (main.py:96-98)
                param_group["lr"] = Settings.Learning_Rate

    Data = Data_Loader(DataSet_Name=Settings.DataSet_Name, Device=Settings.Device, Mode=Settings.Mode,
                       Data_Dir=os.path.join(root, "Data", "DataSets"))
    print("Setup took %fs." % Setup_Timer.Stop())

    Main_Timer = Timer()
    Main_Timer.Start()
    if Settings.Mode == "Discovery":
        for t in range(Epochs):
            Train_Colloc_Points = Generate_Points(Data.Input_Bounds, Settings.Num_Train_Colloc_Points, torch.float32,
                                                  Settings.Device)
            Discovery_Training(Sol_NN, PDE_NN, m, n, Train_Colloc_Points, Data.Train_Inputs, Data.Train_Targets,
                               Optimizer, torch.float32, Settings.Device)
            if t % Settings.Epochs_Between_Prints == 0 or t == Epochs - 1:
                Test_Colloc_Points = Generate_Points(Data.Input_Bounds, Settings.Num_Test_Colloc_Points, torch.float32,
                                                     Settings.Device)
                Test_Coll, Test_Data = Discovery_Testing(Sol_NN, PDE_NN, m, n, Test_Colloc_Points, Data.Test_Inputs,
                                                         Data.Test_Targets, torch.float32, Settings.Device)
                Train_Coll, Train_Data = Discovery_Testing(Sol_NN, PDE_NN, m, n, Train_Colloc_Points,
                                                           Data.Train_Inputs, Data.Train_Targets, torch.float32,
                                                           Settings.Device)
                print("Epoch #%-4d | Test: \t Collocation = %.7f\t Data = %.7f\t Total = %.7f"
                      % (t, Test_Coll, Test_Data, Test_Coll + Test_Data))
                print("            | Train:\t Collocation = %.7f\t Data = %.7f\t Total = %.7f"
                      % (Train_Coll, Train_Data, Train_Coll + Train_Data))
            else:
                print("Epoch #%-4d | " % t)
    elif Settings.Mode == "Extraction":
        Extraction_Points = Generate_Points(Data.Input_Bounds, Settings.Num_Extraction_Points, torch.float32,
                                            Settings.Device)
        (X, Residual, X_Ranked, Residual_Ranked, Residual_Change,
         num_multi_indices, multi_indices_list, _) = Extract_PDE(Sol_NN, PDE_NN, m, n, Extraction_Points,
                                                                 Settings.Extracted_Term_Degree)
        Num_Cols = X.shape[1]
        for i in range(max(Num_Cols - 5, 0), Num_Cols):
            print("The #%u most likely PDE gives a residual of %.4lf (%.2lf%% better than the next sparsest PDE)."
                  % (Num_Cols - i, Residual_Ranked[i], Residual_Change[i] * 100))
            Print_Extracted_PDE(X_Ranked[:, i], m, num_multi_indices, multi_indices_list)
    else:
        print("Mode is %s. It should be one of \"Discovery\", \"Extraction\"." % Settings.Mode)
        print("Something went wrong. Aborting. Thrown by main.")
        sys.exit()

    torch.cuda.synchronize()
    Main_Time = Main_Timer.Stop()
    if Settings.Mode == "Discovery":
        Minutes = int(Main_Time) // 60
        print("Running %d epochs took %um,%.2fs (%fs)." % (Epochs, Minutes, Main_Time - 60 * Minutes, Main_Time))
        if Epochs > 0:
            print("That's an average of %fs per epoch!" % (Main_Time / Epochs))
    else:
        print("Extraction took %fs." % Main_Time)

    if Settings.Mode == "Discovery" and Settings.Save_To_File:
        os.makedirs(os.path.join(root, "Saves"), exist_ok=True)
        torch.save({"Sol_Network_State": Sol_NN.state_dict(),
                    "PDE_Network_State": PDE_NN.state_dict(),
                    "Optimizer_State": Optimizer.state_dict()},
                   os.path.join(root, "Saves", Settings.Save_File_Name))


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--root", default="..", help="checkout root holding Settings.txt, Saves/ and Data/DataSets/")
    ap.add_argument("--settings", default=None, help="explicit path of the settings file")
    a = ap.parse_args()
    main(a.root, a.settings)
