"""Data_Loader with the reference's interface (reference Code/Data_Setup.py:12-72): reads
``<Data_Dir>/<DataSet_Name>.npz`` (keys Train_Inputs, Train_Targets, Test_Inputs, Test_Targets,
Input_Bounds) into tensors on Device.  I/O only, once per run."""
from __future__ import annotations

import os

import numpy
import torch

DATA_DIR = "../Data/DataSets/"      # the reference's cwd-relative location (Data_Setup.py:48)


class Data_Container:
    pass


def Data_Loader(DataSet_Name: str, Device: torch.device, Mode: str, Data_Dir: str | None = None) -> Data_Container:
    path = os.path.join(Data_Dir if Data_Dir is not None else DATA_DIR, DataSet_Name + ".npz")
    DataSet = numpy.load(path)
    Container = Data_Container()
    for key in ("Train_Inputs", "Train_Targets", "Test_Inputs", "Test_Targets"):
        setattr(Container, key, torch.from_numpy(DataSet[key]).to(device=Device))
    Container.Input_Bounds = DataSet["Input_Bounds"]
    return Container
