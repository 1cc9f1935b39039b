"""Settings.txt reader with the reference's file format and result object (reference
Code/Settings_Reader.py:222-379).

Format (reference Settings.txt:1-11): blank lines and lines starting with '#' are ignored; a
setting line is ``<Setting Name> [<type>] : <value>``; the settings must appear in the fixed order
below because the file is scanned forward exactly once; names match case-insensitively; enumerated
values are recognised by their first letter.
"""
from __future__ import annotations

import os

import torch


class Read_Error(Exception):
    pass


class Settings_Container:
    pass


def Index_After_Phrase(Line_In: str, Phrase_In: str, Case_Sensitive: bool = False) -> int:
    """Index of the first character after the first occurrence of Phrase_In in Line_In, -1 if absent."""
    line, phrase = (Line_In, Phrase_In) if Case_Sensitive else (Line_In.lower(), Phrase_In.lower())
    at = line.find(phrase)
    return -1 if at < 0 else at + len(phrase)


def Read_Line_After(File, Phrase: str, Comment_Char: str = "#", Case_Sensitive: bool = False) -> str:
    """Scan forward from the current file position for a non-comment line containing Phrase and return
    what follows it on that line (an inline comment is cut off first)."""
    for Line in iter(File.readline, ""):
        if Line.startswith(Comment_Char):
            continue
        cut = Line.find(Comment_Char, 1)
        if cut > 0:
            Line = Line[:cut]
        idx = Index_After_Phrase(Line, Phrase, Case_Sensitive)
        if idx >= 0:
            return Line[idx:]
    raise Read_Error("Could not find \"" + Phrase + "\" in File.")


def Read_Setting(File, Setting_Name: str) -> str:
    Buffer = Read_Line_After(File, Setting_Name).strip()
    if len(Buffer) == 0:
        raise Read_Error("Missing Setting Value: You need to specify the \"%s\" setting" % Setting_Name)
    return Buffer


def Read_Bool_Setting(File, Setting_Name: str) -> bool:
    Buffer = Read_Setting(File, Setting_Name)
    first = Buffer[0].lower()
    if first == "t":
        return True
    if first == "f":
        return False
    raise Read_Error("Invalid Setting Value: \"%s\" should be \"True\" or \"False\". Got %s" % (Setting_Name, Buffer))


def _enum(File, Setting_Name: str, table: dict) -> str:
    Buffer = Read_Setting(File, Setting_Name)
    key = Buffer[0].lower()
    if key not in table:
        raise Read_Error("\"%s\" should be one of %s. Got %s" % (Setting_Name, sorted(set(table.values())), Buffer))
    return table[key]


_ACTIVATIONS = {"r": "Rational", "s": "Sin", "t": "Tanh"}


def Settings_Reader(Settings_Path: str | None = None) -> Settings_Container:
    """Read ../Settings.txt (the reference's cwd-relative default) or Settings_Path."""
    path = Settings_Path if Settings_Path is not None else "../Settings.txt"
    S = Settings_Container()
    with open(path, "r") as File:
        S.Load_Sol_Network_State = Read_Bool_Setting(File, "Load Sol Network State [bool] :")
        S.Load_PDE_Network_State = Read_Bool_Setting(File, "Load PDE Network State [bool] :")
        S.Load_Optimizer_State = Read_Bool_Setting(File, "Load Optimizer State [bool] :")
        if S.Load_Sol_Network_State or S.Load_PDE_Network_State or S.Load_Optimizer_State:
            S.Load_File_Name = Read_Setting(File, "Load File Name [str] :")
        S.Save_To_File = Read_Bool_Setting(File, "Save State [bool] :")
        if S.Save_To_File:
            S.Save_File_Name = Read_Setting(File, "Save File Name [str] :")
        S.Mode = _enum(File, "Discovery, or Extraction mode [str] :", {"d": "Discovery", "e": "Extraction"})
        S.DataSet_Name = Read_Setting(File, "DataSet Name [str] :")
        # mode-specific settings are only read in their mode, as in the reference (:287-294)
        if S.Mode == "Discovery":
            S.Num_Train_Colloc_Points = int(Read_Setting(File, "Number of Training Collocation Points [int] :"))
            S.Num_Test_Colloc_Points = int(Read_Setting(File, "Number of Testing Collocation Points [int] :"))
        else:
            S.Extracted_Term_Degree = int(Read_Setting(File, "Extracted PDE maximum term degree [int] :"))
            S.Num_Extraction_Points = int(Read_Setting(File, "Number of Extraction Points [int] :"))
        where = _enum(File, "Train on CPU or GPU [GPU, CPU] :", {"g": "GPU", "c": "CPU"})
        if where == "GPU" and torch.cuda.is_available():
            S.Device = torch.device("cuda")
        else:
            if where == "GPU":
                print("You requested a GPU, but cuda is not available on this machine. Switching to CPU")
            S.Device = torch.device("cpu")
        S.PDE_Time_Derivative_Order = int(Read_Setting(File, "PDE - Time Derivative Order [int] :"))
        S.PDE_Spatial_Derivative_Order = int(Read_Setting(File, "PDE - Spatial Derivative Order [int] :"))
        S.Sol_Num_Hidden_Layers = int(Read_Setting(File, "Sol Network - Number of Hidden Layers [int] :"))
        S.Sol_Neurons_Per_Layer = int(Read_Setting(File, "Sol Network - Neurons per Hidden Layer [int] :"))
        # (the reference's reader only accepts an upper-case 'S' for Sin, :331/:349; both cases work here)
        S.Sol_Activation_Function = _enum(File, "Sol Network - Activation Function [str] :", _ACTIVATIONS)
        S.PDE_Normalize_Inputs = Read_Bool_Setting(File, "PDE Network - Normalize Inputs [bool] :")
        S.PDE_Num_Hidden_Layers = int(Read_Setting(File, "PDE Network - Number of Hidden Layers [int] :"))
        S.PDE_Neurons_Per_Layer = int(Read_Setting(File, "PDE Network - Neurons per Hidden Layer [int] :"))
        S.PDE_Activation_Function = _enum(File, "PDE Network - Activation Function [str] :", _ACTIVATIONS)
        S.Optimizer = _enum(File, "Optimizer [Adam, LBFGS, SGD] :", {"a": "Adam", "l": "LBFGS", "s": "SGD"})
        S.Epochs = int(Read_Setting(File, "Number of Epochs [int] :"))
        S.Learning_Rate = float(Read_Setting(File, "Learning Rate [float] :"))
        S.Epochs_Between_Prints = int(Read_Setting(File, "Epochs between testing [int] :"))
    return S
