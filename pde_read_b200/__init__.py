"""pde_read_b200: the data-parallel hot path of PDE-READ on B200 (sm_100a).

Module and function names mirror the reference's ``Code/`` directory (Network, PDE_Residual,
Loss_Functions, Test_Train, Extraction, Points, Settings_Reader, Data_Setup, Timing, main) so a
user of the reference can switch by changing the import path.  The arithmetic of the hot path runs
in libpderead_b200.so (hand-written CUDA, C ABI in include/pderead_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
