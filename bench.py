#!/usr/bin/env python
"""Benchmark of the PDE-READ hot path on B200: collocation points/s for one Collocation_Loss
forward + backward (BASELINE.json metric), and extraction points/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic
collocation points: Collocation_Loss(...) + .backward() through the public Python API, i.e. through
the C ABI of libpderead_b200.so (extraction: Extract_PDE, coords -> ranked candidate PDEs).

Without --config the line is the C2 workload (heat, 100 k points per GPU: the configuration BASELINE.json quotes
on one GPU) and carries an "also" block with the other BASELINE configurations measured in the same invocation
with the same event timing: C3 (KdV, 1 M points), C4 (Kuramoto-Sivashinsky, 4 M points) and C5 (extraction,
10 M points); under --gpus N > 1 these are the strong-scaling values (points sharded over the ranks).
See DESIGN.md section "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------
# workloads.  Domains come from the reference's Matlab/Data/*.mat files.
# ------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: act, (LU, WU), (LN, WN), m, n, points per GPU, box, reference chunk, reference sample per step
    "C1": dict(desc="Burgers shape: U 5x50 Rational, N 2x100 Rational, m=1 n=2, 5000 collocation pts",
               act="Rational", LU=5, WU=50, LN=2, WN=100, m=1, n=2, M=5_000, box=[[0.0, 10.0], [-8.0, 7.9375]],
               ref_chunk=5000, ref_sample=5000, scaling="weak", steps=50),
    "C2": dict(desc="Heat equation: U 5x50 Tanh, N 5x50 Tanh, m=1 n=2, 100k collocation pts per GPU",
               act="Tanh", LU=5, WU=50, LN=5, WN=50, m=1, n=2, M=100_000, box=[[0.0, 10.0], [0.0, 9.95]],
               ref_chunk=5000, ref_sample=20000, scaling="weak", steps=50),
    "C3": dict(desc="KdV: U 5x50 Rational, N 5x50 Rational, m=1 n=3, 1M collocation pts sharded over the GPUs",
               act="Rational", LU=5, WU=50, LN=5, WN=50, m=1, n=3, M=1_000_000, box=[[0.0, 40.0], [-20.0, 19.84]],
               ref_chunk=5000, ref_sample=5000, scaling="strong", steps=10),
    "C4": dict(desc="Kuramoto-Sivashinsky: U 5x100 Rational, N 5x100 Rational, m=1 n=4, 4M collocation pts sharded over the GPUs",
               act="Rational", LU=5, WU=100, LN=5, WN=100, m=1, n=4, M=4_000_000, box=[[0.0, 5.0], [-5.0, 4.96]],
               ref_chunk=2000, ref_sample=2000, scaling="strong", steps=4),
    "C4T": dict(desc="Kuramoto-Sivashinsky shape with Tanh nets (width 100, n=4), 4M pts sharded over the GPUs",
                act="Tanh", LU=5, WU=100, LN=5, WN=100, m=1, n=4, M=4_000_000, box=[[0.0, 5.0], [-5.0, 4.96]],
                ref_chunk=2000, ref_sample=4000, scaling="strong", steps=4),
    "C5": dict(desc="Extraction only: U 5x50 Rational, N 2x100 Rational, n=2, degree K=5 library (C=56) over 10M points sharded over the GPUs, fused Gram + RFE",
               act="Rational", LU=5, WU=50, LN=2, WN=100, m=1, n=2, K=5, M=10_000_000, box=[[0.0, 10.0], [-8.0, 7.9375]],
               ref_chunk=1000, ref_sample=5000, scaling="strong", extraction=True, steps=10),
    "C5S": dict(desc="Extraction stress shape: U 5x100 Rational, N 5x100 Rational, n=4, degree K=5 library (C=252) over 10M points sharded over the GPUs",
                act="Rational", LU=5, WU=100, LN=5, WN=100, m=1, n=4, K=5, M=10_000_000, box=[[0.0, 5.0], [-5.0, 4.96]],
                ref_chunk=1000, ref_sample=1000, ref_steps=2, scaling="strong", extraction=True, steps=3),
}
ALSO = ("C3", "C4", "C5", "C5S")
TAGS = {0: "other", 1: "pack_weights", 2: "jet_fwd<U>", 3: "mlp_fwd<N>", 4: "mlp_bwd<N>", 5: "jet_bwd<U>",
        6: "reduce_partials", 7: "library_gram", 8: "rfe", 9: "optimizer", 10: "input_norm"}


def flops_per_point(c):
    """Algorithmic FLOPs (dense contractions only), SURVEY.md 8d."""
    J = 1 + c["n"] + c["m"]
    FU = 2 * (2 * c["WU"] + (c["LU"] - 1) * c["WU"] ** 2 * J + c["WU"] * J)
    FN = 2 * ((c["n"] + 1) * c["WN"] + (c["LN"] - 1) * c["WN"] ** 2 + c["WN"])
    return FU, FN


def library_columns(n, K):
    from math import comb
    return comb(n + 1 + K, K)


def build_nets(c, device, seed=1234):
    """Random-init weights of the named architecture through the reference-compatible constructor."""
    from pde_read_b200.Network import Neural_Network
    torch.manual_seed(seed)
    U = Neural_Network(c["LU"], c["WU"], 2, 1, Data_Type=torch.float32, Device=torch.device("cpu"),
                       Activation_Function=c["act"])
    N = Neural_Network(c["LN"], c["WN"], c["n"] + 1, 1, Data_Type=torch.float32, Device=torch.device("cpu"),
                       Activation_Function=c["act"])
    # (the initial Rational denominator 1 + 2.383 x^2 has no real root, so random-init nets have no poles)
    sd_u = {k: v.clone() for k, v in U.state_dict().items()}
    sd_n = {k: v.clone() for k, v in N.state_dict().items()}
    if device.type == "cuda":
        U, N = U.to(device), N.to(device)
    return U, N, sd_u, sd_n


def make_points_cpu(c, M, seed):
    g = torch.Generator().manual_seed(seed)
    box = torch.tensor(c["box"], dtype=torch.float32)
    return torch.rand((M, 2), generator=g) * (box[:, 1] - box[:, 0]) + box[:, 0]


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
_REASON_BITS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
_BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in _REASON_BITS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# The reference arm: the UNMODIFIED reference (byte-compiled into oracle/_ref by oracle/build_ref.py) when it is
# there, else the oracle port of the same algorithm.  Either runs on the host cores (or, as a context line, on the
# GPU through the reference's own Device argument).
# ------------------------------------------------------------------------------------------------
_REF = None


def reference_modules():
    global _REF
    if _REF is None:
        try:
            from oracle import build_ref
            mods = build_ref.import_reference()
        except Exception:
            mods = None
        _REF = mods if mods is not None else False
    return _REF or None


def reference_nets(c, sd_u, sd_n, device):
    Network = reference_modules()[0]
    U = Network.Neural_Network(Num_Hidden_Layers=c["LU"], Neurons_Per_Layer=c["WU"], Input_Dim=2, Output_Dim=1,
                               Data_Type=torch.float32, Device=device, Activation_Function=c["act"])
    N = Network.Neural_Network(Num_Hidden_Layers=c["LN"], Neurons_Per_Layer=c["WN"], Input_Dim=c["n"] + 1, Output_Dim=1,
                               Data_Type=torch.float32, Device=device, Activation_Function=c["act"])
    U.load_state_dict(sd_u)
    N.load_state_dict(sd_n)
    return U.to(device), N.to(device)


def reference_step(c, nets, P, device):
    """One step of the reference's own code (Loss_Functions.py:11-65 + backward, or Extraction.py:178-572).  The
    reference holds all points of a call in one autograd graph and runs out of memory beyond a few thousand
    (SURVEY 5), so collocation points go through it in chunks with the gradients accumulated."""
    _, _, Loss_Functions, Extraction, _ = reference_modules()
    U, N = nets
    if c.get("extraction"):
        b, Lib, counts, mil = Extraction.Generate_Library(U, N, c["m"], c["n"], P, c["K"], device)
        X, Res = Extraction.Recursive_Feature_Elimination(Lib, b)
        Extraction.Rank_Candidate_Solutions(X, Res)
        return float(Res[0])
    for p in list(U.parameters()) + list(N.parameters()):
        p.grad = None
    M, total = P.shape[0], 0.0
    for lo in range(0, M, c["ref_chunk"]):
        chunk = P[lo:lo + c["ref_chunk"]].clone()
        loss = Loss_Functions.Collocation_Loss(U, N, c["m"], c["n"], chunk, torch.float32, device) * (chunk.shape[0] / M)
        loss.backward()
        total += float(loss.item())
    return total


def port_step(c, sd_u, sd_n, P):
    from oracle import pde_read_oracle as orc
    if c.get("extraction"):
        b, Lib, counts, mil = orc.generate_library(sd_u, c["act"], sd_n, c["act"], c["m"], c["n"], P, c["K"])
        X, Res, order = orc.recursive_feature_elimination(Lib, b)
        orc.rank_candidate_solutions(X, Res)
        return float(Res[0])
    loss, gs, gp = orc.collocation_loss_and_grads(sd_u, c["act"], sd_n, c["act"], c["m"], c["n"], P,
                                                  chunk=c["ref_chunk"])
    return loss


def time_cpu_reference(c, sd_u, sd_n, sample_points, steps, warmup, budget_s=150.0):
    """Times the reference path on all host cores over a bounded sample of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = make_points_cpu(c, sample_points, seed=999)
    real = reference_modules() is not None
    cpu = torch.device("cpu")
    nets = reference_nets(c, sd_u, sd_n, cpu) if real else None

    def one(Pp):
        return reference_step(c, nets, Pp, cpu) if real else port_step(c, sd_u, sd_n, Pp)

    t0 = time.perf_counter()
    one(P)                                          # first warm-up doubles as a cost probe
    t1 = time.perf_counter() - t0
    if (steps + warmup) * t1 > budget_s and sample_points > c["ref_chunk"]:
        scale = budget_s / ((steps + warmup) * t1)
        sample_points = max(c["ref_chunk"], int(sample_points * scale) // c["ref_chunk"] * c["ref_chunk"])
        P = P[:sample_points].contiguous()
    for _ in range(max(0, warmup - 1)):
        one(P)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one(P)
        ts.append(time.perf_counter() - t0)
    total = float(np.sum(ts))
    kind = "reference" if real else "port"
    what = ("unmodified reference (oracle/_ref, byte-compiled from /root/reference/Code)" if real
            else "oracle/pde_read_oracle.py (port; its RFE is vectorised, i.e. faster than the reference's Python loops)")
    sample = "%d points per step x %d steps, autograd graphs of %d points, %d host threads, %s" % (
        sample_points, steps, c["ref_chunk"], cores, what)
    return dict(value=sample_points * steps / total, cores=cores, sample_points=sample_points,
                ms_per_step=1e3 * total / steps, kind=kind, sample=sample)


def time_reference_on_gpu(c, sd_u, sd_n, dev, sample_points, steps=3):
    """Context only: the same unmodified reference Python with Device = cuda ('optional
    stronger comparator'): thousands of small ATen launches per step."""
    if reference_modules() is None:
        return None
    try:
        nets = reference_nets(c, sd_u, sd_n, dev)
        P = make_points_cpu(c, sample_points, seed=999).to(dev)
        reference_step(c, nets, P, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            reference_step(c, nets, P, dev)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": sample_points * steps / dt, "unit": "points/s",
                "sample": "%d points per step x %d steps, autograd graphs of %d points, reference Python with Device=cuda" % (
                    sample_points, steps, c["ref_chunk"])}
    except Exception as e:      # e.g. out of memory inside the reference's autograd graph
        return {"value": None, "error": "%s: %s" % (type(e).__name__, str(e)[:120])}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def dram_traffic(name, kernel):
    """Per-launch DRAM bytes of a kernel from this round's `ncu --set full` capture (profiles/dram_traffic.json lists
    the capture each number comes from); None when the kernel of this build has not been captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            ent = json.load(fh).get("%s:%s" % (name, kernel))
    except Exception:
        return None, None
    if isinstance(ent, dict):
        return ent.get("bytes"), ent.get("source")
    return ent, None


# ------------------------------------------------------------------------------------------------
# our arm: one configuration
# ------------------------------------------------------------------------------------------------
class Env:
    pass


def run_ours(name, c, env, steps, warmup, with_cpu, peaks):
    dev, world, rank, local_rank, lib, dist = env.dev, env.world, env.rank, env.local_rank, env.lib, env.dist
    from pde_read_b200 import functional as F
    from pde_read_b200.Extraction import Extract_PDE
    from pde_read_b200.Loss_Functions import Collocation_Loss
    from pde_read_b200.parallel import shard_bounds
    extraction = bool(c.get("extraction"))
    metric = "extraction_points_per_second" if extraction else "collocation_points_per_second_fwd_bwd"
    unit = "points/s"
    FU, FN = flops_per_point(c)
    U, N, sd_u, sd_n = build_nets(c, dev)
    if c["scaling"] == "weak":
        M_local, M_total = c["M"], c["M"] * world
    else:
        lo, hi = shard_bounds(c["M"], rank, world)
        M_local, M_total = hi - lo, c["M"]
    P = F.generate_points(np.asarray(c["box"], dtype=np.float64), M_local, dev, seed=1000 + rank)
    P_host = P.cpu().pin_memory()
    params = list(U.parameters()) + list(N.parameters())
    flush = env.flush

    def zero_grads():
        for p in params:
            p.grad = None

    if extraction:
        def step(Pd):
            return Extract_PDE(U, N, c["m"], c["n"], Pd, c["K"], reducer=env.reducer)
    else:
        def step(Pd):
            zero_grads()
            loss = Collocation_Loss(U, N, c["m"], c["n"], Pd, torch.float32, dev)
            loss.backward()
            return loss

    MAXEV = 320

    def new_events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()
        return evs

    for _ in range(warmup):
        step(P)
    torch.cuda.synchronize()

    def timed_region():
        starts, ends = new_events(steps), new_events(steps)
        stage = [new_events(MAXEV) for _ in range(steps)]
        recorded, tags = [], None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with ClockSampler(local_rank) as cs:
            for i in range(steps):
                flush.zero_()                                            # L2 flush between timed iterations
                arr = (ctypes.c_void_p * MAXEV)(*[e.cuda_event for e in stage[i]])
                lib.pderead_debug_stage_events(arr, MAXEV)
                starts[i].record()
                step(P)
                ends[i].record()
                recorded.append(lib.pderead_debug_stage_events_recorded())
                if tags is None:
                    buf = (ctypes.c_int32 * MAXEV)()
                    nt_ = lib.pderead_debug_stage_tags(buf, MAXEV)
                    tags = [int(buf[k]) for k in range(nt_)]
                lib.pderead_debug_stage_events(None, 0)
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        nrec = min(recorded) if recorded else 0
        per_stage = np.zeros((steps, nrec))
        for i in range(steps):
            prev = starts[i]
            for k in range(nrec):
                per_stage[i, k] = prev.elapsed_time(stage[i][k])
                prev = stage[i][k]
        return ms, per_stage, cs.summary(), nrec, (tags or [])[:nrec]

    ms, per_stage, clocks, nrec, tags = timed_region()
    if _BAD & set(clocks["reasons"]):
        ms, per_stage, clocks, nrec, tags = timed_region()       # rejected by throttling: re-measure once
        clocks["remeasured"] = True
    total_ms = float(np.sum(ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    value = M_total * steps / (total_ms * 1e-3)

    # ------------------------------------------------ end-to-end: host buffers in, loss/result out.  Same loop shape
    # as the timed region above (L2 flush before every step, outside the clock); the clock is the host's, around
    # "copy the step's points from pinned host memory -> step -> result on the host".
    def e2e_step():
        Pd = P_host.to(dev, non_blocking=True)
        out = step(Pd)
        if extraction:
            return float(out[1][0])             # Residual already on the host (numpy)
        return out.item()                       # D2H read of the loss
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    e2e_steps = max(3, min(steps, 50))
    e2e_s = 0.0
    for _ in range(e2e_steps):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = M_total * e2e_steps / e2e_s
    C = library_columns(c["n"], c["K"]) if extraction else 0
    d2h = 4 if not extraction else 4 * (C * C + (C + 1) + 3 * C)

    # comm share of the step (multi-GPU): events around the step's exchange -- the reducer's in-place sum of a buffer of
    # the step's payload size (peer-memory all-reduce on one box, NCCL otherwise) -- and around NCCL's for comparison
    comm_ms, comm_nccl_ms, comm_kind = None, None, None
    if world > 1 and not extraction:
        from pde_read_b200 import Loss_Functions as LF
        red = LF.get_reducer() if hasattr(LF, "get_reducer") else None
        nflat = sum((p.numel() + 3) // 4 * 4 for p in params) + 4
        buf = torch.zeros(nflat, dtype=torch.float32, device=dev)

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / 10

        comm_nccl_ms = timed(lambda: dist.all_reduce(buf))
        if red is not None and hasattr(red, "_sum_flat"):
            comm_ms = timed(lambda: red._sum_flat(buf))
            comm_kind = "peer memory (pderead_peer_allreduce)" if getattr(red, "_peer", None) is not None and red._peer.ok else "nccl"
        else:
            comm_ms, comm_kind = comm_nccl_ms, "nccl"

    if rank != 0:
        return None

    # ------------------------------------------------ roofline of the dominant kernel (per-kernel CUDA events)
    stage_ms = per_stage.mean(axis=0) if nrec else np.zeros(0)
    by_tag, launches = {}, {}
    for k in range(nrec):
        nm = TAGS.get(tags[k] if k < len(tags) else 0, "other")
        by_tag[nm] = by_tag.get(nm, 0.0) + float(stage_ms[k])
        launches[nm] = launches.get(nm, 0) + 1
    peak = peaks["fp32"]
    mp = measured_peaks()
    if not extraction:
        key, flops, label = "jet_bwd<U>", 2.0 * FU * M_local, "reverse kernel of the U network (jets)"
        step_flops = 3.0 * (FU + FN) * M_local
        abytes = 12
    else:
        Jx = 1 + c["n"]
        FUx = 2 * (2 * c["WU"] + (c["LU"] - 1) * c["WU"] ** 2 * Jx + c["WU"] * Jx)
        key, flops, label = "jet_fwd<U>", float(FUx) * M_local, "forward kernel of the U network (x-series only)"
        step_flops = float(FUx + FN) * M_local
        abytes = 8
    roof = None
    if by_tag.get(key, 0.0) > 0.0:
        dur_ms, nl = by_tag[key], launches[key]
        ach = flops / (dur_ms * 1e-3) / 1e12
        traffic, tsrc = dram_traffic(name, key)
        roof = {"bound": "fp32_fma", "kernel": "%s, %d launch(es) per step" % (label, nl), "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peaks["source"], "traffic": traffic,
                "traffic_source": tsrc, "kernel_ms": dur_ms / nl, "algorithmic_flops_per_launch": flops / nl}
    else:
        ach = step_flops / (ms_per_step * 1e-3) / 1e12
        roof = {"bound": "fp32_fma", "kernel": "whole step (no per-kernel events)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peaks["source"], "traffic": None}
    roof["stage_ms"] = by_tag
    roof["launches_per_step"] = launches
    ws = step_flops / (ms_per_step * 1e-3) / 1e12
    roof["whole_step"] = {"achieved": ws, "frac": ws / peak}
    roof["hbm"] = {"algorithmic_bytes_per_point": abytes, "achieved_gbs": abytes * M_local / (ms_per_step * 1e-3) / 1e9,
                   "peak_gbs": mp.get("hbm_gbs")}
    if extraction and by_tag.get("library_gram", 0.0) > 0.0:
        gflops = float(C * (C + 1) + 2 * C) * M_local
        gach = gflops / (by_tag["library_gram"] * 1e-3) / 1e12
        roof["gram"] = {"bound": "fp64_fma", "kernel": "gram_kernel (library + upper-triangular [A|b]^T[A|b])",
                        "achieved": gach, "peak": peaks["fp64"], "unit": "TFLOP/s", "frac": gach / peaks["fp64"],
                        "kernel_ms": by_tag["library_gram"] / launches["library_gram"],
                        "algorithmic_flops_per_point": C * (C + 1) + 2 * C,
                        "peak_source": "pderead_fma_peak_probe variant 2 (DFMA chains), measured in this run"}
        roof["rfe_share_of_step"] = by_tag.get("rfe", 0.0) / ms_per_step
    roof["bound_note"] = ("bound by FP32 FMA issue on the CUDA cores; whole step = %.4f of the measured HBM peak "
                          "(%.1f GB/s) on algorithmic bytes and %.4f of the measured bf16 tensor peak (%.1f TFLOP/s) on "
                          "algorithmic FLOPs" % (roof["hbm"]["achieved_gbs"] / mp.get("hbm_gbs", 6538.9), mp.get("hbm_gbs", 6538.9),
                                                 ws / mp.get("bf16_tflops", 1650.7), mp.get("bf16_tflops", 1650.7)))

    # ------------------------------------------------ CPU baseline beside it (bounded sample, rank 0, N = 1 only)
    cpu = None
    if with_cpu and world == 1:
        r = time_cpu_reference(c, sd_u, sd_n, c["ref_sample"], steps=c.get("ref_steps", 3), warmup=1, budget_s=30.0)
        cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": "f32" if not extraction else "f32 (jets, library) + f64 (Gram, RFE)", "data": "synthetic",
        "config": {"workload": name + ": " + c["desc"], "points_per_step_total": M_total,
                   "points_per_gpu": M_local, "weights": "random init (reference initialiser), seed 1234",
                   "l2": "256 MiB buffer written between timed iterations (L2 flush), for `value` and for `e2e`",
                   "parallelism": "points sharded over %d GPU(s), weights replicated" % world,
                   "tensor_cores": os.environ.get("PDEREAD_TC", "auto")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(M_local * 2 * 4),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "clock": "host perf_counter around H2D copy + step + D2H of the result, device idle before each step"},
        "gpu_launches": int(nrec * steps),
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    if comm_ms is not None:
        line["comm_ms"] = comm_ms
        line["comm"] = {"exchange": comm_kind, "ms": comm_ms, "nccl_all_reduce_ms": comm_nccl_ms,
                        "note": "CUDA events around 10 back-to-back exchanges of the step's payload"}
    return line


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="run this workload only (default: C2 with the other BASELINE configurations in an 'also' block)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--points", type=int, default=None, help="override points per step (whole job)")
    args = ap.parse_args()
    primary = args.config or "C2"
    also = [] if (args.config or args.no_also) else [a for a in ALSO]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def cfg(name):
        c = dict(CONFIGS[name])
        if args.points and name == primary:
            c["M"] = args.points
        return c

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        warm = max(args.warmup if args.warmup is not None else 1, 1)

        def ref_line(name, steps, warm_):
            c = cfg(name)
            metric = "extraction_points_per_second" if c.get("extraction") else "collocation_points_per_second_fwd_bwd"
            _, _, sd_u, sd_n = build_nets(c, torch.device("cpu"))
            r = time_cpu_reference(c, sd_u, sd_n, c["ref_sample"], steps, warm_, budget_s=60.0 if name == primary else 25.0)
            return {"impl": "reference", "metric": metric, "value": r["value"], "unit": "points/s", "n_gpus": args.gpus,
                    "steps": steps, "warmup": warm_, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                    "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": name + ": " + c["desc"], "reference_sample": r["sample"]},
                    "cpu_baseline": {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                                     "sample": r["sample"]},
                    "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
        line = ref_line(primary, args.steps or 3, min(warm, 2))
        if also:
            line["also"] = {a: ref_line(a, 2, 1) for a in also if a != "C5S"}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    real_stdout = None
    if world > 1:
        # NCCL may print its version banner on stdout: keep stdout clean for the ONE JSON line (fd 1 -> stderr until then)
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the exchange is one 84..330 KB all-reduce per step: latency-bound, and NCCL's default choice for it on
        # 8 NVSwitch GPUs (ring, LL) is slower than the tree (measured: C2 step 1.71 -> 1.68 ms, e2e +6 %)
        os.environ.setdefault("NCCL_ALGO", "Tree")
        dist.init_process_group("nccl", device_id=dev)
    from pde_read_b200 import _lib, functional as F
    from pde_read_b200 import Loss_Functions
    from pde_read_b200.parallel import ShardReducer
    env = Env()
    env.dev, env.world, env.rank, env.local_rank, env.dist = dev, world, rank, local_rank, dist
    env.lib = _lib.load()
    env.reducer = ShardReducer() if world > 1 else None
    Loss_Functions.set_reducer(env.reducer)
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    peak_ffma = F.fma_peak_probe(dev, 0, 8192)
    peak_ffma2 = F.fma_peak_probe(dev, 1, 8192)
    peak_dfma = F.fma_peak_probe(dev, 2, 2048)
    peaks = {"fp32": max(peak_ffma, peak_ffma2), "fp64": peak_dfma,
             "source": ("pderead_fma_peak_probe measured in this run (FFMA %.1f, FFMA2 %.1f, DFMA %.1f TFLOP/s); "
                        "MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only, and this path is bound by FP32 FMA "
                        "issue (FP64 for the Gram kernel)" % (peak_ffma, peak_ffma2, peak_dfma))}

    warm = max(args.warmup, 3)
    c = cfg(primary)
    line = run_ours(primary, c, env, args.steps or c["steps"], warm, not args.no_cpu_baseline, peaks)
    extra = {}
    for a in also:
        ca = cfg(a)
        F.release_workspaces()
        torch.cuda.empty_cache()
        try:
            extra[a] = run_ours(a, ca, env, ca["steps"], 3, not args.no_cpu_baseline, peaks)
        except Exception as e:                      # an "also" workload must never take the headline line down
            extra[a] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0:
        if extra:
            line["also"] = extra
        if world == 1 and not args.no_cpu_baseline and not args.config:
            _, _, sd_u, sd_n = build_nets(c, torch.device("cpu"))
            line["reference_on_b200"] = time_reference_on_gpu(c, sd_u, sd_n, dev, 20000)
        if real_stdout is not None:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
