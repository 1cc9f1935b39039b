#!/usr/bin/env python
"""Benchmark of the PDE-READ hot path on B200: collocation points/s for one Collocation_Loss
forward + backward (BASELINE.json metric), or extraction points/s (--config C5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic
collocation points: Collocation_Loss(...) + .backward() through the public Python API, i.e. through
the C ABI of libpderead_b200.so.  See DESIGN.md section "Measurement" for what each key means.
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
               ref_chunk=5000, ref_sample=5000, scaling="weak"),
    "C2": dict(desc="Heat equation: U 5x50 Tanh, N 5x50 Tanh, m=1 n=2, 100k collocation pts per GPU",
               act="Tanh", LU=5, WU=50, LN=5, WN=50, m=1, n=2, M=100_000, box=[[0.0, 10.0], [0.0, 9.95]],
               ref_chunk=5000, ref_sample=20000, scaling="weak"),
    "C3": dict(desc="KdV: U 5x50 Rational, N 5x50 Rational, m=1 n=3, 1M collocation pts sharded over the GPUs",
               act="Rational", LU=5, WU=50, LN=5, WN=50, m=1, n=3, M=1_000_000, box=[[0.0, 40.0], [-20.0, 19.84]],
               ref_chunk=5000, ref_sample=5000, scaling="strong"),
    "C4": dict(desc="Kuramoto-Sivashinsky: U 5x100 Rational, N 5x100 Rational, m=1 n=4, 4M collocation pts sharded over the GPUs",
               act="Rational", LU=5, WU=100, LN=5, WN=100, m=1, n=4, M=4_000_000, box=[[0.0, 5.0], [-5.0, 4.96]],
               ref_chunk=2000, ref_sample=2000, scaling="strong"),
    "C4T": dict(desc="Kuramoto-Sivashinsky shape with Tanh nets (width 100, n=4), 4M pts sharded over the GPUs",
                act="Tanh", LU=5, WU=100, LN=5, WN=100, m=1, n=4, M=4_000_000, box=[[0.0, 5.0], [-5.0, 4.96]],
                ref_chunk=2000, ref_sample=4000, scaling="strong"),
    "C5": dict(desc="Extraction only: U 5x50 Rational, N 2x100 Rational, n=2, degree K=5 library (C=56) over 10M points sharded over the GPUs, fused Gram + RFE",
               act="Rational", LU=5, WU=50, LN=2, WN=100, m=1, n=2, K=5, M=10_000_000, box=[[0.0, 10.0], [-8.0, 7.9375]],
               ref_chunk=1000, ref_sample=5000, scaling="strong", extraction=True),
}


def flops_per_point(c):
    """Algorithmic FLOPs (dense contractions only), SURVEY.md 8d."""
    J = 1 + c["n"] + c["m"]
    FU = 2 * (2 * c["WU"] + (c["LU"] - 1) * c["WU"] ** 2 * J + c["WU"] * J)
    FN = 2 * ((c["n"] + 1) * c["WN"] + (c["LN"] - 1) * c["WN"] ** 2 + c["WN"])
    return FU, FN


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
# CPU reference (oracle port of the reference's PyTorch-autograd path), timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(c, sd_u, sd_n, P):
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
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = make_points_cpu(c, sample_points, seed=999)
    t0 = time.perf_counter()
    cpu_reference_step(c, sd_u, sd_n, P)            # first warm-up doubles as a cost probe
    t1 = time.perf_counter() - t0
    if (steps + warmup) * t1 > budget_s and sample_points > c["ref_chunk"]:
        scale = budget_s / ((steps + warmup) * t1)
        sample_points = max(c["ref_chunk"], int(sample_points * scale) // c["ref_chunk"] * c["ref_chunk"])
        P = P[:sample_points].contiguous()
    for _ in range(max(0, warmup - 1)):
        cpu_reference_step(c, sd_u, sd_n, P)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(c, sd_u, sd_n, P)
        ts.append(time.perf_counter() - t0)
    total = float(np.sum(ts))
    return dict(value=sample_points * steps / total, cores=cores, sample_points=sample_points,
                ms_per_step=1e3 * total / steps)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--points", type=int, default=None, help="override points per step (whole job)")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.points:
        c["M"] = args.points
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "extraction_points_per_second" if c.get("extraction") else "collocation_points_per_second_fwd_bwd"
    unit = "points/s"
    FU, FN = flops_per_point(c)

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        U, N, sd_u, sd_n = build_nets(c, torch.device("cpu"))
        r = time_cpu_reference(c, sd_u, sd_n, c["ref_sample"], args.steps, args.warmup)
        sample = "%d points per step, autograd graphs of %d points, %d host threads" % (
            r["sample_points"], c["ref_chunk"], r["cores"])
        line = {
            "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.config + ": " + c["desc"], "reference_sample": sample},
            "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the exchange is one 84..330 KB all-reduce per step: latency-bound, and NCCL's default choice for it on
        # 8 NVSwitch GPUs (ring, LL) is slower than the tree (measured: C2 step 1.71 -> 1.68 ms, e2e +6 %)
        os.environ.setdefault("NCCL_ALGO", "Tree")
        dist.init_process_group("nccl", device_id=dev)
    from pde_read_b200 import _lib, functional as F
    from pde_read_b200 import Loss_Functions
    from pde_read_b200.Extraction import Extract_PDE
    from pde_read_b200.Loss_Functions import Collocation_Loss
    from pde_read_b200.parallel import ShardReducer, shard_bounds
    lib = _lib.load()

    U, N, sd_u, sd_n = build_nets(c, dev)
    reducer = ShardReducer() if world > 1 else None
    Loss_Functions.set_reducer(reducer)
    if c["scaling"] == "weak":
        M_local, M_total = c["M"], c["M"] * world
    else:
        lo, hi = shard_bounds(c["M"], rank, world)
        M_local, M_total = hi - lo, c["M"]
    P = F.generate_points(np.asarray(c["box"], dtype=np.float64), M_local, dev, seed=1000 + rank)
    P_host = P.cpu().pin_memory()
    params = list(U.parameters()) + list(N.parameters())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def zero_grads():
        for p in params:
            p.grad = None

    if c.get("extraction"):
        def step(Pd):
            return Extract_PDE(U, N, c["m"], c["n"], Pd, c["K"], reducer=reducer)
    else:
        def step(Pd):
            zero_grads()
            loss = Collocation_Loss(U, N, c["m"], c["n"], Pd, torch.float32, dev)
            loss.backward()
            return loss

    # stage events: one per kernel launch of the library, for every timed step
    MAXEV = 160
    def new_events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()
        return evs
    for _ in range(args.warmup):
        step(P)
    torch.cuda.synchronize()

    def timed_region():
        starts = new_events(args.steps)
        ends = new_events(args.steps)
        stage = [new_events(MAXEV) for _ in range(args.steps)]
        recorded = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with ClockSampler(local_rank) as cs:
            for i in range(args.steps):
                flush.zero_()                                            # L2 flush between timed iterations
                arr = (ctypes.c_void_p * MAXEV)(*[e.cuda_event for e in stage[i]])
                lib.pderead_debug_stage_events(arr, MAXEV)
                starts[i].record()
                step(P)
                ends[i].record()
                recorded.append(lib.pderead_debug_stage_events_recorded())
                lib.pderead_debug_stage_events(None, 0)
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        # per-kernel durations: stage event k-1 -> k (the first kernel is measured from the step start)
        nrec = min(recorded) if recorded else 0
        per_stage = np.zeros((args.steps, nrec))
        for i in range(args.steps):
            prev = starts[i]
            for k in range(nrec):
                per_stage[i, k] = prev.elapsed_time(stage[i][k])
                prev = stage[i][k]
        return ms, per_stage, cs.summary(), nrec

    ms, per_stage, clocks, nrec = timed_region()
    if _BAD & set(clocks["reasons"]):
        ms, per_stage, clocks, nrec = timed_region()       # rejected by throttling: re-measure once
        clocks["remeasured"] = True
    total_ms = float(np.sum(ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = M_total * args.steps / (total_ms * 1e-3)

    # ------------------------------------------------ end-to-end: host buffers in, loss/result out
    def e2e_step():
        Pd = P_host.to(dev, non_blocking=True)
        out = step(Pd)
        if c.get("extraction"):
            return float(out[1][0])             # Residual already on the host (numpy)
        return out.item()                       # D2H read of the loss
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = M_total * e2e_steps / e2e_s
    d2h = 4 if not c.get("extraction") else 4 * (56 * 56 + 57 + 56 * 3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------ roofline of the dominant kernel
    peak_ffma = F.fma_peak_probe(dev, 0, 8192)
    peak_ffma2 = F.fma_peak_probe(dev, 1, 8192)
    peak = max(peak_ffma, peak_ffma2)
    peak_src = ("pderead_fma_peak_probe measured in this run (FFMA %.1f, FFMA2 %.1f TFLOP/s); MEASURED_PEAKS.json "
                "carries HBM and bf16-tensor peaks only, and this path is bound by FP32 FMA issue" % (peak_ffma, peak_ffma2))
    stage_ms = per_stage.mean(axis=0) if nrec else np.zeros(0)
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            traffic_tab = json.load(fh)
    except Exception:
        pass
    roof = None
    if not c.get("extraction") and nrec >= 8 and (nrec - 4) % 4 == 0:
        # launches of one fused step: pack(U) pack(N) | per chunk of <= 131072 points: U-jets fwd, N fwd + residual,
        # N reverse, U reverse | reduce(U) reduce(N).  The U reverse kernel dominates; its launches are summed.
        nchunk = (nrec - 4) // 4
        per_chunk = stage_ms[2:2 + 4 * nchunk].reshape(nchunk, 4).sum(axis=0)
        names = ["mlp_jet_fwd<U>", "mlp_jet_fwd<N>", "mlp_jet_bwd<N>", "mlp_jet_bwd<U>"]
        dur_ms = float(per_chunk[3])
        flops = 2.0 * FU * M_local                           # reverse pass of U = 2x its forward
        ach = flops / (dur_ms * 1e-3) / 1e12
        tkey = "%s:mlp_jet_bwd<U>" % args.config
        traffic = traffic_tab.get(tkey)
        roof = {"bound": "fp32_fma", "kernel": "mlp_jet_bwd_kernel (U network), %d launch(es) per step" % nchunk,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src,
                "traffic": traffic, "kernel_ms": dur_ms / nchunk, "algorithmic_flops_per_launch": flops / nchunk,
                "stage_ms": dict({"pack_U": float(stage_ms[0]), "pack_N": float(stage_ms[1])},
                                 **{n_: float(v) for n_, v in zip(names, per_chunk)},
                                 **{"reduce_partials<U>": float(stage_ms[-2]), "reduce_partials<N>": float(stage_ms[-1])}),
                "whole_step": {"achieved": 3.0 * (FU + FN) * M_local / (ms_per_step * 1e-3) / 1e12,
                               "frac": 3.0 * (FU + FN) * M_local / (ms_per_step * 1e-3) / 1e12 / peak},
                "hbm": {"algorithmic_bytes_per_point": 12, "achieved_gbs": 12.0 * M_local / (ms_per_step * 1e-3) / 1e9,
                        "peak_gbs": measured_peaks().get("hbm_gbs")}}
    elif c.get("extraction") and nrec >= 6 and (nrec - 3) % 3 == 0:
        # pack(U) pack(N) | per chunk: U-jets fwd (no t-series), N fwd, fused library + Gram | RFE + ranking
        nchunk = (nrec - 3) // 3
        per_chunk = stage_ms[2:2 + 3 * nchunk].reshape(nchunk, 3).sum(axis=0)
        Jx = 1 + c["n"]
        FUx = 2 * (2 * c["WU"] + (c["LU"] - 1) * c["WU"] ** 2 * Jx + c["WU"] * Jx)
        dur_ms = float(per_chunk[0])
        flops = float(FUx) * M_local
        ach = flops / (dur_ms * 1e-3) / 1e12
        roof = {"bound": "fp32_fma", "kernel": "mlp_jet_fwd_kernel (U network, x-series only), %d launch(es) per step" % nchunk,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src,
                "traffic": traffic_tab.get("%s:mlp_jet_fwd<U>" % args.config), "kernel_ms": dur_ms / nchunk,
                "algorithmic_flops_per_launch": flops / nchunk,
                "stage_ms": {"pack_U": float(stage_ms[0]), "pack_N": float(stage_ms[1]), "mlp_jet_fwd<U>": float(per_chunk[0]),
                             "fwd<N>": float(per_chunk[1]), "library_gram": float(per_chunk[2]), "rfe": float(stage_ms[-1])},
                "whole_step": {"achieved": float(FUx + FN) * M_local / (ms_per_step * 1e-3) / 1e12,
                               "frac": float(FUx + FN) * M_local / (ms_per_step * 1e-3) / 1e12 / peak},
                "hbm": {"algorithmic_bytes_per_point": 8, "achieved_gbs": 8.0 * M_local / (ms_per_step * 1e-3) / 1e9,
                        "peak_gbs": measured_peaks().get("hbm_gbs")}}
    if roof is None:
        flops_step = (3.0 * (FU + FN) if not c.get("extraction") else float(FU + FN)) * M_local
        ach = flops_step / (ms_per_step * 1e-3) / 1e12
        roof = {"bound": "fp32_fma", "kernel": "whole step (stage events: %d)" % nrec, "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "stage_ms": [float(v) for v in stage_ms]}

    # the two standard rooflines, for context (neither binds: 12 algorithmic bytes and ~3e5 fp32 FLOP per point,
    # with 3xTF32 accuracy too low for the gradient path -- DESIGN.md section 4)
    mp = measured_peaks()
    step_flops = (3.0 * (FU + FN) if not c.get("extraction") else float(FU + FN)) * M_local / (ms_per_step * 1e-3)
    step_bytes = (12.0 if not c.get("extraction") else 8.0) * M_local / (ms_per_step * 1e-3)
    roof["bound_note"] = ("kernel is bound by FP32 FMA issue on the CUDA cores; whole step = %.4f of the measured HBM peak "
                          "(%.1f GB/s) on algorithmic bytes and %.4f of the measured bf16 tensor peak (%.1f TFLOP/s) on "
                          "algorithmic FLOPs" % (step_bytes / 1e9 / mp.get("hbm_gbs", 6538.9), mp.get("hbm_gbs", 6538.9),
                                                 step_flops / 1e12 / mp.get("bf16_tflops", 1650.7), mp.get("bf16_tflops", 1650.7)))

    # ------------------------------------------------ CPU baseline beside it (bounded sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = time_cpu_reference(c, sd_u, sd_n, c["ref_sample"], steps=3, warmup=1, budget_s=40.0)
        cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
               "sample": "%d points per pass x 3 passes, autograd graphs of %d points (oracle/pde_read_oracle.py)" % (
                   r["sample_points"], c["ref_chunk"])}

    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": "f32" if not c.get("extraction") else "f32 (jets, library) + f64 (Gram, RFE)", "data": "synthetic",
        "config": {"workload": args.config + ": " + c["desc"], "points_per_step_total": M_total,
                   "points_per_gpu": M_local, "weights": "random init (reference initialiser), seed 1234",
                   "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                   "parallelism": "points sharded over %d GPU(s), weights replicated" % world,
                   "tensor_cores": os.environ.get("PDEREAD_TC", "auto") +
                                   " (auto: tcgen05 kernels for forward-only calls of width > 64 and depth >= 3; CUDA cores on the gradient path)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(M_local * 2 * 4),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(nrec * args.steps),
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
