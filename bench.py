#!/usr/bin/env python3
"""bench.py — headline benchmark of fatode_b200 (contract: see the task brief / DESIGN.md §6).

Metric (BASELINE.json): cell-integrations/sec (forward + adjoint) for CBM-IV ROS_ADJ.
One cell-integration = one complete INTEGRATE_ADJ (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:34): forward
Rodas4 sweep 12h -> 84h with trajectory tape, ADJINIT, discrete adjoint sweep for NADJ = 32 cost
functionals and NP = 29 rate-coefficient parameters (the reference driver's problem,
EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_adj_dr.F90:167-292), on synthetic per-cell initial states.

A "step" is one pass of the hot path over one batch of cells per GPU: one C-ABI call.  Default batch: 2^18 cells when the
whole run (warm-up + timed steps + the e2e pass) fits ten minutes, else 2^17 (default_cells(); the >= 1 M-cell workload of
BASELINE.json is processed in waves of ~8.3 k cells whatever the batch size, so throughput does not depend on it;
`--cells 1048576` runs the full-size batch in about a minute per step; `--scaling strong` splits a fixed 2^20 cells over N GPUs).
  value : cells/s with y0 and all outputs resident in HBM (device pointers through the C ABI)
  e2e   : cells/s through the same C-ABI call with pinned HOST buffers (H2D of y0, D2H of y, Lambda,
          Mu, ISTATUS, RSTATUS, IERR inside the timed region)
  roofline    : dominant kernel (k_ros_adj_pair, the backward sweep): ALGORITHMIC FP64 flops of SURVEY.md 8d (the dense
                FULL_ALGEBRA count, evaluated on the returned counters) / its CUDA-event time, against the DFMA peak
                measured on this GPU in this run (MEASURED_PEAKS.json has no FP64 figure); the same for the forward
                kernel; EXECUTED fractions (the kernels run the sparse patterns) only from an ncu capture of the very
                library this run loads (profiles/executed_counts.json, matched by SHA-256) -- nothing is hard-coded
  cpu_baseline: the CPU oracle (C++ restatement of the reference, OpenMP over cells) on a bounded sample
  side        : (N = 1) one timed call each of the other BASELINE configs (small_strato FWD/ROS 2^20 cells, cbm4 RK_ADJ and
                SDIRK_TLM, the shallow-water ensemble) with algorithmic roofline fraction and an oracle bit check
`--impl reference` times that CPU restatement alone (the reference Fortran cannot be built: no Fortran compiler in the
image) on every host thread and prints the same JSON line with "impl": "reference"; it loads oracle/liboracle.so only.
Multi-GPU (torchrun, one rank per GPU): cells are sharded, no data-path collective; one NCCL
all-reduce of the 29x32 Mu sums per step.  Time = max over ranks, barrier on both sides.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N, NP, NADJ, S = 32, 29, 32, 6
T0 = 12.0 * 3600.0
T1 = T0 + 72.0 * 3600.0
F01 = float(np.float32(0.1))


def tolerances():
    # rtol(i) = 0.1*1e-4, atol(i) = 0.1*1e-6 with a REAL(4) 0.1 (cbm4_ros_adj_dr.F90:181-189)
    return np.full(N, F01 * 1.0e-4), np.full(N, F01 * 1.0e-6)


def icntrl_rodas4():
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    return ic


# ---- algorithmic flop model, evaluated on the returned ISTATUS counters
def flops_adjoint_per_step(nadj=NADJ):
    n, s, p = N, S, NP
    LU, SOL, MV, MVP = (2 * n ** 3) // 3 + n * n, 2 * n * n, 2 * n * n, 2 * n * p
    F_jac, F_hess, F_htv, F_jacp, F_djdr = 636, 129, 1483, 250, 250
    per_col = s * n + 2 * n * s * (s - 1) + s * SOL + s * MV + s * (F_htv + 2 * n) + s * (2 * MVP + 2 * p) + \
        (2 * n * s + MVP + MV + 2 * n)
    return F_jac + LU + s * F_jac + F_hess + s * (F_jacp + F_djdr) + nadj * per_col + \
        (F_jac + 2 * n * n + 2 * F_jacp + 2 * n * p)


def flops_forward(attempts, accepted, singular=0):
    n, s = N, S
    LU, SOL, F_fun, F_jac = (2 * n ** 3) // 3 + n * n, 2 * n * n, 573, 636
    per_attempt = s * SOL + 5 * F_fun + 2 * n * s * (s - 1) + 2 * n * 4 + 4 * n * s + 7 * n
    return (attempts + singular) * LU + attempts * per_attempt + accepted * (2 * F_fun + F_jac + 2 * n)


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def _load_inputs():
    """fatode_b200/inputs.py loaded BY PATH: the reference arm and the CPU baseline must not import the product package
    (its __init__ dlopens libfatode_b200.so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_fatode_inputs", os.path.join(ROOT, "fatode_b200", "inputs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def host_threads():
    """threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which only changes
    OpenMP's default; the oracle takes the count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_initial_state(L, model=3, n=N):
    y = np.zeros(n)
    L.oracle_initial_state(model, y.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return y


def oracle_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    L.oracle_ros_adj_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_double,
                                       ctypes.c_double, dp, dp, ip, dp, ip, dp, ip, ctypes.c_int]
    return L


def cpu_run(L, y0, nthreads):
    """CPU restatement (oracle) over the cells of y0 with OpenMP; returns seconds."""
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    nc = y0.shape[0]
    y = y0.copy()
    rtol, atol = tolerances()
    ic, rc = icntrl_rodas4(), np.zeros(20)
    lam, mu = np.zeros((nc, NADJ, N)), np.zeros((nc, NADJ, NP))
    ist, rst, ierr = np.zeros((nc, 20), dtype=np.int32), np.zeros((nc, 20)), np.zeros(nc, dtype=np.int32)
    d = lambda a: a.ctypes.data_as(dp)
    i = lambda a: a.ctypes.data_as(ip)
    t = time.perf_counter()
    L.oracle_ros_adj_batch(3, nc, NP, NADJ, d(y), d(lam), d(mu), T0, T1, d(atol), d(rtol), i(ic), d(rc), i(ist), d(rst), i(ierr),
                           nthreads)
    dt = time.perf_counter() - t
    assert (ierr == 1).all()
    return dt


def cpu_baseline(sample_cells=None):
    """the CPU oracle on a bounded sample of the GPU arm's own first cells (about 20-30 s of CPU work)"""
    perturbed_states = _load_inputs().perturbed_states
    L = oracle_lib()
    cores = host_threads()
    nc = sample_cells or max(64, 48 * cores)
    y0 = perturbed_states(oracle_initial_state(L), nc)
    dt = cpu_run(L, y0, cores)
    return {"value": nc / dt, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": f"{nc} cells of the same workload (perturbed driver states, cells 0..{nc-1} = the first cells of the GPU "
                      f"arm's first step), {dt:.1f} s wall, OpenMP schedule(dynamic) over cells on {cores} threads, C++ "
                      f"restatement of ADJ/ROS_ADJ (oracle/), strict IEEE"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (C++ restatement; no Fortran compiler in this image) on all host
    threads.  Loads oracle/liboracle.so only — never the product package or library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    perturbed_states = _load_inputs().perturbed_states
    L = oracle_lib()
    cores = host_threads()
    nsteps = args.warmup + args.steps
    # bounded sample per step: >= 4096 cells over the timed steps when that fits ~5 minutes of CPU time (1.1 - 1.9 cells/s
    # per thread measured for this workload on the hosts seen so far), never fewer than 4 cells per thread
    budget_cells = int(300.0 * 1.5 * cores / max(nsteps, 1))
    nc = max(4 * cores, min(max(-(-4096 // max(args.steps, 1)), 4 * cores), budget_cells))
    gpu_cells = default_cells(args)
    base = oracle_initial_state(L)
    times = []
    for s in range(nsteps):
        y0 = perturbed_states(base, nc, first_cell=s * gpu_cells * args.gpus)     # the first cells of the GPU arm's step s
        dt = cpu_run(L, y0, cores)
        if s >= args.warmup:
            times.append(dt)
    tot = sum(times)
    v = nc * len(times) / tot
    line = {"impl": "reference", "metric": "cell-integrations/sec (fwd + adjoint), CBM-IV ROS_ADJ", "value": v, "unit": "cells/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(gpu_cells, args.gpus,
                                      note=f"CPU arm: same per-cell workload; each step integrates the first {nc} cells of the "
                                           f"GPU arm's step (a bounded sample: the full {gpu_cells * args.gpus}-cell step would "
                                           f"take {gpu_cells * args.gpus / max(v, 1e-9) / 3600.0:.1f} h on this host)"),
            "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                             "sample": f"{nc} cells per step x {len(times)} timed steps, OpenMP over cells on {cores} host threads"},
            "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def default_cells(args):
    """cells per GPU per step.  Default: 2^18 when the whole run (warm-up + timed device steps + the e2e pass) fits ten
    minutes at the measured ~16 k cells/s, else 2^17 (the driver's 25-step runs must end inside its 870 s limit);
    --scaling strong splits a fixed 2^20 cells over the GPUs."""
    if args.scaling == "strong":
        total = args.cells if args.cells > 0 else (1 << 20)
        return max(1, total // max(args.gpus, 1))
    if args.cells > 0:
        return args.cells
    calls = args.warmup + args.steps + (0 if args.no_e2e else args.steps + 1)
    return (1 << 18) if calls * (1 << 18) / 16000.0 <= 600.0 else (1 << 17)


def workload_config(cells, ngpus, note=None):
    c = {"workload": "EXAMPLES/cbm4 ADJ/ROS_ADJ: CBM-IV N=32, Rodas4 (ICNTRL(3)=5), NADJ=32, NP=29, t=12h..84h, "
                     "rtol=0.1f*1e-4, atol=0.1f*1e-6, y0 = driver state*(1+0.1u) per cell (cell 0 unperturbed)",
         "cells_per_gpu_per_step": int(cells), "global_cells_per_step": int(cells * ngpus), "parallelism": f"cells sharded x{ngpus}",
         "l2_policy": "working set >> L2 (trajectory tape ~27 MB per cell streamed through HBM; no explicit flush needed)"}
    if note:
        c["note"] = note
    return c


SIDE_DEFAULT_CELLS = {"small_strato_fwd": 1 << 20, "cbm4_ros_tlm": 32768, "cbm4_rk_adj": 9472, "cbm4_rk_adj_direct": 9472,
                      "cbm4_rk_adj_rtol1e-5": 4736, "cbm4_rk_adj_rtol1e-4": 4736, "cbm4_rk_tlm": 9472, "cbm4_sdirk_adj": 9472,
                      "cbm4_sdirk_tlm": 9472, "cbm4_sdirk_tlm_direct": 2368, "cbm4_sdirk_tlm_trunc": 2368, "swe_erk": 65536, "swe_erk_tlm": 16384, "swe_erk_tlm_dense": 16384, "swe_erk_adj": 16384, "cbm4_rk": 18944, "cbm4_sdirk": 18944}


def dense_flops_from_counters(ist, n, f_fun, f_jac, complex_pairs=False, sens_cols=0):
    """ALGORITHMIC flops of a run from its returned ISTATUS counters, with the dense FULL_ALGEBRA costs the reference pays
    (SURVEY.md section 8d): LU = 2/3 n^3 + n^2, solve = 2 n^2; an implicit Runge-Kutta 'decomposition' is a real + a complex
    LU (4 x the real cost) and its Nsol counts real and complex solves alike (average 5 n^2); every sensitivity solve comes
    with at least one dense Jacobian mat-vec (2 n^2)."""
    I = ist.astype(np.float64)
    nfun, njac, ndec, nsol = I[:, 0].sum(), I[:, 1].sum(), I[:, 5].sum(), I[:, 6].sum()
    lu = 2.0 * n ** 3 / 3.0 + n * n
    sol = 2.0 * n * n
    if complex_pairs:
        lu, sol = 5.0 * lu, 5.0 * n * n
    f = nfun * f_fun + njac * f_jac + ndec * lu + nsol * sol
    if sens_cols:
        f += nsol * (sens_cols / (sens_cols + 1.0)) * 2.0 * n * n
    return float(f)


def side_measure(w, nc, steps, warmup, fp64_peak=None, hbm_peak=None):
    """One side measurement (1 GPU, host buffers through the C ABI) of another BASELINE config / entry point, next to the CPU
    oracle on a bounded sample with a bit-for-bit check of state and counters.  Returns the JSON-able line."""
    import torch  # noqa: F401  (initialises CUDA the same way the main arm does)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fatode_b200 as F
    import oracle_lib as O
    from fatode_b200.inputs import perturbed_states
    flops = None
    if w == "small_strato_fwd":
        rtol, atol = np.full(5, 1.0e-1), np.full(5, 1.0)
        for _ in range(5):                      # tolerances of the driver's fifth sweep (small_strato_dr.F90:27-38)
            rtol, atol = F01 * rtol, F01 * atol
        t0, t1 = T0, T0 + 30 * 24 * 3600.0
        y0 = perturbed_states(F.initial_state("small_strato"), nc)
        run = lambda y: F.integrate("small_strato", t0, t1, y, rtol, atol)
        cpu = lambda y: O.ros_fwd("small_strato", y, t0, t1, atol, rtol)
        desc = "EXAMPLES/small_strato FWD/ROS (Ros4, REAL(4) tableau), N=5, 12h..30d, rtol 1e-6 atol 1e-5 (REAL(4) 0.1 products) (BASELINE config 2)"
        ncpu = 256
        # FWD/ROS does not count Ndec/Nsol: one LU + S = 4 solves per attempt (Nstp), 3 new-function stages
        flops = lambda r: float((r.istatus[:, 2].astype(np.float64) * (2 * 125 / 3.0 + 25 + 4 * 50 + 2 * 5 * 4 * 3 + 4 * 5 * 4 + 7 * 5)
                                 + r.istatus[:, 0].astype(np.float64) * 42 + r.istatus[:, 1].astype(np.float64) * 52).sum())
    elif w == "cbm4_ros_tlm":
        rtol, atol = tolerances()
        ic = icntrl_rodas4()
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        eye = np.eye(32)
        run = lambda y: F.integrate_tlm("cbm4", T0, T1, y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), rtol, atol, icntrl_u=ic)
        cpu = lambda y: O.ros_tlm("cbm4", y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), T0, T1, atol, rtol, ic)
        desc = "EXAMPLES/cbm4 TLM/ROS_TLM: Rodas4, NTLM=32, Y_tlm(t0)=I, ICNTRL(12)=0, 12h..84h, driver tolerances of ROS_ADJ"
        ncpu = 32
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, sens_cols=32)
    elif w in ("cbm4_rk_adj", "cbm4_rk_adj_direct", "cbm4_rk_adj_rtol1e-5", "cbm4_rk_adj_rtol1e-4"):
        # the two looser-tolerance variants measure the general-pivoting second pass (systems that leave the static pivot
        # pattern are integrated again from t0 with dense factors): rtol 0.1f*1e-4 -> about half of the systems, 0.1f*1e-3 -> all
        scale = {"cbm4_rk_adj_rtol1e-5": 10.0, "cbm4_rk_adj_rtol1e-4": 100.0}.get(w, 1.0)
        rtol, atol = np.full(N, F01 * 1.0e-5 * scale), np.full(N, F01 * 1.0e-7 * scale)
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = 1
        if w == "cbm4_rk_adj_direct":
            ic[6] = 2                           # LSS_Decomp_Big / LSS_Solve_Big: one 96 x 96 factorisation per accepted step
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        run = lambda y: F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic, family="rk")
        cpu = lambda y: O.rk_adj("cbm4", y, 32, T0, T1, atol, rtol, ic)
        desc = ("EXAMPLES/cbm4 ADJ/RK_ADJ: Radau-2A (ICNTRL(3)=1), NADJ=32, NP=29, "
                + ("direct adjoint solve (ICNTRL(7)=2, 3N x 3N DGETRF per step)" if w == "cbm4_rk_adj_direct" else "fixed-sweep discrete adjoint")
                + f", 12h..84h, rtol 0.1f*{1.0e-5 * scale:g}, atol 0.1f*{1.0e-7 * scale:g} (cbm4_rk_adj_dr.F90"
                + (") (BASELINE config 4)" if scale == 1.0 else "; looser tolerances: general-pivoting second pass)"))
        ncpu = 32
        if w == "cbm4_rk_adj_direct":           # + one (2/3 (3n)^3 + (3n)^2) factorisation and NADJ x 2 (3n)^2 solves per accepted step
            flops = lambda r: (dense_flops_from_counters(r.istatus, 32, 573, 636, complex_pairs=True)
                               + float(r.istatus[:, 3].astype(np.float64).sum()) * (2.0 * 96 ** 3 / 3.0 + 96 ** 2 - 5.0 * (2.0 * 32 ** 3 / 3.0 + 32 ** 2)
                                                                                     + 32 * (2.0 * 96 ** 2 - 5.0 * 32 ** 2)))
        else:
            flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, complex_pairs=True, sens_cols=32)
    elif w == "cbm4_rk_tlm":
        rtol, atol = np.full(N, F01 * 1.0e-5), np.full(N, F01 * 1.0e-7)
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = 1
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        eye = np.eye(32)
        run = lambda y: F.integrate_tlm("cbm4", T0, T1, y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), rtol, atol, icntrl_u=ic, family="rk")
        cpu = lambda y: O.rk_tlm("cbm4", y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), T0, T1, atol, rtol, ic)
        desc = ("EXAMPLES/cbm4 TLM/RK_TLM: Radau-2A (ICNTRL(3)=1), NTLM=32, Y_tlm(t0)=I, preset direct TLM solve (3N x 3N DGETRF per accepted "
                "step), 12h..84h, rtol 0.1f*1e-5, atol 0.1f*1e-7")
        ncpu = 32
        # state sweep (real + complex factorisations) + per accepted step one (2/3 (3n)^3 + (3n)^2) factorisation, 3 Jacobians and
        # NTLM x (3 dense mat-vecs + 2 (3n)^2 solve)
        flops = lambda r: (float(r.istatus[:, 0].astype(np.float64).sum()) * 573 + float((r.istatus[:, 1] - 3 * r.istatus[:, 3]).astype(np.float64).sum()) * 636
                           + float((r.istatus[:, 5] - r.istatus[:, 3]).astype(np.float64).sum()) * 5.0 * (2.0 * 32 ** 3 / 3.0 + 32 ** 2)
                           + float((r.istatus[:, 6] - 32 * r.istatus[:, 3]).astype(np.float64).sum()) * 5.0 * 32 ** 2
                           + float(r.istatus[:, 3].astype(np.float64).sum()) * (2.0 * 96 ** 3 / 3.0 + 96 ** 2 + 3 * 636 + 32 * (3 * 2.0 * 32 ** 2 + 2.0 * 96 ** 2)))
    elif w == "cbm4_sdirk_adj":
        rtol, atol = tolerances()
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        run = lambda y: F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, family="sdirk")
        cpu = lambda y: O.sdirk_adj("cbm4", y, 32, T0, T1, atol, rtol)
        desc = ("EXAMPLES/cbm4 ADJ/SDIRK_ADJ: Sdirk4a, NADJ=32, NP=29, direct adjoint solve (presets), 12h..84h, driver tolerances "
                "of ROS_ADJ")
        ncpu = 32
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, sens_cols=32)
    elif w == "cbm4_sdirk_tlm":
        rtol, atol = tolerances()
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        eye = np.eye(32)
        run = lambda y: F.integrate_tlm("cbm4", T0, T1, y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), rtol, atol, family="sdirk")
        cpu = lambda y: O.sdirk_tlm("cbm4", y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), T0, T1, atol, rtol)
        desc = ("EXAMPLES/cbm4 TLM/SDIRK_TLM: Sdirk4a, NTLM=32, Y_tlm(t0)=I, modified-Newton TLM (presets), 12h..84h, "
                "driver tolerances of ROS_ADJ (BASELINE config 4)")
        ncpu = 32
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, sens_cols=32)
    elif w in ("cbm4_sdirk_tlm_direct", "cbm4_sdirk_tlm_trunc"):
        # the modes in which the columns feed back into the step sequence: lock-step kernel (sdirk_tlm_lock.cu), one system per warp
        rtol, atol = tolerances()
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        eye = np.eye(32)
        ic = np.zeros(20, dtype=np.int32)
        direct = w.endswith("direct")
        ic[6 if direct else 11] = 1
        tt = np.full((32, 32), 1.0e-5)     # unit-sized columns: ATOL_TLM = RTOL_TLM = 1e-5
        run = lambda y: F.integrate_tlm("cbm4", T0, T1, y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), rtol, atol, icntrl_u=ic,
                                        atol_tlm=tt, rtol_tlm=tt, family="sdirk")
        cpu = lambda y: O.sdirk_tlm("cbm4", y, np.broadcast_to(eye, (y.shape[0], 32, 32)).copy(), T0, T1, atol, rtol, ic, atol_tlm=tt, rtol_tlm=tt)
        desc = ("EXAMPLES/cbm4 TLM/SDIRK_TLM: Sdirk4a, NTLM=32, Y_tlm(t0)=I, " +
                ("direct TLM solve (ICNTRL(7)=1: a second factorisation per stage)" if direct else
                 "column truncation errors in the step control (ICNTRL(12)=1, ATOL_TLM=RTOL_TLM=1e-5)") +
                ", 12h..84h, driver tolerances of ROS_ADJ; state and columns in lock-step, general pivoting")
        ncpu = 32
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, sens_cols=32)
    elif w in ("swe_erk_tlm_dense", "swe_erk_adj"):
        from fatode_b200.inputs import swe_members
        dt = float(np.float32(2.1573758800627289e-003))
        tol = np.full(4800, 1.0e-4)
        t1 = 5 * dt
        y0 = swe_members(nc)
        e1 = np.zeros((1, 4800))
        e1[0, 0] = 1.0
        icd = np.zeros(20, dtype=np.int32)
        icd[4] = 1
        if w == "swe_erk_adj":
            run = lambda y: F.integrate_erk_adj("swe", 0.0, t1, y, 1, tol, tol)
            cpu = lambda y: O.erk_adj("swe", y, 1, 0.0, t1, tol, tol)
            desc = ("EXAMPLES/swe ADJ/ERK_ADJ: Dopri5, NADJ=1 (Lambda(1,1)=1), discrete adjoint with the example's JAC callback (compute_JacF) "
                    "applied matrix-free, t=0..5 dt, rtol=atol=1e-4, Gaussian bump a=30(1+0.1u) per member (swe2D_erk_adj_dr.F90)")
        else:
            run = lambda y: F.integrate_erk("swe", 0.0, t1, y, tol, tol, icntrl_u=icd, var_tlm=np.broadcast_to(e1, (y.shape[0], 1, 4800)).copy())
            cpu = lambda y: O.erk("swe", y, 0.0, t1, tol, tol, icd, y_tlm=np.broadcast_to(e1, (y.shape[0], 1, 4800)).copy())
            desc = ("EXAMPLES/swe TLM/ERK_TLM: Dopri5, NTLM=1, Y_tlm=e_1, dense-Jacobian mode (ICNTRL(5)=1: JAC + MATMUL per stage) with the "
                    "example's JAC callback applied matrix-free, t=0..5 dt, rtol=atol=1e-4, Gaussian bump a=30(1+0.1u) per member")
        ncpu = 128
        # per stage: FUN (~0.51 Mflop) + the 27 x 4800 entries of JAC (~3 Mflop as the formula export writes them) + one
        # 27-entry product per column (0.26 Mflop); the dense MATMUL / DGEMV of the reference would be 46 Mflop per product
        flops = lambda r: float(r.istatus[:, 0].astype(np.float64).sum()) * 0.51e6 + float(r.istatus[:, 1].astype(np.float64).sum()) * 3.3e6
    elif w in ("swe_erk", "swe_erk_tlm"):
        from fatode_b200.inputs import swe_members
        tlm = w == "swe_erk_tlm"
        dt = float(np.float32(2.1573758800627289e-003))
        tol = np.full(4800, 1.0e-4)
        t1 = (5 if tlm else 50) * dt
        y0 = swe_members(nc)
        e1 = np.zeros((1, 4800))
        e1[0, 0] = 1.0
        if tlm:
            run = lambda y: F.integrate_erk("swe", 0.0, t1, y, tol, tol, var_tlm=np.broadcast_to(e1, (y.shape[0], 1, 4800)).copy())
            cpu = lambda y: O.erk("swe", y, 0.0, t1, tol, tol, y_tlm=np.broadcast_to(e1, (y.shape[0], 1, 4800)).copy())
            desc = ("EXAMPLES/swe TLM/ERK_TLM: Dopri5, NTLM=1, Y_tlm=e_1, finite-difference Jacobian-vector products, t=0..5 dt, "
                    "rtol=atol=1e-4, Gaussian bump a=30(1+0.1u) per member (BASELINE config 5)")
        else:
            run = lambda y: F.integrate_erk("swe", 0.0, t1, y, tol, tol)
            cpu = lambda y: O.erk("swe", y, 0.0, t1, tol, tol)
            desc = ("EXAMPLES/swe FWD/ERK: Dopri5, 40x40 grid (N=4800), t=0..50 dt, rtol=atol=1e-4, Gaussian bump a=30(1+0.1u) per "
                    "member (BASELINE config 5)")
        ncpu = 128
        # ~0.51 Mflop per FUN (1600 grid cells x ~320 flop)
        flops = lambda r: float(r.istatus[:, 0].astype(np.float64).sum()) * 0.51e6
    elif w == "cbm4_rk":
        rtol, atol = np.full(N, F01 * 1.0e-5), np.full(N, F01 * 1.0e-7)
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = 1
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        run = lambda y: F.integrate_rk("cbm4", T0, T1, y, rtol, atol, icntrl_u=ic)
        cpu = lambda y: O.rk_fwd("cbm4", y, T0, T1, atol, rtol, ic)
        desc = "EXAMPLES/cbm4 FWD/RK: Radau-2A (ICNTRL(3)=1), 12h..84h, rtol 0.1f*1e-5, atol 0.1f*1e-7 (cbm4_rk_dr.F90)"
        ncpu = 256
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636, complex_pairs=True)
    else:
        rtol, atol = np.full(N, F01 * 1.0e-5), np.full(N, F01 * 1.0e-7)
        y0 = perturbed_states(F.initial_state("cbm4"), nc)
        run = lambda y: F.integrate_sdirk("cbm4", T0, T1, y, rtol, atol)
        cpu = lambda y: O.sdirk_fwd("cbm4", y, T0, T1, atol, rtol)
        desc = "EXAMPLES/cbm4 FWD/SDIRK: Sdirk4a, 12h..84h, rtol 0.1f*1e-5, atol 0.1f*1e-7 (cbm4_sdirk_dr.F90)"
        ncpu = 128
        flops = lambda r: dense_flops_from_counters(r.istatus, 32, 573, 636)
    ncpu = min(ncpu, nc)
    for _ in range(max(warmup, 1)):
        run(y0[: max(nc // 8, min(64, nc))])
    times, dev_ms = [], []
    for _ in range(steps):
        t = time.perf_counter()
        r = run(y0)
        times.append(time.perf_counter() - t)
        st = F.last_stats()
        dev_ms.append(st.ms_forward + st.ms_forward_tape + st.ms_adjoint)
        phases = {"forward_ms": float(st.ms_forward), "preparation_ms": float(st.ms_forward_tape), "sensitivity_ms": float(st.ms_adjoint),
                  "cells_general": int(st.cells_general)}
        assert (r.ierr == 1).all()
    t = time.perf_counter()
    cores = host_threads()
    ref = cpu(y0[:ncpu])
    tc = time.perf_counter() - t
    same = bool(np.array_equal(ref["y"], r.y[:ncpu]) and np.array_equal(ref["istatus"], r.istatus[:ncpu]))
    sec = float(np.mean(dev_ms)) * 1e-3
    line = {"metric": "cell-integrations/sec, " + w, "value": nc / sec, "unit": "cells/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": float(np.mean(dev_ms)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "cells_per_gpu_per_step": int(nc), "note": "side measurement, not the driver's metric; "
                       "value = CUDA-event kernel time, e2e = wall clock of the C-ABI call with pageable host buffers"},
            "e2e": {"value": nc / float(np.mean(times)), "unit": "cells/s"},
            "cpu_baseline": {"value": ncpu / tc, "unit": "cells/s", "cores": cores, "kind": "port",
                             "sample": f"{ncpu} cells, {tc:.1f} s, OpenMP over cells"},
            "bitwise_equal_to_oracle_on_sample": same, "phases_last_step": phases}
    if flops is not None:
        f = flops(r)
        roof = {"bound": "fp64", "achieved": f / sec / 1e12, "unit": "TFLOP/s", "peak": fp64_peak,
                "frac": (f / sec / 1e12 / fp64_peak) if fp64_peak else None,
                "note": "ALGORITHMIC flops (dense FULL_ALGEBRA costs of the reference, SURVEY.md 8d) from the returned ISTATUS "
                        "counters / summed CUDA-event kernel time; the kernels execute the sparse pattern"}
        line["roofline"] = roof
    if w.startswith("swe_"):
        nfun = float(r.istatus[:, 0].sum())
        line["unit"] = line["e2e"]["unit"] = line["cpu_baseline"]["unit"] = "members/s"
        line["fun_evaluations_per_s"] = nfun / sec
        line["config"]["layout"] = ("one member per CTA: Y and the stage vector in shared memory, K_1..K_S in an L2-resident per-CTA "
                                    "scratch; HBM traffic = Y in + Y out (76.8 kB per member)")
        if hbm_peak:
            line["roofline"]["hbm"] = {"achieved_GBs": nc * 76.8e3 / sec / 1e9, "peak_GBs": hbm_peak,
                                       "note": "mandatory traffic only (Y in + Y out): the stage vectors stay in L2"}
    return line


def run_side_workload(args):
    import fatode_b200 as F
    nc = args.cells if args.cells > 0 else SIDE_DEFAULT_CELLS[args.workload]
    peak = F.measure_fp64_peak(20000)
    print(json.dumps(side_measure(args.workload, nc, args.steps, args.warmup, peak, measured_hbm_peak())), flush=True)


def executed_counts():
    """profiles/executed_counts.json: executed FP64 work per unit from an ncu capture (tools/ncu_executed.py), valid only for
    the build it names -- compared by the SHA-256 of the library this process uses."""
    import hashlib
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "executed_counts.json")))
        h = hashlib.sha256(open(os.path.join(ROOT, "fatode_b200", "lib", "libfatode_b200.so"), "rb").read()).hexdigest()
        if d.get("library_sha256") != h:
            return None
        d["source"] = f"profiles/executed_counts.json (ncu capture {d.get('capture', '?')} of library sha256 {h[:16]})"
        return d
    except Exception:
        return None


def measured_hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs_burst", "hbm_GBs", "hbm_gbs", "hbm_copy_gbs"):
            if k in d:
                return float(d[k])
        for k, v in d.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=0,
                    help="cells per GPU per step (weak scaling) or in total (--scaling strong); 0 = default, see default_cells()")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --cells per GPU (default); strong: a fixed total (default 2^20) split over the GPUs")
    ap.add_argument("--impl", default="fatode_b200", choices=["fatode_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side block (other BASELINE configs, ~2 minutes, N=1 only)")
    ap.add_argument("--workload", default="cbm4_ros_adj", choices=["cbm4_ros_adj"] + sorted(SIDE_DEFAULT_CELLS),
                    help="cbm4_ros_adj = the headline metric (default); the others are side measurements of the other "
                         "BASELINE configs / entry points (single GPU, one JSON line each, not the driver's metric)")
    args = ap.parse_args()
    if args.workload != "cbm4_ros_adj":
        return run_side_workload(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = F.lib()
    dev = torch.device("cuda", local)
    # multi-GPU: the Mu sum over all cells is all-reduced INSIDE the library (fatode_comm_*, NCCL resolved by dlopen); torch's
    # process group only carries the 128-byte NCCL id and the benchmark's own barriers / timing reduction
    lib_comm = False
    if dist is not None:
        idbuf = (ctypes.c_ubyte * 128)()
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0 and L.fatode_comm_unique_id(idbuf) != 0:
            ok.zero_()
        idt = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        dist.broadcast(ok, 0)
        if int(ok.item()) == 1:
            idbuf = (ctypes.c_ubyte * 128)(*idt.cpu().tolist())
            rc_comm = L.fatode_comm_init(world, rank, idbuf)
            ok.fill_(1 if rc_comm == 0 else 0)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        lib_comm = int(ok.item()) == 1
        if not lib_comm:
            L.fatode_comm_finalize()
    nc = default_cells(args)
    rtol, atol = tolerances()
    ic, rc = icntrl_rodas4(), np.zeros(20)
    base = F.initial_state("cbm4")
    fp64_peak = F.measure_fp64_peak(20000) if rank == 0 else None

    # device-resident buffers (torch = allocator + stream plumbing only)
    y0_all = [torch.from_numpy(perturbed_states(base, nc, first_cell=(s * world + rank) * nc)).to(dev)
              for s in range(args.warmup + args.steps)]
    y = torch.empty((nc, N), dtype=torch.float64, device=dev)
    lam = torch.empty((nc, NADJ, N), dtype=torch.float64, device=dev)
    mu = torch.empty((nc, NADJ, NP), dtype=torch.float64, device=dev)
    ist = torch.empty((nc, 20), dtype=torch.int32, device=dev)
    rst = torch.empty((nc, 20), dtype=torch.float64, device=dev)
    ierr = torch.empty((nc,), dtype=torch.int32, device=dev)
    musum = torch.empty((NADJ, NP), dtype=torch.float64, device=dev)
    vp = ctypes.c_void_p
    np_ptr = lambda a: vp(a.ctypes.data)
    stream = torch.cuda.current_stream()

    def step_device(s):
        y.copy_(y0_all[s])
        r = L.fatode_ros_integrate_adj_batch(F.MODEL_CBM4, nc, N, NP, NADJ, vp(y.data_ptr()), vp(lam.data_ptr()), vp(mu.data_ptr()),
                                             T0, T1, None, None, np_ptr(atol), np_ptr(rtol), np_ptr(ic), np_ptr(rc),
                                             vp(ist.data_ptr()), vp(rst.data_ptr()), vp(ierr.data_ptr()), vp(musum.data_ptr()),
                                             None, F.MEM_DEVICE, vp(stream.cuda_stream))
        if r != 0:
            raise RuntimeError(L.fatode_last_error().decode())
        if dist is not None and not lib_comm:
            dist.all_reduce(musum)      # fall-back: the path's only exchange (29x32 Mu sums) through torch's communicator
        return F.last_stats()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for s in range(args.warmup):
        step_device(s)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    n_adj_launches = 0
    general_cells = 0
    ms_adj = ms_fwd = ms_tape = 0.0
    acc_steps = attempts = 0
    for s in range(args.warmup, args.warmup + args.steps):
        st = step_device(s)
        launches += st.launches + 1 + (1 if (dist is not None and not lib_comm) else 0)   # + the y0 copy (+ torch's NCCL kernel; the library counts its own)
        n_adj_launches += st.launches_adjoint
        general_cells += st.cells_general
        ms_adj += st.ms_adjoint
        ms_fwd += st.ms_forward
        ms_tape += st.ms_forward_tape
        acc_steps += st.accepted_steps
        ok = bool((ierr == 1).all().item())
        attempts += int((ist[:, 2].sum() - ist[:, 3].sum()).item())       # Nstp counts attempts + backward steps
        if not ok:
            raise RuntimeError("a cell failed (ierr != 1)")
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sync_all()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    cells_total = nc * world * args.steps
    value = cells_total / (ms * 1e-3)

    # ---- e2e: same call, pinned host buffers, H2D/D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        hy0 = [torch.from_numpy(perturbed_states(base, nc, first_cell=(s * world + rank) * nc)).pin_memory()
               for s in range(args.warmup, args.warmup + args.steps)]
        hy = torch.empty((nc, N), dtype=torch.float64).pin_memory()
        hlam = torch.empty((nc, NADJ, N), dtype=torch.float64).pin_memory()
        hmu = torch.empty((nc, NADJ, NP), dtype=torch.float64).pin_memory()
        hist = torch.empty((nc, 20), dtype=torch.int32).pin_memory()
        hrst = torch.empty((nc, 20), dtype=torch.float64).pin_memory()
        hierr = torch.empty((nc,), dtype=torch.int32).pin_memory()
        hmusum = torch.empty((NADJ, NP), dtype=torch.float64).pin_memory()

        def step_host(k):
            hy.copy_(hy0[k])
            r = L.fatode_ros_integrate_adj_batch(F.MODEL_CBM4, nc, N, NP, NADJ, vp(hy.data_ptr()), vp(hlam.data_ptr()),
                                                 vp(hmu.data_ptr()), T0, T1, None, None, np_ptr(atol), np_ptr(rtol), np_ptr(ic),
                                                 np_ptr(rc), vp(hist.data_ptr()), vp(hrst.data_ptr()), vp(hierr.data_ptr()),
                                                 vp(hmusum.data_ptr()), None, F.MEM_HOST, vp(stream.cuda_stream))
            if r != 0:
                raise RuntimeError(L.fatode_last_error().decode())
            if dist is not None and not lib_comm:
                g = hmusum.to(dev, non_blocking=True)
                dist.all_reduce(g)
                hmusum.copy_(g)
        step_host(0)
        sync_all()
        e0.record(stream)
        for k in range(args.steps):
            step_host(k)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1)
        t = torch.tensor([ms2], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
        h2d = nc * (N * 8 + NADJ * N * 8 + NADJ * NP * 8) + 2 * N * 8     # y0 + Lambda/Mu (INOUT in the reference API)
        d2h = nc * (N * 8 + NADJ * N * 8 + NADJ * NP * 8 + 20 * 4 + 20 * 8 + 4) + NADJ * NP * 8
        e2e = {"value": cells_total / (ms2 * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    # totals over ranks for the flop model
    tot = torch.tensor([acc_steps, attempts, ms_adj, ms_fwd, ms_tape, n_adj_launches, general_cells], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tot)
    if rank == 0:
        acc_t, att_t, ms_adj_t, ms_fwd_t, ms_tape_t, n_adj_t, general_t = [float(x) for x in tot.tolist()]
        f_adj = acc_t * flops_adjoint_per_step()
        f_fwd = flops_forward(att_t, acc_t)
        # dominant kernel: k_ros_adj_pair; per-rank kernel time = sum of its launches on that rank (ranks run concurrently)
        adj_s = ms_adj_t / world * 1e-3
        fwd_s = ms_fwd_t / world * 1e-3
        adj_tflops = f_adj / adj_s / 1e12
        peak = fp64_peak * world
        ex = executed_counts()       # ncu-measured executed FP64 work per unit, only if captured from THIS binary
        roof = {"bound": "fp64", "kernel": "k_ros_adj_pair<Cbm4Warp> (discrete adjoint sweep: solver + accumulator warp per cell, lane = adjoint column)",
                "achieved": adj_tflops, "peak": peak, "unit": "TFLOP/s", "frac": adj_tflops / peak,
                "achieved_is": "ALGORITHMIC flops (SURVEY.md 8d: the dense FULL_ALGEBRA count of the reference, evaluated on the "
                               "returned counters) / CUDA-event time of the kernel's launches in this run; the kernel executes "
                               "the 276/367/128-nnz sparse patterns, so this can exceed 1 -- see executed_*",
                "peak_source": "DFMA micro-benchmark run by this bench on this GPU (fatode_measure_fp64_peak); "
                               "MEASURED_PEAKS.json carries no FP64 figure",
                "algorithmic_flops_per_accepted_step": flops_adjoint_per_step(),
                "launches": int(n_adj_t), "avg_launch_ms": ms_adj_t / max(n_adj_t, 1.0),
                "share_of_step_time": (ms_adj_t / world) / ms,
                "traffic": (ex["adjoint"]["dram_bytes_per_accepted_step"] * acc_t / max(n_adj_t, 1.0)) if ex else None,
                "whole_job_algorithmic_tflops": (f_adj + f_fwd) / (ms * 1e-3) / 1e12,
                "whole_job_frac_of_fp64_peak": (f_adj + f_fwd) / (ms * 1e-3) / 1e12 / peak,
                "forward_ms_per_step": ms_fwd_t / (args.steps * world), "expand_ms_per_step": ms_tape_t / (args.steps * world),
                "adjoint_ms_per_step": ms_adj_t / (args.steps * world),
                "forward_kernel": {"kernel": "k_ros_fwd_cell_tm<Cbm4Cell> (forward sweep + tape, one cell per thread, two warps per SM)",
                                   "achieved": f_fwd / fwd_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                                   "frac": f_fwd / fwd_s / 1e12 / peak, "share_of_step_time": (ms_fwd_t / world) / ms},
                "cells_on_general_path": int(general_t)}
        if ex:
            e_adj = ex["adjoint"]["fp64_flops_per_accepted_step"] * acc_t
            e_fwd = ex["forward"]["fp64_flops_per_attempt"] * att_t
            roof["executed_tflops"] = e_adj / adj_s / 1e12
            roof["executed_frac"] = e_adj / adj_s / 1e12 / peak
            roof["forward_kernel"]["executed_tflops"] = e_fwd / fwd_s / 1e12
            roof["forward_kernel"]["executed_frac"] = e_fwd / fwd_s / 1e12 / peak
            roof["executed_source"] = ex["source"]
        else:
            roof["executed_source"] = ("not reported: profiles/executed_counts.json is absent or was captured from another build of "
                                       "libfatode_b200.so than the one this run loaded")
        line = {"metric": "cell-integrations/sec (fwd + adjoint), CBM-IV ROS_ADJ", "value": value, "unit": "cells/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(nc, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "allreduce": (None if dist is None else ("in-library NCCL (fatode_comm_*): mu_sum returned already summed over ranks"
                                                         if lib_comm else "torch.distributed all_reduce of mu_sum (library communicator unavailable)")),
                "roofline": roof,
                "accepted_steps_per_cell": acc_t / (nc * world * args.steps)}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        if not args.no_side and world == 1:
            # the other BASELINE configs, one timed call each (config 2: small_strato 2^20 cells; config 4: RK_ADJ and
            # SDIRK_TLM; config 5: the shallow-water ensemble), each with its algorithmic roofline fraction and a bit check
            # against the CPU oracle on a sample
            del y0_all, y, lam, mu
            torch.cuda.empty_cache()
            L.fatode_release_workspace()
            side = {}
            hbm = measured_hbm_peak()
            for w_, n_ in (("small_strato_fwd", 1 << 20), ("cbm4_rk_adj", 9472), ("cbm4_sdirk_tlm", 9472), ("swe_erk", 32768), ("swe_erk_adj", 8192)):
                t_ = time.perf_counter()
                try:
                    sl = side_measure(w_, n_, 1, 1, fp64_peak, hbm)
                    side[w_] = {k: sl[k] for k in ("value", "unit", "ms_per_step", "e2e", "cpu_baseline", "roofline",
                                                   "bitwise_equal_to_oracle_on_sample") if k in sl}
                    side[w_]["workload"] = sl["config"]["workload"]
                    side[w_]["cells"] = n_
                except Exception as exc:       # a side case must never cost the headline line
                    side[w_] = {"error": repr(exc)}
                side[w_]["wall_s"] = time.perf_counter() - t_
                L.fatode_release_workspace()
            line["side"] = side
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        if lib_comm:
            L.fatode_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
