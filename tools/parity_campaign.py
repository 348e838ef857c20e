"""Wider GPU-vs-oracle parity sweep of the implicit-family sensitivity entry points and the explicit family (more systems and
tolerances than tests/ can afford): prints one line per case; exits non-zero on any mismatch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states, swe_members
F01 = float(np.float32(0.1)); T0 = 43200.0
bad = 0
gaps = 0


def sens_err(a, b):
    sc = np.max(np.abs(b), axis=2); er = np.max(np.abs(a - b), axis=2)
    fl = 1e-10 * np.max(sc, axis=1, keepdims=True)
    return float(np.max(er / np.maximum(sc, fl)))


def report(name, res, ref, sens=None, tol=1e-9):
    global bad
    handed_back = res.ierr == -20          # no family hands systems back any more; kept as a tripwire
    if handed_back.any():
        global gaps
        gaps += int(handed_back.sum())
        keep = ~handed_back
        print(f"{name}: {int(handed_back.sum())} of {len(res.ierr)} systems handed back with IERR -20 (KNOWN GAP); comparing the other {int(keep.sum())}")
        if not keep.any():
            return
        import types
        res = types.SimpleNamespace(ierr=res.ierr[keep], istatus=res.istatus[keep], rstatus=res.rstatus[keep], y=res.y[keep])
        ref = {k2: (v[keep] if isinstance(v, np.ndarray) and len(v) == len(keep) else v) for k2, v in ref.items()}
        sens = [(l, a[keep], b[keep]) for (l, a, b) in (sens or [])]
    ok = (np.array_equal(res.ierr, ref["ierr"]) and np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.rstatus, ref["rstatus"])
          and np.array_equal(res.y, ref["y"]))
    msg = f"{name}: state/ISTATUS/RSTATUS bitwise {ok}; ierr codes {dict(zip(*np.unique(res.ierr, return_counts=True)))}"
    for label, a, b in (sens or []):
        e = sens_err(a, b)
        msg += f"; {label} max col err {e:.2e} bitwise {np.array_equal(a, b)}"
        ok = ok and e <= tol
    print(msg, flush=True)
    if not ok:
        bad += 1


nc = int(sys.argv[1]) if len(sys.argv) > 1 else 192
first = 1000
for k in (1e-2, 1e-3, 1e-4, 1e-5):
    rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
    y0 = perturbed_states(F.initial_state("cbm4"), nc, first_cell=first)
    first += nc
    t1 = T0 + 30 * 3600.0
    yt = np.tile(np.eye(32), (nc, 1, 1))
    r = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, family="sdirk")
    o = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol)
    report(f"SDIRK_TLM rtol {F01*k:.0e} n {nc}", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])])
    nl = max(16, nc // 4)                                  # lock-step modes: direct solve, column Newton / truncation control
    tt = np.full((32, 32), 10.0 * F01 * k)
    for d7, e9, t12 in ((1, 0, 0), (0, 1, 1), (1, 0, 1)):
        ic = np.zeros(20, dtype=np.int32); ic[6], ic[8], ic[11] = d7, e9, t12
        r = F.integrate_tlm("cbm4", T0, t1, y0[:nl], yt[:nl], rtol, atol, icntrl_u=ic, atol_tlm=tt, rtol_tlm=tt, family="sdirk")
        o = O.sdirk_tlm("cbm4", y0[:nl], yt[:nl], T0, t1, atol, rtol, ic, atol_tlm=tt, rtol_tlm=tt)
        report(f"SDIRK_TLM lock-step ICNTRL(7,9,12)=({d7},{e9},{t12}) rtol {F01*k:.0e} n {nl}", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])])
    r = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, family="sdirk")
    o = O.sdirk_adj("cbm4", y0, 32, T0, t1, atol, rtol)
    report(f"SDIRK_ADJ rtol {F01*k:.0e} n {nc}", r, o, [("Lambda", r.lambda_, o["lambda_"]), ("Mu", r.mu, o["mu"])])
for k in (1e-2, 1e-3, 1e-4, 1e-5):   # Runge-Kutta big-system modes: TLM/RK_TLM (preset direct solve) and RK_ADJ with ICNTRL(7) = 2
    rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
    nb = max(16, nc // 4)
    y0 = perturbed_states(F.initial_state("cbm4"), nb, first_cell=first)
    first += nb
    t1 = T0 + 20 * 3600.0
    ic = np.zeros(20, dtype=np.int32); ic[2] = 1
    yt = np.tile(np.eye(32), (nb, 1, 1))
    r = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, family="rk")
    o = O.rk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    report(f"RK_TLM rtol {F01*k:.0e} n {nb} (general path cells {F.last_stats().cells_general})", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])])
    ic[6] = 2
    r = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic, family="rk")
    o = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic)
    report(f"RK_ADJ direct rtol {F01*k:.0e} n {nb} (general path cells {F.last_stats().cells_general})", r, o, [("Lambda", r.lambda_, o["lambda_"]), ("Mu", r.mu, o["mu"])], 1e-7)
for k in (1e-2, 1e-3, 1e-4):
    rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
    y0 = perturbed_states(F.initial_state("cbm4"), nc, first_cell=first)
    first += nc
    t1 = T0 + 30 * 3600.0
    ic = np.zeros(20, dtype=np.int32); ic[2] = 5
    r = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic)
    o = O.ros_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic)
    loose = 1e-9 if k <= 1e-5 else 1e-6   # the 1e-9 contract is stated at RTOL = 1e-6; big steps amplify rounding differences
    report(f"ROS_ADJ rtol {F01*k:.0e} n {nc} (general path cells {F.last_stats().cells_general})", r, o, [("Lambda", r.lambda_, o["lambda_"]), ("Mu", r.mu, o["mu"])], loose)
    yt = np.tile(np.eye(32), (nc, 1, 1))
    r = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic)
    o = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    report(f"ROS_TLM rtol {F01*k:.0e} n {nc} (general pass cells {F.last_stats().cells_general})", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])], loose)
    # the reference's default call: ICNTRL_U absent -> ICNTRL(12) = 1, columns in the error control (lock-step kernel, bit for bit)
    at, rt = np.full((8, 32), 1e-6), np.full((8, 32), 10 * F01 * k)
    r = F.integrate_tlm("cbm4", T0, t1, y0[:32], yt[:32, :8], rtol, atol, atol_tlm=at, rtol_tlm=rt)
    o = O.ros_tlm("cbm4", y0[:32], yt[:32, :8], T0, t1, atol, rtol, None, atol_tlm=at, rtol_tlm=rt)
    report(f"ROS_TLM preset (ICNTRL(12)=1) rtol {F01*k:.0e} n 32 ntlm 8 bitwise Y_tlm {np.array_equal(r.y_tlm, o['y_tlm'])}", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])], 1e-12)
    ic = np.zeros(20, dtype=np.int32); ic[2] = 1
    r = F.integrate_rk("cbm4", T0, t1, y0, rtol, atol, icntrl_u=ic)
    o = O.rk_fwd("cbm4", y0, T0, t1, atol, rtol, ic)
    report(f"FWD/RK rtol {F01*k:.0e} n {nc} (general pass cells {F.last_stats().cells_general})", r, o)
    r = F.integrate_sdirk("cbm4", T0, t1, y0, rtol, atol)
    o = O.sdirk_fwd("cbm4", y0, T0, t1, atol, rtol)
    report(f"FWD/SDIRK rtol {F01*k:.0e} n {nc} (general pass cells {F.last_stats().cells_general})", r, o)
    r = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic, family="rk")
    o = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic)
    report(f"RK_ADJ rtol {F01*k:.0e} n {nc} (general pass cells {F.last_stats().cells_general})", r, o, [("Lambda", r.lambda_, o["lambda_"]), ("Mu", r.mu, o["mu"])], loose)
DT = float(np.float32(2.1573758800627289e-003))
for tol, method in ((1e-2, 0), (1e-3, 5), (1e-5, 0), (1e-6, 6)):
    ic = np.zeros(20, dtype=np.int32); ic[2] = method
    y0 = swe_members(48, first_member=first); first += 48
    tv = np.full(4800, tol)
    r = F.integrate_erk("swe", 0.0, 50 * DT, y0, tv, tv, icntrl_u=ic)
    o = O.erk("swe", y0, 0.0, 50 * DT, tv, tv, ic)
    report(f"ERK method {method} tol {tol:.0e} n 48 (attempts/member {r.istatus[:,2].mean():.1f})", r, o)
    v = np.random.default_rng(int(1e6 * tol) + method).standard_normal((48, 2, 4800))
    r = F.integrate_erk("swe", 0.0, 5 * DT, y0, tv, tv, icntrl_u=ic, var_tlm=v)
    o = O.erk("swe", y0, 0.0, 5 * DT, tv, tv, ic, y_tlm=v)
    report(f"ERK_TLM method {method} tol {tol:.0e} n 48 ntlm 2", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])])
    icd = ic.copy(); icd[4] = 1                                # dense-Jacobian mode: MATMUL(JAC, S_tlm) with the example's compute_JacF
    r = F.integrate_erk("swe", 0.0, 5 * DT, y0, tv, tv, icntrl_u=icd, var_tlm=v)
    o = O.erk("swe", y0, 0.0, 5 * DT, tv, tv, icd, y_tlm=v)
    report(f"ERK_TLM dense-Jacobian method {method} tol {tol:.0e} n 48 ntlm 2", r, o, [("Y_tlm", r.y_tlm, o["y_tlm"])])
    r = F.integrate_erk_adj("swe", 0.0, 20 * DT, y0, 2, tv, tv, icntrl_u=ic)
    o = O.erk_adj("swe", y0, 2, 0.0, 20 * DT, tv, tv, ic)
    report(f"ERK_ADJ method {method} tol {tol:.0e} n 48 nadj 2 (steps/member {r.istatus[:,3].mean():.1f})", r, o, [("Lambda", r.lambda_, o["lambda_"])])
print("MISMATCHES", bad, "| systems handed back (known gap):", gaps)
sys.exit(1 if bad else 0)
