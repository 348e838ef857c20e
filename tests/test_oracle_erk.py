"""CPU checks of the oracle's explicit Runge-Kutta restatement (oracle/erk.hpp: FWD/ERK/ERK_f90_Integrator.F90, TLM/ERK_TLM/
ERK_TLM_f90_Integrator.F90) and of the shallow-water model (EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90) against the reference's
golden file tests/golden/swe_lsode_sol.txt (= EXAMPLES/swe/SWE_ERK/swe_lsode_sol.txt, the LSODE solution at t = 50 dt that
the driver's `rerror` line is computed against, swe2D_erk_dr.F90:16-52)."""
import os

import numpy as np

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swe_lsode_sol.txt")
DT = float(np.float32(2.1573758800627289e-003))      # the driver's REAL(4) literal (swe2D_erk_dr.F90:106)


def rerr(y, gold):
    return np.sqrt(np.sum((y - gold) ** 2) / np.sum(gold ** 2))


def test_swe_erk_converges_to_the_golden_solution():
    gold = np.loadtxt(GOLD)
    y0 = O.swe_initial_state()
    assert y0.shape == (4800,) and abs(y0[:1600].max() - 130.0) < 0.5 and np.all(y0[1600:] == 2.0)
    errs = []
    for tol in (1e-2, 1e-4, 1e-6):                   # entries 1, 5, 9 of the driver's tolerance sweep (:66)
        r = O.erk("swe", y0, 0.0, 50 * DT, np.full(4800, tol), np.full(4800, tol))
        assert r["ierr"][0] == 1
        I = r["istatus"][0]
        assert I[0] == 7 * I[2] and I[1] == 0 and I[5] == 0 and I[6] == 0      # Dopri5: 7 FUN per attempt, no FSAL
        errs.append(rerr(r["y"][0], gold))
    assert errs[0] < 2e-2 and errs[1] < 5e-4 and errs[2] < 3e-6, errs
    assert errs[2] < errs[1] < errs[0]


def test_swe_erk_all_methods_agree():
    gold = np.loadtxt(GOLD)
    y0 = O.swe_initial_state()
    tol = np.full(4800, 1e-5)
    stages = {1: 3, 2: 4, 3: 5, 4: 7, 5: 8, 6: 12}
    for method, S in stages.items():
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        r = O.erk("swe", y0, 0.0, 50 * DT, tol, tol, ic)
        assert r["ierr"][0] == 1 and r["istatus"][0][0] == S * r["istatus"][0][2]
        assert rerr(r["y"][0], gold) < (2e-3 if method <= 3 else 2e-4), (method, rerr(r["y"][0], gold))


def test_swe_erk_tlm_matches_finite_differences_of_the_state():
    y0 = O.swe_initial_state()
    tol = np.full(4800, float(np.float32(0.1)) * float(np.float32(0.1)) ** 4)        # the TLM driver's 4th sweep value
    v = np.zeros((1, 1, 4800))
    v[0, 0, 0] = 1.0                                                                   # Y_tlm = e_1 (swe2D_erk_tlm_dr.F90:113-116)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 4
    r = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic, y_tlm=v)
    assert r["ierr"][0] == 1
    I = r["istatus"][0]
    assert I[0] == 7 * 2 * I[2]                                                        # one extra FUN per column per stage
    eps = 1e-4
    p = O.erk("swe", y0 + eps * v[0, 0], 0.0, 5 * DT, tol, tol, ic)["y"][0]
    m = O.erk("swe", y0 - eps * v[0, 0], 0.0, 5 * DT, tol, tol, ic)["y"][0]
    fd = (p - m) / (2 * eps)
    tl = r["y_tlm"][0, 0]
    assert np.max(np.abs(fd - tl)) < 2e-3 * np.max(np.abs(tl)), np.max(np.abs(fd - tl)) / np.max(np.abs(tl))
    # the state does not depend on the columns in this mode
    s = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic)
    assert np.array_equal(s["y"], r["y"]) and np.array_equal(s["istatus"][0][2:5], I[2:5])
