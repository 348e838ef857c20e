"""CPU checks of the oracle's SDIRK_TLM restatement (oracle/sdirk_tlm.hpp, TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:31-960).
The reference ships no CBM-IV TLM driver and no golden file for this family ("parity unpinned" by reference artefacts,
SURVEY.md §8c), so the restatement is cross-validated: (i) against the ROS_TLM restatement (itself validated through the
adjoint-TLM duality against the golden-pinned adjoint oracle) — two different discretisations of the same sensitivity
equation must agree to the integration tolerance and converge towards each other as it is tightened; (ii) against finite
differences of its own state sweep; (iii) linearity in Y_tlm(t0); (iv) the ISTATUS closed forms of SURVEY.md §8a."""
import numpy as np

import oracle_lib as O

F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0


def col_err(a, b):
    sc = np.max(np.abs(b), axis=-1)
    return np.max(np.max(np.abs(a - b), axis=-1) / np.maximum(sc, 1e-12 * sc.max()))


def test_sdirk_tlm_agrees_with_ros_tlm_and_converges():
    y0 = O.initial_state("cbm4")
    t1 = T0 + 6 * 3600.0
    errs = []
    for k in (1e-4, 1e-6):
        rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
        ic = np.zeros(20, dtype=np.int32)
        sd = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
        ic[2] = 5
        ro = O.ros_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
        assert sd["ierr"][0] == 1 and ro["ierr"][0] == 1
        errs.append(col_err(sd["y_tlm"][0], ro["y_tlm"][0]))
    assert errs[0] < 5e-3 and errs[1] < 5e-5 and errs[1] < 0.1 * errs[0], errs


def test_sdirk_tlm_matches_finite_differences_and_is_linear():
    rtol, atol = np.full(32, F01 * 1e-6), np.full(32, F01 * 1e-8)
    y0 = O.initial_state("cbm4") * (1.0 + 0.05 * np.cos(np.arange(32.0)))
    t1 = T0 + 2 * 3600.0
    rng = np.random.default_rng(3)
    v = rng.standard_normal((2, 32)) * y0
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 4
    base = O.sdirk_tlm("cbm4", y0, v[None], T0, t1, atol, rtol, ic)
    assert base["ierr"][0] == 1
    eps = 1e-5
    for j in range(2):
        p = O.sdirk_tlm("cbm4", y0 + eps * v[j], v[None], T0, t1, atol, rtol, ic)["y"][0]
        m = O.sdirk_tlm("cbm4", y0 - eps * v[j], v[None], T0, t1, atol, rtol, ic)["y"][0]
        fd = (p - m) / (2 * eps)
        tl = base["y_tlm"][0][j]
        assert np.max(np.abs(fd - tl)) < 2e-4 * np.max(np.abs(tl)), np.max(np.abs(fd - tl)) / np.max(np.abs(tl))
    comb = 0.7 * v[0] - 1.9 * v[1]
    c = O.sdirk_tlm("cbm4", y0, comb[None, None], T0, t1, atol, rtol, ic)
    lin = 0.7 * base["y_tlm"][0][0] - 1.9 * base["y_tlm"][0][1]
    assert np.array_equal(c["y"], base["y"]) and np.array_equal(c["istatus"][0][:6], base["istatus"][0][:6])
    assert np.max(np.abs(c["y_tlm"][0][0] - lin)) < 1e-9 * np.max(np.abs(lin))


def test_sdirk_tlm_counters_all_methods():
    """Ndec = Njac-decomps + re-decomps; Njac = (decomps with a new Jacobian) + (completed stages); the state does not depend on NTLM."""
    rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
    y0 = O.initial_state("cbm4")
    t1 = T0 + 3 * 3600.0
    for method in (1, 2, 3, 4, 5, 9):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        a = O.sdirk_tlm("cbm4", y0, np.eye(32)[None, :3], T0, t1, atol, rtol, ic)
        b = O.sdirk_tlm("cbm4", y0, np.eye(32)[None, :7], T0, t1, atol, rtol, ic)
        assert a["ierr"][0] == 1 and b["ierr"][0] == 1
        assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["y_tlm"][0], b["y_tlm"][0][:3])
        Ia, Ib = a["istatus"][0].astype(np.int64), b["istatus"][0].astype(np.int64)
        assert np.array_equal(Ia[[0, 1, 2, 3, 4, 5, 7]], Ib[[0, 1, 2, 3, 4, 5, 7]])
        # Nsol = It + completed attempts + NTLM * (TLM iterations): the NTLM-dependent part is linear in NTLM
        per_col = (Ib[6] - Ia[6]) // 4
        assert (Ib[6] - Ia[6]) == 4 * per_col and Ia[6] - 3 * per_col == Ia[0] + Ia[2]
    # an ICNTRL(3) outside 0..5 selects Sdirk2a in this family (:262-275)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 1
    c = O.sdirk_tlm("cbm4", y0, np.eye(32)[None, :3], T0, t1, atol, rtol, ic)
    assert np.array_equal(c["y"], a["y"]) and np.array_equal(c["y_tlm"], a["y_tlm"])


def test_sdirk_tlm_direct_solve_and_controlled_modes():
    """ICNTRL(7) = 1 (:659-705): the same state sweep, one more factorisation and NTLM solves per stage, Y_tlm equal to the
    modified-Newton columns within their Newton tolerance and closer to finite differences; ICNTRL(9) / ICNTRL(12) change
    the step sequence only through the columns."""
    y0 = O.initial_state("cbm4")
    t1 = T0 + 6 * 3600.0
    rtol, atol = np.full(32, F01 * 1e-6), np.full(32, F01 * 1e-8)
    for method in (0, 2, 5):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        a = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
        ic[6] = 1
        b = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
        assert a["ierr"][0] == 1 and b["ierr"][0] == 1
        assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["rstatus"], b["rstatus"])
        Ia, Ib = a["istatus"][0].astype(np.int64), b["istatus"][0].astype(np.int64)
        assert np.array_equal(Ia[[0, 1, 2, 3, 4, 7]], Ib[[0, 1, 2, 3, 4, 7]])
        stages = Ib[5] - Ia[5]                       # one LSS_Decomp_TLM per completed stage in the direct flavour
        assert stages >= (Ia[1] - Ia[5]) and stages <= Ia[1] - Ia[3]   # one LSS_Jac1 each; the rest of Njac: PrepareMatrix
        assert Ib[6] == Ia[0] + Ia[2] + 32 * stages  # Nsol: state iterations + error solves + one solve per column and stage
        assert col_err(b["y_tlm"][0], a["y_tlm"][0]) < 1e-7
        ic[6] = 2                                    # only the value 1 selects the direct solve (:332)
        c = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
        assert np.array_equal(c["y_tlm"], a["y_tlm"]) and np.array_equal(c["istatus"], a["istatus"])
    # columns in the error control: unit columns, tolerances in their scale
    rtol, atol = np.full(32, F01 * 1e-3), np.full(32, F01 * 1e-5)
    at, rt = np.full((32, 32), 1e-4), np.full((32, 32), 1e-4)
    ic = np.zeros(20, dtype=np.int32)
    base = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
    ic[11] = 1
    tr = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic, atol_tlm=at, rtol_tlm=rt)
    assert tr["ierr"][0] == 1 and tr["istatus"][0][3] > base["istatus"][0][3]     # smaller steps for the columns' sake
    fine = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol * 1e-4, rtol * 1e-4, np.zeros(20, dtype=np.int32))
    assert col_err(tr["y_tlm"][0], fine["y_tlm"][0]) < col_err(base["y_tlm"][0], fine["y_tlm"][0])
