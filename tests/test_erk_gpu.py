"""GPU parity for the explicit Runge-Kutta family on the shallow-water ensemble: FWD/ERK INTEGRATE and
TLM/ERK_TLM INTEGRATE_TLM (finite-difference mode).  The kernel performs the reference's floating-point operations in the
reference's order, so state, ISTATUS, RSTATUS and Y_tlm are compared bit for bit with the CPU oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
DT = float(np.float32(2.1573758800627289e-003))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swe_lsode_sol.txt")


def check(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    if ref.get("y_tlm") is not None:
        assert np.array_equal(res.y_tlm, ref["y_tlm"])


def test_swe_model_and_initial_state_match_the_oracle():
    import fatode_b200 as F
    assert F.model_dims("swe") == (4800, 0)
    assert np.array_equal(F.initial_state("swe"), O.swe_initial_state())
    assert np.array_equal(F.swe_initial_state(27.5, 1.0, -0.5), O.swe_initial_state(27.5, 1.0, -0.5))


def test_erk_ensemble_matches_oracle_and_golden():
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(24)
    tol = np.full(4800, 1e-4)                                # 5th entry of the driver's sweep
    res = F.integrate_erk("swe", 0.0, 50 * DT, y0, tol, tol)
    ref = O.erk("swe", y0, 0.0, 50 * DT, tol, tol)
    assert (res.ierr == 1).all()
    check(res, ref)
    I = res.istatus.astype(np.int64)
    assert np.all(I[:, 0] == 7 * I[:, 2]) and np.all(I[:, 2] == I[:, 3] + I[:, 4] + (I[:, 2] - I[:, 3] - I[:, 4]))
    assert np.all(I[:, [1, 5, 6, 7]] == 0)
    gold = np.loadtxt(GOLD)                                   # member 0 is the driver's problem
    assert np.sqrt(np.sum((res.y[0] - gold) ** 2) / np.sum(gold ** 2)) < 5e-4


def test_erk_methods_options_and_failures():
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(5)
    tol = np.full(4800, 1e-3)
    for method, scalar in ((1, 0), (2, 1), (3, 0), (5, 0), (6, 1), (9, 0)):
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[2] = scalar, method
        res = F.integrate_erk("swe", 0.0, 20 * DT, y0, tol, tol, icntrl_u=ic)
        ref = O.erk("swe", y0, 0.0, 20 * DT, tol, tol, ic)
        check(res, ref)
    rc = np.zeros(20)
    rc[1], rc[2], rc[3], rc[4], rc[6] = 4 * DT, 1e-4, 0.5, 3.0, 0.8       # Hmax, Hstart, FacMin, FacMax, FacSafe
    check(F.integrate_erk("swe", 0.0, 20 * DT, y0, tol, tol, rcntrl_u=rc), O.erk("swe", y0, 0.0, 20 * DT, tol, tol, None, rc))
    ic = np.zeros(20, dtype=np.int32)
    ic[3] = 6                                                 # Max_no_steps: IERR -6 with the state of the last accepted step
    res = F.integrate_erk("swe", 0.0, 50 * DT, y0, tol, tol, icntrl_u=ic)
    ref = O.erk("swe", y0, 0.0, 50 * DT, tol, tol, ic)
    assert (ref["ierr"] == -6).all()
    check(res, ref)
    bad = tol.copy()
    bad[17] = 0.0                                             # improper tolerance: IERR -5 before integrating, Y untouched
    res = F.integrate_erk("swe", 0.0, 20 * DT, y0, tol, bad)
    assert (res.ierr == -5).all() and np.array_equal(res.y, y0)
    back = F.integrate_erk("swe", 20 * DT, 0.0, y0[:2], tol, tol)          # backwards in time
    check(back, O.erk("swe", y0[:2], 20 * DT, 0.0, tol, tol))


def test_erk_tlm_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(6)
    tol = np.full(4800, float(np.float32(0.1)) * float(np.float32(0.1)) ** 3)
    v = np.zeros((6, 1, 4800))
    v[:, 0, 0] = 1.0                                          # NTLM = 1, Y_tlm = e_1 (swe2D_erk_tlm_dr.F90:113-116)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 4
    res = F.integrate_erk("swe", 0.0, 5 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v)
    ref = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic, y_tlm=v)
    assert (res.ierr == 1).all()
    check(res, ref)
    assert np.all(res.istatus[:, 0] == 14 * res.istatus[:, 2])
    # several columns (the reference's FDJACV leaves the earlier columns' perturbations in the stage vector), other method,
    # user-chosen increment
    rng = np.random.default_rng(5)
    v3 = rng.standard_normal((6, 3, 4800))
    ic[2] = 3
    rc = np.zeros(20)
    rc[11] = 1e-6
    res = F.integrate_erk("swe", 0.0, 5 * DT, y0, tol, tol, icntrl_u=ic, rcntrl_u=rc, var_tlm=v3)
    ref = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic, rc, y_tlm=v3)
    check(res, ref)


def test_erk_tlm_dense_jacobian_mode_matches_oracle():
    """ICNTRL(5) /= 0 (TLM/ERK_TLM :460-474): K_tlm = MATMUL(JAC, S_tlm) with the example's JAC callback (compute_JacF), applied
    matrix-free on the GPU; the state sweep is bit-identical, Y_tlm within 1e-9 (row sums ordered by column except at the
    periodic seams, pow(x, 1.5) from CUDA's libm); the columns are the exact duals of ERK_ADJ's Lambda."""
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(6)
    tol = np.full(4800, 1.0e-4)
    rng = np.random.default_rng(7)
    v3 = rng.standard_normal((6, 3, 4800))
    for method in (0, 1, 5):
        ic = np.zeros(20, dtype=np.int32)
        ic[2], ic[4] = method, 1
        res = F.integrate_erk("swe", 0.0, 5 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v3)
        ref = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic, y_tlm=v3)
        assert (ref["ierr"] == 1).all() and np.array_equal(res.ierr, ref["ierr"])
        assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.rstatus, ref["rstatus"]) and np.array_equal(res.y, ref["y"])
        assert np.all(res.istatus[:, 1] == res.istatus[:, 0])          # one JAC per FUN, no extra FUN per column
        sc = np.max(np.abs(ref["y_tlm"]), axis=2)
        er = np.max(np.abs(res.y_tlm - ref["y_tlm"]), axis=2)
        assert np.all(er <= 1e-9 * sc), np.max(er / sc)
    adj = F.integrate_erk_adj("swe", 0.0, 5 * DT, y0, 2, tol, tol, icntrl_u=ic)
    for k in range(2):
        d = np.einsum("mn,mcn->mc", adj.lambda_[:, k, :], v3)
        assert np.all(np.abs(d - res.y_tlm[:, :, k]) <= 1e-11 * np.max(np.abs(res.y_tlm[:, :, k])))
    # with the columns in the step control (ICNTRL(12)): the step sequence follows the columns' error norms
    ic = np.zeros(20, dtype=np.int32)
    ic[4], ic[11] = 1, 1
    tt = np.full((3, 4800), 1.0e-4)
    res = F.integrate_erk("swe", 0.0, 5 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v3, atol_tlm=tt, rtol_tlm=tt)
    ref = O.erk("swe", y0, 0.0, 5 * DT, tol, tol, ic, y_tlm=v3, atol_tlm=tt, rtol_tlm=tt)
    assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.ierr, ref["ierr"])
    assert np.max(np.abs(res.y - ref["y"])) <= 1e-9 * np.max(np.abs(ref["y"]))
    sc = np.max(np.abs(ref["y_tlm"]), axis=2)
    assert np.all(np.max(np.abs(res.y_tlm - ref["y_tlm"]), axis=2) <= 1e-9 * sc)


def test_erk_tlm_truncation_error_control():
    """ICNTRL(12) /= 0: the columns' error norms (scaled with ATOL_TLM / RTOL_TLM and the error vector itself) enter the step
    control, so the step sequence changes with the columns; state, counters and Y_tlm still match the oracle bit for bit."""
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(5)
    tol = np.full(4800, 1e-3)
    rng = np.random.default_rng(9)
    v = 50.0 * rng.standard_normal((5, 2, 4800))
    at, rt = np.full((2, 4800), 1e-5), np.full((2, 4800), 1e-4)
    for method, scalar in ((0, 0), (3, 1)):
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[2], ic[11] = scalar, method, 1
        res = F.integrate_erk("swe", 0.0, 10 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v, atol_tlm=at, rtol_tlm=rt)
        ref = O.erk("swe", y0, 0.0, 10 * DT, tol, tol, ic, y_tlm=v, atol_tlm=at, rtol_tlm=rt)
        check(res, ref)
        ic[11] = 0
        free = F.integrate_erk("swe", 0.0, 10 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v)
        assert (res.istatus[:, 2] > free.istatus[:, 2]).all()      # the columns' error does bind here
    with pytest.raises(F.FatodeError):                        # ICNTRL(12) without the TLM tolerances
        ic = np.zeros(20, dtype=np.int32)
        ic[11] = 1
        F.integrate_erk("swe", 0.0, 10 * DT, y0, tol, tol, icntrl_u=ic, var_tlm=v)


def test_swe_is_rejected_by_the_implicit_families_and_erk_rejects_the_chemistry_models():
    import fatode_b200 as F
    y0 = F.initial_state("swe")
    tol = np.full(4800, 1e-3)
    with pytest.raises(F.FatodeError):
        F.integrate("swe", 0.0, DT, y0, tol, tol)
    with pytest.raises(F.FatodeError):
        F.integrate_sdirk("swe", 0.0, DT, y0, tol, tol)
    with pytest.raises(F.FatodeError):
        F.integrate_erk("cbm4", 0.0, 60.0, F.initial_state("cbm4"), np.full(32, 1e-3), np.full(32, 1e-3))


def test_erk_full_size_ensemble_properties():
    """4096 members: every member succeeds, the counters obey the closed forms, total height is conserved by the periodic
    upwind scheme to round-off, and a second run reproduces the first bit for bit."""
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    n = 4096
    y0 = swe_members(n)
    tol = np.full(4800, 1e-4)
    a = F.integrate_erk("swe", 0.0, 50 * DT, y0, tol, tol)
    assert (a.ierr == 1).all()
    I = a.istatus.astype(np.int64)
    assert np.all(I[:, 0] == 7 * I[:, 2]) and np.all(I[:, 3] >= 10) and np.all(I[:, 3] + I[:, 4] <= I[:, 2])
    assert np.all(a.rstatus[:, 0] == 50 * DT)
    b = F.integrate_erk("swe", 0.0, 50 * DT, y0, tol, tol)
    assert np.array_equal(a.y, b.y) and np.array_equal(a.istatus, b.istatus)
    sub = np.arange(0, n, 517)
    ref = O.erk("swe", y0[sub], 0.0, 50 * DT, tol, tol)
    assert np.array_equal(a.y[sub], ref["y"]) and np.array_equal(a.istatus[sub], ref["istatus"])
