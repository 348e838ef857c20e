"""GPU parity for INTEGRATE_TLM of TLM/SDIRK_TLM on CBM-IV in the reference's preset mode (modified
Newton TLM, columns outside the step control): state, ISTATUS and RSTATUS bit-identical to the CPU oracle; the column sweep
performs the reference's operations in the reference's order, so Y_tlm is compared within 1e-9 per column (the contract)
and is expected to be bit-identical."""
import ctypes

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
RTOL_PARITY = 1e-9
NOISE_FLOOR = 1e-10


def tols(k=1.0e-4):
    return np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)


def check(res, ref, bitwise=True):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    a, b = res.y_tlm, ref["y_tlm"]
    sc = np.max(np.abs(b), axis=2)
    er = np.max(np.abs(a - b), axis=2)
    floor = NOISE_FLOOR * np.max(sc, axis=1, keepdims=True)
    assert np.all(er <= RTOL_PARITY * np.maximum(sc, floor)), np.max(er / np.maximum(sc, floor))
    if bitwise:
        assert np.array_equal(a, b), "the column sweep follows the reference's operation order"


def test_sdirk_tlm_batch_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 40)
    yt = np.tile(np.eye(32), (40, 1, 1))                # NTLM = N = 32, Y_tlm(t0) = I
    t1 = T0 + 24 * 3600.0
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, family="sdirk")       # default Sdirk4a
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol)
    assert (res.ierr == 1).all()
    check(res, ref)
    st = F.last_stats()
    assert st.accepted_steps == int(res.istatus[:, 3].sum())
    # consistency with the other TLM family (different discretisation of the same sensitivity equation)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    ro = F.integrate_tlm("cbm4", T0, t1, y0[:4], yt[:4], rtol, atol, icntrl_u=ic)
    sc = np.max(np.abs(ro.y_tlm), axis=2)
    er = np.max(np.abs(ro.y_tlm - res.y_tlm[:4]), axis=2) / np.maximum(sc, 1e-9 * sc.max(axis=1, keepdims=True))
    assert er.max() < 2e-2, er.max()


def test_sdirk_tlm_methods_fewer_columns_scalar_tolerances_and_options():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 7)
    rng = np.random.default_rng(11)
    yt = rng.standard_normal((7, 5, 32))
    t1 = T0 + 8 * 3600.0
    for method, scalar, maxit, zero_start in ((1, 0, 0, 0), (2, 0, 0, 0), (3, 1, 0, 0), (4, 0, 5, 1), (5, 0, 0, 0), (9, 0, 0, 0)):
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[2], ic[4], ic[5] = scalar, method, maxit, zero_start
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, family="sdirk")
        ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
        check(res, ref)
    rc = np.zeros(20)
    rc[2], rc[1] = 30.0, 600.0                           # Hstart, Hmax
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, rcntrl_u=rc, family="sdirk")
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, None, rc)
    check(res, ref)


def test_sdirk_tlm_failed_cells_keep_the_columns_of_their_last_accepted_step():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 5)
    yt = np.tile(np.eye(32)[:3], (5, 1, 1))
    ic = np.zeros(20, dtype=np.int32)
    ic[3] = 25                                            # Max_no_steps: every cell ends with IERR -6 after some accepted steps
    t1 = T0 + 24 * 3600.0
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, family="sdirk")
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    assert (ref["ierr"] == -6).all() and (ref["istatus"][:, 3] > 0).all()
    check(res, ref)


def test_sdirk_tlm_tape_slot_overflow_takes_the_second_pass():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 12)
    yt = np.tile(np.eye(32)[:4], (12, 1, 1))
    t1 = T0 + 10 * 3600.0
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol)
    L = F.lib()
    try:
        L.fatode_set_sdirk_tape_slot(int(np.median(ref["istatus"][:, 3])))     # about half of the systems overflow their slot
        L.fatode_set_wave_cells(5)
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, family="sdirk")
    finally:
        L.fatode_set_sdirk_tape_slot(0)
        L.fatode_set_wave_cells(0)
    check(res, ref)
    assert F.last_stats().waves > 3


def test_sdirk_tlm_option_errors():
    import fatode_b200 as F
    rtol, atol = tols()
    y0 = F.initial_state("cbm4")
    for k in (8, 11):                                     # ICNTRL(9), ICNTRL(12) read ATOL_TLM / RTOL_TLM
        ic = np.zeros(20, dtype=np.int32)
        ic[k] = 1
        with pytest.raises(F.FatodeError):
            F.integrate_tlm("cbm4", T0, T0 + 3600.0, y0, np.eye(32)[None], rtol, atol, icntrl_u=ic, family="sdirk")
    rc = np.zeros(20)
    rc[3], rc[4] = 2.0, 1.0                               # FacMin > FacMax is accepted; a bad FacSafe is not: IERR -4 for every cell
    rc[6] = 5.0
    ref = O.sdirk_tlm("cbm4", y0, np.eye(32)[None], T0, T0 + 3600.0, atol, rtol, None, rc)
    res = F.integrate_tlm("cbm4", T0, T0 + 3600.0, y0, np.eye(32)[None], rtol, atol, rcntrl_u=rc, family="sdirk")
    assert res.ierr[0] == ref["ierr"][0]
    if ref["ierr"][0] < 0:
        assert np.array_equal(res.y[0], y0)


def test_sdirk_tlm_general_pivoting_pass_at_loose_tolerances():
    """rtol 1e-3: every system leaves the static pivot pattern in its forward factorisations and takes the second pass with
    netlib's general partial pivoting (dense DGETF2 / DGETRS in the state sweep, the preparation and the column sweep)."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols(1.0e-2)
    y0 = perturbed_states(F.initial_state("cbm4"), 20)
    yt = np.tile(np.eye(32)[:9], (20, 1, 1))
    t1 = T0 + 30 * 3600.0
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, family="sdirk")
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol)
    assert (res.ierr == 1).all() and F.last_stats().cells_general > 0
    check(res, ref)
    # a mixed batch: intermediate tolerance, some systems on each path
    rtol, atol = tols(3.0e-3)
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, family="sdirk")
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol)
    check(res, ref)


def column_tols(ntlm, k):
    """(N, NTLM) tolerances of the columns, different per column and per component"""
    j = np.arange(ntlm)[:, None]
    i = np.arange(32)[None, :]
    return F01 * k * (1.0 + 0.25 * ((i + j) % 3)), F01 * k * 1e-2 * (1.0 + 0.5 * ((i + 2 * j) % 2))


@pytest.mark.parametrize("direct,est,trunc", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)])
def test_sdirk_tlm_lock_step_modes_match_oracle(direct, est, trunc):
    """ICNTRL(7) = 1 (direct TLM solve, a second factorisation per stage), ICNTRL(9) (column Newton control), ICNTRL(12)
    (column truncation errors in the step control) and their combinations: lock-step kernel, everything bit-identical."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols(1.0e-3)
    ncell = 6
    y0 = perturbed_states(F.initial_state("cbm4"), ncell)
    rng = np.random.default_rng(11)
    yt = np.repeat(np.eye(32)[None], ncell, axis=0) + 0.01 * rng.standard_normal((ncell, 32, 32))
    t1 = T0 + 4 * 3600.0
    rt, at = column_tols(32, 1.0e-3)
    at = at * 1.0e4                                       # unit-sized columns: absolute tolerances in their scale
    for method in ((0, 2, 5) if (direct, est, trunc) in ((1, 0, 0), (0, 1, 1)) else (0,)):
        ic = np.zeros(20, dtype=np.int32)
        ic[2], ic[6], ic[8], ic[11] = method, direct, est, trunc
        ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=at, rtol_tlm=rt)
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=at, rtol_tlm=rt, family="sdirk")
        assert (ref["ierr"] == 1).all()
        check(res, ref)
    if trunc or est:                                      # the columns do change the step sequence
        ic0 = np.zeros(20, dtype=np.int32)
        ic0[6] = direct
        base = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic0)
        assert not np.array_equal(base["istatus"], ref["istatus"])


def test_sdirk_tlm_lock_step_fewer_columns_scalar_tolerances_failures():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols(1.0e-4)
    ncell = 5
    y0 = perturbed_states(F.initial_state("cbm4"), ncell)
    yt = np.repeat(np.eye(32)[None, :7], ncell, axis=0)
    t1 = T0 + 2 * 3600.0
    rt, at = column_tols(7, 1.0e-4)
    ic = np.zeros(20, dtype=np.int32)
    ic[1], ic[2], ic[8], ic[11] = 1, 3, 1, 1              # scalar tolerances: the first entry of every column's set
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=at * 1e4, rtol_tlm=rt)
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=at * 1e4, rtol_tlm=rt, family="sdirk")
    check(res, ref)
    # columns whose Newton iterations fail end the attempt (:795-800): few iterations allowed, tight column tolerances
    for maxit, sc in ((3, 1.0e-6), (4, 1.0e-10)):
        ic = np.zeros(20, dtype=np.int32)
        ic[4], ic[8] = maxit, 1
        tt = np.full((7, 32), sc)
        ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=tt, rtol_tlm=tt)
        ic[8] = 0
        base = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
        assert (ref["istatus"][:, 5] > base["istatus"][:, 5]).all()          # more factorisations: rejected attempts
        ic[8] = 1
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=tt, rtol_tlm=tt, family="sdirk")
        check(res, ref)
    # column tolerances no step can meet: the step size collapses, IERR -7 with the columns of the last accepted step
    ic = np.zeros(20, dtype=np.int32)
    ic[6], ic[11] = 1, 1
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=at * 1e-12, rtol_tlm=rt * 1e-9)
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=at * 1e-12, rtol_tlm=rt * 1e-9, family="sdirk")
    assert (ref["ierr"] < 0).all()
    check(res, ref)
    # Max_no_steps reached in the direct mode
    ic = np.zeros(20, dtype=np.int32)
    ic[3], ic[6] = 12, 1
    ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, family="sdirk")
    assert (ref["ierr"] == -6).all()
    check(res, ref)
