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


def test_sdirk_tlm_unsupported_modes_and_option_errors():
    import fatode_b200 as F
    rtol, atol = tols()
    y0 = F.initial_state("cbm4")
    for k in (6, 8, 11):                                  # ICNTRL(7) direct solve, ICNTRL(9), ICNTRL(12)
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
