"""GPU parity for INTEGRATE_TLM (TLM/ROS_TLM) on CBM-IV with forward error control (ICNTRL(12) = 0):
state and ISTATUS bit-identical to the CPU oracle, tangent-linear sensitivities within 1e-9 per column."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
T1 = T0 + 72.0 * 3600.0
RTOL_PARITY = 1e-9
NOISE_FLOOR = 1e-10


def tols():
    return np.full(32, F01 * 1.0e-4), np.full(32, F01 * 1.0e-6)


def check(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    a, b = res.y_tlm, ref["y_tlm"]
    sc = np.max(np.abs(b), axis=2)                       # [cell, column]
    er = np.max(np.abs(a - b), axis=2)
    floor = NOISE_FLOOR * np.max(sc, axis=1, keepdims=True)
    assert np.all(er <= RTOL_PARITY * np.maximum(sc, floor)), np.max(er / np.maximum(sc, floor))


def test_tlm_batch_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 24)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5                                           # Rodas4, vector tolerances, ICNTRL(12) = 0
    yt = np.tile(np.eye(32), (24, 1, 1))                # NTLM = N = 32, Y_tlm(t0) = I
    res = F.integrate_tlm("cbm4", T0, T1, y0, yt, rtol, atol, icntrl_u=ic)
    ref = O.ros_tlm("cbm4", y0, yt, T0, T1, atol, rtol, ic)
    assert (res.ierr == 1).all()
    check(res, ref)
    # ISTATUS closed forms of SURVEY §8a: att attempts, A accepted, n_newf = 5, S = 6, non-autonomous
    I = res.istatus.astype(np.int64)
    A, att = I[:, 3], I[:, 2]
    assert np.all(I[:, 0] == 2 * A + 5 * att) and np.all(I[:, 1] == 2 * A + 5 * att)
    assert np.all(I[:, 5] == att + I[:, 7]) and np.all(I[:, 6] == att * 6 * 33)
    # same step sequence as the adjoint family's forward sweep (both control the forward error only)
    adj = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=ic, want_mu=False)
    assert np.array_equal(adj.y, res.y) and np.array_equal(adj.istatus[:, 3], res.istatus[:, 3])


def test_tlm_fewer_columns_methods_autonomous_and_scalar_tolerances():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 6)
    rng = np.random.default_rng(7)
    yt = rng.standard_normal((6, 5, 32))
    t1 = T0 + 8 * 3600.0
    for method, auto, scalar in ((0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 1), (4, 1, 0), (5, 1, 1)):   # Ros2: some cells end with IERR -7
        ic = np.zeros(20, dtype=np.int32)
        ic[0], ic[1], ic[2] = auto, scalar, method
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic)
        ref = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
        check(res, ref)


def test_tlm_is_the_transpose_of_the_discrete_adjoint_when_autonomous():
    """Adjoint-TLM duality <lambda(t0), dy0> = <lambda(T), dy(T)> on the same step sequence: with the time-derivative
    terms off (ICNTRL(1) = 1) ROS_TLM and ROS_ADJ are exact transposes, Y_tlm(T)(k, j) = Lambda(t0)(j, k)."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 4)
    ic = np.zeros(20, dtype=np.int32)
    ic[0], ic[2] = 1, 5
    t1 = T0 + 24 * 3600.0
    tl = F.integrate_tlm("cbm4", T0, t1, y0, np.tile(np.eye(32), (4, 1, 1)), rtol, atol, icntrl_u=ic)
    ad = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic, want_mu=False)
    assert np.array_equal(tl.y, ad.y)
    lam = ad.lambda_                                       # [cell, m, i] = d y_m(T) / d y_i(t0)
    ytl = np.transpose(tl.y_tlm, (0, 2, 1))                # [cell, k, j] -> compare Y_tlm(T)(k, j) with Lambda(j, k)
    sc = np.max(np.abs(lam), axis=2, keepdims=True)
    ok = sc[:, :, 0] > 1e-10 * sc.max(axis=1)
    err = (np.max(np.abs(ytl - lam), axis=2) / np.maximum(sc[:, :, 0], 1e-300))[ok]
    assert err.max() < 1e-9, err.max()


def test_tlm_default_call_runs_with_truncation_error_control():
    """INTEGRATE_TLM with ICNTRL_U absent presets ICNTRL(12) = 1 (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:70-72): the columns'
    error norms (ros_ErrorNorm_tlm :745-769) enter accept/reject, so state and columns advance in lock-step; the lock-step kernel
    follows the reference's operation order, so Y_tlm is compared bit for bit too."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 12)
    yt = np.tile(np.eye(32)[:6], (12, 1, 1))
    at = np.full((6, 32), 1.0e-6)
    rt = np.full((6, 32), 1.0e-4)
    t1 = T0 + 24 * 3600.0
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, atol_tlm=at, rtol_tlm=rt)      # the reference's default call
    ref = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, None, atol_tlm=at, rtol_tlm=rt)
    assert (res.ierr == 1).all()
    check(res, ref)
    assert np.array_equal(res.y_tlm, ref["y_tlm"])
    # the columns really decide: with forward-only control the step counts differ
    ic = np.zeros(20, dtype=np.int32)
    ic[1], ic[2] = 1, 5
    fwd = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic)
    assert (res.istatus[:, 2] > fwd.istatus[:, 2]).any()
    with pytest.raises(F.FatodeError):                     # the preset needs the column tolerances
        F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 9                                              # unknown method: the reference's IERR -2 for every cell
    r = F.integrate_tlm("cbm4", T0, T0 + 3600.0, y0[:1], np.eye(32)[None], rtol, atol, icntrl_u=ic)
    assert r.ierr[0] == -2 and np.array_equal(r.y[0], y0[0])


def test_tlm_truncation_error_control_methods_and_tolerance_kinds():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 6)
    rng = np.random.default_rng(11)
    yt = rng.standard_normal((6, 32, 32))                  # NTLM = 32: every lane carries a column
    at = 1.0e-6 * (1.0 + rng.random((32, 32)))
    rt = 1.0e-4 * (1.0 + rng.random((32, 32)))
    t1 = T0 + 6 * 3600.0
    for method, auto, scalar in ((1, 0, 0), (2, 0, 0), (3, 0, 1), (4, 1, 0), (5, 1, 1), (5, 0, 0)):
        ic = np.zeros(20, dtype=np.int32)
        ic[0], ic[1], ic[2], ic[11] = auto, scalar, method, 1
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=at, rtol_tlm=rt)
        ref = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=at, rtol_tlm=rt)
        check(res, ref)
        assert np.array_equal(res.y_tlm, ref["y_tlm"]), (method, auto, scalar)


def test_tlm_general_pivoting_pass_at_loose_tolerances():
    """Systems whose factorisations leave the static pivot pattern of the per-thread kernels (almost all at rtol 1e-3) are
    integrated by the lock-step kernel with netlib's partial pivoting: no IERR -20 hand-backs any more."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    y0 = perturbed_states(F.initial_state("cbm4"), 24)
    yt = np.tile(np.eye(32)[:4], (24, 1, 1))
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    t1 = T0 + 24 * 3600.0
    general = 0
    for rt_ in (1.0e-2, 1.0e-3):
        rtol, atol = np.full(32, F01 * rt_), np.full(32, F01 * rt_ * 1.0e-2)
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic)
        ref = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
        assert not (res.ierr == -20).any()
        check(res, ref)
        general += F.last_stats().cells_general
    assert general > 0
