"""ADJ/RK_ADJ (adjoint part): discrete adjoint of the 3-stage implicit Runge-Kutta methods.
CPU: the oracle restatement (oracle/rk_adj.hpp) against the reference's golden files
EXAMPLES/cbm4/CBM4_RK_ADJ/{cbm4_rk_sol.txt, cbm4_output_sens.txt} (driver cbm4_rk_adj_dr.F90: Radau-2A, rtol = 0.1*1e-5,
atol = 0.1*1e-7, NADJ = 32, NP = 29, 12 h .. 84 h).  No combination of estimator / controller / starting values of the
shipped integrator reproduces that run's step sequence (the files predate it), so the pin is at the level of two valid
runs: state 2e-8, dominant sensitivity columns 1e-7.  GPU: forward sweep bit-identical, Lambda/Mu within 1e-9."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
T1 = T0 + 72.0 * 3600.0


def driver_tols():
    return np.full(32, F01 * 1.0e-5), np.full(32, F01 * 1.0e-7)


def radau():
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 1
    return ic


def test_oracle_against_reference_golden_files_and_counters():
    g = json.load(open(os.path.join(G, "cbm4_rk_adj.json")))
    gy, glam, gmu = np.array(g["y"]), np.array(g["lambda"]).reshape(32, 32), np.array(g["mu"]).reshape(32, 29)
    rtol, atol = driver_tols()
    r = O.rk_adj("cbm4", O.initial_state("cbm4"), 32, T0, T1, atol, rtol, radau(), nthreads=1)
    assert r["ierr"][0] == 1
    y, lam, mu = r["y"][0], r["lambda_"][0], r["mu"][0]
    assert np.max(np.abs(y - gy)) <= 1e-7 * np.max(np.abs(gy))                 # observed 3.4e-9
    cs = np.max(np.abs(glam), axis=1)
    err = np.max(np.abs(lam - glam), axis=1) / np.maximum(cs, 1e-300)
    dominant = cs > 1e-6                         # columns within six decades of the largest one
    assert dominant.sum() >= 18 and np.all(err[dominant] <= 1e-5) and np.median(err[dominant]) <= 1e-7, err
    rs = np.max(np.abs(gmu), axis=0)
    rerr = np.max(np.abs(mu - gmu), axis=0) / rs
    noise = {15, 21, 22, 23, 24, 25, 26}         # parameters that multiply identically-zero species (as for ROS_ADJ)
    good = [p for p in range(29) if p not in noise]
    assert np.all(rerr[good] <= 1e-5), rerr      # observed <= 3e-7
    # counters: forward passes F (incl. Newton failures), accepted A, backward adds per step Njac 4, Ndec 1,
    # Nstp 1 and Nsol 2 * sweeps * NADJ with sweeps = min(8, max(NewIt + 2, 7))
    I = r["istatus"][0]
    fwd = O.rk_adj("cbm4", O.initial_state("cbm4"), 32, T0, T1, atol, rtol, np.array([0, 0, 1, 0, 0, 0, 0, 0, 1] + [0] * 11, dtype=np.int32),
                   nthreads=1)                   # ICNTRL(9) = 1: forward sweep only
    Fw = fwd["istatus"][0]
    A = I[3]
    assert np.array_equal(fwd["y"][0], y) and Fw[3] == A and Fw[4] == I[4] and Fw[0] == I[0] and Fw[7] == I[7]
    assert I[1] == Fw[1] + 4 * A and I[5] == Fw[5] + A and I[2] == Fw[2] + A
    per_step = (I[6] - Fw[6]) / (64.0 * A)
    assert 7.0 <= per_step <= 8.0
    # the adjoint of another family agrees to integration accuracy on the dominant columns (independent check)
    ic5 = np.zeros(20, dtype=np.int32)
    ic5[2] = 5
    q = O.ros_adj("cbm4", O.initial_state("cbm4"), 32, T0, T1, atol, rtol, ic5, nthreads=1)
    ql = q["lambda_"][0]
    cq = np.max(np.abs(ql), axis=1)
    e2 = np.max(np.abs(lam - ql), axis=1) / np.maximum(cq, 1e-300)
    assert np.all(e2[cq > 1e-3] <= 1e-4), e2


def test_oracle_options():
    rtol, atol = driver_tols()
    y0 = O.initial_state("cbm4")
    ic = radau()
    ic[6] = 3                                     # adaptive adjoint solve: not restated
    assert O.rk_adj("cbm4", y0, 2, T0, T0 + 600.0, atol, rtol, ic)["ierr"][0] == -100
    ic = radau()
    ic[2] = 5                                     # no Lobatto-3A in RK_ADJ
    assert O.rk_adj("cbm4", y0, 2, T0, T0 + 600.0, atol, rtol, ic)["ierr"][0] == -13
    ic = radau()
    ic[8] = 3                                     # continuous adjoints are not implemented upstream either
    assert O.rk_adj("cbm4", y0, 2, T0, T0 + 600.0, atol, rtol, ic)["ierr"][0] == -9


def test_oracle_direct_adjoint_solve_agrees_with_the_fixed_sweeps():
    """ICNTRL(7) = 2 (lapack_decomp_big / lapack_solve_big: one 3N x 3N DGETRF per step, one DGETRS('T') per column) and the
    preset fixed number of transposed Newton sweeps solve the same linear systems: two independent code paths of the
    restatement agree to the sweeps' convergence level; the counters follow ADJ/RK_ADJ :1331, :1472."""
    rtol, atol = driver_tols()
    y0 = O.initial_state("cbm4")
    t1 = T0 + 6 * 3600.0
    fx = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, radau(), nthreads=1)
    ic = radau()
    ic[6] = 2
    dr = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic, nthreads=1)
    assert dr["ierr"][0] == 1 and np.array_equal(dr["y"], fx["y"])
    sc = np.max(np.abs(fx["lambda_"][0]))
    assert np.max(np.abs(dr["lambda_"][0] - fx["lambda_"][0])) <= 1e-8 * sc            # observed 1.9e-10
    assert np.max(np.abs(dr["mu"][0] - fx["mu"][0])) <= 1e-8 * np.max(np.abs(fx["mu"][0]))
    A = fx["istatus"][0][3]
    assert dr["istatus"][0][5] == fx["istatus"][0][5] + A                             # one more decomposition per step
    fwd_ic = radau()
    fwd_ic[8] = 1
    fwd = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, fwd_ic, nthreads=1)
    assert dr["istatus"][0][6] == fwd["istatus"][0][6] + 32 * A                       # one big solve per column and step


@pytest.mark.gpu
def test_rk_adj_direct_solve_matches_oracle():
    """ICNTRL(7) = 2 on the GPU (k_rk_adj_big: block-cooperative 96 x 96 DGETF2 + DGETRS('T') per column): state, counters and
    Lambda bit-identical to the restatement, Mu to rounding; static-pattern and general-pivoting records."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    from test_ros_adj_gpu import check_against_oracle
    y0 = perturbed_states(F.initial_state("cbm4"), 12)
    t1 = T0 + 10 * 3600.0
    for (rt, at, nadj, with_mu) in ((1.0e-5, 1.0e-7, 32, True), (1.0e-5, 1.0e-7, 5, False), (1.0e-3, 1.0e-5, 32, True)):
        rtol, atol = np.full(32, F01 * rt), np.full(32, F01 * at)
        ic = radau()
        ic[6] = 2
        res = F.integrate_adj("cbm4", T0, t1, y0, nadj, rtol, atol, icntrl_u=ic, want_mu=with_mu, family="rk")
        ref = O.rk_adj("cbm4", y0, nadj, T0, t1, atol, rtol, ic, with_mu=with_mu)
        assert (res.ierr == 1).all() and np.array_equal(res.y, ref["y"])
        assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.rstatus, ref["rstatus"])
        assert np.array_equal(res.lambda_, ref["lambda_"]), "same operations in the same order as DGETF2 / DGETRS('T')"
        check_against_oracle(res, ref, nadj)
    assert F.last_stats().cells_general > 0          # the loose-tolerance case went through the dense preparation records


@pytest.mark.gpu
def test_cbm4_rk_adj_batch_matches_oracle():
    """Forward sweep bit-identical (state, RSTATUS, every ISTATUS counter incl. the backward sweep's), Lambda and Mu within
    1e-9 with the same per-column / per-parameter scaling as the ROS_ADJ parity test."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    from test_ros_adj_gpu import check_against_oracle
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 48)
    res = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=radau(), family="rk")
    ref = O.rk_adj("cbm4", y0, 32, T0, T1, atol, rtol, radau())
    assert (res.ierr == 1).all()
    assert np.array_equal(res.y, ref["y"]) and np.array_equal(res.rstatus, ref["rstatus"])
    check_against_oracle(res, ref, 32)
    g = json.load(open(os.path.join(G, "cbm4_rk_adj.json")))
    glam = np.array(g["lambda"]).reshape(32, 32)
    cs = np.max(np.abs(glam), axis=1)
    err = np.max(np.abs(res.lambda_[0] - glam), axis=1) / np.maximum(cs, 1e-300)
    assert np.median(err[cs > 1e-6]) <= 1e-7                      # cell 0 = the driver's problem vs the reference's file


@pytest.mark.gpu
def test_rk_adj_variants():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    from test_ros_adj_gpu import check_against_oracle
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 12)
    t1 = T0 + 8 * 3600.0
    for method, nadj, with_mu in ((0, 32, True), (4, 5, True), (1, 32, False), (3, 7, True)):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        te = T0 + 3600.0 if method == 3 else t1
        res = F.integrate_adj("cbm4", T0, te, y0, nadj, rtol, atol, icntrl_u=ic, want_mu=with_mu, family="rk")
        ref = O.rk_adj("cbm4", y0, nadj, T0, te, atol, rtol, ic, with_mu=with_mu)
        assert np.array_equal(res.y, ref["y"])
        check_against_oracle(res, ref, nadj)
    # forward sweep only (ICNTRL(9) = 1), tiny tape slots (overflow -> second group), Gustafsson controller not selectable
    ic = radau()
    ic[8] = 1
    res = F.integrate_adj("cbm4", T0, t1, y0, 4, rtol, atol, icntrl_u=ic, family="rk")
    ref = O.rk_adj("cbm4", y0, 4, T0, t1, atol, rtol, ic)
    assert np.array_equal(res.y, ref["y"]) and np.array_equal(res.istatus, ref["istatus"])
    ic = radau()
    ic[3] = 20                                                    # Max_no_steps = 20: IERR -9 like the reference
    res = F.integrate_adj("cbm4", T0, T1, y0, 4, rtol, atol, icntrl_u=ic, family="rk")
    ref = O.rk_adj("cbm4", y0, 4, T0, T1, atol, rtol, ic)
    assert np.array_equal(res.ierr, ref["ierr"]) and (res.ierr == -9).all()
    ic = radau()
    ic[6] = 3                                                     # adaptive adjoint solve: not built
    with pytest.raises(Exception):
        F.integrate_adj("cbm4", T0, t1, y0, 4, rtol, atol, icntrl_u=ic, family="rk")


@pytest.mark.gpu
def test_rk_adj_waves_batches_and_tape_overflow_are_invisible():
    import ctypes
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 40)
    t1 = T0 + 30 * 3600.0
    base = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=radau(), family="rk")
    L = F.lib()
    L.fatode_set_rk_tape_slot.argtypes = [ctypes.c_int]
    L.fatode_set_tape_budget.argtypes = [ctypes.c_uint64]
    L.fatode_set_wave_cells.argtypes = [ctypes.c_int64]
    try:
        L.fatode_set_wave_cells(7)                    # six waves
        L.fatode_set_tape_budget(1 << 30)             # ~0.75 GB of records per backward batch
        L.fatode_set_rk_tape_slot(100)                # most systems overflow their slot and take the second pass
        res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=radau(), family="rk")
        st = F.last_stats()
    finally:
        L.fatode_set_wave_cells(0)
        L.fatode_set_tape_budget(0)
        L.fatode_set_rk_tape_slot(0)
        L.fatode_release_workspace()
    assert st.waves > 6
    assert (base.istatus[:, 3] > 100).any()
    for a, b in ((res.y, base.y), (res.istatus, base.istatus), (res.rstatus, base.rstatus), (res.lambda_, base.lambda_),
                 (res.mu, base.mu), (res.ierr, base.ierr)):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_rk_adj_general_pivoting_pass_at_loose_tolerances():
    """rtol 1e-4: every system leaves the static pivot pattern in its real or complex factorisations and takes the second pass
    with netlib's general partial pivoting (dense DGETF2/ZGETF2 forward, dense DGETRS('T')/ZGETRS('T') in the backward sweep)."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-3), np.full(32, F01 * 1.0e-5)
    y0 = perturbed_states(F.initial_state("cbm4"), 12)
    t1 = T0 + 20 * 3600.0
    res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=radau(), family="rk")
    ref = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, radau())
    assert (res.ierr == 1).all() and F.last_stats().cells_general > 0
    assert np.array_equal(res.ierr, ref["ierr"]) and np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    assert np.array_equal(res.lambda_, ref["lambda_"]), "the backward sweep follows the reference's operation order"
    sc = np.max(np.abs(ref["mu"]), axis=2)
    er = np.max(np.abs(res.mu - ref["mu"]), axis=2)
    assert np.all(er <= 1e-7 * np.maximum(sc, 1e-10 * sc.max(axis=1, keepdims=True)))
