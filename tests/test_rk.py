"""FWD/RK (forward part): fully implicit 3-stage Runge-Kutta methods.
CPU: the oracle restatement (oracle/rk.hpp) is pinned to the reference's golden file
EXAMPLES/cbm4/CBM4_RK/cbm4_rk_sol.txt (driver cbm4_rk_dr.F90: Radau-2A, rtol = 0.1*1e-5, atol = 0.1*1e-7 with a
REAL(4) 0.1, 12 h .. 84 h).  GPU: the per-thread kernel is bit-identical to the oracle (state, RSTATUS, all counters)."""
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


def test_oracle_matches_reference_golden_file():
    g = np.array(json.load(open(os.path.join(G, "cbm4_rk.json")))["y"])
    rtol, atol = driver_tols()
    old = O.lib().oracle_set_libm(1)          # the golden file was produced with the platform libm's pow
    try:
        r = O.rk_fwd("cbm4", O.initial_state("cbm4"), T0, T1, atol, rtol, radau())
    finally:
        O.lib().oracle_set_libm(old)
    assert r["ierr"][0] == 1
    assert np.max(np.abs(r["y"][0] - g)) <= 1e-10 * np.max(np.abs(g))          # observed 1.1e-12
    # portable elementary kernels: another valid run of the same method (step sequence differs by a few steps)
    p = O.rk_fwd("cbm4", O.initial_state("cbm4"), T0, T1, atol, rtol, radau())
    assert np.max(np.abs(p["y"][0] - g)) <= 1e-7 * np.max(np.abs(g))
    I = p["istatus"][0]
    # Nstp counts passes that got past the decomposition (Newton failures included); Ndec <= passes; RK_Solve counts 2
    assert I[2] >= I[3] + I[4] and I[1] <= I[5] and I[7] == 0 and I[6] > 2 * I[2]


def test_oracle_option_errors_and_methods():
    rtol, atol = driver_tols()
    y0 = O.initial_state("cbm4")
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 7
    assert O.rk_fwd("cbm4", y0, T0, T1, atol, rtol, ic)["ierr"][0] == -13
    rc = np.zeros(20)
    rc[9] = 1.5                                   # Qmin > 1
    assert O.rk_fwd("cbm4", y0, T0, T1, atol, rtol, radau(), rc)["ierr"][0] == -7
    ref = O.rk_fwd("cbm4", y0, T0, T0 + 6 * 3600.0, atol, rtol, radau())["y"][0]
    ref1 = O.rk_fwd("cbm4", y0, T0, T0 + 3600.0, atol, rtol, radau())["y"][0]
    for method in (0, 3, 4, 5):                   # Lobatto-3C (default), Gauss, Radau-1A, Lobatto-3A agree with Radau-2A
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        # (the Gauss method, not stiffly accurate, collapses to ~0.02 s steps later in the afternoon: one hour only)
        r = O.rk_fwd("cbm4", y0, T0, T0 + (3600.0 if method == 3 else 6 * 3600.0), atol, rtol, ic)
        assert r["ierr"][0] == 1
        want = ref1 if method == 3 else ref
        assert np.max(np.abs(r["y"][0] - want)) <= 2e-5 * np.max(np.abs(want))
    # the classic error estimator is only reachable by calling RungeKutta directly (raw): same solution to tolerance
    icr = np.zeros(20, dtype=np.int32)
    icr[2] = 1
    r = O.rk_fwd("cbm4", y0, T0, T0 + 6 * 3600.0, atol, rtol, icr, raw=True)
    assert r["ierr"][0] == 1 and np.max(np.abs(r["y"][0] - ref)) <= 2e-5 * np.max(np.abs(ref))


def assert_bitwise(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"]), "state must be bit-identical to the oracle"


@pytest.mark.gpu
def test_cbm4_radau_batch_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 96)
    res = F.integrate_rk("cbm4", T0, T1, y0, rtol, atol, icntrl_u=radau())
    ref = O.rk_fwd("cbm4", y0, T0, T1, atol, rtol, radau())
    assert (res.ierr == 1).all()
    assert_bitwise(res, ref)
    g = np.array(json.load(open(os.path.join(G, "cbm4_rk.json")))["y"])
    assert np.max(np.abs(res.y[0] - g)) <= 1e-7 * np.max(np.abs(g))       # cell 0 = the driver's problem


@pytest.mark.gpu
def test_rk_methods_options_and_small_models():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 16)
    t1 = T0 + 10 * 3600.0
    for method in (0, 1, 3, 4, 5):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        rc = np.zeros(20)
        if method == 3:
            ic[5] = 1                      # zero starting values for the Newton iterations
            rc[8] = 1.0e-2                 # NewtonTol
        if method == 4:
            ic[10] = 1                     # classic step-size controller
        te = T0 + 3600.0 if method == 3 else t1    # Gauss: see test_oracle_option_errors_and_methods
        assert_bitwise(F.integrate_rk("cbm4", T0, te, y0, rtol, atol, icntrl_u=ic, rcntrl_u=rc),
                       O.rk_fwd("cbm4", y0, T0, te, atol, rtol, ic, rc))
    # small_strato (dense real + complex LU with pivoting per thread) and roberts
    ys = perturbed_states(F.initial_state("small_strato"), 64)
    rs, as_ = np.full(5, 1e-4), np.full(5, 1e-2)
    for method in (0, 1, 5):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        assert_bitwise(F.integrate_rk("small_strato", T0, T0 + 5 * 86400.0, ys, rs, as_, icntrl_u=ic),
                       O.rk_fwd("small_strato", ys, T0, T0 + 5 * 86400.0, as_, rs, ic))
    yr = np.array([[1.0, 0.0, 0.0]])
    rr, ar = np.full(3, 1e-6), np.array([1e-8, 1e-14, 1e-6])
    rc = np.zeros(20)
    rc[2] = 1e-3
    assert_bitwise(F.integrate_rk("roberts", 0.0, 4.0e4, yr, rr, ar, icntrl_u=radau(), rcntrl_u=rc),
                   O.rk_fwd("roberts", yr, 0.0, 4.0e4, ar, rr, radau(), rc))
    # option errors come back per cell like the reference's IERR
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 9
    assert (F.integrate_rk("cbm4", T0, t1, y0, rtol, atol, icntrl_u=ic).ierr == -13).all()


@pytest.mark.gpu
def test_general_pivoting_second_pass_rk_and_sdirk():
    """At loose tolerances some CBM-IV systems need interchanges outside the static pattern (e.g. column 26 with row 28):
    the first pass hands them back and the dense general-pivoting pass must reproduce the oracle bit for bit."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-4), np.full(32, F01 * 1.0e-6)
    y0 = perturbed_states(F.initial_state("cbm4"), 192)
    res = F.integrate_rk("cbm4", T0, T1, y0, rtol, atol, icntrl_u=radau())
    st = F.last_stats()
    assert 0 < st.cells_general < 192, "this case is known to need the second pass for a few systems"
    assert_bitwise(res, O.rk_fwd("cbm4", y0, T0, T1, atol, rtol, radau()))
    assert (res.ierr == 1).all()
    rtol, atol = np.full(32, 1.0e-3), np.full(32, 1.0e-5)
    res = F.integrate_sdirk("cbm4", T0, T1, y0, rtol, atol)
    assert_bitwise(res, O.sdirk_fwd("cbm4", y0, T0, T1, atol, rtol))
    print("SDIRK systems on the general-pivoting pass:", F.last_stats().cells_general)
