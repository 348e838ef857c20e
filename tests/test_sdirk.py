"""FWD/SDIRK (forward part).
CPU: the oracle restatement (oracle/sdirk.hpp) is pinned to the reference's golden file
EXAMPLES/cbm4/CBM4_SDIRK/cbm4_sdirk_sol.txt (driver cbm4_sdirk_dr.F90: Sdirk4a, rtol = 0.1*1e-5, atol = 0.1*1e-7 with a
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


def test_oracle_matches_reference_golden_file():
    g = np.array(json.load(open(os.path.join(G, "cbm4_sdirk.json")))["y"])
    rtol, atol = driver_tols()
    old = O.lib().oracle_set_libm(1)          # the golden file was produced with the platform libm's pow
    try:
        r = O.sdirk_fwd("cbm4", O.initial_state("cbm4"), T0, T1, atol, rtol)
    finally:
        O.lib().oracle_set_libm(old)
    assert r["ierr"][0] == 1
    assert np.max(np.abs(r["y"][0] - g)) <= 1e-10 * np.max(np.abs(g))
    p = O.sdirk_fwd("cbm4", O.initial_state("cbm4"), T0, T1, atol, rtol)   # portable kernels: a second valid run
    assert np.max(np.abs(p["y"][0] - g)) <= 1e-9 * np.max(np.abs(g))
    I = p["istatus"][0]
    assert I[2] == I[3] + I[4] and I[5] >= I[1] and I[6] > I[0] - 1 and I[7] == 0   # Nstp = Nacc + Nrej; Ndec >= Njac


def assert_bitwise(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"]), "state must be bit-identical to the oracle"


@pytest.mark.gpu
def test_cbm4_sdirk_batch_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = driver_tols()
    y0 = perturbed_states(F.initial_state("cbm4"), 96)
    res = F.integrate_sdirk("cbm4", T0, T1, y0, rtol, atol)
    ref = O.sdirk_fwd("cbm4", y0, T0, T1, atol, rtol)
    assert (res.ierr == 1).all()
    assert_bitwise(res, ref)
    g = np.array(json.load(open(os.path.join(G, "cbm4_sdirk.json")))["y"])
    assert np.max(np.abs(res.y[0] - g)) <= 1e-9 * np.max(np.abs(g))       # cell 0 = the driver's problem


@pytest.mark.gpu
def test_sdirk_methods_options_and_small_models():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-4), np.full(32, F01 * 1.0e-6)
    y0 = perturbed_states(F.initial_state("cbm4"), 8)
    t1 = T0 + 10 * 3600.0
    for method in (1, 2, 3, 4, 5):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        if method == 5:
            ic[5] = 1                      # zero starting values for the Newton iterations
        rc = np.zeros(20)
        if method == 3:
            rc[8] = 1.0e-2                 # NewtonTol
        assert_bitwise(F.integrate_sdirk("cbm4", T0, t1, y0, rtol, atol, icntrl_u=ic, rcntrl_u=rc),
                       O.sdirk_fwd("cbm4", y0, T0, t1, atol, rtol, ic, rc))
    # small_strato (dense LU with pivoting per thread) and roberts
    ys = perturbed_states(F.initial_state("small_strato"), 64)
    rs, as_ = np.full(5, 1e-4), np.full(5, 1e-2)
    for method in (0, 1, 3):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        assert_bitwise(F.integrate_sdirk("small_strato", T0, T0 + 5 * 24 * 3600.0, ys, rs, as_, icntrl_u=ic),
                       O.sdirk_fwd("small_strato", ys, T0, T0 + 5 * 24 * 3600.0, as_, rs, ic))
    yr = perturbed_states(np.array([1.0, 0.0, 0.0]), 32)
    rr, ar = np.full(3, 1e-4), np.full(3, 1e-8)
    assert_bitwise(F.integrate_sdirk("roberts", 0.0, 4.0e5, yr, rr, ar), O.sdirk_fwd("roberts", yr, 0.0, 4.0e5, ar, rr))
    # option errors follow the reference: the last failing check wins, Y untouched
    bad = atol.copy()
    bad[2] = 0.0
    r = F.integrate_sdirk("cbm4", T0, t1, y0[:1], rtol, bad)
    assert r.ierr[0] == -5 and np.array_equal(r.y[0], y0[0])
