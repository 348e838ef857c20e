"""ADJ/SDIRK_ADJ: discrete adjoint of the SDIRK family.
CPU: the oracle restatement (oracle/sdirk_adj.hpp) against the reference's golden files
EXAMPLES/cbm4/CBM4_SDIRK_ADJ/{cbm4_sdirk_sol.txt, cbm4_output_sens.txt} (driver cbm4_sdirk_adj_dr.F90: Sdirk4a, rtol = 0.1*1e-6,
atol = 0.1*1e-8, ICNTRL(4) = 100000, direct adjoint solve, NADJ = 32, NP = 29, 12 h .. 84 h).
GPU: forward sweep and every counter bit-identical to the oracle, Lambda/Mu within 1e-9 per column."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
T1 = T0 + 72.0 * 3600.0
RTOL_PARITY = 1e-9
NOISE_FLOOR = 1e-10


def driver_tols():
    return np.full(32, F01 * 1.0e-6), np.full(32, F01 * 1.0e-8)


def driver_icntrl():
    ic = np.zeros(20, dtype=np.int32)
    ic[3] = 100000
    return ic


def test_oracle_against_reference_golden_files_and_counters():
    g = json.load(open(os.path.join(G, "cbm4_sdirk_adj.json")))
    gy, glam, gmu = np.array(g["y"]), np.array(g["lambda"]).reshape(32, 32), np.array(g["mu"]).reshape(32, 29)
    rtol, atol = driver_tols()
    r = O.sdirk_adj("cbm4", O.initial_state("cbm4"), 32, T0, T1, atol, rtol, driver_icntrl(), nthreads=1)
    assert r["ierr"][0] == 1
    y, lam, mu = r["y"][0], r["lambda_"][0], r["mu"][0]
    assert np.max(np.abs(y - gy)) <= 1e-9 * np.max(np.abs(gy))                 # observed 1.0e-11
    cs = np.max(np.abs(glam), axis=1)
    err = np.max(np.abs(lam - glam), axis=1) / np.maximum(cs, 1e-300)
    dominant = cs > 1e-6                         # columns within six decades of the largest one
    assert dominant.sum() >= 18 and np.all(err[dominant] <= 1e-6) and np.median(err[dominant]) <= 2e-8, err   # observed <= 7.4e-8
    rs = np.max(np.abs(gmu), axis=0)
    rerr = np.max(np.abs(mu - gmu), axis=0) / np.maximum(rs, 1e-300)
    noise = {15, 21, 22, 23, 24, 25, 26}         # parameters that multiply identically-zero species (as for ROS_ADJ)
    good = [p for p in range(29) if p not in noise]
    assert np.all(rerr[good] <= 1e-5), rerr
    # counters: forward sweep = the FWD/SDIRK integrator's (same loop + Push); the backward sweep adds per popped step
    # Nstp 1, Njac 1 + S, Ndec 1 + S, Nsol S * NADJ (direct mode)
    I = r["istatus"][0].astype(np.int64)
    ic = driver_icntrl()
    f = O.sdirk_fwd("cbm4", O.initial_state("cbm4"), T0, T1, atol, rtol, ic)
    Fw = f["istatus"][0].astype(np.int64)
    A = I[3]
    assert np.array_equal(f["y"][0], y) and Fw[3] == A and Fw[4] == I[4] and Fw[0] == I[0] and Fw[7] == I[7]
    assert I[2] == Fw[2] + A and I[1] == Fw[1] + 6 * A and I[5] == Fw[5] + 6 * A and I[6] == Fw[6] + 5 * 32 * A


def test_oracle_newton_mode_agrees_with_direct_mode_and_with_ros_adj():
    rtol, atol = np.full(32, F01 * 1.0e-5), np.full(32, F01 * 1.0e-7)
    y0 = O.initial_state("cbm4")
    t1 = T0 + 6 * 3600.0
    ic = np.zeros(20, dtype=np.int32)
    d = O.sdirk_adj("cbm4", y0, 8, T0, t1, atol, rtol, ic)
    ic[6] = 2                                     # ICNTRL(7) /= 1: simplified Newton iterations on the adjoint stages
    n = O.sdirk_adj("cbm4", y0, 8, T0, t1, atol, rtol, ic)
    assert d["ierr"][0] == 1 and n["ierr"][0] == 1 and np.array_equal(d["y"], n["y"])
    sc = np.max(np.abs(d["lambda_"][0]), axis=1)
    e = np.max(np.abs(d["lambda_"][0] - n["lambda_"][0]), axis=1) / sc
    assert e.max() < 1e-3, e
    ic5 = np.zeros(20, dtype=np.int32)
    ic5[2] = 5
    ro = O.ros_adj("cbm4", y0, 8, T0, t1, atol, rtol, ic5)
    sc = np.max(np.abs(ro["lambda_"][0]), axis=1)
    e = np.max(np.abs(d["lambda_"][0] - ro["lambda_"][0]), axis=1) / sc      # two discretisations of the same adjoint equation
    assert np.median(e) < 5e-5 and e.max() < 1e-2, e
    # methods and the Lambda-only entry (SDIRKADJ2)
    for method in (1, 2, 3, 5):
        icm = np.zeros(20, dtype=np.int32)
        icm[2] = method
        a = O.sdirk_adj("cbm4", y0, 3, T0, T0 + 3600.0, atol, rtol, icm)
        b = O.sdirk_adj("cbm4", y0, 3, T0, T0 + 3600.0, atol, rtol, icm, with_mu=False)
        assert a["ierr"][0] == 1 and np.array_equal(a["lambda_"], b["lambda_"])


def check(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    for a, b in ((res.lambda_, ref["lambda_"]), (res.mu, ref["mu"])):
        if b is None:
            continue
        sc = np.max(np.abs(b), axis=2)
        er = np.max(np.abs(a - b), axis=2)
        floor = NOISE_FLOOR * np.max(sc, axis=1, keepdims=True)
        assert np.all(er <= RTOL_PARITY * np.maximum(sc, floor)), np.max(er / np.maximum(sc, floor))


@pytest.mark.gpu
def test_cbm4_sdirk_adj_batch_matches_oracle():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-5), np.full(32, F01 * 1.0e-7)
    y0 = perturbed_states(F.initial_state("cbm4"), 24)
    t1 = T0 + 24 * 3600.0
    res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, family="sdirk")
    ref = O.sdirk_adj("cbm4", y0, 32, T0, t1, atol, rtol)
    assert (res.ierr == 1).all()
    check(res, ref)
    assert np.array_equal(res.lambda_, ref["lambda_"]), "the backward sweep follows the reference's operation order"


@pytest.mark.gpu
def test_cbm4_sdirk_adj_driver_problem_against_golden():
    import fatode_b200 as F
    g = json.load(open(os.path.join(G, "cbm4_sdirk_adj.json")))
    gy, glam = np.array(g["y"]), np.array(g["lambda"]).reshape(32, 32)
    rtol, atol = driver_tols()
    res = F.integrate_adj("cbm4", T0, T1, F.initial_state("cbm4"), 32, rtol, atol, icntrl_u=driver_icntrl(), family="sdirk")
    assert res.ierr[0] == 1
    assert np.max(np.abs(res.y[0] - gy)) <= 1e-9 * np.max(np.abs(gy))
    cs = np.max(np.abs(glam), axis=1)
    err = np.max(np.abs(res.lambda_[0] - glam), axis=1) / np.maximum(cs, 1e-300)
    assert np.all(err[cs > 1e-6] <= 1e-6)


@pytest.mark.gpu
def test_sdirk_adj_variants():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-4), np.full(32, F01 * 1.0e-6)
    y0 = perturbed_states(F.initial_state("cbm4"), 6)
    t1 = T0 + 8 * 3600.0
    for method, nadj, with_mu, scalar in ((1, 5, True, 0), (2, 32, False, 0), (3, 7, True, 1), (5, 32, True, 0)):
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[2] = scalar, method
        res = F.integrate_adj("cbm4", T0, t1, y0, nadj, rtol, atol, icntrl_u=ic, want_mu=with_mu, family="sdirk")
        ref = O.sdirk_adj("cbm4", y0, nadj, T0, t1, atol, rtol, ic, with_mu=with_mu)
        check(res, ref)
    # a forward sweep that fails (Max_no_steps) still runs the adjoint over the taped steps and returns 1, as the reference
    ic = np.zeros(20, dtype=np.int32)
    ic[3] = 20
    res = F.integrate_adj("cbm4", T0, t1, y0, 4, rtol, atol, icntrl_u=ic, family="sdirk")
    ref = O.sdirk_adj("cbm4", y0, 4, T0, t1, atol, rtol, ic)
    assert (ref["ierr"] == 1).all() and (ref["rstatus"][:, 0] < t1).all()
    check(res, ref)
    ic = np.zeros(20, dtype=np.int32)
    ic[6] = 2                                    # Newton mode of the adjoint solve reads ATOL_adj / RTOL_adj
    with pytest.raises(F.FatodeError):
        F.integrate_adj("cbm4", T0, t1, y0, 4, rtol, atol, icntrl_u=ic, family="sdirk")


def newton_check(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"]) and np.array_equal(res.y, ref["y"])
    assert np.array_equal(res.lambda_, ref["lambda_"]), "the sweep follows the reference's operation order"
    if ref["mu"] is not None:
        sc = np.max(np.abs(ref["mu"]), axis=2, keepdims=True)
        floor = 1e-10 * np.max(sc, axis=1, keepdims=True)
        assert np.all(np.abs(res.mu - ref["mu"]) <= 1e-9 * np.maximum(sc, floor))


@pytest.mark.gpu
def test_sdirk_adj_simplified_newton_adjoint_stages_match_oracle():
    """ICNTRL(7) /= 1 (:890-951): one factorisation per step, the adjoint stages by simplified Newton iterations measured in the
    columns' own tolerances; columns that do not converge take the fall-back (new shared factors for the columns after them)."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    y0 = perturbed_states(F.initial_state("cbm4"), 6)
    t1 = T0 + 5 * 3600.0
    j = np.arange(32)[:, None]
    i = np.arange(32)[None, :]
    for k, method, nadj, with_mu, scalar in ((1.0e-4, 0, 32, True, 0), (1.0e-5, 2, 7, True, 0), (1.0e-3, 5, 32, False, 1), (1.0e-4, 3, 5, True, 1),
                                              (1.0e-2, 0, 32, False, 0)):  # the last one: forward sweep on the general-pivoting pass (Lambda only: the fall-backs blow the
                                              # sensitivities up to 1e39 there, Mu's regrouped sums then differ by more than 1e-9)
        rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
        ata = (F01 * k * 1e-2 * (1.0 + 0.5 * ((i + j) % 3)))[:nadj]
        rta = (F01 * k * (1.0 + 0.25 * ((i + 2 * j) % 2)))[:nadj]
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[2], ic[6] = scalar, method, 2
        ref = O.sdirk_adj("cbm4", y0, nadj, T0, t1, atol, rtol, ic, with_mu=with_mu, atol_adj=ata, rtol_adj=rta)
        res = F.integrate_adj("cbm4", T0, t1, y0, nadj, rtol, atol, icntrl_u=ic, want_mu=with_mu, family="sdirk", atol_adj=ata, rtol_adj=rta)
        assert (ref["ierr"] == 1).all()
        newton_check(res, ref)
        icd = ic.copy()
        icd[6] = 1
        d = O.sdirk_adj("cbm4", y0, nadj, T0, t1, atol, rtol, icd, with_mu=with_mu)
        assert not np.array_equal(d["istatus"], ref["istatus"])        # fewer factorisations, more solves than the direct mode
    # few iterations allowed: columns fail and take the fall-back path (more factorisations than one per step)
    rtol, atol = np.full(32, F01 * 1.0e-4), np.full(32, F01 * 1.0e-6)
    ic = np.zeros(20, dtype=np.int32)
    ic[4], ic[6] = 4, 2
    tt = np.full((32, 32), 1.0e-12)
    ref = O.sdirk_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic, atol_adj=tt, rtol_adj=tt)
    icf = ic.copy()
    icf[8] = 1                                    # forward sweep only (ICNTRL(9)): its counters
    res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic, family="sdirk", atol_adj=tt, rtol_adj=tt)
    assert np.array_equal(res.ierr, ref["ierr"]) and np.array_equal(res.istatus, ref["istatus"])
    steps = ref["istatus"][:, 3]
    assert (ref["istatus"][:, 5] > 2 * steps + ref["istatus"][:, 4]).any(), "no fall-back factorisation happened: the case does not test it"
    ok = ref["ierr"] == 1
    assert np.array_equal(res.lambda_[ok], ref["lambda_"][ok])


@pytest.mark.gpu
def test_sdirk_adj_general_pivoting_pass_at_loose_tolerances():
    """rtol 1e-3: every system leaves the static pivot pattern and takes the second pass with general partial pivoting."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(32, F01 * 1.0e-2), np.full(32, F01 * 1.0e-4)
    y0 = perturbed_states(F.initial_state("cbm4"), 20)
    t1 = T0 + 30 * 3600.0
    res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, family="sdirk")
    ref = O.sdirk_adj("cbm4", y0, 32, T0, t1, atol, rtol)
    assert (res.ierr == 1).all() and F.last_stats().cells_general > 0
    check(res, ref)
    assert np.array_equal(res.lambda_, ref["lambda_"])
    rtol, atol = np.full(32, F01 * 3.0e-3), np.full(32, F01 * 3.0e-5)
    res = F.integrate_adj("cbm4", T0, t1, y0, 7, rtol, atol, family="sdirk", want_mu=False)
    ref = O.sdirk_adj("cbm4", y0, 7, T0, t1, atol, rtol, with_mu=False)
    check(res, ref)
