"""TLM/RK_TLM (first half): tangent linear model of the 3-stage implicit Runge-Kutta methods in the
reference's preset mode (direct TLM solve: one 3N x 3N DGETRF per accepted step, TLM/RK_TLM/RK_TLM_f90_Integrator.F90:761-793).
PARITY UNPINNED by reference artefacts (no CBM-IV driver, no golden file for this family): the CPU part cross-checks the
restatement (oracle/rk_tlm.hpp) against its siblings; the GPU part holds the product to the restatement bit for bit."""
import numpy as np
import pytest

import oracle_lib as O

F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0


def tols(rt=1.0e-5, at=1.0e-7):
    return np.full(32, F01 * rt), np.full(32, F01 * at)


def method(m):
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = m
    return ic


def test_oracle_state_sweep_counters_and_duality_with_the_direct_adjoint():
    rtol, atol = tols()
    y0 = O.initial_state("cbm4")
    t1 = T0 + 6 * 3600.0
    eye = np.eye(32)[None]
    r = O.rk_tlm("cbm4", y0, eye, T0, t1, atol, rtol, method(1), nthreads=1)
    assert r["ierr"][0] == 1
    # the state sweep is the RK_ADJ forward sweep without LU reuse: when no step reuses the LU the two are the same run
    fwd_ic = method(1)
    fwd_ic[8] = 1
    f = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, fwd_ic, nthreads=1)
    I, Fw = r["istatus"][0], f["istatus"][0]
    A = I[3]
    assert A == Fw[3] and I[4] == Fw[4] and I[2] == Fw[2] and np.array_equal(r["y"], f["y"])
    # counters (TLM/RK_TLM :768, 773, 782, 1120-1134): Njac += 3 A, Ndec += A, Nsol += NTLM A, Nfun += 3 per Newton iteration
    assert I[1] == Fw[1] + 3 * A and I[5] == Fw[5] + A and I[6] == Fw[6] + 32 * A
    assert I[0] > Fw[0] and (I[0] - Fw[0]) % 3 == 0
    # duality: the discrete adjoint with the direct solve is the exact transpose of the direct TLM of the same steps
    dir_ic = method(1)
    dir_ic[6] = 2
    a = O.rk_adj("cbm4", y0, 32, T0, t1, atol, rtol, dir_ic, nthreads=1)
    yt, lam = r["y_tlm"][0], a["lambda_"][0]
    assert np.max(np.abs(yt - lam.T)) <= 1e-12 * np.max(np.abs(lam))
    # exact linearity in Y_tlm(t0)
    v = np.sin(1.0 + np.arange(32.0))[None, None, :]
    rv = O.rk_tlm("cbm4", y0, v, T0, t1, atol, rtol, method(1), nthreads=1)
    ref = np.einsum("j,jk->k", v[0, 0], yt)
    assert np.max(np.abs(rv["y_tlm"][0, 0] - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_oracle_against_finite_differences_and_the_sdirk_tlm():
    rtol, atol = tols(1.0e-6, 1.0e-8)
    y0 = O.initial_state("cbm4")
    t1 = T0 + 2 * 3600.0
    eye = np.eye(32)[None]
    r = O.rk_tlm("cbm4", y0, eye, T0, t1, atol, rtol, method(1), nthreads=1)
    s = O.sdirk_tlm("cbm4", y0, eye, T0, t1, atol, rtol, nthreads=1)
    yt, st = r["y_tlm"][0], s["y_tlm"][0]
    sc = np.max(np.abs(yt))
    assert np.max(np.abs(yt - st)) <= 2e-4 * sc                     # two methods, integration accuracy
    j = 25                                                           # a species every other one reacts to
    d = 1e-6 * max(abs(y0[j]), 1e-12)
    yp, ym = y0.copy(), y0.copy()
    yp[j] += d
    ym[j] -= d
    fwd_ic = method(1)
    fwd_ic[8] = 1
    fp = O.rk_adj("cbm4", yp, 1, T0, t1, atol, rtol, fwd_ic, nthreads=1)["y"][0]
    fm = O.rk_adj("cbm4", ym, 1, T0, t1, atol, rtol, fwd_ic, nthreads=1)["y"][0]
    fd = (fp - fm) / (2 * d)
    assert np.max(np.abs(fd - yt[j])) <= 5e-3 * np.max(np.abs(yt[j]))


def test_oracle_options():
    rtol, atol = tols()
    y0 = O.initial_state("cbm4")
    eye = np.eye(32)[None, :2]
    assert O.rk_tlm("cbm4", y0, eye, T0, T0 + 600.0, atol, rtol, method(5))["ierr"][0] == -13      # no Lobatto-3A
    ic = method(1)
    ic[11] = 1
    assert O.rk_tlm("cbm4", y0, eye, T0, T0 + 600.0, atol, rtol, ic)["ierr"][0] == -100           # ICNTRL(12): not restated
    ic = method(1)
    ic[3] = 3                                                                                   # Max_no_steps = 3: IERR -9, columns advanced
    r = O.rk_tlm("cbm4", y0, eye, T0, T0 + 7200.0, atol, rtol, ic)
    assert r["ierr"][0] == -9 and not np.array_equal(r["y_tlm"][0], eye[0])


@pytest.mark.gpu
def test_cbm4_rk_tlm_batch_matches_oracle():
    """State, RSTATUS, every ISTATUS counter and Y_tlm bit-identical to the restatement: all four methods, NTLM = 32 and 5,
    the general-pivoting pass at loose tolerances, a failing run (Max_no_steps)."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    y0 = perturbed_states(F.initial_state("cbm4"), 16)
    eye = np.broadcast_to(np.eye(32), (16, 32, 32)).copy()
    for (m, hours, rt, at, ntlm) in ((1, 10.0, 1.0e-5, 1.0e-7, 32), (0, 4.0, 1.0e-5, 1.0e-7, 32), (3, 1.0, 1.0e-5, 1.0e-7, 5),
                                     (4, 4.0, 1.0e-5, 1.0e-7, 7), (1, 10.0, 1.0e-3, 1.0e-5, 32)):
        rtol, atol = tols(rt, at)
        t1 = T0 + hours * 3600.0
        yt0 = eye[:, :ntlm].copy()
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt0, rtol, atol, icntrl_u=method(m), family="rk")
        ref = O.rk_tlm("cbm4", y0, yt0, T0, t1, atol, rtol, method(m))
        assert np.array_equal(res.ierr, ref["ierr"]) and (res.ierr == 1).all(), (m, res.ierr)
        assert np.array_equal(res.y, ref["y"]) and np.array_equal(res.rstatus, ref["rstatus"])
        assert np.array_equal(res.istatus, ref["istatus"]), (m, res.istatus[0], ref["istatus"][0])
        assert np.array_equal(res.y_tlm, ref["y_tlm"]), (m, np.max(np.abs(res.y_tlm - ref["y_tlm"])))
    assert F.last_stats().cells_general > 0
    rtol, atol = tols()
    ic = method(1)
    ic[3] = 4                                                     # Max_no_steps = 4: IERR -9 with the columns of the accepted steps
    res = F.integrate_tlm("cbm4", T0, T0 + 7200.0, y0, eye, rtol, atol, icntrl_u=ic, family="rk")
    ref = O.rk_tlm("cbm4", y0, eye, T0, T0 + 7200.0, atol, rtol, ic)
    assert (res.ierr == -9).all() and np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.y, ref["y"]) and np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.y_tlm, ref["y_tlm"])
    ic = method(1)
    ic[11] = 1
    with pytest.raises(Exception):
        F.integrate_tlm("cbm4", T0, T0 + 600.0, y0, eye, rtol, atol, icntrl_u=ic, family="rk")
