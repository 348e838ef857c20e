"""Continuous Rosenbrock adjoints of ADJ/ROS_ADJ, as the reference BEHAVES:
  * ICNTRL(7) = 3 (ros_CadjInt, ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1447): INTEGRATE_ADJ calls it as
    ros_CadjInt(NADJ, Lambda, Tend, Tstart, ...) (:588-592), so H = Direction*H is negative and the first statement of its time
    loop, IF (((T+0.1d0*H) == T).OR.(H <= Roundoff)) (:1517), returns IERR = -7 before anything is integrated: the caller gets
    the forward solution, ADJINIT's Lambda / Mu, the forward counters plus the last checkpoint's FUN / JAC, RSTATUS(Nhexit) = 0.
  * ICNTRL(7) = 4 (ros_SimpleCadjInt :1675): the first stage multiplies by the (negated) Jacobian itself (LSS_Mul_Jac :1741) where
    the later stages use its transpose (LSS_Mul_Jactr1 :1781); the recursion overflows within tens of steps.  Restated in the
    oracle (simple_cadjoint) to show it; the product returns FATODE_EUNSUPPORTED for it.
The quadrature path is not restated: its weights ros_W come from a loop that never executes."""
import numpy as np
import pytest

import oracle_lib as O

T0 = 12.0 * 3600.0
F01 = float(np.float32(0.1))


def ic_of(method, adjtype, cadj_method=0):
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = method
    ic[5] = cadj_method
    ic[6] = adjtype
    return ic


def test_oracle_continuous_adjoint_stops_with_minus_seven_after_adjinit():
    y0 = O.initial_state("small_strato")
    rtol, atol = np.full(5, 1e-6), np.full(5, 1e-2)
    t1 = T0 + 6 * 3600.0
    fwd = O.ros_adj("small_strato", y0, 5, T0, t1, atol, rtol, ic_of(3, 1), with_mu=False, nthreads=1)      # ICNTRL(7) = 1: no adjoint
    c = O.ros_adj("small_strato", y0, 5, T0, t1, atol, rtol, ic_of(3, 3), with_mu=False, nthreads=1)
    assert fwd["ierr"][0] == 1 and c["ierr"][0] == -7
    assert np.array_equal(c["y"], fwd["y"]) and np.array_equal(c["lambda_"][0], np.eye(5))
    I, Fw = c["istatus"][0], fwd["istatus"][0]
    assert I[0] == Fw[0] + 2 and I[1] == Fw[1] + 1 and np.array_equal(I[2:], Fw[2:])       # non-autonomous: FUN + time derivative, JAC
    assert c["rstatus"][0][1] == 0.0 and c["rstatus"][0][0] == fwd["rstatus"][0][0]
    r = O.ros_adj("roberts", O.initial_state("roberts"), 3, 0.0, 40.0, np.array([1e-8, 1e-14, 1e-6]), np.full(3, 1e-4), ic_of(5, 3), nthreads=1)
    y = r["y"][0]
    assert r["ierr"][0] == -7 and np.array_equal(np.diag(r["lambda_"][0]), y)              # roberts' ADJINIT: Lambda = diag(y)


def test_oracle_simplified_continuous_adjoint_overflows_like_the_reference_routine():
    y0 = O.initial_state("small_strato")
    rtol, atol = np.full(5, 1e-6), np.full(5, 1e-2)
    d = O.ros_adj("small_strato", y0, 5, T0, T0 + 180.0, atol, rtol, ic_of(3, 2), with_mu=False, nthreads=1)
    s = O.ros_adj("small_strato", y0, 5, T0, T0 + 180.0, atol, rtol, ic_of(3, 4), with_mu=False, nthreads=1)
    assert d["ierr"][0] == 1 and s["ierr"][0] == 1 and np.array_equal(d["y"], s["y"])
    assert np.max(np.abs(d["lambda_"][0])) < 10.0
    big = np.abs(s["lambda_"][0])
    assert (not np.all(np.isfinite(big))) or np.max(big) > 1e50, "J instead of J^T in the first stage (:1741): the recursion explodes"
    # the checkpoints and counters of the mode are what the Fortran produces: one JAC per step + one per new-function stage, one
    # decomposition per step, S solves per column and step
    I, D = s["istatus"][0], d["istatus"][0]
    A = I[3]
    assert A == D[3] and I[5] == (D[5] - A) + A and I[6] == (D[6] - A * 4 * 5) + A * 4 * 5


@pytest.mark.gpu
def test_gpu_continuous_adjoint_modes_behave_like_the_reference():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    # cbm4, ICNTRL(7) = 3
    y0 = perturbed_states(F.initial_state("cbm4"), 8)
    rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
    t1 = T0 + 6 * 3600.0
    res = F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic_of(5, 3))
    ref = O.ros_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic_of(5, 3))
    assert (res.ierr == -7).all() and np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.y, ref["y"]) and np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.lambda_, ref["lambda_"]) and np.array_equal(res.mu, ref["mu"])
    # small models
    ys = perturbed_states(F.initial_state("small_strato"), 6)
    rs, as_ = np.full(5, 1e-6), np.full(5, 1e-2)
    res = F.integrate_adj("small_strato", T0, t1, ys, 5, rs, as_, icntrl_u=ic_of(3, 3), want_mu=False)
    ref = O.ros_adj("small_strato", ys, 5, T0, t1, as_, rs, ic_of(3, 3), with_mu=False)
    assert (res.ierr == -7).all() and np.array_equal(res.y, ref["y"]) and np.array_equal(res.istatus, ref["istatus"])
    assert np.array_equal(res.rstatus, ref["rstatus"]) and np.array_equal(res.lambda_, ref["lambda_"])
    yr = np.tile(F.initial_state("roberts"), (4, 1))
    rr, ar = np.full(3, 1e-4), np.array([1e-8, 1e-14, 1e-6])
    res = F.integrate_adj("roberts", 0.0, 40.0, yr, 3, rr, ar, icntrl_u=ic_of(5, 3))
    ref = O.ros_adj("roberts", yr, 3, 0.0, 40.0, ar, rr, ic_of(5, 3))
    assert (res.ierr == -7).all() and np.array_equal(res.y, ref["y"]) and np.array_equal(res.istatus, ref["istatus"])
    assert np.array_equal(res.lambda_, ref["lambda_"]) and np.array_equal(res.mu, ref["mu"])
    with pytest.raises(Exception):
        F.integrate_adj("cbm4", T0, t1, y0, 32, rtol, atol, icntrl_u=ic_of(5, 4))
