"""CPU check of the oracle's ROS_TLM restatement (oracle/ros.hpp::tlm, TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707).
The reference ships no CBM-IV TLM driver and no golden file, so the restatement is validated against the golden-pinned
adjoint oracle through the adjoint-TLM duality: on the same step sequence (forward error control) and with the
time-derivative terms off, the discrete adjoint is the exact transpose of the tangent linear model."""
import numpy as np

import oracle_lib as O

F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0


def test_tlm_is_transpose_of_adjoint_autonomous():
    rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
    y0 = O.initial_state("cbm4") * (1.0 + 0.05 * np.cos(np.arange(32.0)))
    ic = np.zeros(20, dtype=np.int32)
    ic[0], ic[2] = 1, 5
    t1 = T0 + 5 * 3600.0
    tl = O.ros_tlm("cbm4", y0, np.eye(32)[None], T0, t1, atol, rtol, ic)
    ad = O.ros_adj("cbm4", y0, 32, T0, t1, atol, rtol, ic, with_mu=False)
    assert tl["ierr"][0] == 1 and np.array_equal(tl["y"], ad["y"])
    lam, ytl = ad["lambda_"][0], tl["y_tlm"][0].T
    sc = np.max(np.abs(lam), axis=1)
    ok = sc > 1e-10 * sc.max()
    err = (np.max(np.abs(ytl - lam), axis=1) / np.maximum(sc, 1e-300))[ok]
    assert err.max() < 1e-10, err.max()
    # ISTATUS closed forms (autonomous): Nfun = A + att*n_newf, Njac = A + att*n_newf, Nsol = att*S*(1+NTLM)
    I = tl["istatus"][0].astype(np.int64)
    A, att = I[3], I[2]
    assert I[0] == A + 5 * att and I[1] == A + 5 * att and I[5] == att and I[6] == att * 6 * 33


def test_tlm_truncation_error_control_changes_the_step_sequence():
    """ICNTRL(12) = 1 (the preset when ICNTRL_U is absent): the TLM columns enter the error norm."""
    rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
    y0 = O.initial_state("cbm4")
    t1 = T0 + 2 * 3600.0
    yt = np.eye(32)[None]
    tt = np.full((32, 32), 1e-4), np.full((32, 32), 1e-3)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    a = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    ic[11] = 1
    b = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=tt[1], rtol_tlm=tt[0])
    assert a["ierr"][0] == 1 and b["ierr"][0] == 1
    assert b["istatus"][0][2] >= a["istatus"][0][2]
