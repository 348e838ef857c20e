"""GPU parity for the small example models on the per-thread kernel (BASELINE configs 0 and 1):
FWD/ROS INTEGRATE for roberts (N = 3) and small_strato (N = 5), one system per thread, dense DGETF2/DGETRS with partial
pivoting per thread.  Bar: state, RSTATUS and all ISTATUS counters bit-identical to the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
T1_STRATO = T0 + 30 * 24 * 3600.0


def strato_tols(level):
    """small_strato_dr.F90:27-38: RTOL = 0.1, ATOL = 1 multiplied `level` times by a REAL(4) 0.1"""
    rtol, atol = np.full(5, 1.0e-1), np.full(5, 1.0)
    for _ in range(level):
        rtol, atol = F01 * rtol, F01 * atol
    return rtol, atol


def assert_bitwise(res, ref):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"]), "state must be bit-identical to the oracle"


@pytest.mark.parametrize("level,method", [(3, 0), (5, 0), (4, 5), (3, 2)])
def test_small_strato_forward_batch_matches_oracle(level, method):
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = strato_tols(level)
    y0 = perturbed_states(F.initial_state("small_strato"), 192)
    assert np.array_equal(F.initial_state("small_strato"), O.initial_state("small_strato"))
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = method
    res = F.integrate("small_strato", T0, T1_STRATO, y0, rtol, atol, icntrl_u=ic)
    ref = O.ros_fwd("small_strato", y0, T0, T1_STRATO, atol, rtol, ic)
    assert_bitwise(res, ref)
    if method != 2:
        assert (res.ierr == 1).all()
    else:   # Ros3 runs into "step size too small" (IERR -7) for some of these cells in the reference algorithm too
        assert (ref["ierr"] == -7).sum() > 0
    assert np.all(res.istatus[:, 5] == 0) and np.all(res.istatus[:, 6] == 0)   # FWD/ROS never counts Ndec/Nsol


def test_roberts_forward_matches_oracle():
    """BASELINE config 0: y0 = (1,0,0), tend = 4e7, Rodas4, autonomous, Hstart = 1e-3 (REAL(4)), float-rounded tolerances;
    plus a batch of perturbed systems and the other tableaux."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    rtol, atol = np.full(3, float(np.float32(1e-4))), np.full(3, float(np.float32(1e-6)))
    rc = np.zeros(20)
    rc[2] = float(np.float32(1e-3))
    y0 = perturbed_states(np.array([1.0, 0.0, 0.0]), 64)
    y0[1:, 1] = 1e-6 * np.abs(y0[1:, 0] - 1.0)          # some systems start off the invariant manifold
    for method in (5, 0, 1, 2, 3, 4):
        ic = np.zeros(20, dtype=np.int32)
        ic[0] = 1
        ic[2] = method
        res = F.integrate("roberts", 0.0, 4.0e7, y0, rtol, atol, icntrl_u=ic, rcntrl_u=rc)
        ref = O.ros_fwd("roberts", y0, 0.0, 4.0e7, atol, rtol, ic, rc)
        assert_bitwise(res, ref)
    assert res.ierr[0] == 1


def test_small_strato_one_million_cells():
    """BASELINE config 1 at full size: 2^20 cells on one GPU.  The second half of the batch repeats the first half, so
    equal inputs must give equal bits wherever they sit; a strided sample is compared with the oracle."""
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    n = 1 << 20
    rtol, atol = strato_tols(3)
    half = perturbed_states(F.initial_state("small_strato"), n // 2)
    y0 = np.concatenate([half, half])
    res = F.integrate("small_strato", T0, T0 + 3 * 24 * 3600.0, y0, rtol, atol)
    st = F.last_stats()
    assert (res.ierr == 1).all()
    assert np.array_equal(res.y[: n // 2], res.y[n // 2:]) and np.array_equal(res.istatus[: n // 2], res.istatus[n // 2:])
    idx = np.arange(0, n // 2, 4099)
    ref = O.ros_fwd("small_strato", y0[idx], T0, T0 + 3 * 24 * 3600.0, atol, rtol)
    assert np.array_equal(res.y[idx], ref["y"]) and np.array_equal(res.istatus[idx], ref["istatus"])
    print(f"small_strato 2^20 cells x 3 days: forward kernel {st.ms_forward:.1f} ms -> {n / st.ms_forward * 1e3:.3g} cells/s")


def test_small_strato_adjoint_matches_oracle_and_reference_matrix():
    """INTEGRATE_ADJ (RosenbrockADJ2: Lambda only) on small_strato, the reference's small_ros_adj example (NADJ = 5, default
    Rodas3, rtol 1e-4, atol 1e-3, 3 days; golden matrix rkadj_ADJ_results.m, which pins the path to integration-tolerance level
    only — see tests/test_oracle_golden.py).  One system per thread with the reference's operation order: state, ISTATUS, RSTATUS
    and Lambda are compared bit for bit with the CPU restatement."""
    import json
    import os
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    t1 = T0 + 3 * 24 * 3600.0
    y0 = perturbed_states(F.initial_state("small_strato"), 200)
    rtol, atol = np.full(5, 1e-4), np.full(5, 1e-3)
    res = F.integrate_adj("small_strato", T0, t1, y0, 5, rtol, atol, want_mu=False)
    ref = O.ros_adj("small_strato", y0, 5, T0, t1, atol, rtol, with_mu=False)
    assert (res.ierr == 1).all()
    assert_bitwise(res, ref)
    assert np.array_equal(res.lambda_, ref["lambda_"])
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_strato.json")))
    glam = np.array(g["lambda"]).reshape(5, 5)
    err = np.max(np.abs(res.lambda_[0] - glam), axis=1) / np.max(np.abs(glam), axis=1)
    assert np.all(err < 5e-5), err
    # every tableau, autonomous flag, fewer columns, tight tolerances (long tapes)
    for method, auto, nadj in ((1, 0, 5), (2, 1, 3), (3, 0, 1), (4, 0, 5), (5, 1, 2)):
        ic = np.zeros(20, dtype=np.int32)
        ic[0], ic[2] = auto, method
        r2 = F.integrate_adj("small_strato", T0, T0 + 12 * 3600.0, y0[:40], nadj, rtol * 1e-2, atol * 1e-2, icntrl_u=ic, want_mu=False)
        o2 = O.ros_adj("small_strato", y0[:40], nadj, T0, T0 + 12 * 3600.0, atol * 1e-2, rtol * 1e-2, ic, with_mu=False)
        assert_bitwise(r2, o2)
        assert np.array_equal(r2.lambda_, o2["lambda_"]), (method, auto, nadj)


def test_roberts_adjoint_with_parameter_sensitivities_matches_oracle():
    """INTEGRATE_ADJ (RosenbrockADJ1: Lambda and Mu, NP = 3) on Roberts, EXAMPLES/roberts/ROBERTS_ROS_ADJ without the quadrature
    terms: autonomous Rodas4; failed forward sweeps return before ADJINIT with the caller's Lambda / Mu untouched."""
    import fatode_b200 as F
    y0 = np.tile(np.array([1.0, 0.0, 0.0]), (16, 1)) * (1.0 + 0.01 * np.arange(16))[:, None]
    rtol, atol = np.full(3, 1e-5), np.array([1e-8, 1e-14, 1e-6])
    ic = np.zeros(20, dtype=np.int32)
    ic[0], ic[2] = 1, 5
    rc = np.zeros(20)
    rc[2] = float(np.float32(1e-3))
    for tend in (40.0, 4.0e3):
        res = F.integrate_adj("roberts", 0.0, tend, y0, 3, rtol, atol, icntrl_u=ic, rcntrl_u=rc, want_mu_sum=True)
        ref = O.ros_adj("roberts", y0, 3, 0.0, tend, atol, rtol, ic, rc)
        assert (res.ierr == 1).all()
        assert_bitwise(res, ref)
        assert np.array_equal(res.lambda_, ref["lambda_"]) and np.array_equal(res.mu, ref["mu"])
        assert np.allclose(res.mu_sum, res.mu.sum(axis=0), rtol=1e-12, atol=0.0)
    ic[3] = 5                                              # Max_no_steps = 5: IERR -6 for every system
    res = F.integrate_adj("roberts", 0.0, 4.0e3, y0, 3, rtol, atol, icntrl_u=ic, rcntrl_u=rc)
    ref = O.ros_adj("roberts", y0, 3, 0.0, 4.0e3, atol, rtol, ic, rc)
    assert (res.ierr == -6).all()
    assert_bitwise(res, ref)
    assert not res.lambda_.any()
