"""Unit-level GPU parity: the device model functions and the warp LU/solve must reproduce the
oracle's arithmetic bit for bit (this is what makes the forward sweep's ISTATUS reproducible)."""
import ctypes

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def d(a):
    return a.ctypes.data_as(dp)


def i(a):
    return a.ctypes.data_as(ip)


def test_cbm4_fun_jac_hessian_bitwise():
    import fatode_b200 as F
    from fatode_b200.inputs import perturbed_states
    L, OL = F.lib(), O.lib()
    ys = perturbed_states(F.initial_state("cbm4"), 12)
    # times spanning night, sunrise, noon, sunset (SUN = 0, ramp, 1, ramp)
    times = [43200.0, 50000.123, 70000.0, 100000.5, 16200.0 + 86400, 16200.1 + 86400, 70199.9, 70200.0, 90000.0,
             129600.0, 200000.7, 302399.0]
    for y, t in zip(ys, times):
        f, J, H = np.zeros(32), np.zeros(1024), np.zeros(258)
        assert L.fatode_probe_cbm4(ctypes.c_double(t), d(y), d(f), d(J), d(H)) == 0, L.fatode_last_error()
        fo, Jo, Ho = np.zeros(32), np.zeros(1024), np.zeros(258)
        OL.oracle_cbm4_fun_jac(ctypes.c_double(t), d(y), d(fo), d(Jo), d(Ho))
        assert np.array_equal(f, fo), t
        assert np.array_equal(J, Jo), t
        assert np.array_equal(H, Ho), t


def _lu_both(A, b):
    import fatode_b200 as F
    L, OL = F.lib(), O.lib()
    Af = np.ascontiguousarray(A.ravel(order="F"))
    lu, who, x, info = np.zeros(1024), np.zeros(32, dtype=np.int32), np.zeros(32), np.zeros(1, dtype=np.int32)
    assert L.fatode_probe_lu32(d(Af), d(b), d(lu), i(who), d(x), i(info)) == 0, L.fatode_last_error()
    a = Af.copy()
    ipiv = np.zeros(32, dtype=np.int32)
    xn, xt = np.zeros(32), np.zeros(32)
    inf = OL.oracle_dgetf2_solve(32, d(a), i(ipiv), d(b), d(xn), d(xt))
    return lu.reshape(32, 32), x, info[0], a.reshape(32, 32).T, xn, inf, ipiv


def test_warp_lu_matches_netlib_dgetf2_dgetrs_bitwise():
    rng = np.random.default_rng(5)
    for trial in range(6):
        A = rng.standard_normal((32, 32))
        if trial % 2:
            A[rng.random((32, 32)) < 0.7] = 0.0          # sparse, forces real pivoting
            A += np.diag(rng.standard_normal(32) * 0.1)
        if trial == 4:
            A[5, :] = A[9, :]                             # ties in the pivot search
        b = rng.standard_normal(32)
        lu, x, info, lu_ref, x_ref, info_ref, ipiv = _lu_both(A, b)
        assert info == info_ref
        assert np.array_equal(lu, lu_ref)
        if info == 0:
            assert np.array_equal(x, x_ref)


def test_warp_lu_reports_singular_like_dgetf2():
    A = np.eye(32)
    A[7, 7] = 0.0
    b = np.ones(32)
    lu, x, info, lu_ref, x_ref, info_ref, _ = _lu_both(A, b)
    assert info == info_ref == 8
