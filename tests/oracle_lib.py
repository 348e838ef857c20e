"""ctypes view of the CPU oracle (oracle/liboracle.so) — checker only, used by tests."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
MODEL = {"roberts": 1, "small_strato": 2, "cbm4": 3, "swe": 4}
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
        L = ctypes.CDLL(so)
        L.oracle_ros_fwd_batch.argtypes = [ctypes.c_int, ctypes.c_long, _dp, ctypes.c_double, ctypes.c_double, _dp, _dp,
                                           _ip, _dp, _ip, _dp, _ip, ctypes.c_int]
        L.oracle_ros_adj_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp,
                                           ctypes.c_double, ctypes.c_double, _dp, _dp, _ip, _dp, _ip, _dp, _ip,
                                           ctypes.c_int]
        L.oracle_rk_adj_batch.argtypes = L.oracle_ros_adj_batch.argtypes
        L.oracle_sdirk_adj_batch.argtypes = L.oracle_ros_adj_batch.argtypes
        L.oracle_rk_fwd_batch.argtypes = [ctypes.c_int, ctypes.c_long, _dp, ctypes.c_double, ctypes.c_double, _dp, _dp,
                                          _ip, _dp, _ip, _dp, _ip, ctypes.c_int, ctypes.c_int]
        L.oracle_sdirk_fwd_batch.argtypes = [ctypes.c_int, ctypes.c_long, _dp, ctypes.c_double, ctypes.c_double, _dp, _dp,
                                             _ip, _dp, _ip, _dp, _ip, ctypes.c_int]
        L.oracle_ros_tlm_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                           _dp, _dp, _dp, _dp, _ip, _dp, _ip, _dp, _ip, ctypes.c_int]
        L.oracle_sdirk_tlm_batch.argtypes = L.oracle_ros_tlm_batch.argtypes
        L.oracle_rk_tlm_batch.argtypes = L.oracle_ros_tlm_batch.argtypes
        L.oracle_erk_batch.argtypes = L.oracle_ros_tlm_batch.argtypes
        L.oracle_swe_initial_state.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp]
        L.oracle_swe_fun.argtypes = [_dp, _dp]
        L.oracle_swe_jac_mul.argtypes = [_dp, _dp, ctypes.c_int, _dp]
        L.oracle_erk_adj_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                           _dp, _dp, ctypes.POINTER(ctypes.c_int), _dp, ctypes.POINTER(ctypes.c_int), _dp,
                                           ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def dims(model):
    n, p = ctypes.c_int(), ctypes.c_int()
    lib().oracle_model_dims(MODEL[model], ctypes.byref(n), ctypes.byref(p))
    return n.value, p.value


def initial_state(model):
    n, _ = dims(model)
    y = np.zeros(n)
    lib().oracle_initial_state(MODEL[model], _d(y))
    return y


def _ctrl(icntrl, rcntrl):
    ic = np.zeros(20, dtype=np.int32) if icntrl is None else np.ascontiguousarray(icntrl, dtype=np.int32)
    rc = np.zeros(20) if rcntrl is None else np.ascontiguousarray(rcntrl, dtype=np.float64)
    return ic, rc


def ros_fwd(model, y, tin, tout, atol, rtol, icntrl=None, rcntrl=None, nthreads=0):
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc = y.shape[0]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lib().oracle_ros_fwd_batch(MODEL[model], nc, _d(y), tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst),
                               _i(ierr), nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr)


def ros_adj(model, y, nadj, tin, tout, atol, rtol, icntrl=None, rcntrl=None, with_mu=True, nthreads=0):
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    _, npar = dims(model)
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    use_mu = with_mu and npar > 0
    mu = np.zeros((nc, nadj, max(npar, 1)))
    lib().oracle_ros_adj_batch(MODEL[model], nc, npar if use_mu else 0, nadj, _d(y), _d(lam), _d(mu) if use_mu else None,
                               tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst), _i(ierr), nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr, lambda_=lam, mu=mu[:, :, :npar] if use_mu else None)


def ros_tlm(model, y, y_tlm, tin, tout, atol, rtol, icntrl=None, rcntrl=None, atol_tlm=None, rtol_tlm=None, nthreads=0):
    """INTEGRATE_TLM (TLM/ROS_TLM); y_tlm: [ncells, NTLM, N].  icntrl=None means ICNTRL_U absent (presets apply)."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    yt = np.array(y_tlm, dtype=np.float64, order="C").reshape(nc, -1, n)
    ntlm = yt.shape[1]
    ic = None if icntrl is None else np.ascontiguousarray(icntrl, dtype=np.int32)
    rc = np.zeros(20) if rcntrl is None else np.ascontiguousarray(rcntrl, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    at = None if atol_tlm is None else np.ascontiguousarray(atol_tlm, dtype=np.float64)
    rt = None if rtol_tlm is None else np.ascontiguousarray(rtol_tlm, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    r = lib().oracle_ros_tlm_batch(MODEL[model], nc, ntlm, _d(y), _d(yt), tin, tout, None if at is None else _d(at),
                                   None if rt is None else _d(rt), _d(atol), _d(rtol), None if ic is None else _i(ic), _d(rc),
                                   _i(ist), _d(rst), _i(ierr), nthreads)
    assert r == 0
    return dict(y=y, y_tlm=yt, istatus=ist, rstatus=rst, ierr=ierr)


def sdirk_fwd(model, y, tin, tout, atol, rtol, icntrl=None, rcntrl=None, nthreads=0):
    """FWD/SDIRK INTEGRATE looped over cells."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc = y.shape[0]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lib().oracle_sdirk_fwd_batch(MODEL[model], nc, _d(y), tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst),
                                 _i(ierr), nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr)


def rk_fwd(model, y, tin, tout, atol, rtol, icntrl=None, rcntrl=None, raw=False, nthreads=0):
    """FWD/RK INTEGRATE looped over cells (raw=True: RungeKutta called directly, no INTEGRATE presets)."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc = y.shape[0]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lib().oracle_rk_fwd_batch(MODEL[model], nc, _d(y), tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst),
                              _i(ierr), 1 if raw else 0, nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr)


def rk_adj(model, y, nadj, tin, tout, atol, rtol, icntrl=None, rcntrl=None, with_mu=True, nthreads=0):
    """INTEGRATE_ADJ of ADJ/RK_ADJ looped over cells; lambda_: [ncells, NADJ, N], mu: [ncells, NADJ, NP]."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    _, npar = dims(model)
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    use_mu = with_mu and npar > 0
    mu = np.zeros((nc, nadj, max(npar, 1)))
    lib().oracle_rk_adj_batch(MODEL[model], nc, npar if use_mu else 0, nadj, _d(y), _d(lam), _d(mu) if use_mu else None,
                              tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst), _i(ierr), nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr, lambda_=lam, mu=mu[:, :, :npar] if use_mu else None)


def sdirk_tlm(model, y, y_tlm, tin, tout, atol, rtol, icntrl=None, rcntrl=None, atol_tlm=None, rtol_tlm=None, nthreads=0):
    """INTEGRATE_TLM (TLM/SDIRK_TLM); y_tlm: [ncells, NTLM, N]."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    yt = np.array(y_tlm, dtype=np.float64, order="C").reshape(nc, -1, n)
    ntlm = yt.shape[1]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    at = None if atol_tlm is None else np.ascontiguousarray(atol_tlm, dtype=np.float64)
    rt = None if rtol_tlm is None else np.ascontiguousarray(rtol_tlm, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    r = lib().oracle_sdirk_tlm_batch(MODEL[model], nc, ntlm, _d(y), _d(yt), tin, tout, None if at is None else _d(at),
                                     None if rt is None else _d(rt), _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst),
                                     _i(ierr), nthreads)
    assert r == 0
    return dict(y=y, y_tlm=yt, istatus=ist, rstatus=rst, ierr=ierr)


def rk_tlm(model, y, y_tlm, tin, tout, atol, rtol, icntrl=None, rcntrl=None, nthreads=0):
    """INTEGRATE_TLM (TLM/RK_TLM, preset direct TLM solve); y_tlm: [ncells, NTLM, N]."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    yt = np.array(y_tlm, dtype=np.float64, order="C").reshape(nc, -1, n)
    ntlm = yt.shape[1]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    r = lib().oracle_rk_tlm_batch(MODEL[model], nc, ntlm, _d(y), _d(yt), tin, tout, None, None, _d(atol), _d(rtol), _i(ic), _d(rc),
                                  _i(ist), _d(rst), _i(ierr), nthreads)
    assert r == 0
    return dict(y=y, y_tlm=yt, istatus=ist, rstatus=rst, ierr=ierr)


SWE_N = 4800


def swe_initial_state(amplitude=30.0, u0=2.0, v0=2.0):
    """Gaussian bump of the SWE drivers (EXAMPLES/swe/SWE_ERK/swe2D_erk_dr.F90:100-103)."""
    y = np.zeros(SWE_N)
    lib().oracle_swe_initial_state(float(amplitude), float(u0), float(v0), _d(y))
    return y


def swe_fun(y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    p = np.zeros(SWE_N)
    lib().oracle_swe_fun(_d(y), _d(p))
    return p


def swe_jac_mul(y, x, trans=False):
    """JAC(y) x or JAC(y)^T x with the JAC callback of the shallow-water drivers (compute_JacF, swe2D_erk_adj_dr.F90:2-13)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.zeros(SWE_N)
    lib().oracle_swe_jac_mul(_d(y), _d(x), 1 if trans else 0, _d(o))
    return o


def erk_adj(model, y, nadj, tin, tout, atol, rtol, icntrl=None, rcntrl=None, nthreads=0):
    """INTEGRATE_ADJ of ADJ/ERK_ADJ (no parameters) looped over members; lambda_: [ncells, NADJ, N], ADJINIT: Lambda(k,k) = 1."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    r = lib().oracle_erk_adj_batch(MODEL[model], nc, nadj, _d(y), _d(lam), tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist),
                                   _d(rst), _i(ierr), nthreads)
    assert r == 0
    return dict(y=y, lambda_=lam, istatus=ist, rstatus=rst, ierr=ierr)


def erk(model, y, tin, tout, atol, rtol, icntrl=None, rcntrl=None, y_tlm=None, atol_tlm=None, rtol_tlm=None, nthreads=0):
    """INTEGRATE (FWD/ERK) or, with y_tlm [ncells, NTLM, N], INTEGRATE_TLM (TLM/ERK_TLM, finite-difference mode)."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    yt = None if y_tlm is None else np.array(y_tlm, dtype=np.float64, order="C").reshape(nc, -1, n)
    ntlm = 0 if yt is None else yt.shape[1]
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    at = None if atol_tlm is None else np.ascontiguousarray(atol_tlm, dtype=np.float64)
    rt = None if rtol_tlm is None else np.ascontiguousarray(rtol_tlm, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    r = lib().oracle_erk_batch(MODEL[model], nc, ntlm, _d(y), None if yt is None else _d(yt), tin, tout,
                               None if at is None else _d(at), None if rt is None else _d(rt), _d(atol), _d(rtol), _i(ic), _d(rc),
                               _i(ist), _d(rst), _i(ierr), nthreads)
    assert r == 0
    return dict(y=y, y_tlm=yt, istatus=ist, rstatus=rst, ierr=ierr)


def sdirk_adj(model, y, nadj, tin, tout, atol, rtol, icntrl=None, rcntrl=None, with_mu=True, nthreads=0, atol_adj=None, rtol_adj=None):
    """INTEGRATE_ADJ of ADJ/SDIRK_ADJ looped over cells; lambda_: [ncells, NADJ, N], mu: [ncells, NADJ, NP].
    atol_adj / rtol_adj [NADJ, N]: read by the simplified-Newton adjoint stages (default: atol / rtol for every column, as the drivers)."""
    y = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = y.shape
    _, npar = dims(model)
    ic, rc = _ctrl(icntrl, rcntrl)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    use_mu = with_mu and npar > 0
    mu = np.zeros((nc, nadj, max(npar, 1)))
    if atol_adj is not None:
        ata = np.ascontiguousarray(atol_adj, dtype=np.float64).reshape(nadj, n)
        rta = np.ascontiguousarray(rtol_adj, dtype=np.float64).reshape(nadj, n)
        L = lib()
        L.oracle_sdirk_adj_batch_tol.restype = ctypes.c_int
        L.oracle_sdirk_adj_batch_tol.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3 + \
            [ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 9 + [ctypes.c_int]
        L.oracle_sdirk_adj_batch_tol(MODEL[model], nc, npar if use_mu else 0, nadj, _d(y), _d(lam), _d(mu) if use_mu else None,
                                     tin, tout, _d(ata), _d(rta), _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst), _i(ierr), nthreads)
        return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr, lambda_=lam, mu=mu[:, :, :npar] if use_mu else None)
    lib().oracle_sdirk_adj_batch(MODEL[model], nc, npar if use_mu else 0, nadj, _d(y), _d(lam), _d(mu) if use_mu else None,
                                 tin, tout, _d(atol), _d(rtol), _i(ic), _d(rc), _i(ist), _d(rst), _i(ierr), nthreads)
    return dict(y=y, istatus=ist, rstatus=rst, ierr=ierr, lambda_=lam, mu=mu[:, :, :npar] if use_mu else None)
