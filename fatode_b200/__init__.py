"""fatode_b200 — host-side mirror of FATODE's INTEGRATE / INTEGRATE_ADJ interfaces over the
B200-native batched engine (C ABI in include/fatode_b200.h, CUDA in fatode_b200/csrc).

The functions keep the reference's argument names and meaning (TIN, TOUT, VAR/Y, RTOL, ATOL,
ICNTRL_U, RCNTRL_U -> ISTATUS_U, RSTATUS_U, IERR; Lambda(N,NADJ), Mu(NP,NADJ)) with one extra leading
axis for the cells of a batch.  The EXTERNAL callbacks of the Fortran API are replaced by a model id.
There is no CPU implementation behind this module: if the CUDA library cannot be loaded or no GPU
is present the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfatode_b200.so")

MODEL_ROBERTS, MODEL_SMALL_STRATO, MODEL_CBM4, MODEL_SWE = 1, 2, 3, 4
MODELS = {"roberts": MODEL_ROBERTS, "small_strato": MODEL_SMALL_STRATO, "cbm4": MODEL_CBM4, "swe": MODEL_SWE}
MEM_HOST, MEM_DEVICE = 0, 1

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


class FatodeError(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_int64), ("waves", ctypes.c_int64), ("ms_forward", ctypes.c_double),
                ("ms_forward_tape", ctypes.c_double), ("ms_adjoint", ctypes.c_double), ("ms_other", ctypes.c_double),
                ("tape_bytes_peak", ctypes.c_uint64), ("accepted_steps", ctypes.c_int64), ("attempts", ctypes.c_int64),
                ("cells_general", ctypes.c_int64), ("cells_deferred", ctypes.c_int64),
                ("launches_forward", ctypes.c_int64), ("launches_expand", ctypes.c_int64), ("launches_adjoint", ctypes.c_int64)]


_lib = None


def lib():
    """Loads the CUDA library (never a substitute): raises FatodeError if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FatodeError(f"{LIB_PATH} is not built; run `python -m fatode_b200.build` "
                              "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.fatode_last_error.restype = ctypes.c_char_p
        L.fatode_version.restype = ctypes.c_char_p
        L.fatode_measure_fp64_peak.restype = ctypes.c_double
        L.fatode_measure_fp64_peak.argtypes = [ctypes.c_int]
        L.fatode_alloc_pinned.restype = ctypes.c_void_p
        L.fatode_alloc_pinned.argtypes = [ctypes.c_uint64]
        L.fatode_free_pinned.argtypes = [ctypes.c_void_p]
        L.fatode_set_tape_budget.argtypes = [ctypes.c_uint64]
        L.fatode_set_wave_cells.argtypes = [ctypes.c_int64]
        L.fatode_set_path.argtypes = [ctypes.c_int]
        L.fatode_set_pipeline.argtypes = [ctypes.c_int]
        L.fatode_debug_irregular_every.argtypes = [ctypes.c_int]
        L.fatode_get_last_stats.argtypes = [ctypes.POINTER(Stats)]
        L.fatode_model_dims.argtypes = [ctypes.c_int, _ip, _ip]
        L.fatode_model_initial_state.argtypes = [ctypes.c_int, _dp]
        vp = ctypes.c_void_p
        L.fatode_ros_integrate_batch.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, vp, ctypes.c_double,
                                                 ctypes.c_double, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.fatode_ros_integrate_adj_batch.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int, vp, vp, vp, ctypes.c_double, ctypes.c_double,
                                                     vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.fatode_sdirk_integrate_batch.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, vp, ctypes.c_double,
                                                   ctypes.c_double, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.fatode_rk_integrate_batch.argtypes = L.fatode_sdirk_integrate_batch.argtypes
        L.fatode_rk_integrate_adj_batch.argtypes = L.fatode_ros_integrate_adj_batch.argtypes
        L.fatode_sdirk_integrate_adj_batch.argtypes = L.fatode_ros_integrate_adj_batch.argtypes
        L.fatode_ros_integrate_tlm_batch.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, vp, vp,
                                                     ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                     ctypes.c_int, vp]
        L.fatode_sdirk_integrate_tlm_batch.argtypes = L.fatode_ros_integrate_tlm_batch.argtypes
        L.fatode_rk_integrate_tlm_batch.argtypes = L.fatode_ros_integrate_tlm_batch.argtypes
        L.fatode_set_sdirk_tape_slot.argtypes = [ctypes.c_int]
        L.fatode_set_erk_tape_slot.argtypes = [ctypes.c_int]
        L.fatode_erk_integrate_adj_batch.restype = ctypes.c_int
        L.fatode_erk_integrate_adj_batch.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int] + \
            [ctypes.c_void_p] * 3 + [ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 8 + [ctypes.c_int, ctypes.c_void_p]
        L.fatode_erk_integrate_batch.argtypes = L.fatode_sdirk_integrate_batch.argtypes
        L.fatode_erk_integrate_tlm_batch.argtypes = L.fatode_ros_integrate_tlm_batch.argtypes
        L.fatode_swe_initial_state.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise FatodeError(f"fatode_b200 error {rc}: {lib().fatode_last_error().decode()}")


def model_id(model) -> int:
    return MODELS[model] if isinstance(model, str) else int(model)


def model_dims(model):
    n, p = ctypes.c_int(), ctypes.c_int()
    _check(lib().fatode_model_dims(model_id(model), ctypes.byref(n), ctypes.byref(p)))
    return n.value, p.value


def initial_state(model) -> np.ndarray:
    n, _ = model_dims(model)
    y = np.zeros(n)
    _check(lib().fatode_model_initial_state(model_id(model), y.ctypes.data_as(_dp)))
    return y


def last_stats() -> Stats:
    s = Stats()
    lib().fatode_get_last_stats(ctypes.byref(s))
    return s


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _ctrl(icntrl_u, rcntrl_u):
    ic = None if icntrl_u is None else np.ascontiguousarray(icntrl_u, dtype=np.int32)
    rc = None if rcntrl_u is None else np.ascontiguousarray(rcntrl_u, dtype=np.float64)
    if ic is not None and ic.shape != (20,):
        raise ValueError("ICNTRL_U must have 20 entries")
    if rc is not None and rc.shape != (20,):
        raise ValueError("RCNTRL_U must have 20 entries")
    return ic, rc


@dataclass
class Result:
    y: np.ndarray            # [ncells, N]  state at TOUT (or at exit time on failure)
    istatus: np.ndarray      # [ncells, 20] ISTATUS_U
    rstatus: np.ndarray      # [ncells, 20] RSTATUS_U
    ierr: np.ndarray         # [ncells]     IERR (1 = success)
    lambda_: np.ndarray | None = None   # [ncells, NADJ, N]  = Lambda(N,NADJ) per cell (Fortran order)
    mu: np.ndarray | None = None        # [ncells, NADJ, NP] = Mu(NP,NADJ) per cell
    mu_sum: np.ndarray | None = None    # [NADJ, NP] deterministic sum over cells
    y_tlm: np.ndarray | None = None     # [ncells, NTLM, N] = Y_tlm(N,NTLM) per cell


def integrate(model, tin, tout, var, rtol, atol, icntrl_u=None, rcntrl_u=None, model_params=None) -> Result:
    """Batched FWD/ROS INTEGRATE (FWD/ROS/ROS_f90_Integrator.F90:32). var: [ncells, N] or [N]."""
    mid = model_id(model)
    y = np.atleast_2d(np.array(var, dtype=np.float64, order="C"))
    nc, n = y.shape
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    mp = None if model_params is None else np.ascontiguousarray(model_params, dtype=np.float64)
    _check(lib().fatode_ros_integrate_batch(mid, nc, n, _ptr(y), float(tin), float(tout), _ptr(atol), _ptr(rtol),
                                            _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), _ptr(mp), MEM_HOST, None))
    return Result(y, ist, rst, ierr)


def integrate_sdirk(model, tin, tout, var, rtol, atol, icntrl_u=None, rcntrl_u=None, model_params=None) -> Result:
    """Batched FWD/SDIRK INTEGRATE (FWD/SDIRK/SDIRK_f90_Integrator.F90:25). var: [ncells, N] or [N]."""
    mid = model_id(model)
    y = np.atleast_2d(np.array(var, dtype=np.float64, order="C"))
    nc, n = y.shape
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    mp = None if model_params is None else np.ascontiguousarray(model_params, dtype=np.float64)
    _check(lib().fatode_sdirk_integrate_batch(mid, nc, n, _ptr(y), float(tin), float(tout), _ptr(atol), _ptr(rtol),
                                              _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), _ptr(mp), MEM_HOST, None))
    return Result(y, ist, rst, ierr)


def integrate_rk(model, tin, tout, var, rtol, atol, icntrl_u=None, rcntrl_u=None, model_params=None) -> Result:
    """Batched FWD/RK INTEGRATE (FWD/RK/RK_f90_Integrator.F90:32). var: [ncells, N] or [N]."""
    mid = model_id(model)
    y = np.atleast_2d(np.array(var, dtype=np.float64, order="C"))
    nc, n = y.shape
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    mp = None if model_params is None else np.ascontiguousarray(model_params, dtype=np.float64)
    _check(lib().fatode_rk_integrate_batch(mid, nc, n, _ptr(y), float(tin), float(tout), _ptr(atol), _ptr(rtol),
                                           _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), _ptr(mp), MEM_HOST, None))
    return Result(y, ist, rst, ierr)


def integrate_erk(model, tin, tout, var, rtol, atol, icntrl_u=None, rcntrl_u=None, var_tlm=None, atol_tlm=None, rtol_tlm=None) -> Result:
    """Batched FWD/ERK INTEGRATE (FWD/ERK/ERK_f90_Integrator.F90:28) or, with var_tlm [ncells, NTLM, N], TLM/ERK_TLM
    INTEGRATE_TLM (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:28). var: [ncells, N] or [N]; model "swe"."""
    mid = model_id(model)
    y = np.atleast_2d(np.array(var, dtype=np.float64, order="C"))
    nc, n = y.shape
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    if var_tlm is None:
        _check(lib().fatode_erk_integrate_batch(mid, nc, n, _ptr(y), float(tin), float(tout), _ptr(atol), _ptr(rtol),
                                                _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), None, MEM_HOST, None))
        return Result(y, ist, rst, ierr)
    yt = np.array(var_tlm, dtype=np.float64, order="C").reshape(nc, -1, n)
    at = None if atol_tlm is None else np.ascontiguousarray(atol_tlm, dtype=np.float64)
    rt = None if rtol_tlm is None else np.ascontiguousarray(rtol_tlm, dtype=np.float64)
    _check(lib().fatode_erk_integrate_tlm_batch(mid, nc, n, yt.shape[1], _ptr(y), _ptr(yt), float(tin), float(tout), _ptr(at),
                                                _ptr(rt), _ptr(atol), _ptr(rtol), _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst),
                                                _ptr(ierr), None, MEM_HOST, None))
    return Result(y, ist, rst, ierr, y_tlm=yt)


def swe_initial_state(amplitude=30.0, u0=2.0, v0=2.0) -> np.ndarray:
    """Gaussian bump of the shallow-water drivers (EXAMPLES/swe/SWE_ERK/swe2D_erk_dr.F90:100-103), grid2vec order."""
    y = np.zeros(4800)
    _check(lib().fatode_swe_initial_state(float(amplitude), float(u0), float(v0), y.ctypes.data_as(_dp)))
    return y


def integrate_adj(model, tin, tout, y, nadj, rtol, atol, np_=None, icntrl_u=None, rcntrl_u=None, want_mu=True,
                  want_mu_sum=False, model_params=None, family="ros", atol_adj=None, rtol_adj=None) -> Result:
    """Batched INTEGRATE_ADJ: family="ros" ADJ/ROS_ADJ (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:34), family="rk"
    ADJ/RK_ADJ (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:20), family="sdirk" ADJ/SDIRK_ADJ (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:29).
    y: [ncells, N] or [N]."""
    mid = model_id(model)
    yy = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = yy.shape
    _, npm = model_dims(mid)
    npar = npm if np_ is None else int(np_)
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    with_mu = want_mu and npar > 0
    mu = np.zeros((nc, nadj, npar)) if with_mu else None
    mus = np.zeros((nadj, npar)) if (with_mu and want_mu_sum) else None
    mp = None if model_params is None else np.ascontiguousarray(model_params, dtype=np.float64)
    # ATOL_adj / RTOL_adj (N, NADJ): read by ADJ/SDIRK_ADJ's simplified-Newton adjoint stages (ICNTRL(7) /= 1) only
    ata = None if atol_adj is None else np.ascontiguousarray(atol_adj, dtype=np.float64).reshape(nadj, n)
    rta = None if rtol_adj is None else np.ascontiguousarray(rtol_adj, dtype=np.float64).reshape(nadj, n)
    entry = {"rk": lib().fatode_rk_integrate_adj_batch, "sdirk": lib().fatode_sdirk_integrate_adj_batch}.get(
        family, lib().fatode_ros_integrate_adj_batch)
    _check(entry(mid, nc, n, npar if with_mu else 0, nadj, _ptr(yy), _ptr(lam), _ptr(mu),
                                                float(tin), float(tout), _ptr(ata), _ptr(rta), _ptr(atol), _ptr(rtol), _ptr(ic),
                                                _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), _ptr(mus), _ptr(mp),
                                                MEM_HOST, None))
    return Result(yy, ist, rst, ierr, lam, mu, mus)


def integrate_erk_adj(model, tin, tout, y, nadj, rtol, atol, icntrl_u=None, rcntrl_u=None) -> Result:
    """Batched INTEGRATE_ADJ of ADJ/ERK_ADJ (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:33) on the shallow-water ensemble, no parameters;
    y: [members, 4800]; result.lambda_: [members, nadj, 4800] (ADJINIT: Lambda(k,k) = 1)."""
    mid = model_id(model)
    yy = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = yy.shape
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    lam = np.zeros((nc, nadj, n))
    _check(lib().fatode_erk_integrate_adj_batch(mid, nc, n, 0, nadj, _ptr(yy), _ptr(lam), None, float(tin), float(tout), _ptr(atol),
                                                _ptr(rtol), _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), None, MEM_HOST, None))
    return Result(yy, ist, rst, ierr, lam, None, None)


def integrate_tlm(model, tin, tout, y, y_tlm, rtol, atol, icntrl_u=None, rcntrl_u=None, atol_tlm=None, rtol_tlm=None,
                  model_params=None, family="ros") -> Result:
    """Batched INTEGRATE_TLM: family="ros" TLM/ROS_TLM (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:35), family="sdirk"
    TLM/SDIRK_TLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:31), family="rk" TLM/RK_TLM (TLM/RK_TLM/RK_TLM_f90_Integrator.F90:33,
    preset direct TLM solve). y: [ncells, N]; y_tlm: [ncells, NTLM, N]."""
    mid = model_id(model)
    yy = np.atleast_2d(np.array(y, dtype=np.float64, order="C"))
    nc, n = yy.shape
    yt = np.array(y_tlm, dtype=np.float64, order="C")
    yt = yt if yt.ndim == 3 else yt.reshape(nc, -1, n)      # (an empty batch cannot infer NTLM from a flat array)
    ntlm = yt.shape[1]
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    ic, rc = _ctrl(icntrl_u, rcntrl_u)
    at = None if atol_tlm is None else np.ascontiguousarray(atol_tlm, dtype=np.float64)
    rt = None if rtol_tlm is None else np.ascontiguousarray(rtol_tlm, dtype=np.float64)
    ist = np.zeros((nc, 20), dtype=np.int32)
    rst = np.zeros((nc, 20))
    ierr = np.zeros(nc, dtype=np.int32)
    mp = None if model_params is None else np.ascontiguousarray(model_params, dtype=np.float64)
    entry = {"sdirk": lib().fatode_sdirk_integrate_tlm_batch, "rk": lib().fatode_rk_integrate_tlm_batch}.get(
        family, lib().fatode_ros_integrate_tlm_batch)
    _check(entry(mid, nc, n, ntlm, _ptr(yy), _ptr(yt), float(tin), float(tout), _ptr(at), _ptr(rt),
                 _ptr(atol), _ptr(rtol), _ptr(ic), _ptr(rc), _ptr(ist), _ptr(rst), _ptr(ierr), _ptr(mp), MEM_HOST, None))
    return Result(yy, ist, rst, ierr, y_tlm=yt)


def measure_fp64_peak(iters=20000) -> float:
    v = lib().fatode_measure_fp64_peak(int(iters))
    if v < 0:
        raise FatodeError(lib().fatode_last_error().decode())
    return v
