"""Python mirror of the batch container / text I/O of the C ABI (include/fatode_b200.h, csrc/fatode_io.cu).

save_batch / load_batch: the "FATODEB1" structure-of-arrays container for the arrays of a whole batch (y, Lambda, Mu, ISTATUS,
RSTATUS, IERR) in the layouts the *_batch entry points use.  write_solution_text / write_sens_text / read_*: the reference's own
one-number-per-line E24.16 files of a single system (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_adj_dr.F90:134-164)."""
import ctypes

import numpy as np

from . import FatodeError, lib, model_id

_NAMES = ("y", "lambda", "mu", "istatus", "rstatus", "ierr")


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _sig():
    L = lib()
    vp, cp, i64, ci, cd = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    ip, dp, lp = ctypes.POINTER(ci), ctypes.POINTER(cd), ctypes.POINTER(i64)
    L.fatode_file_write.argtypes = [cp, ci, i64, ci, ci, ci, cd, cd, vp, vp, vp, vp, vp, vp]
    L.fatode_file_info.argtypes = [cp, ip, lp, ip, ip, ip, dp, dp, ip]
    L.fatode_file_read.argtypes = [cp, cp, i64, i64, vp]
    L.fatode_text_write_solution.argtypes = [cp, ci, vp]
    L.fatode_text_write_sens.argtypes = [cp, ci, ci, ci, vp, vp]
    L.fatode_text_read_solution.argtypes = [cp, ci, vp]
    L.fatode_text_read_sens.argtypes = [cp, ci, ci, ci, vp, vp]
    return L


def save_batch(path, model, tin, tout, y, lambda_=None, mu=None, istatus=None, rstatus=None, ierr=None):
    y = np.ascontiguousarray(np.atleast_2d(y), dtype=np.float64)
    nc, n = y.shape
    lam = None if lambda_ is None else np.ascontiguousarray(lambda_, dtype=np.float64).reshape(nc, -1, n)
    nadj = 0 if lam is None else lam.shape[1]
    m = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64).reshape(nc, nadj, -1)
    npar = 0 if m is None else m.shape[2]
    ist = None if istatus is None else np.ascontiguousarray(istatus, dtype=np.int32).reshape(nc, 20)
    rst = None if rstatus is None else np.ascontiguousarray(rstatus, dtype=np.float64).reshape(nc, 20)
    ie = None if ierr is None else np.ascontiguousarray(ierr, dtype=np.int32).reshape(nc)
    mid = model_id(model) if isinstance(model, str) else int(model)
    rc = _sig().fatode_file_write(str(path).encode(), mid, nc, n, npar, nadj, float(tin), float(tout), _p(y), _p(lam), _p(m),
                                  _p(ist), _p(rst), _p(ie))
    if rc != 0:
        raise FatodeError(f"fatode_file_write({path}) failed: {rc}")


def batch_info(path):
    L = _sig()
    model, nvar, npar, nadj = (ctypes.c_int() for _ in range(4))
    nc = ctypes.c_int64()
    tin, tout = ctypes.c_double(), ctypes.c_double()
    has = (ctypes.c_int * 6)()
    rc = L.fatode_file_info(str(path).encode(), model, nc, nvar, npar, nadj, tin, tout, has)
    if rc != 0:
        raise FatodeError(f"{path}: not a FATODEB1 container ({rc})")
    return dict(model=model.value, ncells=nc.value, nvar=nvar.value, np=npar.value, nadj=nadj.value, tin=tin.value, tout=tout.value,
                arrays=[n for n, h in zip(_NAMES, has) if h])


def load_batch(path, first_cell=0, ncells=None):
    """dict of the arrays present, cells [first_cell, first_cell + ncells) — a rank reads only its own shard"""
    L = _sig()
    info = batch_info(path)
    nc = info["ncells"] - first_cell if ncells is None else ncells
    n, npar, nadj = info["nvar"], info["np"], info["nadj"]
    shapes = {"y": ((nc, n), np.float64), "lambda": ((nc, nadj, n), np.float64), "mu": ((nc, nadj, npar), np.float64),
              "istatus": ((nc, 20), np.int32), "rstatus": ((nc, 20), np.float64), "ierr": ((nc,), np.int32)}
    out = dict(info)
    for name in info["arrays"]:
        a = np.empty(*shapes[name])
        if L.fatode_file_read(str(path).encode(), name.encode(), first_cell, nc, _p(a)) != 0:
            raise FatodeError(f"{path}: cannot read array {name}")
        out[name] = a
    return out


def write_solution_text(path, y):
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    if _sig().fatode_text_write_solution(str(path).encode(), len(y), _p(y)) != 0:
        raise FatodeError(f"cannot write {path}")


def write_sens_text(path, lambda_, mu=None):
    lam = np.ascontiguousarray(lambda_, dtype=np.float64)
    nadj, n = lam.shape
    m = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64).reshape(nadj, -1)
    if _sig().fatode_text_write_sens(str(path).encode(), n, 0 if m is None else m.shape[1], nadj, _p(lam), _p(m)) != 0:
        raise FatodeError(f"cannot write {path}")


def read_solution_text(path, nvar):
    y = np.empty(nvar)
    if _sig().fatode_text_read_solution(str(path).encode(), nvar, _p(y)) != 0:
        raise FatodeError(f"cannot read {nvar} numbers from {path}")
    return y


def read_sens_text(path, nvar, npar, nadj):
    lam = np.empty((nadj, nvar))
    mu = np.empty((nadj, npar)) if npar > 0 else None
    if _sig().fatode_text_read_sens(str(path).encode(), nvar, npar, nadj, _p(lam), _p(mu)) != 0:
        raise FatodeError(f"cannot read {path}")
    return lam, mu
