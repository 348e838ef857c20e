import os, sys, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states
L = F.lib(); OL = O.lib()
dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int)
d = lambda a: a.ctypes.data_as(dp); i = lambda a: a.ctypes.data_as(ip)
y0 = perturbed_states(F.initial_state("cbm4"), 4)
for c, t in ((0, 43200.0), (1, 50000.123), (2, 70000.0), (3, 100000.5)):
    y = y0[c].copy()
    f, J, H = np.zeros(32), np.zeros(1024), np.zeros(258)
    rc = L.fatode_probe_cbm4(ctypes.c_double(t), d(y), d(f), d(J), d(H)); assert rc == 0, L.fatode_last_error()
    fo, Jo, Ho = np.zeros(32), np.zeros(1024), np.zeros(258)
    OL.oracle_cbm4_fun_jac(ctypes.c_double(t), d(y), d(fo), d(Jo), d(Ho))
    print(f"t={t}: fun bitwise {np.array_equal(f, fo)} maxrel {np.max(np.abs(f-fo)/np.maximum(np.abs(fo),1e-300)):.3e} | "
          f"jac bitwise {np.array_equal(J, Jo)} maxrel {np.max(np.abs(J-Jo)/np.maximum(np.abs(Jo),1e-300)):.3e} nnz {np.count_nonzero(J)} {np.count_nonzero(Jo)} | "
          f"hess bitwise {np.array_equal(H, Ho)} maxrel {np.max(np.abs(H-Ho)/np.maximum(np.abs(Ho),1e-300)):.3e}")
    if not np.array_equal(J, Jo):
        bad = np.nonzero(J != Jo)[0][:10]; print("   jac diffs at (row,col):", [(b % 32, b // 32, J[b], Jo[b]) for b in bad])
    if not np.array_equal(f, fo):
        bad = np.nonzero(f != fo)[0][:10]; print("   fun diffs:", [(b, f[b], fo[b]) for b in bad])
    # LU of E = ghinv I - J
    E = -Jo.copy(); E[np.arange(32) * 33] += 1.0 / (0.25 * 100.0)
    b = np.linspace(1.0, 2.0, 32) * 1e3
    lu, who, x, info = np.zeros(1024), np.zeros(32, dtype=np.int32), np.zeros(32), np.zeros(1, dtype=np.int32)
    rc = L.fatode_probe_lu32(d(E), d(b), d(lu), i(who), d(x), i(info)); assert rc == 0
    a = E.copy(); ipiv = np.zeros(32, dtype=np.int32); xn, xt = np.zeros(32), np.zeros(32)
    inf = OL.oracle_dgetf2_solve(32, d(a), i(ipiv), d(b), d(xn), d(xt))
    lu_ref = a.reshape(32, 32).T   # row-major view of column-major LU
    print(f"   LU bitwise {np.array_equal(lu.reshape(32,32), lu_ref)} info {info[0]} {inf} solve bitwise {np.array_equal(x, xn)} maxrel {np.max(np.abs(x-xn)/np.abs(xn)):.3e} swaps {np.sum(ipiv != np.arange(1,33))} who_ident {np.array_equal(who, np.arange(32))}")
# random dense matrix with real pivoting
rng = np.random.default_rng(1)
A = rng.standard_normal((32, 32)); A[3, :] = A[7, :] * 0.5 + 1e-3 * rng.standard_normal(32)
Af = np.asfortranarray(A).ravel(order="F").copy()
b = rng.standard_normal(32)
lu, who, x, info = np.zeros(1024), np.zeros(32, dtype=np.int32), np.zeros(32), np.zeros(1, dtype=np.int32)
L.fatode_probe_lu32(d(Af), d(b), d(lu), i(who), d(x), i(info))
a = Af.copy(); ipiv = np.zeros(32, dtype=np.int32); xn, xt = np.zeros(32), np.zeros(32)
OL.oracle_dgetf2_solve(32, d(a), i(ipiv), d(b), d(xn), d(xt))
print("random: LU bitwise", np.array_equal(lu.reshape(32, 32), a.reshape(32, 32).T), "solve bitwise", np.array_equal(x, xn), "residual", np.max(np.abs(A @ x - b)))
