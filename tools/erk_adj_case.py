"""Small fixed case for profiling the shallow-water adjoint / dense-Jacobian kernels: 296 members (one per CTA slot), 5 dt."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fatode_b200 as F
from fatode_b200.inputs import swe_members
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mode = sys.argv[2] if len(sys.argv) > 2 else "adj"
DT = float(np.float32(2.1573758800627289e-003)); tol = np.full(4800, 1e-4)
m = swe_members(n)
if mode == "adj":
    r = F.integrate_erk_adj("swe", 0.0, 5 * DT, m, 1, tol, tol)
else:
    ic = np.zeros(20, dtype=np.int32); ic[4] = 1
    v = np.zeros((n, 1, 4800)); v[:, 0, 0] = 1.0
    r = F.integrate_erk("swe", 0.0, 5 * DT, m, tol, tol, icntrl_u=ic, var_tlm=v)
s = F.last_stats()
print(f"{mode} members {n}: kernel {s.ms_forward:.1f} ms, Nfun {r.istatus[:, 0].mean():.1f} Njac {r.istatus[:, 1].mean():.1f} ok {(r.ierr == 1).all()}")
