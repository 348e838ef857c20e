import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
y0 = F.initial_state("cbm4")
def run(n, fam):
    ic = np.zeros(20, dtype=np.int32); ic[2] = 5; ic[3] = n; ic[6] = 1   # adjoint type: none (forward only)
    if fam == "adj":
        g = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=ic); r = O.ros_adj("cbm4", y0, 32, T0, T1, atol, rtol, ic, nthreads=1)
    else:
        g = F.integrate("cbm4", T0, T1, y0, rtol, atol, icntrl_u=ic); r = O.ros_fwd("cbm4", y0, T0, T1, atol, rtol, ic, nthreads=1)
    return g, r
for fam in ("adj", "fwd"):
    lo = None
    for n in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 100, 150, 200, 300, 400, 600, 800, 1000):
        g, r = run(n, fam)
        same = np.array_equal(g.y, r["y"])
        print(fam, n, "ierr", g.ierr[0], r["ierr"][0], "y bitwise", same, "maxrel %.2e" % np.max(np.abs(g.y - r["y"]) / np.maximum(np.abs(r["y"]), 1e-300)),
              "rstat", g.rstatus[0][:3], r["rstatus"][0][:3], "ist", g.istatus[0][:5], r["istatus"][0][:5], flush=True)
