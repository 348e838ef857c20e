"""Throughput of FWD/ERK INTEGRATE (+ ERK_TLM) on the shallow-water ensemble (Dopri5, rtol = atol = 1e-4, t = 0 .. 50 dt)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import swe_members
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
tlm = len(sys.argv) > 2 and sys.argv[2] == "tlm"
DT = float(np.float32(2.1573758800627289e-003))
tol = np.full(4800, 1e-4)
y = swe_members(nm)
v = None
if tlm:
    v = np.zeros((nm, 1, 4800)); v[:, 0, 0] = 1.0
for rep in range(2):
    t = time.time(); r = F.integrate_erk("swe", 0.0, (5 if tlm else 50) * DT, y, tol, tol, var_tlm=v); dt = time.time() - t
    s = F.last_stats()
    att = int(r.istatus[:, 2].sum()); nfun = int(r.istatus[:, 0].sum())
    print(f"ERK{'_TLM' if tlm else ''} members {nm} rep {rep}: wall {dt:.3f}s kernel {s.ms_forward:.1f} ms -> {nm/(s.ms_forward*1e-3):.0f} members/s, "
          f"{att/nm:.1f} attempts/member, {nfun/(s.ms_forward*1e-3)/1e6:.2f} M FUN/s, ok {(r.ierr==1).all()}", flush=True)
