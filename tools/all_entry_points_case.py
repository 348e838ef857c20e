"""Tiny run of every entry point (small tape budget: several waves, hand-back path) — a quick end-to-end sanity case."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 1.5 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
F.lib().fatode_set_tape_budget(256 << 20)
y = perturbed_states(F.initial_state("cbm4"), 40)
r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic, want_mu_sum=True); print("adj", (r.ierr == 1).all(), F.last_stats().waves)
F.lib().fatode_debug_irregular_every(7)
r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic); print("adj+handback", (r.ierr == 1).all(), F.last_stats().cells_general)
F.lib().fatode_debug_irregular_every(0)
r = F.integrate_tlm("cbm4", T0, T1, y[:12], np.tile(np.eye(32), (12, 1, 1)), rtol, atol, icntrl_u=ic); print("tlm", (r.ierr == 1).all())
r = F.integrate("cbm4", T0, T1, y, rtol, atol); print("fwd", (r.ierr == 1).all())
r = F.integrate_sdirk("cbm4", T0, T1, y, rtol, atol); print("sdirk", (r.ierr == 1).all())
ys = perturbed_states(F.initial_state("small_strato"), 70)
r = F.integrate("small_strato", T0, T0 + 86400.0, ys, np.full(5, 1e-4), np.full(5, 1e-2)); print("strato", (r.ierr == 1).all())
r = F.integrate_sdirk("roberts", 0.0, 40.0, np.array([[1.0, 0, 0]] * 5), np.full(3, 1e-4), np.full(3, 1e-8)); print("roberts sdirk", (r.ierr == 1).all())
icr = np.zeros(20, dtype=np.int32); icr[2] = 1
r = F.integrate_rk("cbm4", T0, T1, y, rtol, atol, icntrl_u=icr); print("rk", (r.ierr == 1).all(), F.last_stats().cells_general)
r = F.integrate_adj("cbm4", T0, T1, y[:12], 32, rtol, atol, icntrl_u=icr, family="rk"); print("rk_adj", (r.ierr == 1).all(), F.last_stats().cells_general)
r = F.integrate_adj("cbm4", T0, T1, y[:12], 32, rtol, atol, family="sdirk"); print("sdirk_adj", (r.ierr == 1).all(), F.last_stats().cells_general)
r = F.integrate_tlm("cbm4", T0, T1, y[:12], np.tile(np.eye(32), (12, 1, 1)), rtol, atol, family="sdirk"); print("sdirk_tlm", (r.ierr == 1).all())
r = F.integrate_tlm("cbm4", T0, T1, y[:12], np.tile(np.eye(32), (12, 1, 1)), rtol, atol, icntrl_u=icr, family="rk"); print("rk_tlm (direct)", (r.ierr == 1).all())
icd = icr.copy(); icd[6] = 2
r = F.integrate_adj("cbm4", T0, T1, y[:12], 32, rtol, atol, icntrl_u=icd, family="rk"); print("rk_adj direct", (r.ierr == 1).all())
loose_r, loose_a = np.full(32, F01 * 1e-2), np.full(32, F01 * 1e-4)      # general-pivoting passes
r = F.integrate_adj("cbm4", T0, T1, y[:12], 8, loose_r, loose_a, family="sdirk"); print("sdirk_adj general pass", (r.ierr == 1).all(), F.last_stats().cells_general)
r = F.integrate_adj("cbm4", T0, T1, y[:12], 8, loose_r, loose_a, icntrl_u=icr, family="rk"); print("rk_adj general pass", (r.ierr == 1).all(), F.last_stats().cells_general)
from fatode_b200.inputs import swe_members
DT = float(np.float32(2.1573758800627289e-003)); tol = np.full(4800, 1e-3)
m = swe_members(9)
r = F.integrate_erk("swe", 0.0, 10 * DT, m, tol, tol); print("swe erk", (r.ierr == 1).all(), r.istatus[:, 2].mean())
v = np.zeros((9, 1, 4800)); v[:, 0, 0] = 1.0
r = F.integrate_erk("swe", 0.0, 5 * DT, m, tol, tol, var_tlm=v); print("swe erk_tlm", (r.ierr == 1).all())
icd5 = np.zeros(20, dtype=np.int32); icd5[4] = 1
r = F.integrate_erk("swe", 0.0, 5 * DT, m, tol, tol, icntrl_u=icd5, var_tlm=v); print("swe erk_tlm dense Jacobian", (r.ierr == 1).all(), r.istatus[:, 1].mean())
r = F.integrate_erk_adj("swe", 0.0, 5 * DT, m, 2, tol, tol); print("swe erk_adj", (r.ierr == 1).all(), r.istatus[:, 1].mean())
tt = np.full((32, 32), 1.0e-4)
for name, k in (("direct", 6), ("column Newton control", 8), ("column truncation control", 11)):
    icl = np.zeros(20, dtype=np.int32); icl[k] = 1
    r = F.integrate_tlm("cbm4", T0, T1, y[:12], np.tile(np.eye(32), (12, 1, 1)), rtol, atol, icntrl_u=icl, atol_tlm=tt, rtol_tlm=tt, family="sdirk")
    print("sdirk_tlm lock-step:", name, (r.ierr == 1).all(), F.last_stats().cells_general)
icn = np.zeros(20, dtype=np.int32); icn[6] = 2
r = F.integrate_adj("cbm4", T0, T1, y[:12], 8, rtol, atol, icntrl_u=icn, family="sdirk", atol_adj=np.tile(atol, (8, 1)), rtol_adj=np.tile(rtol, (8, 1)))
print("sdirk_adj simplified-Newton adjoint stages", (r.ierr == 1).all(), r.istatus[:, 5].mean())
