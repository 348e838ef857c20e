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
