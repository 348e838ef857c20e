"""Small fixed case for profiling: 444 cbm4 cells (one adjoint round on 148 SMs x 3 warps)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 444
hours = float(sys.argv[2]) if len(sys.argv) > 2 else 72.0
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + hours * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y = perturbed_states(F.initial_state("cbm4"), nc)
if os.environ.get("PROF_TAPE_GB"):
    F.lib().fatode_set_tape_budget(int(float(os.environ["PROF_TAPE_GB"]) * 2**30))
r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic)
s = F.last_stats()
if os.environ.get("PROF_COUNTS"):
    import json
    json.dump({"cells": nc, "hours": hours, "accepted_steps": int(s.accepted_steps), "attempts": int(r.istatus[:, 2].sum() - r.istatus[:, 3].sum())},
              open(os.environ["PROF_COUNTS"], "w"))
print(f"ncells {nc}: fwd {s.ms_forward:.1f} fwd_tape {s.ms_forward_tape:.1f} adj {s.ms_adjoint:.1f} ms steps {s.accepted_steps} ok {(r.ierr==1).all()}")
