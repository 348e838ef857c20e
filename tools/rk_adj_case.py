"""Throughput of ADJ/RK_ADJ INTEGRATE_ADJ (cbm4, Radau-2A, driver tolerances, 72 h, NADJ = 32, NP = 29)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-5), np.full(32, F01 * 1e-7)
ic = np.zeros(20, dtype=np.int32); ic[2] = 1
y = perturbed_states(F.initial_state("cbm4"), nc)
for rep in range(2):
    r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic, family="rk")
    s = F.last_stats()
    tot = s.ms_forward + s.ms_forward_tape + s.ms_adjoint
    print(f"RK_ADJ ncells {nc}: fwd {s.ms_forward:.1f} prep {s.ms_forward_tape:.1f} adj {s.ms_adjoint:.1f} ms -> {nc / tot * 1e3:.0f} cells/s; "
          f"acc/cell {r.istatus[:,3].mean():.0f} sol/cell {r.istatus[:,6].mean():.0f} ierr {np.unique(r.ierr)}", flush=True)
