"""Throughput of FWD/SDIRK INTEGRATE (cbm4, Sdirk4a, driver tolerances, 72 h) on the per-thread kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-5), np.full(32, F01 * 1e-7)
y = perturbed_states(F.initial_state("cbm4"), nc)
for rep in range(2):
    r = F.integrate_sdirk("cbm4", T0, T1, y, rtol, atol)
    s = F.last_stats()
    print(f"SDIRK ncells {nc}: kernel {s.ms_forward:.1f} ms -> {nc / s.ms_forward * 1e3:.0f} cells/s; steps/cell {r.istatus[:,2].mean():.0f} "
          f"fun/cell {r.istatus[:,0].mean():.0f} ok {(r.ierr==1).all()}", flush=True)
