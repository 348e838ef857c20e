"""Throughput of INTEGRATE_ADJ of ADJ/SDIRK_ADJ (cbm4, NADJ = 32, NP = 29, Sdirk4a, direct adjoint solve, 72 h) through the C ABI."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * k), np.full(32, F01 * k * 1e-2)
y = perturbed_states(F.initial_state("cbm4"), nc)
for rep in range(2):
    t = time.time(); r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, family="sdirk"); dt = time.time() - t
    s = F.last_stats()
    dev = (s.ms_forward + s.ms_forward_tape + s.ms_adjoint) * 1e-3
    print(f"SDIRK_ADJ ncells {nc} rep {rep}: wall {dt:.3f}s | forward {s.ms_forward:.0f} prep {s.ms_forward_tape:.0f} backward {s.ms_adjoint:.0f} ms "
          f"-> {nc/dev:.0f} cells/s (device time), {s.accepted_steps/nc:.0f} accepted steps per cell, waves {s.waves}, ok {(r.ierr==1).all()}", flush=True)
