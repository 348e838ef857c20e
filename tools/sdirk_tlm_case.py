"""Throughput of INTEGRATE_TLM of TLM/SDIRK_TLM (cbm4, NTLM = 32, Sdirk4a, 72 h) through the C ABI with host buffers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
y = perturbed_states(F.initial_state("cbm4"), nc)
yt = np.tile(np.eye(32), (nc, 1, 1))
for rep in range(2):
    t = time.time(); r = F.integrate_tlm("cbm4", T0, T1, y, yt, rtol, atol, family="sdirk"); dt = time.time() - t
    s = F.last_stats()
    dev = (s.ms_forward + s.ms_forward_tape + s.ms_adjoint) * 1e-3
    print(f"SDIRK_TLM ncells {nc} rep {rep}: wall {dt:.3f}s | state sweep {s.ms_forward:.0f} prep {s.ms_forward_tape:.0f} column sweep {s.ms_adjoint:.0f} ms "
          f"-> {nc/dev:.0f} cells/s (device time), {s.accepted_steps/nc:.0f} accepted steps per cell, waves {s.waves}, ok {(r.ierr==1).all()}", flush=True)
