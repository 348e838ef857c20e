"""Wall-clock throughput of integrate_adj through the C ABI (host buffers), a few repetitions."""
import time, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y = perturbed_states(F.initial_state("cbm4"), nc)
for rep in range(reps):
    t = time.time(); r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic); dt = time.time() - t
    s = F.last_stats()
    print(f"ncells {nc} rep {rep}: wall {dt:.3f}s -> {nc/dt:.0f} cells/s | fwd {s.ms_forward:.0f} expand {s.ms_forward_tape:.0f} adj {s.ms_adjoint:.0f} ms waves {s.waves} ok {(r.ierr==1).all()}", flush=True)
