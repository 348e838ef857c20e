"""Forward-only profiling case (FWD/ROS family INTEGRATE, per-thread kernel MODE 0)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
hours = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + hours * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y = perturbed_states(F.initial_state("cbm4"), nc)
r = F.integrate("cbm4", T0, T1, y, rtol, atol, icntrl_u=ic)
s = F.last_stats()
print(f"ncells {nc}: fwd {s.ms_forward:.2f} ms attempts/cell {r.istatus[:,2].mean():.1f} ok {(r.ierr==1).all()} general {s.cells_general}")
