"""Small fixed case for profiling the lock-step SDIRK_TLM kernel and the Newton-mode SDIRK_ADJ sweep: 296 systems (one per warp slot)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mode = sys.argv[2] if len(sys.argv) > 2 else "direct"
hours = float(sys.argv[3]) if len(sys.argv) > 3 else 12.0
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + hours * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
y = perturbed_states(F.initial_state("cbm4"), n)
ic = np.zeros(20, dtype=np.int32)
if mode == "newton_adj":
    ic[6] = 2
    r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic, family="sdirk", atol_adj=np.tile(atol, (32, 1)), rtol_adj=np.tile(rtol, (32, 1)))
else:
    ic[{"direct": 6, "est": 8, "trunc": 11}[mode]] = 1
    tt = np.full((32, 32), 1.0e-5)
    r = F.integrate_tlm("cbm4", T0, T1, y, np.tile(np.eye(32), (n, 1, 1)), rtol, atol, icntrl_u=ic, atol_tlm=tt, rtol_tlm=tt, family="sdirk")
s = F.last_stats()
print(f"{mode} systems {n}: fwd {s.ms_forward:.1f} prep {s.ms_forward_tape:.1f} sweep {s.ms_adjoint:.1f} ms, steps {r.istatus[:, 3].mean():.1f} ok {(r.ierr == 1).all()}")
