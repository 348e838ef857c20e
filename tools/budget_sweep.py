"""Phase times of the headline call (CBM-IV ROS_ADJ, 72 h) against the tape budget: does a smaller working set (fewer 2 MB pages
under the scattered stage-block writes of k_expand_tape) change the per-step cost?  usage: budget_sweep.py [cells] [GB ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fatode_b200 as F
from fatode_b200.inputs import perturbed_states
# arguments: waves-per-run, then budget[:wave_cells] pairs (0 = the library's default)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfgs = [tuple(float(v) for v in (x.split(":") + ["0"])[:2]) for x in sys.argv[2:]] or [(0.0, 0.0), (64.0, 0.0), (32.0, 0.0)]
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
for gb, wave in cfgs:
    nc = nw * (int(wave) if wave > 0 else 8584)
    y = perturbed_states(F.initial_state("cbm4"), nc)
    F.lib().fatode_set_tape_budget(int(gb * 2**30))
    F.lib().fatode_set_wave_cells(int(wave))
    for rep in range(2):
        r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic)
        s = F.last_stats()
    tot = s.ms_forward + s.ms_forward_tape + s.ms_adjoint
    print(f"budget {gb:6.1f} GB wave {int(wave)} cells {nc}: fwd {s.ms_forward:.0f} expand {s.ms_forward_tape:.0f} adj {s.ms_adjoint:.0f} sum {tot:.0f} ms "
          f"waves {s.waves} launches {s.launches} -> {nc / tot * 1e3:.0f} cells/s ok {(r.ierr == 1).all()}", flush=True)
