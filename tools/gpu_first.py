"""First-contact script for the GPU box: FP64 probe, 1-cell parity vs oracle, small batch timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states

print("fp64 peak TFLOP/s:", F.measure_fp64_peak(20000), flush=True)
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y0 = F.initial_state("cbm4")
t = time.time(); res = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=ic); print("gpu 1 cell s", time.time() - t, flush=True)
ref = O.ros_adj("cbm4", y0, 32, T0, T1, atol, rtol, ic, nthreads=1)
print("ierr", res.ierr, ref["ierr"]); print("istatus gpu", res.istatus[0][:8]); print("istatus ref", ref["istatus"][0][:8])
print("rstatus", res.rstatus[0][:3], ref["rstatus"][0][:3])
nw = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
print("y normwise", nw(res.y, ref["y"]), "bitwise equal:", np.array_equal(res.y, ref["y"]))
sc = np.max(np.abs(ref["lambda_"][0]), axis=1); er = np.max(np.abs(res.lambda_[0] - ref["lambda_"][0]), axis=1)
print("lambda col err/scale max over cols>1e-30:", np.max((er / np.maximum(sc, 1e-300))[sc > 1e-30]))
sc = np.max(np.abs(ref["mu"][0]), axis=1); er = np.max(np.abs(res.mu[0] - ref["mu"][0]), axis=1)
print("mu col err/scale max:", np.max((er / np.maximum(sc, 1e-300))[sc > 1e-30]))
for nc in (444, 4096):
    y = perturbed_states(y0, nc)
    t = time.time(); r = F.integrate_adj("cbm4", T0, T1, y, 32, rtol, atol, icntrl_u=ic); dt = time.time() - t
    s = F.last_stats()
    print(f"ncells {nc}: wall {dt:.3f}s cells/s {nc/dt:.1f} | fwd {s.ms_forward:.1f} ms fwd_tape {s.ms_forward_tape:.1f} ms adj {s.ms_adjoint:.1f} ms "
          f"waves {s.waves} steps {s.accepted_steps} tape {s.tape_bytes_peak/1e9:.2f} GB ierr_ok {(r.ierr==1).all()}", flush=True)
    print("   nacc min/mean/max", r.istatus[:, 3].min(), r.istatus[:, 3].mean(), r.istatus[:, 3].max())
