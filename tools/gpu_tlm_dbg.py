import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states
F01 = float(np.float32(0.1)); T0 = 43200.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
y0 = perturbed_states(F.initial_state("cbm4"), 6)
rng = np.random.default_rng(7)
yt = rng.standard_normal((6, 5, 32))
t1 = T0 + 8 * 3600.0
for method, auto, scalar in ((0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 1), (4, 1, 0), (5, 1, 1)):
    ic = np.zeros(20, dtype=np.int32); ic[0], ic[1], ic[2] = auto, scalar, method
    res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic)
    ref = O.ros_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic)
    b = ref["y_tlm"]; sc = np.max(np.abs(b), axis=2); er = np.max(np.abs(res.y_tlm - b), axis=2)
    print(method, auto, scalar, "ierr", res.ierr, ref["ierr"], "ist eq", np.array_equal(res.istatus, ref["istatus"]), "y eq", np.array_equal(res.y, ref["y"]),
          "tlm err", np.max(er / np.maximum(sc, 1e-10 * sc.max(axis=1, keepdims=True))))
    if not np.array_equal(res.istatus, ref["istatus"]):
        print("  gpu", res.istatus[0, :8], "\n  ref", ref["istatus"][0, :8])
