import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y0 = perturbed_states(F.initial_state("cbm4"), 8)
g = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=ic); r = O.ros_adj("cbm4", y0, 32, T0, T1, atol, rtol, ic)
print("forward bitwise:", np.array_equal(g.y, r["y"]), np.array_equal(g.istatus, r["istatus"]), np.array_equal(g.rstatus, r["rstatus"]))
np.set_printoptions(precision=2, linewidth=220)
for c in (0, 3):
    mu, rm = g.mu[c], r["mu"][c]      # [col m][p]
    lam, rl = g.lambda_[c], r["lambda_"][c]
    cs = np.max(np.abs(rm), axis=1); ce = np.max(np.abs(mu - rm), axis=1)
    print("cell", c, "Mu col err/colscale log10:", np.round(np.log10(np.maximum(ce, 1e-300) / np.maximum(cs, 1e-300)), 1))
    print("   worst elem per col:", np.argmax(np.abs(mu - rm), axis=1))
    rs = np.max(np.abs(rm), axis=0); re = np.max(np.abs(mu - rm), axis=0)
    print("   Mu row(param) err/rowscale log10:", np.round(np.log10(np.maximum(re, 1e-300) / np.maximum(rs, 1e-300)), 1))
    ls = np.max(np.abs(rl), axis=1); le = np.max(np.abs(lam - rl), axis=1)
    print("   Lambda col err/colscale log10:", np.round(np.log10(np.maximum(le, 1e-300) / np.maximum(ls, 1e-300)), 1))
    print("   Lambda colscale log10:", np.round(np.log10(np.maximum(ls, 1e-300)), 0))
    # elementwise: error relative to |lambda| row scale x col scale?
    m3 = np.delete(mu, 3, axis=1); r3 = np.delete(rm, 3, axis=1)
    cs3 = np.max(np.abs(r3), axis=1); ce3 = np.max(np.abs(m3 - r3), axis=1)
    print("   Mu col err/colscale w/o param 4 log10:", np.round(np.log10(np.maximum(ce3, 1e-300) / np.maximum(cs3, 1e-300)), 1))
