"""GPU parity for INTEGRATE_ADJ of ADJ/ERK_ADJ on the shallow-water ensemble: the forward sweep (state,
ISTATUS, RSTATUS) bit-identical to the CPU oracle, Lambda within 1e-9 per column (the backward sweep has no adaptive decisions; the
sums of a JAC^T product are ordered by row except at the periodic seams and pow(x, 1.5) comes from CUDA's libm)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
N = 4800
DT = float(np.float32(2.1573758800627289e-003))
RTOL_PARITY = 1e-9


def check(res, ref, tol=RTOL_PARITY):
    assert np.array_equal(res.ierr, ref["ierr"])
    assert np.array_equal(res.istatus, ref["istatus"]), "ISTATUS must be bit-exact"
    assert np.array_equal(res.rstatus, ref["rstatus"])
    assert np.array_equal(res.y, ref["y"])
    a, b = res.lambda_, ref["lambda_"]
    sc = np.max(np.abs(b), axis=2)
    er = np.max(np.abs(a - b), axis=2)
    assert np.all(er <= tol * np.maximum(sc, 1e-300)), np.max(er / np.maximum(sc, 1e-300))
    return float(np.max(er / np.maximum(sc, 1e-300)))


def test_erk_adj_matches_oracle_all_methods():
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(6)
    t1 = 5 * DT
    for method, tolv, nadj in ((0, 1.0e-4, 1), (1, 1.0e-3, 2), (2, 1.0e-3, 1), (3, 1.0e-4, 3), (5, 1.0e-5, 1), (6, 1.0e-6, 2)):
        tol = np.full(N, tolv)
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        ref = O.erk_adj("swe", y0, nadj, 0.0, t1, tol, tol, ic)
        res = F.integrate_erk_adj("swe", 0.0, t1, y0, nadj, tol, tol, icntrl_u=ic)
        assert (ref["ierr"] == 1).all()
        check(res, ref)


def test_erk_adj_scalar_tolerances_long_window_tape_slot_and_failures():
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(5)
    tol = np.full(N, 1.0e-4)
    ic = np.zeros(20, dtype=np.int32)
    ic[1] = 1                                             # scalar tolerances
    t1 = 30 * DT
    ref = O.erk_adj("swe", y0, 1, 0.0, t1, tol, tol, ic)
    res = F.integrate_erk_adj("swe", 0.0, t1, y0, 1, tol, tol, icntrl_u=ic)
    check(res, ref)
    # a small tape slot: every member overflows it and is integrated again with Max_no_steps checkpoints
    L = F.lib()
    L.fatode_set_erk_tape_slot(3)
    try:
        ic2 = ic.copy()
        ic2[3] = 200
        ref2 = O.erk_adj("swe", y0, 1, 0.0, t1, tol, tol, ic2)
        res2 = F.integrate_erk_adj("swe", 0.0, t1, y0, 1, tol, tol, icntrl_u=ic2)
    finally:
        L.fatode_set_erk_tape_slot(0)
    check(res2, ref2)
    assert F.last_stats().cells_deferred == 5 and (ref2["istatus"][:, 3] > 3).all()
    # more accepted steps than Max_no_steps checkpoints: the reference STOPs in ERK_Push; IERR -6, Lambda as given
    ic3 = np.zeros(20, dtype=np.int32)
    ic3[3] = 3
    ref3 = O.erk_adj("swe", y0, 1, 0.0, t1, tol, tol, ic3)
    res3 = F.integrate_erk_adj("swe", 0.0, t1, y0, 1, tol, tol, icntrl_u=ic3)
    assert (ref3["ierr"] == -6).all() and np.array_equal(res3.ierr, ref3["ierr"])
    assert np.array_equal(res3.y, ref3["y"]) and np.array_equal(res3.istatus[:, [0, 2, 3, 4]], ref3["istatus"][:, [0, 2, 3, 4]])
    assert not res3.lambda_.any()
    # option error: every member gets the parser's code
    rc = np.zeros(20)
    rc[6] = -1.0                                          # never read: <= 0 entries do not override
    bad = np.full(N, 1.0e-4)
    bad[7] = -1.0
    res4 = F.integrate_erk_adj("swe", 0.0, t1, y0, 1, tol, bad, rcntrl_u=rc)
    assert (res4.ierr == -5).all()


def test_erk_adj_ensemble_duality_with_the_tangent_linear_model():
    """256 members: lambda_k . v against the finite-difference tangent linear model of the same engine (size-independent property)."""
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    y0 = swe_members(256)
    tol = np.full(N, 1.0e-4)
    t1 = 5 * DT
    res = F.integrate_erk_adj("swe", 0.0, t1, y0, 2, tol, tol)
    assert (res.ierr == 1).all()
    rng = np.random.default_rng(5)
    v = rng.standard_normal(N)
    tl = F.integrate_erk("swe", 0.0, t1, y0, tol, tol, var_tlm=np.broadcast_to(v, (256, 1, N)).copy())
    assert np.array_equal(tl.y, res.y)
    for k in range(2):
        d = res.lambda_[:, k, :] @ v
        assert np.all(np.abs(d - tl.y_tlm[:, 0, k]) < 2e-3 * np.max(np.abs(tl.y_tlm[:, 0, :2])))
