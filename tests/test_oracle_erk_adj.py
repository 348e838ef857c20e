"""CPU checks of the oracle's ERK_ADJ restatement (oracle/erk_adj.hpp: ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:33-140, 608-700,
1276-1740) and of the JAC callback of the shallow-water example generated from compute_JacF (tools/gen_swe_jac.py,
EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90:483-2023).  The reference ships no artefact for this family (the driver prints only):
"parity unpinned"; the restatement is cross-validated against finite differences of the golden-pinned state sweep, by the
adjoint / tangent-linear duality and through the counter relations of the Fortran."""
import numpy as np

import oracle_lib as O

N = 4800
DT = float(np.float32(2.1573758800627289e-003))


def test_generated_jacobian_against_finite_differences():
    rng = np.random.default_rng(0)
    y = O.swe_initial_state() * (1.0 + 0.01 * rng.standard_normal(N))
    v, w = rng.standard_normal(N), rng.standard_normal(N)
    jv, jtw = O.swe_jac_mul(y, v), O.swe_jac_mul(y, w, trans=True)
    assert abs(w @ jv - v @ jtw) <= 1e-13 * abs(w @ jv)                 # JAC x and JAC^T x are the same matrix
    eps = 1e-6
    fd = (O.swe_fun(y + eps * v) - O.swe_fun(y - eps * v)) / (2 * eps)
    # the reference's formula export is not the exact derivative: six-digit literals (0.0833333 for 1/12: 4e-7) and a few entries
    # of the u- and v-equations whose first-eigenvalue terms lack a factor (:1183-1216), 3e-5 of the product's size here
    assert np.max(np.abs(fd - jv)) < 1e-4 * np.max(np.abs(jv))
    # entry by entry for one column per field: 27 / 18 / 18 structural entries, the bulk of them at literal precision
    for c, nnz in ((820, 27), (1600 + 820, 18), (3200 + 820, 18)):
        e = np.zeros(N)
        e[c] = 1.0
        h = 1e-4 * abs(y[c])
        fdc = (O.swe_fun(y + h * e) - O.swe_fun(y - h * e)) / (2 * h)
        jc = O.swe_jac_mul(y, e)
        assert np.count_nonzero(jc) == nnz and np.count_nonzero(np.abs(fdc) > 1e-9) == nnz
        rel = np.abs(fdc - jc)[jc != 0] / np.abs(fdc[jc != 0])
        assert np.median(rel) < 1e-6 and np.count_nonzero(rel < 2e-6) >= nnz - 6


def test_erk_adj_forward_sweep_counters_and_sensitivities():
    y0 = O.swe_initial_state()
    t1 = 5 * DT
    tol = np.full(N, 1.0e-4)
    for method, S in ((0, 7), (1, 3), (3, 5)):
        ic = np.zeros(20, dtype=np.int32)
        ic[2] = method
        a = O.erk_adj("swe", y0, 2, 0.0, t1, tol, tol, ic)
        f = O.erk("swe", y0, 0.0, t1, tol, tol, ic)
        assert a["ierr"][0] == 1 and np.array_equal(a["y"], f["y"]) and np.array_equal(a["rstatus"], f["rstatus"])
        Ia, If = a["istatus"][0], f["istatus"][0]
        assert Ia[0] == If[0] and Ia[3] == If[3] and Ia[4] == If[4]
        assert Ia[1] == S * If[3] and Ia[2] == If[2] + If[3]          # one JAC per stage and one Nstp per popped step
    # Lambda(:, k) = d y_k(t1) / d y(t0): against finite differences of the state sweep and against the tangent linear model
    a = O.erk_adj("swe", y0, 3, 0.0, t1, tol, tol)
    lam = a["lambda_"][0]
    rng = np.random.default_rng(1)
    v = rng.standard_normal(N)
    eps = 1e-5
    yp = O.erk("swe", y0 + eps * v, 0.0, t1, tol, tol)["y"][0]
    ym = O.erk("swe", y0 - eps * v, 0.0, t1, tol, tol)["y"][0]
    fd = (yp - ym) / (2 * eps)
    tl = O.erk("swe", y0, 0.0, t1, tol, tol, y_tlm=v[None, None])["y_tlm"][0][0]
    for k in range(3):
        assert abs(lam[k] @ v - fd[k]) < 1e-4 * np.max(np.abs(fd[:3]))
        assert abs(lam[k] @ v - tl[k]) < 1e-3 * np.max(np.abs(tl[:3]))   # the TLM's own finite-difference increment is sqrt(rtol) = 1e-2
    # more accepted steps than Max_no_steps checkpoints: the reference STOPs in ERK_Push (:766-770); restated as IERR -6, Lambda untouched
    ic = np.zeros(20, dtype=np.int32)
    ic[3] = 3
    b = O.erk_adj("swe", y0, 1, 0.0, t1, tol, tol, ic)
    assert b["ierr"][0] == -6 and b["rstatus"][0][0] < t1 and b["istatus"][0][3] == 4 and not b["lambda_"].any()   # the fourth accepted step finds the buffer full


def test_erk_tlm_dense_jacobian_mode_is_the_dual_of_erk_adj():
    """TLM/ERK_TLM with ICNTRL(5) /= 0 (:460-474): JAC + MATMUL per stage; same state sweep as the finite-difference mode, one JAC
    per FUN and no FUN per column; its columns are the exact duals of ERK_ADJ's Lambda (same JAC, same tableau)."""
    y0 = O.swe_initial_state()
    t1 = 5 * DT
    tol = np.full(N, 1.0e-4)
    rng = np.random.default_rng(2)
    v = rng.standard_normal((2, N))
    fdm = O.erk("swe", y0, 0.0, t1, tol, tol, y_tlm=v[None])
    ic = np.zeros(20, dtype=np.int32)
    ic[4] = 1
    den = O.erk("swe", y0, 0.0, t1, tol, tol, ic, y_tlm=v[None])
    assert den["ierr"][0] == 1 and np.array_equal(den["y"], fdm["y"]) and np.array_equal(den["rstatus"], fdm["rstatus"])
    Id, If = den["istatus"][0], fdm["istatus"][0]
    assert Id[0] == Id[1] == 7 * Id[2] and If[0] == 3 * Id[0] and np.array_equal(Id[2:5], If[2:5])
    adj = O.erk_adj("swe", y0, 3, 0.0, t1, tol, tol)
    for k in range(3):
        for c in range(2):
            assert abs(adj["lambda_"][0][k] @ v[c] - den["y_tlm"][0][c][k]) <= 1e-12 * np.max(np.abs(den["y_tlm"][0][c]))
    # first column against the finite-difference mode (its increment is sqrt(rtol) = 1e-2; later columns of that mode carry the
    # earlier columns' perturbations, fact kept from the reference) and against central differences of the state sweep
    assert np.max(np.abs(den["y_tlm"][0][0] - fdm["y_tlm"][0][0])) < 2e-2 * np.max(np.abs(den["y_tlm"][0][0]))
    eps = 1e-6
    yp = O.erk("swe", y0 + eps * v[0], 0.0, t1, tol, tol)["y"][0]
    ym = O.erk("swe", y0 - eps * v[0], 0.0, t1, tol, tol)["y"][0]
    assert np.max(np.abs((yp - ym) / (2 * eps) - den["y_tlm"][0][0])) < 1e-2 * np.max(np.abs(den["y_tlm"][0][0]))


def test_generated_swe_jacobian_headers_are_current():
    """The committed headers are what tools/gen_swe_jac.py produces from the reference (skipped on boxes without it), and the
    product's copy equals the oracle's."""
    import os
    import subprocess
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [os.path.join(root, "oracle", "gen", "swe_jac_gen.h"), os.path.join(root, "fatode_b200", "csrc", "gen", "swe_jac_gen.h")]
    before = [open(p).read() for p in paths]
    assert before[0] == before[1]
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    stats = [os.stat(p) for p in paths]
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_swe_jac.py")], check=True, capture_output=True)
    same = [open(p).read() == b for p, b in zip(paths, before)]
    for p, st, ok in zip(paths, stats, same):
        if ok:
            os.utime(p, (st.st_atime, st.st_mtime))      # identical content: keep the time stamps (no needless rebuild)
    assert all(same)
