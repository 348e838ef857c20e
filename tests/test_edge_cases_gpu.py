"""Edge cases of the entry points added in the second half of round 2, through the C ABI: empty batches, one system with one
column, device-resident buffers (FATODE_MEM_DEVICE), argument errors."""
import ctypes

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
F01 = float(np.float32(0.1))
T0 = 12.0 * 3600.0
DT = float(np.float32(2.1573758800627289e-003))


def test_empty_batches_return_at_once():
    import fatode_b200 as F
    tol32 = np.full(32, 1e-4)
    tol = np.full(4800, 1e-4)
    r = F.integrate_erk_adj("swe", 0.0, 5 * DT, np.zeros((0, 4800)), 1, tol, tol)
    assert r.ierr.shape == (0,) and r.lambda_.shape == (0, 1, 4800)
    ic = np.zeros(20, dtype=np.int32)
    ic[6] = 1
    r = F.integrate_tlm("cbm4", T0, T0 + 60.0, np.zeros((0, 32)), np.zeros((0, 2, 32)), tol32, tol32, icntrl_u=ic, family="sdirk")
    assert r.ierr.shape == (0,)
    ic[6] = 2
    r = F.integrate_adj("cbm4", T0, T0 + 60.0, np.zeros((0, 32)), 2, tol32, tol32, icntrl_u=ic, family="sdirk",
                        atol_adj=np.ones((2, 32)), rtol_adj=np.ones((2, 32)))
    assert r.ierr.shape == (0,)


def test_one_system_one_column():
    import fatode_b200 as F
    rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
    y0 = F.initial_state("cbm4")[None]
    t1 = T0 + 3600.0
    yt = np.zeros((1, 1, 32))
    yt[0, 0, 5] = 1.0
    tt = np.full((1, 32), 1.0e-4)
    for k in (6, 8, 11):
        ic = np.zeros(20, dtype=np.int32)
        ic[1], ic[k] = 1, 1                                  # scalar tolerances
        ref = O.sdirk_tlm("cbm4", y0, yt, T0, t1, atol, rtol, ic, atol_tlm=tt, rtol_tlm=tt)
        res = F.integrate_tlm("cbm4", T0, t1, y0, yt, rtol, atol, icntrl_u=ic, atol_tlm=tt, rtol_tlm=tt, family="sdirk")
        assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.y, ref["y"]) and np.array_equal(res.y_tlm, ref["y_tlm"])
    ic = np.zeros(20, dtype=np.int32)
    ic[6] = 2
    ref = O.sdirk_adj("cbm4", y0, 1, T0, t1, atol, rtol, ic, atol_adj=tt, rtol_adj=tt)
    res = F.integrate_adj("cbm4", T0, t1, y0, 1, rtol, atol, icntrl_u=ic, family="sdirk", atol_adj=tt, rtol_adj=tt)
    assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.lambda_, ref["lambda_"])
    tol = np.full(4800, 1e-3)
    m = O.swe_initial_state()[None]
    ref = O.erk_adj("swe", m, 1, 0.0, 3 * DT, tol, tol)
    res = F.integrate_erk_adj("swe", 0.0, 3 * DT, m, 1, tol, tol)
    assert np.array_equal(res.istatus, ref["istatus"]) and np.array_equal(res.y, ref["y"])
    assert np.max(np.abs(res.lambda_ - ref["lambda_"])) <= 1e-9 * np.max(np.abs(ref["lambda_"]))


def test_erk_adj_device_resident_buffers_and_argument_errors():
    import torch
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    L = F.lib()
    n, nadj = 7, 2
    tol = np.full(4800, 1e-4)
    y0 = swe_members(n)
    ref = F.integrate_erk_adj("swe", 0.0, 5 * DT, y0, nadj, tol, tol)      # host-buffer path (checked against the oracle elsewhere)
    dev = torch.device("cuda", 0)
    y = torch.from_numpy(y0.copy()).to(dev)
    lam = torch.zeros((n, nadj, 4800), dtype=torch.float64, device=dev)
    ist = torch.zeros((n, 20), dtype=torch.int32, device=dev)
    rst = torch.zeros((n, 20), dtype=torch.float64, device=dev)
    ierr = torch.zeros((n,), dtype=torch.int32, device=dev)
    vp = ctypes.c_void_p
    p = lambda a: vp(a.ctypes.data)
    rc = L.fatode_erk_integrate_adj_batch(F.model_id("swe"), n, 4800, 0, nadj, vp(y.data_ptr()), vp(lam.data_ptr()), None, 0.0, 5 * DT,
                                          p(tol), p(tol), None, None, vp(ist.data_ptr()), vp(rst.data_ptr()), vp(ierr.data_ptr()), None,
                                          F.MEM_DEVICE, vp(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0
    assert np.array_equal(y.cpu().numpy(), ref.y) and np.array_equal(lam.cpu().numpy(), ref.lambda_)
    assert np.array_equal(ist.cpu().numpy(), ref.istatus) and (ierr.cpu().numpy() == 1).all()
    # argument errors: parameter sensitivities (no parameters in this model), nadj out of range, wrong dimension, wrong model
    mu = np.zeros((n, nadj, 3))
    hy, hl = y0.copy(), np.zeros((n, nadj, 4800))
    hi, hr, he = np.zeros((n, 20), dtype=np.int32), np.zeros((n, 20)), np.zeros(n, dtype=np.int32)
    call = lambda model, nvar, np_, na, pmu: L.fatode_erk_integrate_adj_batch(model, n, nvar, np_, na, p(hy), p(hl), pmu, 0.0, 5 * DT, p(tol), p(tol),
                                                                              None, None, p(hi), p(hr), p(he), None, F.MEM_HOST, None)
    assert call(F.model_id("swe"), 4800, 3, nadj, p(mu)) == -105
    assert call(F.model_id("swe"), 4800, 0, 0, None) == -101
    assert call(F.model_id("swe"), 32, 0, nadj, None) == -101
    assert call(F.MODEL_CBM4, 32, 0, nadj, None) < 0
