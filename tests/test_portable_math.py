"""The correctly rounded elementary functions behind the step sequence: COS of Update_SUN and Err**(1/ELO)
of the Rosenbrock step controller.  CPU: the oracle's copy against a multi-precision reference (mpmath) — every sample must be
THE correctly rounded double.  GPU: the product's copy evaluated on the device, bit for bit against the oracle's."""
import ctypes

import numpy as np
import pytest

import oracle_lib as O


def _probe_oracle(x, elo):
    L = O.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.oracle_probe_math.argtypes = [ctypes.c_int, dp, ctypes.c_double, dp, dp]
    L.oracle_probe_math.restype = None
    x = np.ascontiguousarray(x, dtype=np.float64)
    c, r = np.empty_like(x), np.empty_like(x)
    L.oracle_probe_math(len(x), x.ctypes.data_as(dp), float(elo), c.ctypes.data_as(dp), r.ctypes.data_as(dp))
    return c, r


def _samples(n, seed):
    rng = np.random.default_rng(seed)
    # the arguments Update_SUN produces: PI * Ttmp, |Ttmp| <= 1, plus the whole supported range and the quadrant boundaries
    x = np.concatenate([3.14159265358979 * rng.uniform(-1.0, 1.0, n), rng.uniform(-3.92, 3.92, n // 4),
                        np.pi / 4 * np.arange(-5, 6) * (1 + rng.uniform(-1e-15, 1e-15, 11)), [0.0, 1e-300, 1e-8, -1e-8]])
    return x


def test_cos_is_correctly_rounded():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    x = _samples(6000, 1)
    c, _ = _probe_oracle(x, 4.0)
    ref = np.array([float(mp.cos(mp.mpf(float(v)))) for v in x])     # mpmath rounds the 200-bit value to nearest
    assert np.array_equal(c, ref), np.flatnonzero(c != ref)[:5]


@pytest.mark.parametrize("elo", [2.0, 3.0, 4.0])
def test_error_root_is_correctly_rounded(elo):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(int(elo))
    # Err ranges from its floor 1e-10 to large values after a failed step; include exact powers
    x = np.concatenate([10.0 ** rng.uniform(-10, 6, 6000), rng.uniform(0.5, 2.0, 2000), [1.0, 16.0, 81.0, 1e-10, 0.0625, 27.0, 8.0]])
    _, r = _probe_oracle(x, elo)
    y = mp.mpf(float(1.0 / elo))                                      # the reference raises to ONE/ros_ELO, a double
    ref = np.array([float(mp.power(mp.mpf(float(v)), y)) for v in x])
    assert np.array_equal(r, ref), np.flatnonzero(r != ref)[:5]


@pytest.mark.gpu
def test_device_kernels_equal_the_oracle_copy_bitwise():
    import fatode_b200 as F
    L = F.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    L.fatode_probe_math.argtypes = [ctypes.c_int, dp, ctypes.c_double, dp, dp]
    L.fatode_probe_math.restype = ctypes.c_int
    x = _samples(200000, 3)
    for elo in (2.0, 3.0, 4.0):
        xx = np.ascontiguousarray(np.concatenate([x, 10.0 ** np.random.default_rng(5).uniform(-10, 6, 200000)]))
        c, r = np.empty_like(xx), np.empty_like(xx)
        # cos is only defined for |x| <= 5 pi / 4: feed the root samples through a clamp for the cos column
        xc = np.clip(xx, -3.92, 3.92)
        assert L.fatode_probe_math(len(xx), xc.ctypes.data_as(dp), elo, c.ctypes.data_as(dp), r.ctypes.data_as(dp)) == 0
        oc, _ = _probe_oracle(xc, elo)
        assert np.array_equal(c, oc)
        assert L.fatode_probe_math(len(xx), xx.ctypes.data_as(dp), elo, c.ctypes.data_as(dp), r.ctypes.data_as(dp)) == 0
        _, orr = _probe_oracle(xx, elo)
        assert np.array_equal(r, orr)
