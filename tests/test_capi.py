"""CPU-side checks of the drop-in boundary: the CUDA library builds, loads, exports every symbol
include/fatode_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fatode_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fatode_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from fatode_b200 import build
    so = build.build()
    L = ctypes.CDLL(so)
    syms = declared_symbols()
    assert "fatode_ros_integrate_adj_batch" in syms and "fatode_ros_integrate_batch" in syms
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/fatode_b200.h but not exported"


def test_model_registry_and_driver_states():
    import fatode_b200 as F
    assert F.model_dims("cbm4") == (32, 29)
    assert F.model_dims("small_strato") == (5, 0)
    assert F.model_dims("roberts") == (3, 3)
    import oracle_lib as O
    for m in ("cbm4", "small_strato", "roberts"):
        assert np.array_equal(F.initial_state(m), O.initial_state(m))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fatode_b200 as F
    with pytest.raises(F.FatodeError):
        F.integrate("cbm4", 0.0, 1.0, F.initial_state("cbm4"), np.full(32, 1e-5), np.full(32, 1e-7))


def test_inputs_generator_is_counter_based():
    from fatode_b200.inputs import perturbed_states
    import fatode_b200 as F
    base = F.initial_state("cbm4")
    a = perturbed_states(base, 64)
    b = perturbed_states(base, 16, first_cell=32)
    assert np.array_equal(a[32:48], b)
    assert np.array_equal(a[0], base)
    assert np.all(a[:, 6] == 0.0) and np.all(np.abs(a[1:, 0] / base[0] - 1.0) <= 0.1)
    assert len({tuple(r) for r in a}) == 64


def test_swe_ensemble_inputs_shard_consistently_and_match_the_oracle():
    """Host-side input generation of the shallow-water ensemble (no GPU): any rank can regenerate its members from the global
    member index, and the library's initial state equals the oracle's bit for bit."""
    import numpy as np
    import fatode_b200 as F
    from fatode_b200.inputs import swe_members
    import oracle_lib as O
    whole = swe_members(6)
    parts = np.concatenate([swe_members(2, first_member=0), swe_members(4, first_member=2)])
    assert np.array_equal(whole, parts)
    assert np.array_equal(whole[0], O.swe_initial_state())          # member 0 is the driver's problem
    assert np.array_equal(F.swe_initial_state(33.0, 1.5, -2.0), O.swe_initial_state(33.0, 1.5, -2.0))
    assert len({float(m[:1600].max()) for m in whole}) == 6         # distinct amplitudes
