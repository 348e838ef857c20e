"""Build-level guard (no GPU needed): the record-writing kernels must keep their 32-byte stores.
ptxas 12.9 once narrowed every `st.global.v4.f64` of the generated jac_values<32> inside k_sdirk_tlm_prep to a 64-bit store of
its first element (three quarters of each stage Jacobian silently missing) when a second kernel started to share that template
instantiation.  The GPU parity tests catch the consequence; this test catches the cause at build time, on the CPU box."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fatode_b200", "lib", "libfatode_b200.so")
# kernel -> minimum number of 256-bit global stores in its SASS (generated jac_values / hess / LU record writers)
KERNELS = {"k_sdirk_tlm_prepE": 75, "k_sdirk_adj_prepE": 84, "k_rk_adj_prepE": 115, "k_expand_tape": 150,
           "k_sdirk_tlm_prep_dense": 70, "k_sdirk_adj_prep_dense": 75, "k_rk_adj_prep_dense": 70}


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_preparation_kernels_keep_their_256_bit_stores():
    import fatode_b200.build as B
    B.build()
    elf = subprocess.run(["cuobjdump", "-elf", SO], capture_output=True, text=True).stdout
    for k, nmin in KERNELS.items():
        syms = sorted(x for x in set(re.findall(r"_ZN6fatode\d+%s[A-Za-z0-9_]*" % k, elf)) if "_param_" not in x)
        assert syms, k
        for sym in syms:            # every instantiation (k_expand_tape: staged and unstaged inputs)
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
            kinds = re.findall(r"\bSTG[.A-Za-z0-9]*", sass)
            wide = sum(1 for s in kinds if s.endswith(".256"))
            narrow = sum(1 for s in kinds if s.endswith(".64"))
            assert wide >= nmin and narrow == 0, (sym, wide, narrow)
