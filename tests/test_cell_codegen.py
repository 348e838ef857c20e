"""CPU check of the generated per-thread CBM-IV code (fatode_b200/csrc/gen/cbm4_cell_gen.h) against the oracle:
tests/cell_codegen_check.cpp replays every DGETF2 / DGETRS('N') the oracle performs along real trajectories
through the generated static-pattern lu_factor / lu_solve and compares bit for bit (plus FUN, E, Hessian values
and the transposed solve of the backward sweep to rounding); family 2 does the same for the ZGETF2 / ZGETRS('N') calls
of the implicit Runge-Kutta integrator (zlu_factor / zlu_solve, all five methods).  No GPU involved: the header compiles for the host."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cellgen") / "cell_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", exe,
                    os.path.join(HERE, "cell_codegen_check.cpp")], check=True, cwd=HERE)
    return exe


@pytest.mark.parametrize("cells,hours,adj_family", [(16, 72.0, 1), (8, 72.0, 0), (10, 72.0, 2)])
def test_generated_lu_matches_dgetf2_bitwise(checker, cells, hours, adj_family):
    r = subprocess.run([checker, str(cells), str(hours), str(adj_family)], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    f = dict(zip(r.stdout.split()[0::2], r.stdout.split()[1::2]))
    assert int(f["lu_mismatch"]) == 0 and int(f["solve_mismatch"]) == 0 and int(f["fun_mismatch"]) == 0
    if adj_family == 2:   # FWD/RK: every real factorisation is followed by a complex one (ZGETF2 / ZGETRS on two planes)
        z = dict(zip(r.stdout.split("complex:")[1].split()[0::2], r.stdout.split("complex:")[1].split()[1::2]))
        assert int(z["LUs"]) == int(f["LUs"]) > 300 * cells and int(z["lu_mismatch"]) == 0 and int(z["solve_mismatch"]) == 0
        assert int(z["solves"]) > 2 * int(z["LUs"])
    else:
        assert int(f["LUs"]) > 800 * cells and int(f["solves"]) == 6 * (int(f["LUs"]) - int(f["irregular"]))


def test_generated_header_is_current():
    """The committed header is what tools/gen_cbm4_cell.py produces from the reference (skipped on boxes without it)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    path = os.path.join(ROOT, "fatode_b200", "csrc", "gen", "cbm4_cell_gen.h")
    before = open(path).read()
    st = os.stat(path)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_cbm4_cell.py")], check=True, capture_output=True)
    same = open(path).read() == before
    if same:
        os.utime(path, (st.st_atime, st.st_mtime))   # identical content: keep the time stamp so that the library is not rebuilt
    assert same
