"""Batch container ("FATODEB1") and the reference's text formats through the C ABI (host code: runs without a GPU)."""
import json
import os

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_container_round_trip_and_sharded_read(tmp_path):
    from fatode_b200 import io
    rng = np.random.default_rng(0)
    nc = 37
    y, lam, mu = rng.standard_normal((nc, 32)), rng.standard_normal((nc, 5, 32)), rng.standard_normal((nc, 5, 29))
    ist, rst, ierr = rng.integers(0, 1000, (nc, 20)).astype(np.int32), rng.standard_normal((nc, 20)), np.ones(nc, dtype=np.int32)
    p = tmp_path / "batch.fbt"
    io.save_batch(p, "cbm4", 43200.0, 302400.0, y, lam, mu, ist, rst, ierr)
    info = io.batch_info(p)
    assert (info["ncells"], info["nvar"], info["np"], info["nadj"], info["tin"], info["tout"]) == (nc, 32, 29, 5, 43200.0, 302400.0)
    assert info["arrays"] == ["y", "lambda", "mu", "istatus", "rstatus", "ierr"]
    b = io.load_batch(p)
    for k, a in (("y", y), ("lambda", lam), ("mu", mu), ("istatus", ist), ("rstatus", rst), ("ierr", ierr)):
        assert np.array_equal(b[k], a)
    s = io.load_batch(p, first_cell=10, ncells=7)                  # a rank's shard
    assert np.array_equal(s["lambda"], lam[10:17]) and np.array_equal(s["ierr"], ierr[10:17])
    # sections are 64-byte aligned and laid out exactly like the C ABI arrays: y of cell 3 sits at offset + 3 * 32 * 8
    raw = open(p, "rb").read()
    assert raw[:8] == b"FATODEB1"
    off = int.from_bytes(raw[56 + 40:56 + 48], "little")
    assert off % 64 == 0 and np.array_equal(np.frombuffer(raw, dtype=np.float64, count=32, offset=off + 3 * 256), y[3])
    # inputs only (y0 of a batch): the other arrays are simply absent
    io.save_batch(tmp_path / "y0.fbt", "cbm4", 0.0, 1.0, y)
    assert io.batch_info(tmp_path / "y0.fbt")["arrays"] == ["y"]


def test_text_files_have_the_reference_format(tmp_path):
    """cbm4_ros_sol.txt / cbm4_output_sens.txt: 'E24.16' lines and the 'Lambda: column' / 'Mu: column' headings of
    cbm4_ros_adj_dr.F90:134-164; the golden values survive a write/read cycle to the printed 16 digits."""
    from fatode_b200 import io
    g = json.load(open(os.path.join(G, "cbm4_ros_adj.json")))
    y = np.array(g["y"])
    lam, mu = np.array(g["lambda"]).reshape(32, 32), np.array(g["mu"]).reshape(32, 29)
    io.write_solution_text(tmp_path / "sol.txt", y)
    lines = open(tmp_path / "sol.txt").read().splitlines()
    assert len(lines) == 32 and all(len(l) == 24 for l in lines)
    assert lines[0] == "  0.3719123500361147E-01" and lines[1] == "  0.3233107660661603E+12"     # the reference file's first lines
    assert np.array_equal(io.read_solution_text(tmp_path / "sol.txt", 32), y)       # golden values are 16-digit decimals
    io.write_sens_text(tmp_path / "sens.txt", lam, mu)
    txt = open(tmp_path / "sens.txt").read().splitlines()
    assert txt[0] == "Lambda: column  1" and txt[33] == "Lambda: column  2" and txt[32 * 33] == "Mu: column  1"
    l2, m2 = io.read_sens_text(tmp_path / "sens.txt", 32, 29, 32)
    assert np.array_equal(l2, lam) and np.array_equal(m2, mu)
    io.write_solution_text(tmp_path / "z.txt", np.array([0.0, -1.5, 1e-300, -2.5e300]))
    assert open(tmp_path / "z.txt").read().splitlines() == ["  0.0000000000000000E+00", " -0.1500000000000000E+01",
                                                           "  0.1000000000000000-299", " -0.2500000000000000+301"]
