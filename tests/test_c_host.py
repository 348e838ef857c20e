"""A plain C program compiled against include/fatode_b200.h (the header as a C compiler sees it, not the ctypes mirror) and
linked with libfatode_b200.so: builds on the CPU box; on a GPU box it runs the sharded adjoint with the library's own NCCL
all-reduce of the Mu sums — one process per GPU, no torch anywhere."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c", "abi_two_ranks")


def build_exe():
    from fatode_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-D_DEFAULT_SOURCE", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_two_ranks.c"),
           "-o", EXE, "-L" + libdir, "-lfatode_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.run(cmd, check=True)
    return EXE


def test_c_host_compiles_and_links_against_the_header():
    exe = build_exe()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_host_runs_sharded_adjoint_with_in_library_allreduce():
    import torch
    exe = build_exe()
    n = min(torch.cuda.device_count(), 2)
    r = subprocess.run([exe, str(n), "24", "2.0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert f"comm_size {n}" in r.stdout


def test_every_prototype_of_the_header_links_from_plain_c(tmp_path):
    """Strict C99 translation unit that takes the address of every function include/fatode_b200.h declares (so the header has to
    parse as C, not C++, and every declared symbol has to exist with C linkage in the library)."""
    import re
    from fatode_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    hdr = open(os.path.join(ROOT, "include", "fatode_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(fatode_[a-z0-9_]+)\s*\(", body)))
    assert len(names) >= 30 and "fatode_erk_integrate_adj_batch" in names and "fatode_sdirk_integrate_tlm_batch" in names
    src = tmp_path / "all_symbols.c"
    src.write_text('#include "fatode_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n  fn table[] = {\n'
                   + "".join(f"    (fn){n},\n" for n in names)
                   + '  };\n  printf("%d\\n", (int)(sizeof table / sizeof table[0]));\n  return table[0] == (fn)0;\n}\n')
    exe = tmp_path / "all_symbols"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L" + libdir, "-lfatode_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and int(r.stdout.strip()) == len(names)
