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
