"""Builds the CUDA shared library in-tree (fatode_b200/lib/libfatode_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
-fmad=false: a*b+c is never contracted implicitly (the forward sweep reproduces the reference's
operation order); the adjoint kernels spell fma() explicitly.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "fatode_capi.cu")]
OUT = os.path.join(HERE, "lib", "libfatode_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    deps = list(SRC)
    for root, _, files in os.walk(os.path.join(HERE, "csrc")):
        deps += [os.path.join(root, f) for f in files if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "fatode_b200.h"))
    return deps


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
