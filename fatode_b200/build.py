"""Builds the CUDA shared library in-tree (fatode_b200/lib/libfatode_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
-fmad=false: a*b+c is never contracted implicitly (the forward sweep reproduces the reference's
operation order); the adjoint kernels spell fma() explicitly.

Every csrc/*.cu is one translation unit (compiled in parallel into fatode_b200/lib/obj/, rebuilt only when it or a
header it includes — transitively — changed), linked into one shared library.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libfatode_b200.so")
OBJ = os.path.join(HERE, "lib", "obj")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC"]
_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def units():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


SRC = units()


def deps(path, seen=None):
    """the file and every quoted include reachable from it"""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in _INC.findall(open(path).read()):
        deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def sources():
    d = set()
    for u in units():
        d |= deps(u)
    return sorted(d)


def obj_of(unit):
    return os.path.join(OBJ, os.path.basename(unit)[:-3] + ".o")


def stale(unit):
    o = obj_of(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in deps(unit))


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(not stale(u) and os.path.getmtime(obj_of(u)) <= t for u in units())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    todo = [u for u in units() if force or stale(u)]

    def compile_unit(u):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj_of(u), u]
        return u, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for u, cmd, r in ex.map(compile_unit, todo):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
            if verbose:
                sys.stderr.write(r.stderr)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [obj_of(u) for u in units()]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
