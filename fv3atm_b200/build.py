"""Build the C-ABI shared library fv3atm_b200/libfv3tracer.so for sm_100a (in-tree, nvcc cross-compiles
without a GPU).  The library statically links cudart and does not link libcuda, so it loads on a CPU-only
box (every compute entry point then fails loudly with "no CUDA device")."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfv3tracer.so")
SOURCES = ["fv3t_api.cu"]
HEADERS = ["fv3t_common.cuh", "fv3t_ppm.cuh", "fv3t_advect.cuh", "fv3t_remap.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact contract with the FMA-free oracle: no contraction, IEEE division/sqrt (nvcc defaults)
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "fv3tracer.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that lack an OpenMP spec file; nvcc only needs a plain host g++
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfv3tracer.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
