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
# Two translation units: the strict kernels + host orchestration keep the bit-exact contract with the FMA-free oracle
# (no contraction, IEEE division/sqrt); the production kernels (fv3t_fast.cu) are built with FMA contraction on.
# The fast TU is compiled once per precision and the exact TU once per precision and parameter type (seven objects, in parallel):
# the k_advect5 instantiations dominate the build time.
SOURCES = [("fv3t_api.cu", "fv3t_api.o", ["--fmad=false"]),
           ("fv3t_fast.cu", "fv3t_fast_f64.o", ["--fmad=true", "-DFV3T_INST_F64"]),
           ("fv3t_fast.cu", "fv3t_fast_f32.o", ["--fmad=true", "-DFV3T_INST_F32"]),
           ("fv3t_exact.cu", "fv3t_exact_f64.o", ["--fmad=false", "-DFV3T_INST_F64", "-DFV3T_INST_WHOLE"]),
           ("fv3t_exact.cu", "fv3t_exact_f32.o", ["--fmad=false", "-DFV3T_INST_F32", "-DFV3T_INST_WHOLE"]),
           ("fv3t_exact.cu", "fv3t_exact_sub_f64.o", ["--fmad=false", "-DFV3T_INST_F64", "-DFV3T_INST_SUB"]),
           ("fv3t_exact.cu", "fv3t_exact_sub_f32.o", ["--fmad=false", "-DFV3T_INST_F32", "-DFV3T_INST_SUB"])]
HEADERS = ["fv3t_common.cuh", "fv3t_ppm.cuh", "fv3t_advect.cuh", "fv3t_advect2.cuh", "fv3t_advect3.cuh", "fv3t_advect4.cuh", "fv3t_advect5.cuh", "fv3t_advect5_launch.cuh", "fv3t_deln.cuh", "fv3t_tp2d.cuh", "fv3t_remap4.cuh", "fv3t_remap.cuh",
           "fv3t_remap2.cuh", "fv3t_remap3.cuh", "fv3t_remap5.cuh", "fv3t_fast.h", "fv3t_fast.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "--extended-lambda", "-Xcompiler", "-fPIC",
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
    deps = [os.path.join(CSRC, f) for f in [x[0] for x in SOURCES] + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "fv3tracer.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    env = dict(os.environ)
    procs = []
    dev = ["-DFV3T_A5_DEV"] if os.environ.get("FV3T_A5_DEV") else []  # development builds: k_advect5 for hord 8 / 10 only
    dev += [x for x in os.environ.get("FV3T_EXTRA_DEFS", "").split() if x]  # experiments, e.g. -DFV3T_NO_SUB
    only = [x for x in os.environ.get("FV3T_BUILD_ONLY", "").split(",") if x]  # development: recompile the named objects only
    reuse = []
    for src, objname, extra in SOURCES:
        obj = os.path.join(CSRC, objname)
        if only and not any(o in objname for o in only) and os.path.exists(obj):
            reuse.append(obj)
            continue
        extra = extra + dev
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((obj, subprocess.Popen(cmd, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = list(reuse)
    for obj, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed compiling {obj}")
        objs.append(obj)
    # cudart is linked statically and libcuda is not linked, so the library loads on a CPU-only box
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB] + objs,
                       cwd=CSRC, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libfv3tracer.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
