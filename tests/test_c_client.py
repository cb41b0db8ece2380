"""The plain-C client of the C-ABI (tests/c/test_cabi.c): compiled with gcc against include/fv3tracer.h and linked with
libfv3tracer.so -- the way a C or Fortran host binds the library, without Python in between.  On a box without a CUDA device
the client checks the documented loud failure; on the B200 it calls every f64 entry point (see the file header)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_client(tmp_path):
    from fv3atm_b200 import build
    build.build()
    exe = str(tmp_path / "test_cabi")
    libdir = os.path.join(ROOT, "fv3atm_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "test_cabi.c"),
                    "-o", exe, "-L", libdir, "-lfv3tracer", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def _has_gpu():
    from fv3atm_b200 import lib as L
    return L.load().fv3t_device_count() > 0


def test_c_client_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build_client(tmp_path)
    if _has_gpu():
        pytest.skip("a CUDA device is present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_client_calls_every_entry_point(tmp_path):
    exe = _build_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
