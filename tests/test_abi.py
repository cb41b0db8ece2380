"""CPU checks of the drop-in boundary: libfv3tracer.so loads without a GPU, exports every symbol include/fv3tracer.h
declares (and the binding's list is in sync with the header), and fails loudly -- never silently -- when no CUDA
device exists.  No compute entry point is exercised here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fv3tracer.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    per_prec = set(re.findall(r"fv3t_##P##_(\w+)\s*\(", src))
    common = set(re.findall(r"\b(fv3t_(?!##)\w+)\s*\(", src))
    common = {c for c in common if not c.startswith("fv3t_f64") and not c.startswith("fv3t_f32")}
    out = set(common)
    for p in ("f64", "f32"):
        out |= {f"fv3t_{p}_{f}" for f in per_prec}
    return out


@pytest.fixture(scope="module")
def lib():
    from fv3atm_b200 import build, lib as L
    build.build()
    return L


def test_library_exports_every_declared_symbol(lib):
    l = lib.load()
    declared = _header_symbols()
    assert len(declared) >= 40
    missing = [s for s in sorted(declared) if not hasattr(l, s)]
    assert not missing, missing
    # the python binding's export list is the header's
    assert set(lib.EXPORTS) == declared, set(lib.EXPORTS) ^ declared


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "fv3tracer.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under fv3atm_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "fv3atm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_binding" not in txt and "liboracle" not in txt and "fv3_oracle" not in txt, os.path.join(dp, f)


def test_fails_loudly_without_a_gpu(lib):
    l = lib.load()
    if l.fv3t_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from fv3atm_b200 import cubed_sphere as cs
    from fv3atm_b200.tracer import TracerContext
    g = cs.make_grid(8).astype("float64")
    with pytest.raises(lib.Fv3tError, match="no CUDA device"):
        TracerContext(9, 8, 2, g, dtype=np.float64)


def test_null_context_is_an_error_not_a_crash(lib):
    l = lib.load()
    rc = l.fv3t_f64_tracer_2d_resident(None, 1, 8, 0, C.c_double(1.0), None)
    assert rc != 0 and b"null context" in l.fv3t_last_error()
    assert l.fv3t_sync(None) != 0
    assert l.fv3t_destroy(None) == 0
