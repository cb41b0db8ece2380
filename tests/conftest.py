import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices() -> int:
    try:
        from fv3atm_b200 import build, lib as L
        build.build()
        return int(L.load().fv3t_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing them (the library has no
    CPU fallback, so they cannot pass there); `-m gpu` on the B200 runs them."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU fallback); run with -m gpu on a B200")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding as ob
    ob.build()
    return ob


_CASES = {}


def get_case(n, npz, nq, dtype="float64", courant=0.7, divergent=0.15):
    """Session cache of synthetic cases (grid generation dominates set-up time)."""
    import numpy as np
    from fv3atm_b200 import synthetic as sy
    key = (n, npz, nq, str(np.dtype(dtype)), courant, divergent)
    if key not in _CASES:
        gkey = ("grid", n)
        if gkey not in _CASES:
            from fv3atm_b200 import cubed_sphere as cs
            _CASES[gkey] = cs.make_grid(n)
        _CASES[key] = sy.make_case(n, npz, nq, dtype=dtype, courant=courant, divergent=divergent, grid=_CASES[gkey])
    return _CASES[key]


@pytest.fixture(scope="session")
def case_factory():
    return get_case
