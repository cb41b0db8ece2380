"""ctypes binding of the C-ABI library (include/fv3tracer.h).  There is no CPU fallback: if the library is
missing it must be built (python -m fv3atm_b200.build), and every compute call fails loudly without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfv3tracer.so")

FIELD = {"q": 0, "dp1": 1, "mfx": 2, "mfy": 3, "cx": 4, "cy": 5, "pe": 6, "delp": 7}
KCLASS = {"advect": 0, "remap": 1, "halo": 2, "cmax": 3, "scale": 4}

# every symbol include/fv3tracer.h declares
_PER_PREC = ["create", "tracer_2d", "tracer_2d_1L", "set_damping", "remap_tracers", "tracer_step", "mapn_tracer", "map_scalar", "map1_ppm", "map_field", "fv_tp_2d", "upload", "download", "set_vertical",
             "tracer_2d_resident", "remap_tracers_resident", "remap_prepare", "tracer_2d_begin", "tracer_2d_set_cmax", "halo_local",
             "halo_pack", "halo_unpack", "halo_pack_host", "halo_unpack_host", "halo_gather", "halo_scatter", "tracer_2d_substep", "tracer_2d_finish"]
_COMMON = ["fv3t_last_error", "fv3t_device_count", "fv3t_destroy", "fv3t_sync", "fv3t_device_ptr", "fv3t_halo_strip_elems",
           "fv3t_neighbor", "fv3t_halo_list_create", "fv3t_halo_list_count", "fv3t_halo_local_table", "fv3t_kernel_launches", "fv3t_timer_start", "fv3t_timer_stop_ms", "fv3t_profile_enable",
           "fv3t_profile_get_ms"]
EXPORTS = _COMMON + [f"fv3t_{p}_{f}" for p in ("f64", "f32") for f in _PER_PREC]


class Dims(C.Structure):
    _fields_ = [("npx", C.c_int), ("npz", C.c_int), ("nq_max", C.c_int), ("ntiles", C.c_int), ("tile_id", C.c_int * 6),
                ("sub_layout", C.c_int), ("sub_bi", C.c_int * 6), ("sub_bj", C.c_int * 6)]


class GridPtrs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("area", "rarea", "dx", "dy", "dxa", "dya", "sin_sg")]


class Fv3tError(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Fv3tError(f"{LIB_PATH} is missing: build it with `python -m fv3atm_b200.build` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.fv3t_last_error.restype = C.c_char_p
        _lib.fv3t_device_ptr.restype = C.c_void_p
        _lib.fv3t_halo_strip_elems.restype = C.c_size_t
        _lib.fv3t_kernel_launches.restype = C.c_uint64
    return _lib


def check(rc: int):
    if rc != 0:
        raise Fv3tError(load().fv3t_last_error().decode())


def prec(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(f"unsupported dtype {dtype}")


def fn(dtype, name):
    p, _ = prec(dtype)
    return getattr(load(), f"fv3t_{p}_{name}")


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))
