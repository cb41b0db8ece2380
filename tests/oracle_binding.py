"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by fv3atm_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "fv3_oracle_advect.hpp", "fv3_oracle_remap.hpp")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def halo_offsets(n: int):
    """Flat (dst, src) offsets into the tile-major stack of halo-padded planes, from the package's
    contact-table maps."""
    from fv3atm_b200 import cubed_sphere as cs
    dt, dj, di, st, sj, si = cs.halo_index_table(n)
    nd = n + 6
    dst = (dt.astype(np.int64) * nd + dj) * nd + di
    src = (st.astype(np.int64) * nd + sj) * nd + si
    return np.ascontiguousarray(dst), np.ascontiguousarray(src)


def ppm_line(q1, c, dxa, iord, is_, ie, isd, npx, edges, lim_fac=1.0):
    """q1 on isd..ied, c on is..ie+1 -> flux on is..ie+1"""
    s, ct = _sfx(q1.dtype)
    flux = np.zeros(ie - is_ + 2, dtype=q1.dtype)
    q1 = np.ascontiguousarray(q1)
    c = np.ascontiguousarray(c, dtype=q1.dtype)
    dxa = np.ascontiguousarray(dxa, dtype=q1.dtype)
    getattr(lib(), f"orc_{s}_ppm_line")(_p(flux), _p(q1), _p(c), _p(dxa), int(iord), int(is_), int(ie), int(isd), int(npx),
                                        int(edges), ct(lim_fac))
    return flux


def tracer_2d(case, hord=8, q_split=0, lim_fac=1.0, grid_arrays=None, inplace=False, halo=None):
    """Run the oracle's tracer_2d on (copies of) a synthetic Case.  Returns dict with the post-state of
    q, dp1, mfx, mfy, cx, cy and nsplt, ksplt, cmax.  inplace=True works on the Case's own arrays and `halo` takes a
    precomputed halo_offsets(n) pair, so that a timed call contains nothing but the oracle itself (bench.py)."""
    s, ct = _sfx(case.dtype)
    g = grid_arrays or case.metrics()
    if inplace:
        out = {k: getattr(case, k) for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    else:
        out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    dst, src = halo if halo is not None else halo_offsets(case.n)
    nsplt = C.c_int(0)
    ksplt = np.zeros(case.npz, dtype=np.int32)
    cmax = np.zeros(case.npz, dtype=case.dtype)
    getattr(lib(), f"orc_{s}_tracer_2d")(
        6, case.n, case.npz, case.nq, _p(out["q"]), _p(out["dp1"]), _p(out["mfx"]), _p(out["mfy"]), _p(out["cx"]),
        _p(out["cy"]), _p(g["area"]), _p(g["rarea"]), _p(g["dx"]), _p(g["dy"]), _p(g["dxa"]), _p(g["dya"]), _p(g["sin_sg"]),
        _p(dst), _p(src), C.c_int64(dst.size), int(hord), int(q_split), ct(lim_fac), C.byref(nsplt), _p(ksplt), _p(cmax))
    out["nsplt"] = nsplt.value
    out["ksplt"] = ksplt
    out["cmax"] = cmax
    return out


def tracer_2d_1l(case, hord=8, lim_fac=1.0):
    """The oracle's tracer_2d_1L (fv_tracer2d.F90:92-321) on copies of a synthetic Case; same outputs as tracer_2d
    (ksplt = the per-level sub-step counts, nsplt = their maximum)."""
    s, ct = _sfx(case.dtype)
    g = case.metrics()
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    dst, src = halo_offsets(case.n)
    nsplt = C.c_int(0)
    ksplt = np.zeros(case.npz, dtype=np.int32)
    cmax = np.zeros(case.npz, dtype=case.dtype)
    getattr(lib(), f"orc_{s}_tracer_2d_1l")(
        6, case.n, case.npz, case.nq, _p(out["q"]), _p(out["dp1"]), _p(out["mfx"]), _p(out["mfy"]), _p(out["cx"]),
        _p(out["cy"]), _p(g["area"]), _p(g["rarea"]), _p(g["dx"]), _p(g["dy"]), _p(g["dxa"]), _p(g["dya"]), _p(g["sin_sg"]),
        _p(dst), _p(src), C.c_int64(dst.size), int(hord), ct(lim_fac), C.byref(nsplt), _p(ksplt), _p(cmax))
    out["nsplt"] = nsplt.value
    out["ksplt"] = ksplt
    out["cmax"] = cmax
    return out


def remap_tracers(q, pe, ak, bk, ptop, kord_tr, fill=True, inplace=False, delp=None):
    """q [6, nq, km, n+6, n+6], pe [6, n+2, km+1, n+2] -> (q_out, delp_out [6, km, n+6, n+6]); inplace=True remaps q itself"""
    s, ct = _sfx(q.dtype)
    ntiles, nq, km, nd, _ = q.shape
    n = nd - 6
    qo = q if inplace else np.array(q, copy=True, order="C")
    if delp is None:
        delp = np.zeros((ntiles, km, nd, nd), dtype=q.dtype)
    kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord_tr, dtype=np.int32), (nq,)))
    pe = np.ascontiguousarray(pe, dtype=q.dtype)
    ak = np.ascontiguousarray(ak, dtype=q.dtype)
    bk = np.ascontiguousarray(bk, dtype=q.dtype)
    getattr(lib(), f"orc_{s}_remap_tracers")(int(ntiles), int(n), int(km), int(nq), _p(pe), _p(ak), _p(bk), ct(ptop), _p(qo),
                                            _p(delp), _p(kord), int(bool(fill)))
    return qo, delp


def copy_corners(q2d, n, direction):
    s, _ = _sfx(q2d.dtype)
    out = np.array(q2d, copy=True, order="C")
    getattr(lib(), f"orc_{s}_copy_corners")(_p(out), int(n), int(direction))
    return out


def profile_col(which, a1, delp, iv, kord, qmin=0.0, qs=0.0):
    """which: 0 scalar_profile, 1 cs_profile, 2 ppm_profile.  a1 [km] cell means -> a4 [km, 4]"""
    s, ct = _sfx(a1.dtype)
    km = a1.shape[0]
    a4 = np.zeros((km, 4), dtype=a1.dtype)
    a4[:, 0] = a1
    delp = np.ascontiguousarray(delp, dtype=a1.dtype)
    getattr(lib(), f"orc_{s}_profile_col")(int(which), _p(a4), _p(delp), int(km), int(iv), int(kord), ct(qmin), ct(qs))
    return a4


def fillz_col(q, dp):
    """q [nq, km] -> filled copy"""
    s, _ = _sfx(q.dtype)
    nq, km = q.shape
    out = np.array(q, copy=True, order="C")
    dp = np.ascontiguousarray(dp, dtype=q.dtype)
    getattr(lib(), f"orc_{s}_fillz_col")(int(km), int(nq), _p(out), _p(dp))
    return out


def map_col(which, pe1, pe2, q, kord, q_min=0.0, fill=False):
    """which: 0 mapn_tracer, 1 map1_q2(+fillz).  q [nq, km]"""
    s, ct = _sfx(q.dtype)
    nq, km = q.shape
    out = np.array(q, copy=True, order="C")
    kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord, dtype=np.int32), (nq,)))
    pe1 = np.ascontiguousarray(pe1, dtype=q.dtype)
    pe2 = np.ascontiguousarray(pe2, dtype=q.dtype)
    getattr(lib(), f"orc_{s}_map_col")(int(which), int(km), int(nq), _p(pe1), _p(pe2), _p(out), _p(kord), ct(q_min),
                                      int(bool(fill)))
    return out


def map_field_col(use_cs, pe1, pe2, q, iv, kord, q_min=0.0, qs=0.0):
    """map_scalar (use_cs = False: scalar_profile) / map1_ppm (use_cs = True: cs_profile) for one column q [km], kn = km"""
    s, ct = _sfx(q.dtype)
    km = q.shape[0]
    out = np.array(q, copy=True, order="C")
    pe1 = np.ascontiguousarray(pe1, dtype=q.dtype)
    pe2 = np.ascontiguousarray(pe2, dtype=q.dtype)
    getattr(lib(), f"orc_{s}_map_field_col")(int(bool(use_cs)), int(km), _p(pe1), _p(pe2), _p(out), int(iv), int(kord), ct(q_min), ct(qs))
    return out


def tracer_2d_damp(case, hord, nord_tr, trdm, del6_u, del6_v, da_min, q_split=0, lim_fac=1.0):
    """tracer_2d with tracer damping (trdm2 > 1e-4: deln_flux on the first sub-step, tp_core.F90:1239-1387); del6_u
    [6, n+7, n+6], del6_v [6, n+6, n+7] and da_min are the fv_grid_type members of fv_arrays.F90:124,183."""
    s, ct = _sfx(case.dtype)
    g = case.metrics()
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    dst, src = halo_offsets(case.n)
    nsplt = C.c_int(0)
    ksplt = np.zeros(case.npz, dtype=np.int32)
    cmax = np.zeros(case.npz, dtype=case.dtype)
    d6u = np.ascontiguousarray(del6_u, dtype=case.dtype)
    d6v = np.ascontiguousarray(del6_v, dtype=case.dtype)
    getattr(lib(), f"orc_{s}_tracer_2d_damp")(
        6, case.n, case.npz, case.nq, _p(out["q"]), _p(out["dp1"]), _p(out["mfx"]), _p(out["mfy"]), _p(out["cx"]),
        _p(out["cy"]), _p(g["area"]), _p(g["rarea"]), _p(g["dx"]), _p(g["dy"]), _p(g["dxa"]), _p(g["dya"]), _p(g["sin_sg"]),
        _p(dst), _p(src), C.c_int64(dst.size), int(hord), int(q_split), ct(lim_fac), C.byref(nsplt), _p(ksplt), _p(cmax),
        _p(d6u), _p(d6v), ct(da_min), int(nord_tr), ct(trdm))
    out["nsplt"] = nsplt.value
    out["ksplt"] = ksplt
    return out


def fv_tp_2d(q, crx, cry, hord, xfx, yfx, ra_x, ra_y, area, dxa, dya, rarea=None, del6_u=None, del6_v=None, da_min=0.0,
             lim_fac=1.0, mfx=None, mfy=None, mass=None, nord=-1, damp_c=0.0):
    """fv_tp_2d (tp_core.F90:110-249) for ONE tile and one 2-D field: q [n+6, n+6] (ghosted; returned with the corner view the
    reference leaves in it), crx/xfx [n+6, n+1], cry/yfx [n+1, n+6], ra_x [n+6, n], ra_y [n, n+6]; mfx [n, n+1] / mfy [n+1, n] both
    or neither; mass [n+6, n+6] or None; nord < 0 = absent.  Returns (fx [n, n+1], fy [n+1, n], q)."""
    s, ct = _sfx(q.dtype)
    dt = q.dtype
    n = q.shape[0] - 6
    qq = np.array(q, copy=True, order="C")
    fx = np.zeros((n, n + 1), dtype=dt)
    fy = np.zeros((n + 1, n), dtype=dt)
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dt)
    keep = [c(a) for a in (crx, cry, xfx, yfx, ra_x, ra_y, area, dxa, dya, rarea, del6_u, del6_v, mfx, mfy, mass)]
    pp = [None if a is None else _p(a) for a in keep]
    getattr(lib(), f"orc_{s}_fv_tp_2d_full")(
        n, _p(qq), pp[0], pp[1], int(hord), _p(fx), _p(fy), pp[2], pp[3], pp[4], pp[5], pp[6], pp[7], pp[8], pp[9], pp[10],
        pp[11], ct(da_min), ct(lim_fac), pp[12], pp[13], pp[14], int(nord), ct(damp_c))
    return fx, fy, qq
