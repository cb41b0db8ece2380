"""Host-side mirror of the reference interface for the tracer-transport path, on top of the C-ABI.

Names and argument meaning follow the reference (atmos_cubed_sphere/model/fv_tracer2d.F90:324 tracer_2d,
model/fv_mapz.F90:134 Lagrangian_to_Eulerian (tracer part), :1386 mapn_tracer); arrays use the package
convention: numpy C-order with the reversed Fortran shape and a leading tile axis, i.e. exactly the bytes
of the tile-major stack of Fortran arrays the C-ABI takes.  All compute happens in libfv3tracer.so on the
GPU; nothing here falls back to the CPU."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import lib as L


class TracerContext:
    """Device-resident state of the path for `ntiles` cubed-sphere tiles on one GPU (fv3t_ctx)."""

    def __init__(self, npx: int, npz: int, nq_max: int, grid: dict, dtype=np.float64, tiles=(1, 2, 3, 4, 5, 6),
                 device: int = 0, stream: int | None = None, sub_layout: int = 0, sub_blocks=None):
        """sub_layout = L >= 2: a sub-tile context (fv3t_dims.sub_layout): `tiles` names the tile of every resident sub-domain,
        sub_blocks its (bi, bj) block in the L x L decomposition, npx is the LOCAL extent + 1 and `grid` holds the local metric
        arrays, one leading entry per resident sub-domain."""
        self.dtype = np.dtype(dtype)
        self.npx, self.npz, self.nq_max = int(npx), int(npz), int(nq_max)
        self.n = self.npx - 1
        self.tiles = tuple(int(t) for t in tiles)
        self.nt = len(self.tiles)
        self.sub_layout = int(sub_layout)
        pad = lambda v: (C.c_int * 6)(*(list(v) + [0] * (6 - len(v))))
        blocks = list(sub_blocks) if sub_blocks is not None else []
        if self.sub_layout >= 2 and len(blocks) != self.nt:
            raise ValueError("sub_blocks: one (bi, bj) per resident sub-domain")
        d = L.Dims(self.npx, self.npz, self.nq_max, self.nt, pad(self.tiles), self.sub_layout, pad([b[0] for b in blocks]),
                   pad([b[1] for b in blocks]))
        g = L.GridPtrs()
        self._keep = []
        for k in ("area", "rarea", "dx", "dy", "dxa", "dya", "sin_sg"):
            a = np.ascontiguousarray(grid[k], dtype=self.dtype)
            if a.shape[0] != self.nt and self.sub_layout < 2:
                a = np.ascontiguousarray(a[[t - 1 for t in self.tiles]])
            self._keep.append(a)
            setattr(g, k, a.ctypes.data)
        self._h = C.c_void_p()
        L.check(L.fn(self.dtype, "create")(C.byref(self._h), C.byref(d), C.byref(g), int(device),
                                           C.c_void_p(stream) if stream else None))
        self._ct = L.prec(self.dtype)[1]

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            L.load().fv3t_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _f(self, name):
        return L.fn(self.dtype, name)

    # ---- reference-facing host-array entry points -------------------------------------------------
    def tracer_2d(self, q, dp1, mfx, mfy, cx, cy, hord, q_split=0, nord_tr=0, trdm=0.0, lim_fac=1.0):
        """tracer_2d(q, dp1, mfx, mfy, cx, cy, ..., nq, hord, q_split, ..., nord_tr, trdm, lim_fac): in-place on the
        numpy arrays like the Fortran INTENT(INOUT) dummies.  Returns (nsplt, ksplt)."""
        nq = q.shape[1]
        for a in (q, dp1, mfx, mfy, cx, cy):
            assert a.dtype == self.dtype and a.flags["C_CONTIGUOUS"]
        nsplt = C.c_int(0)
        ksplt = np.zeros(self.npz, dtype=np.int32)
        L.check(self._f("tracer_2d")(self._h, L.ptr(q), L.ptr(dp1), L.ptr(mfx), L.ptr(mfy), L.ptr(cx), L.ptr(cy), int(nq),
                                     int(hord), int(q_split), int(nord_tr), self._ct(trdm), self._ct(lim_fac),
                                     C.byref(nsplt), L.ptr(ksplt)))
        return nsplt.value, ksplt

    def set_damping(self, del6_u, del6_v, da_min, nord_tr=0, trdm=0.0):
        """fv_grid_type%del6_u / del6_v / da_min for deln_flux (tracer damping) and the settings of the resident entries."""
        d6u = np.ascontiguousarray(del6_u, dtype=self.dtype) if del6_u is not None else None
        d6v = np.ascontiguousarray(del6_v, dtype=self.dtype) if del6_v is not None else None
        L.check(self._f("set_damping")(self._h, L.ptr(d6u) if d6u is not None else None, L.ptr(d6v) if d6v is not None else None,
                                       self._ct(da_min), int(nord_tr), self._ct(trdm)))

    def tracer_2d_1L(self, q, dp1, mfx, mfy, cx, cy, hord, nord_tr=0, trdm=0.0, lim_fac=1.0):
        """tracer_2d_1L (fv_tracer2d.F90:92-321, the z_tracer variant): same q / cx / cy / mfx / mfy post-state as tracer_2d, dp1
        advanced only between a level's own sub-steps.  Returns (max over k of the per-level sub-step count, the counts)."""
        nq = q.shape[1]
        for a in (q, dp1, mfx, mfy, cx, cy):
            assert a.dtype == self.dtype and a.flags["C_CONTIGUOUS"]
        nsplt = C.c_int(0)
        ksplt = np.zeros(self.npz, dtype=np.int32)
        L.check(self._f("tracer_2d_1L")(self._h, L.ptr(q), L.ptr(dp1), L.ptr(mfx), L.ptr(mfy), L.ptr(cx), L.ptr(cy), int(nq),
                                        int(hord), 0, int(nord_tr), self._ct(trdm), self._ct(lim_fac), C.byref(nsplt), L.ptr(ksplt)))
        return nsplt.value, ksplt

    def remap_tracers(self, pe, ak, bk, ptop, q, delp, kord_tr, fill=True):
        """Tracer part of Lagrangian_to_Eulerian for every row: q and delp updated in place."""
        nq = q.shape[1]
        kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord_tr, dtype=np.int32), (nq,)))
        ak = np.ascontiguousarray(ak, dtype=self.dtype)
        bk = np.ascontiguousarray(bk, dtype=self.dtype)
        L.check(self._f("remap_tracers")(self._h, L.ptr(pe), L.ptr(ak), L.ptr(bk), self._ct(ptop), L.ptr(q), L.ptr(delp),
                                         int(nq), L.ptr(kord), int(bool(fill))))

    def tracer_step(self, q, dp1, mfx, mfy, cx, cy, pe, ak, bk, ptop, delp, hord, kord_tr, q_split=0, lim_fac=1.0, fill=True, nq=None):
        """tracer_2d + tracer remap on host arrays as one call, pipelined per tracer (fv3t_*_tracer_step).  Arguments are numpy
        arrays or integer addresses of (page-locked) host buffers; in place like the two separate calls.  Returns nsplt."""
        nq = int(nq) if nq is not None else (q.shape[1] if hasattr(q, "shape") else int(self.nq_max))
        kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord_tr, dtype=np.int32), (nq,)))
        ak = np.ascontiguousarray(ak, dtype=self.dtype)
        bk = np.ascontiguousarray(bk, dtype=self.dtype)
        P = lambda a: C.c_void_p(a) if isinstance(a, int) else L.ptr(a)
        nsplt = C.c_int(0)
        L.check(self._f("tracer_step")(self._h, P(q), P(dp1), P(mfx), P(mfy), P(cx), P(cy), P(pe), L.ptr(ak), L.ptr(bk),
                                       self._ct(ptop), P(delp), int(nq), int(hord), int(q_split), self._ct(lim_fac), L.ptr(kord),
                                       int(bool(fill)), C.byref(nsplt)))
        return nsplt.value

    def mapn_tracer(self, nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill):
        """Row-granular entry with the reference's own argument list (one-tile context)."""
        kord = np.ascontiguousarray(kord, dtype=np.int32)
        L.check(self._f("mapn_tracer")(self._h, int(nq), int(km), L.ptr(pe1), L.ptr(pe2), L.ptr(q1), L.ptr(dp2), L.ptr(kord),
                                       int(j), int(i1), int(i2), int(isd), int(ied), int(jsd), int(jed), self._ct(q_min),
                                       int(bool(fill))))

    def map_field(self, q, iv, kord, q_min=0.0, qs=None, use_cs=False):
        """map_scalar (use_cs = False: scalar_profile) / map1_ppm (use_cs = True: cs_profile) of ONE scalar field
        q [nt, npz, n+6, n+6] for every row, in place, from the resident pe onto ak + bk*ps (fv3t_*_map_field)."""
        assert q.dtype == self.dtype and q.flags["C_CONTIGUOUS"]
        qsp = None
        if qs is not None:
            qs = np.ascontiguousarray(qs, dtype=self.dtype)
            qsp = L.ptr(qs)
        L.check(self._f("map_field")(self._h, L.ptr(q), qsp, int(iv), int(kord), self._ct(q_min), int(bool(use_cs))))

    def fv_tp_2d(self, q, crx, cry, hord, xfx, yfx, ra_x, ra_y, lim_fac=1.0, mfx=None, mfy=None, mass=None, nord=-1,
                 damp_c=0.0):
        """fv_tp_2d (tp_core.F90:110-249) for nlev stacked 2-D fields per resident tile: q [nt, nlev, n+6, n+6] (in place: the
        corner halos come back with the dir = 1 corner view), crx/xfx [nt, nlev, n+6, n+1], cry/yfx [nt, nlev, n+1, n+6],
        ra_x [nt, nlev, n+6, n], ra_y [nt, nlev, n, n+6]; mfx [nt, nlev, n, n+1] and mfy [nt, nlev, n+1, n] both or neither;
        mass like q or None; nord < 0 = absent.  Returns (fx [nt, nlev, n, n+1], fy [nt, nlev, n+1, n])."""
        assert q.dtype == self.dtype and q.flags["C_CONTIGUOUS"] and q.ndim == 4
        nt, nlev, n = q.shape[0], q.shape[1], self.n
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=self.dtype)
        arrs = [c(a) for a in (crx, cry, xfx, yfx, ra_x, ra_y, mfx, mfy, mass)]
        pp = [None if a is None else L.ptr(a) for a in arrs]
        fx = np.zeros((nt, nlev, n, n + 1), dtype=self.dtype)
        fy = np.zeros((nt, nlev, n + 1, n), dtype=self.dtype)
        L.check(self._f("fv_tp_2d")(self._h, int(nlev), L.ptr(q), pp[0], pp[1], int(hord), L.ptr(fx), L.ptr(fy), pp[2], pp[3],
                                     pp[4], pp[5], self._ct(lim_fac), pp[6], pp[7], pp[8], int(nord), self._ct(damp_c)))
        return fx, fy

    # ---- device-resident operation -------------------------------------------------------------------
    def upload(self, field: str, host: np.ndarray, nq: int | None = None):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        L.check(self._f("upload")(self._h, L.FIELD[field], L.ptr(host), int(nq if nq is not None else self.nq_max)))
        self.sync()

    def upload_ptr(self, field: str, host_ptr: int, nq: int):
        """Asynchronous upload from (pinned) host memory given by address."""
        L.check(self._f("upload")(self._h, L.FIELD[field], C.c_void_p(host_ptr), int(nq)))

    def download(self, field: str, host: np.ndarray, nq: int | None = None):
        assert host.dtype == self.dtype and host.flags["C_CONTIGUOUS"]
        L.check(self._f("download")(self._h, L.FIELD[field], L.ptr(host), int(nq if nq is not None else self.nq_max)))
        return host

    def download_ptr(self, field: str, host_ptr: int, nq: int):
        L.check(self._f("download")(self._h, L.FIELD[field], C.c_void_p(host_ptr), int(nq)))

    def set_vertical(self, ak, bk, ptop):
        ak = np.ascontiguousarray(ak, dtype=self.dtype)
        bk = np.ascontiguousarray(bk, dtype=self.dtype)
        L.check(self._f("set_vertical")(self._h, L.ptr(ak), L.ptr(bk), self._ct(ptop)))

    def tracer_2d_resident(self, nq, hord, q_split=0, lim_fac=1.0) -> int:
        nsplt = C.c_int(0)
        L.check(self._f("tracer_2d_resident")(self._h, int(nq), int(hord), int(q_split), self._ct(lim_fac), C.byref(nsplt)))
        return nsplt.value

    def remap_prepare(self):
        """Hint: the resident pe is final for this step -> remap coefficients are computed concurrently with tracer_2d."""
        L.check(self._f("remap_prepare")(self._h))

    def remap_tracers_resident(self, nq, kord_tr, fill=True):
        kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord_tr, dtype=np.int32), (nq,)))
        L.check(self._f("remap_tracers_resident")(self._h, int(nq), L.ptr(kord), int(bool(fill))))

    # ---- building blocks for face-sharded contexts ---------------------------------------------------
    def tracer_2d_begin(self, nq, q_split=0):
        cmax = np.zeros(self.npz, dtype=self.dtype)
        L.check(self._f("tracer_2d_begin")(self._h, int(nq), int(q_split), L.ptr(cmax)))
        return cmax

    def tracer_2d_set_cmax(self, cmax, q_split=0) -> int:
        cmax = np.ascontiguousarray(cmax, dtype=self.dtype)
        nsplt = C.c_int(0)
        L.check(self._f("tracer_2d_set_cmax")(self._h, L.ptr(cmax), int(q_split), C.byref(nsplt)))
        return nsplt.value

    def halo_local(self, it):
        L.check(self._f("halo_local")(self._h, int(it)))

    def halo_pack(self, it, local_tile, edge, dev_ptr):
        L.check(self._f("halo_pack")(self._h, int(it), int(local_tile), int(edge), C.c_void_p(dev_ptr)))

    def halo_unpack(self, it, local_tile, edge, dev_ptr):
        L.check(self._f("halo_unpack")(self._h, int(it), int(local_tile), int(edge), C.c_void_p(dev_ptr)))

    # generic exchange by gather list (sub-tile contexts)
    def halo_list_create(self, offsets) -> int:
        offs = np.ascontiguousarray(offsets, dtype=np.int32)
        lid = C.c_int(-1)
        L.check(L.load().fv3t_halo_list_create(self._h, L.ptr(offs), int(offs.size), C.byref(lid)))
        return lid.value

    def halo_local_table(self, dst, src):
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        src = np.ascontiguousarray(src, dtype=np.int32)
        assert dst.size == src.size
        L.check(L.load().fv3t_halo_local_table(self._h, L.ptr(dst), L.ptr(src), int(dst.size)))

    def halo_gather(self, it, local_tile, list_id, dev_ptr, stride=0):
        L.check(self._f("halo_gather")(self._h, int(it), int(local_tile), int(list_id), C.c_void_p(dev_ptr), int(stride)))

    def halo_scatter(self, it, local_tile, list_id, dev_ptr, stride=0):
        L.check(self._f("halo_scatter")(self._h, int(it), int(local_tile), int(list_id), C.c_void_p(dev_ptr), int(stride)))

    def tracer_2d_substep(self, it, hord, lim_fac=1.0):
        L.check(self._f("tracer_2d_substep")(self._h, int(it), int(hord), self._ct(lim_fac)))

    def tracer_2d_finish(self):
        L.check(self._f("tracer_2d_finish")(self._h))

    # ---- misc
    def sync(self):
        L.check(L.load().fv3t_sync(self._h))

    def device_ptr(self, field: str) -> int:
        return int(L.load().fv3t_device_ptr(self._h, L.FIELD[field]) or 0)

    def halo_strip_elems(self) -> int:
        return int(L.load().fv3t_halo_strip_elems(self._h))

    def neighbor(self, global_tile: int, edge: int):
        t, e, r = C.c_int(), C.c_int(), C.c_int()
        L.check(L.load().fv3t_neighbor(self._h, int(global_tile), int(edge), C.byref(t), C.byref(e), C.byref(r)))
        return t.value, e.value, bool(r.value)

    def kernel_launches(self) -> int:
        return int(L.load().fv3t_kernel_launches(self._h))

    def timer_start(self):
        L.check(L.load().fv3t_timer_start(self._h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        L.check(L.load().fv3t_timer_stop_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def profile_enable(self, on=True):
        L.check(L.load().fv3t_profile_enable(self._h, int(bool(on))))

    def profile_get(self, kclass: str):
        ms, nl = C.c_float(), C.c_int()
        L.check(L.load().fv3t_profile_get_ms(self._h, L.KCLASS[kclass], C.byref(ms), C.byref(nl)))
        return float(ms.value), int(nl.value)
