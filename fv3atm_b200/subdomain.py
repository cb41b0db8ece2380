"""Sub-tile decomposition of the cubed sphere: ``layout = L x L`` square sub-domains per tile, as a reference rank sees them when
the model runs with ``layout > 1,1`` (``fv_grid_bounds_type``: ``is..ie`` a proper sub-range of ``1..npx-1``, fv_arrays.F90:1178-1186;
``mpp_define_domains`` with the cubed-sphere mosaic, fv_mp_mod.F90:581-629).

What a sub-domain needs that a whole tile does not:
  * its halo comes from up to eight neighbours: the four sides AND the four diagonal neighbours where the corner of the sub-domain
    lies inside a tile or on a tile edge (``mpp_update_domains`` fills those); only at a TRUE cube corner does ``copy_corners``
    (tp_core.F90:253-330) rebuild the 3 x 3 block, and only those sub-domains carry the corner flags (fv_arrays.F90:181);
  * the tile-edge formulas of ``xppm`` / ``yppm`` apply on the sides that lie on a tile edge only.
This module slices whole-tile arrays into sub-domain arrays and builds the halo exchange as flat gather lists
``dst (sub-domain, plane offset) <- src (sub-domain, plane offset)``; the kernels get the edge / corner flags (``A5Sub`` in
csrc/fv3t_advect5.cuh).  Pure numpy: shared by the multi-GPU driver (partition.py), the CPU tests and the host-simulated kernels."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import cubed_sphere as cs

NG = cs.NG
SW, SE, NE, NW = 1, 2, 4, 8


@dataclass(frozen=True)
class Sub:
    tile: int   # 0-based
    bi: int     # block column / row inside the tile
    bj: int
    i0: int     # global cell index = local cell index + i0
    j0: int


class SubMosaic:
    def __init__(self, n: int, L: int):
        if n % L:
            raise ValueError(f"layout {L}x{L} does not divide n = {n}")
        self.n, self.L, self.m = n, L, n // L
        if self.m < 2 * NG + 2:
            raise ValueError("sub-domains narrower than 8 cells are not supported (the edge stencils of the two sides would overlap)")
        self.subs = [Sub(t, bi, bj, bi * self.m, bj * self.m) for t in range(6) for bj in range(L) for bi in range(L)]
        self._table = None

    def __len__(self):
        return len(self.subs)

    def index(self, tile: int, bi: int, bj: int) -> int:
        return (tile * self.L + bj) * self.L + bi

    # ---- kernel flags: (no_w, no_e, no_s, no_n, cmask) -------------------------------------------------------------------
    def flags(self, s: int):
        sb, L = self.subs[s], self.L
        w, e, so, no = sb.bi == 0, sb.bi == L - 1, sb.bj == 0, sb.bj == L - 1
        cmask = (SW if w and so else 0) | (SE if e and so else 0) | (NE if e and no else 0) | (NW if w and no else 0)
        return (int(not w), int(not e), int(not so), int(not no), cmask)

    # ---- slicing whole-tile arrays [6, ..., rows, cols] ---------------------------------------------------------------------
    def _cut(self, a, s, rows, cols, r_extra, c_extra):
        sb, m = self.subs[s], self.m
        return np.ascontiguousarray(a[sb.tile][..., sb.j0:sb.j0 + m + r_extra, sb.i0:sb.i0 + m + c_extra])

    def cells(self, a, s):      # (isd:ied, jsd:jed): q, dp1, area, rarea, dxa, dya, sin_sg [6, ..., nd, nd]
        return self._cut(a, s, None, None, 2 * NG, 2 * NG)

    def xface(self, a, s):      # (is:ie+1, jsd:jed): cx [6, ..., nd, n+1]
        return self._cut(a, s, None, None, 2 * NG, 1)

    def yface(self, a, s):      # (isd:ied, js:je+1): cy [6, ..., n+1, nd]
        return self._cut(a, s, None, None, 1, 2 * NG)

    def mfx(self, a, s):        # (is:ie+1, js:je) [6, ..., n, n+1]
        return self._cut(a, s, None, None, 0, 1)

    def mfy(self, a, s):        # (is:ie, js:je+1) [6, ..., n+1, n]
        return self._cut(a, s, None, None, 1, 0)

    def dx(self, a, s):         # (isd:ied, jsd:jed+1) [6, nd+1, nd]
        return self._cut(a, s, None, None, 2 * NG + 1, 2 * NG)

    def dy(self, a, s):         # (isd:ied+1, jsd:jed) [6, nd, nd+1]
        return self._cut(a, s, None, None, 2 * NG, 2 * NG + 1)

    def pe(self, a, s):         # (is-1:ie+1, km+1, js-1:je+1) stored [6, n+2 (j), km+1, n+2 (i)]
        sb, m = self.subs[s], self.m
        return np.ascontiguousarray(a[sb.tile][sb.j0:sb.j0 + m + 2, :, sb.i0:sb.i0 + m + 2])

    def metrics(self, g: dict, s: int) -> dict:
        """Sub-domain slices of the whole-tile metric dictionary of ``SyntheticCase.metrics()``."""
        out = {k: self.cells(g[k], s) for k in ("area", "rarea", "dxa", "dya", "sin_sg")}
        out["dx"] = self.dx(g["dx"], s)
        out["dy"] = self.dy(g["dy"], s)
        return out

    def put_interior(self, whole, s, local):
        """whole[tile, ..., interior of sub-domain s] <- local[..., interior]"""
        sb, m = self.subs[s], self.m
        whole[sb.tile][..., sb.j0 + NG:sb.j0 + NG + m, sb.i0 + NG:sb.i0 + NG + m] = local[..., NG:NG + m, NG:NG + m]

    # ---- halo exchange ------------------------------------------------------------------------------------------------------
    def halo_table(self):
        """Gather lists of the scalar halo update of every sub-domain: int64 arrays (dst_sub, dst_off, src_sub, src_off), offsets
        into the (m+6) x (m+6) planes.  Covers the side halos and the diagonal blocks that are not true cube corners."""
        if self._table is not None:
            return self._table
        n, m, L = self.n, self.m, self.L
        nd, md = n + 2 * NG, m + 2 * NG
        # source of every cell of the halo-extended whole tiles, as a flat index into the [6, nd, nd] stack (-1: true corner block)
        G = np.full((6, nd, nd), -1, dtype=np.int64)
        own = np.arange(6 * nd * nd, dtype=np.int64).reshape(6, nd, nd)
        G[:, NG:NG + n, NG:NG + n] = own[:, NG:NG + n, NG:NG + n]
        dt, dj, di, st, sj, si = cs.halo_index_table(n)
        G[dt, dj, di] = (st.astype(np.int64) * nd + sj) * nd + si
        ds, do, ss, so = [], [], [], []
        aj, ai = np.meshgrid(np.arange(md), np.arange(md), indexing="ij")
        halo = ~((aj >= NG) & (aj < NG + m) & (ai >= NG) & (ai < NG + m))
        for s, sb in enumerate(self.subs):
            g = G[sb.tile, sb.j0:sb.j0 + md, sb.i0:sb.i0 + md]
            sel = halo & (g >= 0)
            src = g[sel]
            t2, r2 = np.divmod(src, nd * nd)
            j2, i2 = np.divmod(r2, nd)          # array coordinates in the whole source tile: interior cells
            bj2, bi2 = (j2 - NG) // m, (i2 - NG) // m
            ds.append(np.full(src.size, s, dtype=np.int64))
            do.append((aj[sel] * md + ai[sel]).astype(np.int64))
            ss.append((t2 * L + bj2) * L + bi2)
            so.append((j2 - bj2 * m) * md + (i2 - bi2 * m))
        self._table = tuple(np.concatenate(a) for a in (ds, do, ss, so))
        return self._table

    def pair_lists(self):
        """The exchange grouped by (dst sub-domain, src sub-domain): {(d, s): (dst_off int32[], src_off int32[])}."""
        ds, do, ss, so = self.halo_table()
        out = {}
        key = ds * len(self) + ss
        order = np.argsort(key, kind="stable")
        ks, starts = np.unique(key[order], return_index=True)
        ends = list(starts[1:]) + [key.size]
        for k, a, b in zip(ks, starts, ends):
            idx = order[a:b]
            out[(int(k // len(self)), int(k % len(self)))] = (do[idx].astype(np.int32), so[idx].astype(np.int32))
        return out

    def fill_halos(self, stack: np.ndarray) -> np.ndarray:
        """numpy reference of the exchange: stack [nsub, ..., m+6, m+6] in place."""
        ds, do, ss, so = self.halo_table()
        md = self.m + 2 * NG
        flat = stack.reshape(stack.shape[0], -1, md * md)
        flat[ds, :, do] = flat[ss, :, so]
        return stack
