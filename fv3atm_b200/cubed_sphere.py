"""Cubed-sphere topology and geometry needed by the tracer-transport path (host side, numpy).

This is the part of the reference's *geometry / domain* layer that the hot path consumes read-only:

* the 12-contact mosaic of the six tiles (reference: atmos_cubed_sphere/tools/fv_mp_mod.F90:581-629)
  turned into per-edge affine index maps (scalar, cell-centred halo fill, ``ng = 3``,
  fv_mp_mod.F90:104);
* the equidistant gnomonic grid (``gnomonic_ed``, model/fv_grid_utils.F90:1456-1550) and the tile
  rotations of ``mirror_grid`` (tools/fv_grid_tools.F90:3150-3274);
* the metric terms the path reads: ``dx, dy`` (fv_grid_tools.F90:905-950),
  ``dxa, dya`` (:975-990 incl. the ``fill_corners(..., AGRID)`` corner blocks, fv_mp_mod.F90:1522-1546),
  ``area, rarea`` (grid_area, :2932-3116), ``sin_sg(1:5)`` with the transport fix-ups at the tile
  corners (model/fv_grid_utils.F90:396-434, 653-716).

Array convention (everywhere in this package): numpy C-order arrays whose shape is the *reversed*
Fortran shape, i.e. a Fortran ``a(isd:ied, jsd:jed)`` is ``a[j - jsd, i - isd]`` so that ``i`` is
contiguous exactly as in the Fortran host.  A leading axis of length 6 stacks the tiles.

Not a port: the reference builds these fields with FMS halo updates over MPI; here the halo-extended
corner-point grid is assembled directly from the contact table and every metric is a vectorised
numpy expression.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

NG = 3  # halo width, fv_mp_mod.F90:104
W, E, S, N = 0, 1, 2, 3
EDGE_NAMES = ("W", "E", "S", "N")
TINY_NUMBER = 1.0e-8  # fv_grid_utils.F90 tiny_number used by fill_ghost on sin_sg


# ----------------------------------------------------------------------------------------------
# topology
# ----------------------------------------------------------------------------------------------
def _contacts(n: int):
    """The 12 contact lines of the 6-tile mosaic, 1-based tiles and cells (fv_mp_mod.F90:581-629).

    Each entry: (tile1, tile2, (istart1, iend1, jstart1, jend1), (istart2, iend2, jstart2, jend2)).
    """
    nx = ny = n
    return [
        (1, 2, (nx, nx, 1, ny), (1, 1, 1, ny)),      # 1: tile1 E  - tile2 W
        (1, 3, (1, nx, ny, ny), (1, 1, ny, 1)),      # 2: tile1 N  - tile3 W (reversed)
        (1, 5, (1, 1, 1, ny), (nx, 1, ny, ny)),      # 3: tile1 W  - tile5 N (reversed)
        (1, 6, (1, nx, 1, 1), (1, nx, ny, ny)),      # 4: tile1 S  - tile6 N
        (2, 3, (1, nx, ny, ny), (1, nx, 1, 1)),      # 5: tile2 N  - tile3 S
        (2, 4, (nx, nx, 1, ny), (nx, 1, 1, 1)),      # 6: tile2 E  - tile4 S (reversed)
        (2, 6, (1, nx, 1, 1), (nx, nx, ny, 1)),      # 7: tile2 S  - tile6 E (reversed)
        (3, 4, (nx, nx, 1, ny), (1, 1, 1, ny)),      # 8: tile3 E  - tile4 W
        (3, 5, (1, nx, ny, ny), (1, 1, ny, 1)),      # 9: tile3 N  - tile5 W (reversed)
        (4, 5, (1, nx, ny, ny), (1, nx, 1, 1)),      # 10: tile4 N - tile5 S
        (4, 6, (nx, nx, 1, ny), (nx, 1, 1, 1)),      # 11: tile4 E - tile6 S (reversed)
        (5, 6, (nx, nx, 1, ny), (1, 1, 1, ny)),      # 12: tile5 E - tile6 W
    ]


def _edge_of(line, n):
    """Which tile edge a contact line (is, ie, js, je) lies on, its start point, direction, outward normal."""
    is_, ie, js, je = line
    p0 = np.array([is_, js])
    d = np.array([np.sign(ie - is_), np.sign(je - js)])
    if is_ == ie:  # constant i -> W or E edge
        edge = E if is_ == n else W
        out = np.array([1, 0]) if edge == E else np.array([-1, 0])
    else:
        edge = N if js == n else S
        out = np.array([0, 1]) if edge == N else np.array([0, -1])
    return edge, p0, d, out


@dataclass(frozen=True)
class EdgeMap:
    """Halo cells beyond edge ``edge`` of tile ``tile`` are cells of ``nbr_tile``:
    ``(i', j') = A @ (i, j) + b`` (all 1-based cell indices)."""
    tile: int          # 0-based
    edge: int
    nbr_tile: int      # 0-based
    nbr_edge: int
    A: np.ndarray      # 2x2 signed permutation
    b: np.ndarray      # 2

    def map_cells(self, i, j):
        i = np.asarray(i)
        j = np.asarray(j)
        ip = self.A[0, 0] * i + self.A[0, 1] * j + self.b[0]
        jp = self.A[1, 0] * i + self.A[1, 1] * j + self.b[1]
        return ip, jp

    def map_corners(self, i, j):
        """Same map for corner (B-grid) points: point (i,j) is the SW corner of cell (i,j), i.e. sits
        at position (i-1/2, j-1/2) in cell coordinates."""
        i = np.asarray(i, dtype=np.float64) - 0.5
        j = np.asarray(j, dtype=np.float64) - 0.5
        ip = self.A[0, 0] * i + self.A[0, 1] * j + self.b[0] + 0.5
        jp = self.A[1, 0] * i + self.A[1, 1] * j + self.b[1] + 0.5
        return np.rint(ip).astype(np.int64), np.rint(jp).astype(np.int64)

    @property
    def rotated(self) -> bool:
        return bool(self.A[0, 0] == 0)


def edge_maps(n: int):
    """``maps[tile][edge] -> EdgeMap`` for all 24 directed tile edges."""
    maps = [[None] * 4 for _ in range(6)]
    for (t1, t2, l1, l2) in _contacts(n):
        e1, p1, d1, out1 = _edge_of(l1, n)
        e2, p2, d2, out2 = _edge_of(l2, n)
        for (ta, ea, pa, da, outa, tb, eb, pb, db, outb) in (
            (t1, e1, p1, d1, out1, t2, e2, p2, d2, out2),
            (t2, e2, p2, d2, out2, t1, e1, p1, d1, out1),
        ):
            inb = -outb
            # h = pa + s*da + m*outa (m>=1)  ->  pb + s*db + (m-1)*inb
            A = np.outer(db, da) + np.outer(inb, outa)
            b = pb - inb - A @ pa
            assert maps[ta - 1][ea] is None
            maps[ta - 1][ea] = EdgeMap(ta - 1, ea, tb - 1, eb, A.astype(np.int64), b.astype(np.int64))
    assert all(m is not None for row in maps for m in row)
    return maps


def halo_index_table(n: int, ng: int = NG):
    """Flat gather table for the scalar edge-halo fill of all six tiles.

    Returns int32 arrays ``(dst_tile, dst_j, dst_i, src_tile, src_j, src_i)`` with 0-based
    *array offsets* into halo-padded ``(n+2ng, n+2ng)`` tiles (offset = index - isd).  The four
    ``ng x ng`` corner blocks are not part of the exchange: the hot path rebuilds them with
    ``copy_corners`` (tp_core.F90:253-330)."""
    maps = edge_maps(n)
    out = [[] for _ in range(6)]
    rng = np.arange(1, n + 1)
    for t in range(6):
        for e in range(4):
            m = np.arange(1, ng + 1)
            if e == W:
                i, j = np.meshgrid(1 - m, rng, indexing="ij")
            elif e == E:
                i, j = np.meshgrid(n + m, rng, indexing="ij")
            elif e == S:
                i, j = np.meshgrid(rng, 1 - m, indexing="ij")
            else:
                i, j = np.meshgrid(rng, n + m, indexing="ij")
            i = i.ravel()
            j = j.ravel()
            ip, jp = maps[t][e].map_cells(i, j)
            assert ip.min() >= 1 and ip.max() <= n and jp.min() >= 1 and jp.max() <= n
            out[0].append(np.full(i.size, t))
            out[1].append(j + ng - 1)
            out[2].append(i + ng - 1)
            out[3].append(np.full(i.size, maps[t][e].nbr_tile))
            out[4].append(jp + ng - 1)
            out[5].append(ip + ng - 1)
    return tuple(np.concatenate(a).astype(np.int32) for a in out)


def fill_edge_halos(a: np.ndarray, n: int, ng: int = NG) -> np.ndarray:
    """In-place scalar halo fill of ``a[6, ..., n+2ng, n+2ng]`` (numpy reference of the exchange that
    ``mpp_update_domains`` / the group halo update performs for ``q``, fv_tracer2d.F90:499,561)."""
    dt, dj, di, st, sj, si = halo_index_table(n, ng)
    a[dt, ..., dj, di] = a[st, ..., sj, si]
    return a


def copy_corners_np(q: np.ndarray, n: int, direction: int, ng: int = NG) -> np.ndarray:
    """numpy restatement of copy_corners (tp_core.F90:253-330) for one halo-padded 2-D slab
    ``q[..., j-jsd, i-isd]``; used only to build inputs/tests."""
    npx = npy = n + 1
    o = ng - 1  # offset: index -> array position

    def g(i, j):
        return q[..., j + o, i + o]

    src = q.copy()

    def gs(i, j):
        return src[..., j + o, i + o]

    for j in range(1 - ng, 1):
        for i in range(1 - ng, 1):           # SW
            q[..., j + o, i + o] = gs(j, 1 - i) if direction == 1 else gs(1 - j, i)
        for i in range(npx, npx + ng):       # SE
            q[..., j + o, i + o] = gs(npy - j, i - npx + 1) if direction == 1 else gs(npy + j - 1, npx - i)
    for j in range(npy, npy + ng):
        for i in range(npx, npx + ng):       # NE
            q[..., j + o, i + o] = gs(j, 2 * npx - 1 - i) if direction == 1 else gs(2 * npy - 1 - j, i)
        for i in range(1 - ng, 1):           # NW
            q[..., j + o, i + o] = gs(npy - j, i - 1 + npx) if direction == 1 else gs(j + 1 - npx, npy - i)
    return q


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def _tile1_lonlat(n: int):
    """Corner points of tile 1: equidistant gnomonic projection (gnomonic_ed) centred on lon 0, lat 0
    after the -pi shift of gnomonic_grids (fv_grid_utils.F90:1421-1440)."""
    alpha = np.arcsin(1.0 / np.sqrt(3.0))
    th = -alpha + 2.0 * alpha * np.arange(n + 1) / n
    t = np.sqrt(2.0) * np.tan(th)
    t = 0.5 * (t - t[::-1])          # enforce the 4-fold symmetry that mirror_grid imposes
    ti, tj = np.meshgrid(t, t, indexing="xy")   # [j, i]
    lon = np.arctan2(ti, 1.0)
    lat = np.arctan2(tj, np.sqrt(1.0 + ti * ti))
    return lon, lat


def _rot(axis, x, y, z, ang_deg):
    a = np.deg2rad(ang_deg)
    c, s = np.cos(a), np.sin(a)
    # exact values for the multiples of 90 degrees used by mirror_grid
    c = np.round(c, 15)
    s = np.round(s, 15)
    if axis == 1:
        return x, c * y + s * z, -s * y + c * z
    if axis == 2:
        return c * x - s * z, y, s * x + c * z
    return c * x + s * y, -s * x + c * y, z


def tile_corner_xyz(n: int) -> np.ndarray:
    """Unit vectors (right-handed xyz) of the (n+1)^2 corner points of the six tiles,
    shape (6, n+1, n+1, 3) indexed [tile, j-1, i-1].  Follows mirror_grid's rotation recipe
    (fv_grid_tools.F90:3188-3274), including its left-handed z = -sin(lat) convention
    (spherical_to_cartesian, :2827-2838)."""
    lon, lat = _tile1_lonlat(n)
    x1 = np.cos(lon) * np.cos(lat)
    y1 = np.sin(lon) * np.cos(lat)
    z1 = -np.sin(lat)
    out = np.empty((6, n + 1, n + 1, 3))
    for t in range(6):
        x, y, z = x1, y1, z1
        if t == 1:
            x, y, z = _rot(3, x, y, z, -90.0)
        elif t == 2:
            x, y, z = _rot(3, x, y, z, -90.0)
            x, y, z = _rot(1, x, y, z, 90.0)
        elif t == 3:
            x, y, z = _rot(3, x, y, z, -180.0)
            x, y, z = _rot(1, x, y, z, 90.0)
        elif t == 4:
            x, y, z = _rot(3, x, y, z, 90.0)
            x, y, z = _rot(2, x, y, z, 90.0)
        elif t == 5:
            x, y, z = _rot(2, x, y, z, 90.0)
        out[t, ..., 0] = x
        out[t, ..., 1] = y
        out[t, ..., 2] = -z      # back to a right-handed frame
    out /= np.sqrt((out ** 2).sum(-1, keepdims=True))
    # Points on shared tile edges are made bit-identical by copying from the lower-numbered tile
    # (the reference does the same after mirror_grid, fv_grid_tools.F90:839-854).
    maps = edge_maps(n)
    rng = np.arange(1, n + 2)
    for t in range(6):
        for e in range(4):
            mp = maps[t][e]
            if mp.nbr_tile > t:
                continue
            if e == W:
                i, j = np.full(n + 1, 1), rng
            elif e == E:
                i, j = np.full(n + 1, n + 1), rng
            elif e == S:
                i, j = rng, np.full(n + 1, 1)
            else:
                i, j = rng, np.full(n + 1, n + 1)
            ip, jp = mp.map_corners(i, j)
            out[t, j - 1, i - 1] = out[mp.nbr_tile, jp - 1, ip - 1]
    return out


def _unit(v):
    return v / np.sqrt((v * v).sum(-1, keepdims=True))


def _gc_dist(p, q):
    """Great-circle distance between unit vectors (accurate for small separations)."""
    c = np.cross(p, q)
    return np.arctan2(np.sqrt((c * c).sum(-1)), (p * q).sum(-1))


def _mid(p, q):
    return _unit(p + q)


def _cos_angle(p1, p2, p3):
    """cos of the angle at p1 between the arcs p1->p2 and p1->p3 (fv_grid_utils.F90:3005-3049)."""
    P = np.cross(p1, p2)
    Q = np.cross(p1, p3)
    ddd = np.sqrt((P * P).sum(-1) * (Q * Q).sum(-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        ang = np.where(ddd > 0.0, (P * Q).sum(-1) / np.where(ddd > 0, ddd, 1.0), 1.0)
    return np.clip(ang, -1.0, 1.0)


def _tri_area(a, b, c):
    """Solid angle of a spherical triangle (Van Oosterom & Strackee; stable for tiny cells)."""
    num = np.abs((a * np.cross(b, c)).sum(-1))
    den = 1.0 + (a * b).sum(-1) + (b * c).sum(-1) + (c * a).sum(-1)
    return 2.0 * np.arctan2(num, den)


@dataclass
class Grid:
    """Metric fields of ``fv_grid_type`` that the tracer path reads (fv_arrays.F90:72-210), for all six
    tiles, float64, halo-padded; ``[tile, j-jsd, i-isd]``.  ``dx`` has one extra row (jed+1), ``dy`` one
    extra column (ied+1); ``sin_sg`` is ``[tile, 5, j, i]`` (sub-cell positions 1..5)."""
    n: int
    radius: float
    area: np.ndarray
    rarea: np.ndarray
    dx: np.ndarray
    dy: np.ndarray
    dxa: np.ndarray
    dya: np.ndarray
    sin_sg: np.ndarray
    da_min: float
    corner_xyz: np.ndarray = field(repr=False, default=None)   # (6, n+1+2ng, n+1+2ng, 3) halo-extended corners
    center_xyz: np.ndarray = field(repr=False, default=None)   # (6, n+2ng, n+2ng, 3)

    @property
    def npx(self):
        return self.n + 1

    def astype(self, dtype):
        """The ``real`` copies the model reads (4- or 8-byte build)."""
        return {k: np.ascontiguousarray(getattr(self, k), dtype=dtype)
                for k in ("area", "rarea", "dx", "dy", "dxa", "dya", "sin_sg")}


def damping_metrics(grid: "Grid"):
    """``del6_u (isd:ied, jsd:jed+1)``, ``del6_v (isd:ied+1, jsd:jed)`` and ``da_min`` of ``fv_grid_type`` (fv_arrays.F90:124,183),
    read by ``deln_flux`` only (tracer damping, off by default).  Form of fv_grid_utils.F90:799-819 -- ``del6_v = sin * dy / dxc``,
    ``del6_u = sin * dx / dyc`` with the face sine taken from the two adjacent sub-cell values -- with the C-grid distances
    ``dxc, dyc`` (which this synthetic grid does not carry) replaced by the mean of the two adjacent A-grid widths.  Inputs of the
    path like every other metric: the oracle and the library are handed the same arrays.  Returns ([6, nd+1, nd], [6, nd, nd+1], da_min)."""
    nd = grid.n + 2 * NG
    ss = np.nan_to_num(grid.sin_sg, nan=1.0)
    dxa, dya = np.nan_to_num(grid.dxa, nan=1.0), np.nan_to_num(grid.dya, nan=1.0)
    del6_v = np.zeros((6, nd, nd + 1))
    del6_u = np.zeros((6, nd + 1, nd))
    # x-faces i = isd..ied+1: cells i-1 (east value, position 3) and i (west value, position 1); one-sided at the array ends
    sin_w = np.concatenate([ss[:, 0], ss[:, 0, :, -1:]], axis=2)           # position 1 of cell i
    sin_e = np.concatenate([ss[:, 2, :, :1], ss[:, 2]], axis=2)            # position 3 of cell i-1
    dxc = 0.5 * (np.concatenate([dxa[:, :, :1], dxa], axis=2) + np.concatenate([dxa, dxa[:, :, -1:]], axis=2))
    del6_v[:] = 0.5 * (sin_w + sin_e) * np.nan_to_num(grid.dy, nan=1.0) / dxc
    sin_s = np.concatenate([ss[:, 1], ss[:, 1, -1:, :]], axis=1)           # position 2 of cell j
    sin_n = np.concatenate([ss[:, 3, :1, :], ss[:, 3]], axis=1)            # position 4 of cell j-1
    dyc = 0.5 * (np.concatenate([dya[:, :1, :], dya], axis=1) + np.concatenate([dya, dya[:, -1:, :]], axis=1))
    del6_u[:] = 0.5 * (sin_s + sin_n) * np.nan_to_num(grid.dx, nan=1.0) / dyc
    return del6_u, del6_v, float(grid.da_min)


def extended_corner_points(n: int, ng: int = NG) -> np.ndarray:
    """Corner points on ``(1-ng : n+1+ng)^2`` per tile: own points inside, the neighbouring tile's own
    points in the four edge halos (what the halo update of ``grid`` delivers, fv_grid_tools.F90:886-890).
    The four corner blocks are left NaN: nothing on the tracer path may depend on them except through
    the explicit ``fill_corners``/``sin_sg`` fix-ups reproduced in :func:`make_grid`."""
    P = tile_corner_xyz(n)
    maps = edge_maps(n)
    m = n + 1 + 2 * ng
    ext = np.full((6, m, m, 3), np.nan)
    o = ng - 1
    ext[:, ng:ng + n + 1, ng:ng + n + 1] = P
    rng = np.arange(1, n + 2)
    for t in range(6):
        for e in range(4):
            d = np.arange(1, ng + 1)
            if e == W:
                i, j = np.meshgrid(1 - d, rng, indexing="ij")
            elif e == E:
                i, j = np.meshgrid(n + 1 + d, rng, indexing="ij")
            elif e == S:
                i, j = np.meshgrid(rng, 1 - d, indexing="ij")
            else:
                i, j = np.meshgrid(rng, n + 1 + d, indexing="ij")
            i = i.ravel()
            j = j.ravel()
            ip, jp = maps[t][e].map_corners(i, j)
            assert ip.min() >= 1 and ip.max() <= n + 1 and jp.min() >= 1 and jp.max() <= n + 1
            ext[t, j + o, i + o] = P[maps[t][e].nbr_tile, jp - 1, ip - 1]
    return ext


def make_grid(n: int, radius: float = 6.3712e6, ng: int = NG) -> Grid:
    """Build the metric terms for a global C``n`` cubed sphere."""
    npx = npy = n + 1
    G = extended_corner_points(n, ng)          # [t, j, i, 3] on corner points 1-ng .. n+1+ng
    sw = G[:, :-1, :-1]
    se = G[:, :-1, 1:]
    nw = G[:, 1:, :-1]
    ne = G[:, 1:, 1:]
    with np.errstate(invalid="ignore"):
        dx = _gc_dist(G[:, :, :-1], G[:, :, 1:]) * radius       # (6, m+1 rows, m cols): dx(i,j) i:isd..ied, j:jsd..jed+1
        dy = _gc_dist(G[:, :-1, :], G[:, 1:, :]) * radius       # dy(i,j) i:isd..ied+1, j:jsd..jed
        pw = _mid(sw, nw)
        pe = _mid(se, ne)
        ps = _mid(sw, se)
        pn = _mid(nw, ne)
        dxa = _gc_dist(pw, pe) * radius
        dya = _gc_dist(ps, pn) * radius
        area = (_tri_area(sw, se, ne) + _tri_area(sw, ne, nw)) * radius ** 2
        ctr = _unit(sw + se + nw + ne)                           # cell_center2
        cos1 = _cos_angle(pw, ctr, nw)
        cos2 = _cos_angle(ps, se, ctr)
        cos3 = _cos_angle(pe, ctr, se)
        cos4 = _cos_angle(pn, nw, ctr)
        ec1 = _unit(pe - pw - ctr * ((pe - pw) * ctr).sum(-1, keepdims=True))
        ec2 = _unit(pn - ps - ctr * ((pn - ps) * ctr).sum(-1, keepdims=True))
        cos5 = (ec1 * ec2).sum(-1)
        cos_sg = np.stack([cos1, cos2, cos3, cos4, cos5], axis=1)
        sin_sg = np.minimum(1.0, np.sqrt(np.maximum(0.0, 1.0 - cos_sg ** 2)))

    o = ng - 1

    # --- corner blocks -------------------------------------------------------------------------
    # dxa/dya: fill_corners(dxa, dya, AGRID) with mySign = +1 (fv_mp_mod.F90:1522-1546)
    def ix(i):
        return i + o

    for t in range(6):
        x = dxa[t]
        y = dya[t]
        for j in range(1, ng + 1):
            for i in range(1, ng + 1):
                x[ix(1 - j), ix(1 - i)] = y[ix(i), ix(1 - j)]                          # SW
                x[ix(npy - 1 + j), ix(1 - i)] = y[ix(npy - 1 - i + 1), ix(1 - j)]      # NW
                x[ix(1 - j), ix(npx - 1 + i)] = y[ix(i), ix(npx - 1 + j)]              # SE
                x[ix(npy - 1 + j), ix(npx - 1 + i)] = y[ix(npy - 1 - i + 1), ix(npx - 1 + j)]  # NE
        for j in range(1, ng + 1):
            for i in range(1, ng + 1):
                y[ix(1 - i), ix(1 - j)] = x[ix(1 - j), ix(i)]                          # SW
                y[ix(npy - 1 + i), ix(1 - j)] = x[ix(npy - 1 + j), ix(i)]              # NW
                y[ix(1 - i), ix(npx - 1 + j)] = x[ix(1 - j), ix(npx - 1 - i + 1)]      # SE
                y[ix(npy - 1 + i), ix(npx - 1 + j)] = x[ix(npy - 1 + j), ix(npx - 1 - i + 1)]  # NE

    # sin_sg: fill_ghost(tiny_number) on the corner blocks, then the transport fix-ups
    # (fv_grid_utils.F90:653-716).  sin_sg[t, p-1, j, i].
    for t in range(6):
        s = sin_sg[t]
        for blk_j in (slice(0, ng), slice(ng + n, ng + n + ng)):
            for blk_i in (slice(0, ng), slice(ng + n, ng + n + ng)):
                s[:, blk_j, blk_i] = TINY_NUMBER
        for i in (0, -1, -2):                       # SW
            s[2, ix(i), ix(0)] = s[1, ix(1), ix(i)]           # sin_sg(0,i,3) = sin_sg(i,1,2)
            s[3, ix(0), ix(i)] = s[0, ix(i), ix(1)]           # sin_sg(i,0,4) = sin_sg(1,i,1)
        for i in range(npy, npy + 3):               # NW
            s[2, ix(i), ix(0)] = s[3, ix(npy - 1), ix(npy - i)]   # sin_sg(0,i,3) = sin_sg(npy-i,npy-1,4)
        for i in (0, -1, -2):
            s[1, ix(npy), ix(i)] = s[0, ix(npy - i), ix(1)]       # sin_sg(i,npy,2) = sin_sg(1,npy-i,1)
        for j in (0, -1, -2):                       # SE
            s[0, ix(j), ix(npx)] = s[1, ix(1), ix(npx - j)]       # sin_sg(npx,j,1) = sin_sg(npx-j,1,2)
        for i in range(npx, npx + 3):
            s[3, ix(0), ix(i)] = s[2, ix(npx - i), ix(npx - 1)]   # sin_sg(i,0,4) = sin_sg(npx-1,npx-i,3)
        for i in (0, 1, 2):                         # NE
            s[0, ix(npy + i), ix(npx)] = s[3, ix(npy - 1), ix(npx + i)]   # sin_sg(npx,npy+i,1) = sin_sg(npx+i,npy-1,4)
            s[1, ix(npy), ix(npx + i)] = s[2, ix(npy + i), ix(npx - 1)]   # sin_sg(npx+i,npy,2) = sin_sg(npx-1,npy+i,3)

    # area / dx / dy in the never-read corner blocks: benign positive filler instead of NaN
    def _fill_nan(a, val):
        a[np.isnan(a)] = val
        return a

    area = _fill_nan(area, float(np.nanmean(area)))
    dx = _fill_nan(dx, float(np.nanmean(dx)))
    dy = _fill_nan(dy, float(np.nanmean(dy)))
    dxa = _fill_nan(dxa, float(np.nanmean(dxa)))
    dya = _fill_nan(dya, float(np.nanmean(dya)))
    sin_sg = _fill_nan(sin_sg, TINY_NUMBER)
    rarea = 1.0 / area
    da_min = float(area[:, ng:ng + n, ng:ng + n].min())
    return Grid(n=n, radius=radius, area=area, rarea=rarea, dx=dx, dy=dy, dxa=dxa, dya=dya,
                sin_sg=sin_sg, da_min=da_min, corner_xyz=G, center_xyz=ctr)
