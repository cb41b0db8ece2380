"""Synthetic, self-consistent inputs for the tracer-transport path (host side, numpy).

What the reference would hand to ``tracer_2d`` / ``Lagrangian_to_Eulerian`` after ``dyn_core``
(fv_dynamics.F90:618-762), manufactured analytically so that sharp invariants hold:

* area fluxes come from a stream function evaluated at the (shared, bit-identical) cell-corner points
  of the halo-extended grid -- the recipe of the reference's own ``init_winds``
  (tools/test_cases.F90:364-396) -- plus an optional potential (divergent) part that is symmetric in the
  two cells sharing a face; both are single-valued on tile edges, so global tracer mass is conserved;
* Courant numbers are obtained by inverting ``xfx = cx*dxa*dy*sin_sg`` (fv_tracer2d.F90:392-405);
* mass fluxes are ``mfx = xfx * 0.5*(dp(i-1)+dp(i))`` (cf. sw_core.F90:941-955);
* tracer shapes: cosine bell (test_cases.F90:968-987), Gaussian, slotted cylinder, q == 1
  (free stream), sparse random-positive with exact zeros, signed smooth field, thin layers.

Shapes follow the package convention (reversed Fortran shape, leading tile axis):
  q   [6, nq, npz, n+6, n+6]     dp1 [6, npz, n+6, n+6]
  cx  [6, npz, n+6, n+1]         cy  [6, npz, n+1, n+6]
  mfx [6, npz, n,   n+1]         mfy [6, npz, n+1, n]
  pe  [6, n+2, npz+1, n+2]       (Fortran pe(is-1:ie+1, km+1, js-1:je+1))
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

from . import cubed_sphere as cs

NG = cs.NG


def hybrid_coordinate(npz: int, ptop: float = 100.0, p0: float = 1.0e5):
    """A smooth synthetic hybrid sigma-pressure coordinate ``pe(k) = ak(k) + bk(k)*ps`` with thin layers
    near the top and the surface (the reference's tables live in tools/fv_eta.F90/.h; only their
    qualitative shape matters here)."""
    s = np.linspace(0.0, 1.0, npz + 1)
    eta = 0.5 * (1.0 - np.cos(np.pi * s)) * 0.35 + 0.65 * s ** 2.2      # monotone 0..1
    eta = (eta - eta[0]) / (eta[-1] - eta[0])
    c = 0.25
    bk = (np.maximum(eta - c, 0.0) / (1.0 - c)) ** 1.6
    ak = ptop + (p0 - ptop) * eta - bk * p0
    ak[-1] = 0.0
    bk[-1] = 1.0
    ak[0] = ptop
    bk[0] = 0.0
    assert np.all(np.diff(ak + bk * 0.85e5) > 0) and np.all(np.diff(ak + bk * 1.05e5) > 0)
    return ak, bk, float(ptop)


@dataclass
class Case:
    n: int
    npz: int
    nq: int
    dtype: np.dtype
    grid: cs.Grid
    q: np.ndarray
    dp1: np.ndarray
    cx: np.ndarray
    cy: np.ndarray
    mfx: np.ndarray
    mfy: np.ndarray
    pe: np.ndarray
    ak: np.ndarray
    bk: np.ndarray
    ptop: float
    xfx_full: np.ndarray = None     # float64 area fluxes the Courant numbers were derived from
    yfx_full: np.ndarray = None

    def metrics(self):
        return self.grid.astype(self.dtype)


def _psi_solid_body(xyz, axis):
    return -(xyz * axis).sum(-1)


def area_fluxes(grid: cs.Grid, npz: int, courant: float, divergent: float = 0.15, seed: int = 0):
    """Per-level area fluxes on the halo-extended staggering the path needs:
    ``xfx[6, npz, n+6, n+1]`` on (is:ie+1, jsd:jed) and ``yfx[6, npz, n+1, n+6]`` on (isd:ied, js:je+1).
    Level k rotates about its own axis with its own speed so that per-level Courant maxima differ."""
    n = grid.n
    G = grid.corner_xyz             # [6, n+7, n+7, 3] corner points -2..n+4
    C = grid.center_xyz             # [6, n+6, n+6, 3]
    R = grid.radius
    o = NG - 1
    xfx = np.empty((6, npz, n + 6, n + 1))
    yfx = np.empty((6, npz, n + 1, n + 6))
    dxa_min = np.nanmin(grid.dxa[:, NG:NG + n, NG:NG + n])
    for k in range(npz):
        f = k / max(npz - 1, 1)
        alpha = 0.25 * np.pi + 0.9 * np.pi * f               # axis tilt sweeps with height
        beta = 2.0 * np.pi * f * 1.7
        axis = np.array([np.sin(alpha) * np.cos(beta), np.sin(alpha) * np.sin(beta), np.cos(alpha)])
        speed = 0.35 + 0.65 * (0.5 - 0.5 * np.cos(2.0 * np.pi * f)) if npz > 1 else 1.0
        # stream function scaled so that |flux| / cell_area ~ courant at the fastest point
        amp = courant * speed * dxa_min * R
        psi = amp * _psi_solid_body(G, axis)                  # [6, n+7, n+7]
        # west faces i = 1..n+1, all rows jsd..jed:  -(psi(i,j+1)-psi(i,j))
        xr = -(psi[:, 1:, :] - psi[:, :-1, :])                # [6, n+6, n+7] on corner columns -2..n+4
        yr = (psi[:, :, 1:] - psi[:, :, :-1])                 # [6, n+7, n+6] on corner rows   -2..n+4
        xk = xr[:, :, NG:NG + n + 1].copy()
        yk = yr[:, NG:NG + n + 1, :].copy()
        if divergent != 0.0:
            # potential (divergent) part: gradient of chi(p) = (p.a)^2 + 0.5 (p.b), integrated over each
            # face from its two end points only, hence single-valued wherever the stream-function part is
            axd = np.array([np.cos(1.3 + 2.1 * f), np.sin(1.3 + 2.1 * f) * 0.8, 0.6 * np.sin(0.7 + 3.0 * f)])
            ampd = divergent * courant * speed * dxa_min * R

            def face_flux(A, B):
                mid = cs._unit(A + B)
                pa = (mid * axd).sum(-1, keepdims=True)
                pb = (mid * axis).sum(-1, keepdims=True)
                V = 2.0 * pa * (axd - pa * mid) + 0.5 * (axis - pb * mid)
                nrm = np.cross(A, B)
                ln = np.sqrt((nrm * nrm).sum(-1))
                arc = np.arctan2(ln, (A * B).sum(-1))
                return (V * nrm).sum(-1) / ln * arc * ampd

            with np.errstate(invalid="ignore", divide="ignore"):
                # x faces: from corner (i,j) to (i,j+1); A x B points toward -i for a right-handed (i,j)
                gx = -face_flux(G[:, :-1, :], G[:, 1:, :])          # [6, n+6, n+7]
                gy = face_flux(G[:, :, :-1], G[:, :, 1:])           # [6, n+7, n+6]
            xk += gx[:, :, NG:NG + n + 1]
            yk += gy[:, NG:NG + n + 1, :]
        xfx[:, k] = np.nan_to_num(xk)
        yfx[:, k] = np.nan_to_num(yk)
    return xfx, yfx


def courant_from_flux(grid: cs.Grid, xfx: np.ndarray, yfx: np.ndarray):
    """Invert fv_tracer2d.F90:392-405: cx = xfx / (dxa(upwind) * dy * sin_sg(upwind, 3|1))."""
    n = grid.n
    sl = slice(NG, NG + n + 1)
    # x faces i=1..n+1 at rows jsd..jed: upwind-left cell i-1, right cell i
    dxa_l = grid.dxa[:, :, NG - 1:NG + n]
    dxa_r = grid.dxa[:, :, NG:NG + n + 1]
    s3_l = grid.sin_sg[:, 2, :, NG - 1:NG + n]
    s1_r = grid.sin_sg[:, 0, :, NG:NG + n + 1]
    dy = grid.dy[:, :, sl]
    den_pos = (dxa_l * dy * s3_l)[:, None]
    den_neg = (dxa_r * dy * s1_r)[:, None]
    cx = np.where(xfx > 0.0, xfx / den_pos, xfx / den_neg)
    dya_l = grid.dya[:, NG - 1:NG + n, :]
    dya_r = grid.dya[:, NG:NG + n + 1, :]
    s4_l = grid.sin_sg[:, 3, NG - 1:NG + n, :]
    s2_r = grid.sin_sg[:, 1, NG:NG + n + 1, :]
    dx = grid.dx[:, sl, :]
    den_pos = (dya_l * dx * s4_l)[:, None]
    den_neg = (dya_r * dx * s2_r)[:, None]
    cy = np.where(yfx > 0.0, yfx / den_pos, yfx / den_neg)
    return cx, cy


def surface_pressure(grid: cs.Grid, amp: float = 5.0e3):
    C = grid.center_xyz
    with np.errstate(invalid="ignore"):
        lon = np.arctan2(C[..., 1], C[..., 0])
        lat = np.arcsin(np.clip(C[..., 2], -1, 1))
    ps = 1.0e5 + amp * (np.cos(lat) ** 2 * np.sin(2.0 * lon) + 0.5 * np.sin(lat) * np.cos(3.0 * lon + 0.4))
    return np.nan_to_num(ps, nan=1.0e5)


def tracer_fields(grid: cs.Grid, npz: int, nq: int, seed: int = 20260101):
    """nq tracer initial conditions [6, nq, npz, n+6, n+6] (float64), compute domain + edge halos filled."""
    n = grid.n
    C = np.nan_to_num(grid.center_xyz)
    rng = np.random.default_rng(seed)
    kk = (np.arange(npz) + 0.5) / npz
    vert_smooth = 0.2 + np.exp(-((kk - 0.55) / 0.18) ** 2)
    vert_grad = 0.05 + kk ** 2
    vert_layers = np.where((np.arange(npz) // 3) % 2 == 0, 1.0, 0.0)
    vert_noise = 0.5 + 0.5 * (-1.0) ** np.arange(npz)

    def gc(center):
        c = np.asarray(center, dtype=float)
        c = c / np.linalg.norm(c)
        return np.arccos(np.clip((C * c).sum(-1), -1.0, 1.0))        # angular distance [6, n+6, n+6]

    r0 = 1.0 / 3.0 * 1.0   # a/3 on the unit sphere
    r = gc([1.0, 0.35, 0.2])
    bell = np.where(r < r0, 0.5 * (1.0 + np.cos(np.pi * r / r0)), 0.0)
    gauss = np.exp(-(gc([-0.3, 1.0, 0.5]) / 0.25) ** 2)
    rc = gc([0.57, 0.57, 0.6])
    lon_c = np.arctan2(C[..., 1], C[..., 0])
    slot = np.where((rc < 0.5) & ~((np.abs(lon_c - np.pi / 4) < 0.08) & (C[..., 2] < 0.75)), 1.0, 0.1)
    signed = C[..., 0] * C[..., 1] + 0.3 * C[..., 2]
    ridge = np.maximum(0.0, 1.0 - np.abs(C[..., 2] * 4.0 - 1.0))
    protos = [
        ("cosine_bell", bell[:, None] * vert_smooth[None, :, None, None] * 1.0e-2),
        ("gaussian", gauss[:, None] * vert_grad[None, :, None, None] * 1.0e-3),
        ("slotted_cyl", slot[:, None] * vert_layers[None, :, None, None] * 1.0e-4 + 1.0e-6),
        ("free_stream", np.ones((6, npz, n + 6, n + 6))),
        ("sparse_rand", None),
        ("signed", signed[:, None] * (vert_smooth - 0.6)[None, :, None, None]),
        ("ridge_noise", ridge[:, None] * vert_noise[None, :, None, None] * 1.0e-5),
        ("bell_thin", bell[:, None] * vert_layers[None, :, None, None] * 3.0e-3),
        ("gauss_const", gauss[:, None] * np.ones(npz)[None, :, None, None] * 2.0e-3 + 1.0e-5),
    ]
    q = np.zeros((6, nq, npz, n + 6, n + 6))
    for iq in range(nq):
        name, fld = protos[iq % len(protos)]
        if fld is None:
            fld = rng.random((6, npz, n + 6, n + 6)) * 1.0e-3
            fld[rng.random(fld.shape) < 0.10] = 0.0
        if iq >= len(protos):      # more than 9 tracers: perturb amplitude so fields stay distinct
            fld = fld * (1.0 + 0.05 * (iq // len(protos)))
        q[:, iq] = fld
    # corner blocks are never consumed (copy_corners overwrites them); poison them mildly
    q[..., :NG, :NG] = 0.0
    q[..., :NG, -NG:] = 0.0
    q[..., -NG:, :NG] = 0.0
    q[..., -NG:, -NG:] = 0.0
    cs.fill_edge_halos(q, n)
    return q


def make_case(n: int, npz: int, nq: int, dtype=np.float64, courant: float = 0.7, divergent: float = 0.15,
              seed: int = 20260101, grid: cs.Grid | None = None, lagrangian_perturb: float = 0.3) -> Case:
    """One consistent set of ``tracer_2d`` + tracer-remap inputs for a global C``n`` L``npz`` problem."""
    dtype = np.dtype(dtype)
    grid = grid or cs.make_grid(n)
    ak, bk, ptop = hybrid_coordinate(npz)
    ps = surface_pressure(grid)                                    # [6, n+6, n+6]
    pe_ref = ak[None, :, None, None] + bk[None, :, None, None] * ps[:, None]     # [6, npz+1, n+6, n+6]
    dp1 = np.diff(pe_ref, axis=1)                                  # Eulerian delp before the step
    xfx, yfx = area_fluxes(grid, npz, courant, divergent, seed)
    cx, cy = courant_from_flux(grid, xfx, yfx)
    # mass fluxes on the compute domain (single-valued: symmetric average of the two cells)
    c0, c1 = NG, NG + n
    dpl = dp1[:, :, c0:c1, c0 - 1:c1]        # cells i-1 for faces 1..n+1
    dpr = dp1[:, :, c0:c1, c0:c1 + 1]
    mfx = xfx[:, :, c0:c1, :] * (0.5 * (dpl + dpr))
    dpl = dp1[:, :, c0 - 1:c1, c0:c1]
    dpr = dp1[:, :, c0:c1 + 1, c0:c1]
    mfy = yfx[:, :, :, c0:c1] * (0.5 * (dpl + dpr))
    q = tracer_fields(grid, npz, nq, seed)

    # Lagrangian interface pressures after the step: delp advected with the same mass fluxes, plus a
    # smooth monotone perturbation (<= lagrangian_perturb of a layer) to make the remap non-trivial.
    rarea = grid.rarea[:, None, c0:c1, c0:c1]
    dp_lag = dp1[:, :, c0:c1, c0:c1] + (mfx[..., :-1] - mfx[..., 1:] + mfy[:, :, :-1, :] - mfy[:, :, 1:, :]) * rarea
    pe = lagrangian_pe(dp_lag, ptop, n, npz, lagrangian_perturb, grid)

    def cast(a):
        return np.ascontiguousarray(a, dtype=dtype)

    return Case(n=n, npz=npz, nq=nq, dtype=dtype, grid=grid, q=cast(q), dp1=cast(dp1), cx=cast(cx), cy=cast(cy),
                mfx=cast(mfx), mfy=cast(mfy), pe=cast(pe), ak=cast(ak), bk=cast(bk), ptop=ptop,
                xfx_full=xfx, yfx_full=yfx)


def lagrangian_pe(dp_lag: np.ndarray, ptop: float, n: int, npz: int, perturb: float, grid: cs.Grid) -> np.ndarray:
    """Assemble ``pe(is-1:ie+1, km+1, js-1:je+1)`` ([6, n+2, npz+1, n+2]) from the Lagrangian layer
    thicknesses on the compute domain; top and surface interfaces are pinned (fv_mapz.F90:269-272)."""
    pe_c = np.empty((6, npz + 1, n, n))
    pe_c[:, 0] = ptop
    pe_c[:, 1:] = ptop + np.cumsum(dp_lag, axis=1)
    if perturb != 0.0 and npz > 1:
        C = grid.center_xyz[:, NG:NG + n, NG:NG + n]
        phase = 3.0 * C[..., 0] + 2.0 * C[..., 1] - 4.0 * C[..., 2]
        kk = np.arange(1, npz)[None, :, None, None]
        wob = np.sin(phase[:, None] + 0.9 * kk) * np.sin(np.pi * kk / npz)
        thick = np.minimum(dp_lag[:, :-1], dp_lag[:, 1:])
        pe_c[:, 1:-1] += perturb * 0.5 * thick * wob          # |shift| <= perturb/2 of the thinner neighbour
    assert np.all(np.diff(pe_c, axis=1) > 0)
    pe = np.zeros((6, n + 2, npz + 1, n + 2))
    pe[:, 1:-1, :, 1:-1] = np.transpose(pe_c, (0, 2, 1, 3))
    # the (is-1, ie+1, js-1, je+1) rim is only read by the wind remap (out of scope); keep it benign
    pe[:, 0] = pe[:, 1]
    pe[:, -1] = pe[:, -2]
    pe[:, :, :, 0] = pe[:, :, :, 1]
    pe[:, :, :, -1] = pe[:, :, :, -2]
    return pe


def global_mass(q_c: np.ndarray, dp_c: np.ndarray, area_c: np.ndarray) -> np.ndarray:
    """Sum_k,j,i,tile q*dp*area per tracer in extended precision.  q_c [6,nq,npz,n,n], dp_c [6,npz,n,n]."""
    w = (dp_c.astype(np.longdouble) * area_c[:, None].astype(np.longdouble))[:, None]
    return (q_c.astype(np.longdouble) * w).sum(axis=(0, 2, 3, 4))
