"""Device-side generator of the synthetic workload for large configurations (bench only).

Same recipe as fv3atm_b200/synthetic.py (stream-function + potential area fluxes evaluated at the shared
corner points of the halo-extended grid, inverted to Courant numbers; symmetric mass fluxes; Lagrangian pe
from the advected delp; tracer shapes = horizontal pattern x vertical profile) but evaluated with torch on the
GPU directly into the TracerContext's device mirrors, because a C768 L127 x 9 tracer input set is ~55 GB.
Only input *generation* lives here; nothing of the transport path."""
from __future__ import annotations

import numpy as np
import torch

from . import cubed_sphere as cs
from . import synthetic as sy
from .devarray import field_view

NG = cs.NG


def hybrid(npz: int):
    """(ak, bk, ptop) of the synthetic hybrid coordinate fill_context uses."""
    return sy.hybrid_coordinate(npz)


def fill_context(ctx, grid: cs.Grid, nq: int, courant: float = 0.7, divergent: float = 0.15, seed: int = 20260101,
                 lagrangian_perturb: float = 0.3, device: int = 0, q_first: int = 0):
    """Fill q, dp1, cx, cy, mfx, mfy, pe of `ctx` (its resident tiles; tracers q_first .. q_first+nq-1 of the global
    tracer list) and set ak/bk/ptop.  Returns (ak, bk, ptop)."""
    n, npz = ctx.n, ctx.npz
    dev = torch.device(f"cuda:{device}")
    f64 = torch.float64
    tdt = torch.float64 if ctx.dtype == np.float64 else torch.float32

    tl = [t - 1 for t in ctx.tiles]
    nt = len(tl)

    def T(a):
        a = np.asarray(a)
        if a.ndim >= 3 and a.shape[0] == 6:
            a = a[tl]
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    G = torch.nan_to_num(T(grid.corner_xyz))            # [6, n+7, n+7, 3]
    C = torch.nan_to_num(T(grid.center_xyz))            # [6, n+6, n+6, 3]
    dxa, dya, dx, dy = T(grid.dxa), T(grid.dya), T(grid.dx), T(grid.dy)
    ssg, rarea = T(grid.sin_sg), T(grid.rarea)
    R = grid.radius
    dxa_min = float(np.nanmin(grid.dxa[:, NG:NG + n, NG:NG + n]))
    ak, bk, ptop = sy.hybrid_coordinate(npz)
    ps = T(sy.surface_pressure(grid))                   # [6, n+6, n+6]
    akt, bkt = T(ak), T(bk)

    sl = slice(NG, NG + n + 1)
    den_xp = dxa[:, :, NG - 1:NG + n] * dy[:, :, sl] * ssg[:, 2, :, NG - 1:NG + n]
    den_xn = dxa[:, :, NG:NG + n + 1] * dy[:, :, sl] * ssg[:, 0, :, NG:NG + n + 1]
    den_yp = dya[:, NG - 1:NG + n, :] * dx[:, sl, :] * ssg[:, 3, NG - 1:NG + n, :]
    den_yn = dya[:, NG:NG + n + 1, :] * dx[:, sl, :] * ssg[:, 1, NG:NG + n + 1, :]

    v_cx, v_cy = field_view(ctx, "cx", device=device), field_view(ctx, "cy", device=device)
    v_mfx, v_mfy = field_view(ctx, "mfx", device=device), field_view(ctx, "mfy", device=device)
    v_dp1, v_pe = field_view(ctx, "dp1", device=device), field_view(ctx, "pe", device=device)
    v_q = field_view(ctx, "q", nq, device=device)
    c0, c1 = NG, NG + n

    def unit(v):
        return v / torch.sqrt((v * v).sum(-1, keepdim=True))

    def face_flux(A, B, axd, axis, ampd):
        mid = unit(A + B)
        pa = (mid * axd).sum(-1, keepdim=True)
        pb = (mid * axis).sum(-1, keepdim=True)
        V = 2.0 * pa * (axd - pa * mid) + 0.5 * (axis - pb * mid)
        nrm = torch.linalg.cross(A, B)
        ln = torch.sqrt((nrm * nrm).sum(-1))
        arc = torch.atan2(ln, (A * B).sum(-1))
        return torch.nan_to_num((V * nrm).sum(-1) / ln * arc * ampd)

    pe_run = torch.full((nt, n, n), float(ptop), dtype=f64, device=dev)
    v_pe[:, 1:-1, 0, 1:-1] = pe_run.to(tdt)
    Cc = C[:, c0:c1, c0:c1]
    phase = 3.0 * Cc[..., 0] + 2.0 * Cc[..., 1] - 4.0 * Cc[..., 2]
    dp_prev = None
    for k in range(npz):
        f = k / max(npz - 1, 1)
        alpha = 0.25 * np.pi + 0.9 * np.pi * f
        beta = 2.0 * np.pi * f * 1.7
        axis = torch.tensor([np.sin(alpha) * np.cos(beta), np.sin(alpha) * np.sin(beta), np.cos(alpha)], dtype=f64, device=dev)
        speed = 0.35 + 0.65 * (0.5 - 0.5 * np.cos(2.0 * np.pi * f)) if npz > 1 else 1.0
        amp = courant * speed * dxa_min * R
        psi = -amp * (G * axis).sum(-1)
        xk = -(psi[:, 1:, :] - psi[:, :-1, :])[:, :, NG:NG + n + 1]
        yk = (psi[:, :, 1:] - psi[:, :, :-1])[:, NG:NG + n + 1, :]
        if divergent != 0.0:
            axd = torch.tensor([np.cos(1.3 + 2.1 * f), np.sin(1.3 + 2.1 * f) * 0.8, 0.6 * np.sin(0.7 + 3.0 * f)], dtype=f64, device=dev)
            ampd = divergent * courant * speed * dxa_min * R
            xk = xk - face_flux(G[:, :-1, :], G[:, 1:, :], axd, axis, ampd)[:, :, NG:NG + n + 1]
            yk = yk + face_flux(G[:, :, :-1], G[:, :, 1:], axd, axis, ampd)[:, NG:NG + n + 1, :]
        v_cx[:, k] = torch.where(xk > 0, xk / den_xp, xk / den_xn).to(tdt)
        v_cy[:, k] = torch.where(yk > 0, yk / den_yp, yk / den_yn).to(tdt)
        dpk = (akt[k + 1] - akt[k]) + (bkt[k + 1] - bkt[k]) * ps            # [6, n+6, n+6]
        v_dp1[:, k] = dpk.to(tdt)
        mfx = xk[:, c0:c1, :] * (0.5 * (dpk[:, c0:c1, c0 - 1:c1] + dpk[:, c0:c1, c0:c1 + 1]))
        mfy = yk[:, :, c0:c1] * (0.5 * (dpk[:, c0 - 1:c1, c0:c1] + dpk[:, c0:c1 + 1, c0:c1]))
        v_mfx[:, k] = mfx.to(tdt)
        v_mfy[:, k] = mfy.to(tdt)
        dp_lag = dpk[:, c0:c1, c0:c1] + (mfx[..., :-1] - mfx[..., 1:] + mfy[:, :-1, :] - mfy[:, 1:, :]) * rarea[:, c0:c1, c0:c1]
        # interface k+1 (bottom of layer k): running sum, perturbed except at the surface
        pe_run = pe_run + dp_lag
        if k < npz - 1 and lagrangian_perturb != 0.0:
            dpn = (akt[k + 2] - akt[k + 1]) + (bkt[k + 2] - bkt[k + 1]) * ps[:, c0:c1, c0:c1]
            thick = torch.minimum(dp_lag, dpn) * 0.8
            wob = torch.sin(phase + 0.9 * (k + 1)) * np.sin(np.pi * (k + 1) / npz)
            v_pe[:, 1:-1, k + 1, 1:-1] = (pe_run + lagrangian_perturb * 0.5 * thick * wob).to(tdt)
        else:
            v_pe[:, 1:-1, k + 1, 1:-1] = pe_run.to(tdt)
    v_pe[:, 0] = v_pe[:, 1]
    v_pe[:, -1] = v_pe[:, -2]
    v_pe[:, :, :, 0] = v_pe[:, :, :, 1]
    v_pe[:, :, :, -1] = v_pe[:, :, :, -2]

    # tracers: horizontal pattern x vertical profile (synthetic.tracer_fields recipe)
    kk = (torch.arange(npz, device=dev, dtype=f64) + 0.5) / npz
    ar = torch.arange(npz, device=dev)
    vert = {
        "smooth": 0.2 + torch.exp(-((kk - 0.55) / 0.18) ** 2),
        "grad": 0.05 + kk ** 2,
        "layers": torch.where((ar // 3) % 2 == 0, 1.0, 0.0).to(f64),
        "noise": 0.5 + 0.5 * (-1.0) ** ar.to(f64),
        "one": torch.ones(npz, device=dev, dtype=f64),
    }

    def gc(center):
        c = torch.tensor(center, dtype=f64, device=dev)
        c = c / torch.linalg.norm(c)
        return torch.arccos(torch.clamp((C * c).sum(-1), -1.0, 1.0))

    r0 = 1.0 / 3.0
    r = gc([1.0, 0.35, 0.2])
    bell = torch.where(r < r0, 0.5 * (1.0 + torch.cos(np.pi * r / r0)), torch.zeros_like(r))
    gauss = torch.exp(-(gc([-0.3, 1.0, 0.5]) / 0.25) ** 2)
    rc = gc([0.57, 0.57, 0.6])
    lon_c = torch.atan2(C[..., 1], C[..., 0])
    slot = torch.where((rc < 0.5) & ~((torch.abs(lon_c - np.pi / 4) < 0.08) & (C[..., 2] < 0.75)), 1.0, 0.1).to(f64)
    signed = C[..., 0] * C[..., 1] + 0.3 * C[..., 2]
    ridge = torch.clamp(1.0 - torch.abs(C[..., 2] * 4.0 - 1.0), min=0.0)
    protos = [(bell, vert["smooth"], 1.0e-2, 0.0), (gauss, vert["grad"], 1.0e-3, 0.0), (slot, vert["layers"], 1.0e-4, 1.0e-6),
              (None, None, 1.0, 0.0), ("rand", None, 1.0e-3, 0.0), (signed, vert["smooth"] - 0.6, 1.0, 0.0),
              (ridge, vert["noise"], 1.0e-5, 0.0), (bell, vert["layers"], 3.0e-3, 0.0), (gauss, vert["one"], 2.0e-3, 1.0e-5)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    for iq in range(nq):
        H, V, a, b = protos[(iq + q_first) % len(protos)]
        scale = 1.0 + 0.05 * ((iq + q_first) // len(protos))
        for t in range(nt):
            if H is None:
                v_q[t, iq] = 1.0
            elif isinstance(H, str):
                rnd = torch.rand((npz, n + 6, n + 6), generator=gen, device=dev, dtype=f64) * a
                rnd[torch.rand((npz, n + 6, n + 6), generator=gen, device=dev) < 0.10] = 0.0
                v_q[t, iq] = (rnd * scale).to(tdt)
            else:
                v_q[t, iq] = ((H[t][None] * V[:, None, None] * a + b) * scale).to(tdt)
    torch.cuda.synchronize(dev)
    ctx.set_vertical(ak, bk, ptop)
    return ak, bk, ptop
