"""fv_tp_2d as an operator of its own (tp_core.F90:110-249): fv3t_*_fv_tp_2d against the oracle's
fv_tp_2d on the same ghosted fields -- fluxes fx, fy and the post-state of q -- bit for bit, for every hord, both branches
(with mass fluxes = tracers / pt; without = delp, vorticity), with and without deln_flux damping (both of its forms)."""
import numpy as np
import pytest

from fv3atm_b200 import cubed_sphere as cs
from fv3atm_b200.tracer import TracerContext

pytestmark = pytest.mark.gpu
HORDS = [8, 10, 9, 13, 12, 7, 11, -5, 5, 6, 1, 2, 3, 4]


def ghosted_fields(case, nlev):
    """nlev 2-D fields per tile with their edge halos filled (tracer iq of level 0 -> field iq): [6, nlev, nd, nd]"""
    import oracle_binding as ob
    dst, src = ob.halo_offsets(case.n)
    nd = case.n + 6
    q = np.ascontiguousarray(case.q[:, :nlev, 0]).copy()           # [6, nlev, nd, nd]
    for l in range(nlev):
        stack = np.ascontiguousarray(q[:, l])                      # tile-major stack of planes
        flat = stack.reshape(-1)
        flat[dst] = flat[src]
        q[:, l] = flat.reshape(6, nd, nd)
    return q


def operator_inputs(case, nlev):
    """Courant numbers of level l and flux areas / area ratios in the shapes of fv_tp_2d's dummies (fv_tracer2d.F90:387-405, 449-462;
    the sin_sg factor is dropped: any consistent input serves a parity test)."""
    g = case.metrics()
    n, dt = case.n, case.dtype
    area = np.nan_to_num(g["area"], nan=1.0)
    crx = np.ascontiguousarray(case.cx[:, :nlev])                  # [6, nlev, nd, n+1]
    cry = np.ascontiguousarray(case.cy[:, :nlev])                  # [6, nlev, n+1, nd]
    aw, ae = area[:, None, :, 2:n + 3], area[:, None, :, 3:n + 4]  # cells i-1, i of x-face i = 1..n+1
    xfx = (crx * np.where(crx > 0, aw, ae)).astype(dt)
    as_, an = area[:, None, 2:n + 3, :], area[:, None, 3:n + 4, :]
    yfx = (cry * np.where(cry > 0, as_, an)).astype(dt)
    ra_x = (area[:, None, :, 3:n + 3] + xfx[..., :-1] - xfx[..., 1:]).astype(dt)   # [6, nlev, nd, n]
    ra_y = (area[:, None, 3:n + 3, :] + yfx[..., :-1, :] - yfx[..., 1:, :]).astype(dt)  # [6, nlev, n, nd]
    return crx, cry, xfx, yfx, np.ascontiguousarray(ra_x), np.ascontiguousarray(ra_y)


def oracle_planes(q, crx, cry, hord, xfx, yfx, ra_x, ra_y, g, d6u=None, d6v=None, da_min=0.0, mfx=None, mfy=None, mass=None,
                  nord=-1, damp_c=0.0, lim_fac=1.0):
    import oracle_binding as ob
    nt, nlev = q.shape[:2]
    n = q.shape[-1] - 6
    fx = np.zeros((nt, nlev, n, n + 1), q.dtype)
    fy = np.zeros((nt, nlev, n + 1, n), q.dtype)
    qo = np.zeros_like(q)
    pick = lambda a, t, l: None if a is None else a[t, l]
    for t in range(nt):
        for l in range(nlev):
            fx[t, l], fy[t, l], qo[t, l] = ob.fv_tp_2d(
                q[t, l], crx[t, l], cry[t, l], hord, xfx[t, l], yfx[t, l], ra_x[t, l], ra_y[t, l], g["area"][t], g["dxa"][t],
                g["dya"][t], rarea=g["rarea"][t], del6_u=None if d6u is None else d6u[t], del6_v=None if d6v is None else d6v[t],
                da_min=da_min, lim_fac=lim_fac, mfx=pick(mfx, t, l), mfy=pick(mfy, t, l), mass=pick(mass, t, l), nord=nord,
                damp_c=damp_c)
    return fx, fy, qo


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", HORDS)
@pytest.mark.parametrize("tracer", [False, True])
def test_fv_tp_2d_fluxes(case_factory, hord, dtype, tracer):
    case = case_factory(24, 16, 9, dtype)
    nlev = 3
    g = case.metrics()
    q = ghosted_fields(case, nlev)
    crx, cry, xfx, yfx, ra_x, ra_y = operator_inputs(case, nlev)
    mfx = np.ascontiguousarray(case.mfx[:, :nlev]) if tracer else None
    mfy = np.ascontiguousarray(case.mfy[:, :nlev]) if tracer else None
    rfx, rfy, rq = oracle_planes(q, crx, cry, hord, xfx, yfx, ra_x, ra_y, g, mfx=mfx, mfy=mfy, lim_fac=0.9)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, g, dtype=case.dtype)
    qg = q.copy()
    fx, fy = ctx.fv_tp_2d(qg, crx, cry, hord, xfx, yfx, ra_x, ra_y, lim_fac=0.9, mfx=mfx, mfy=mfy)
    ctx.close()
    assert np.abs(rfx).max() > 0 and np.abs(rfy).max() > 0
    assert np.array_equal(fx, rfx), f"fx: {np.abs(fx - rfx).max() / np.abs(rfx).max()}"
    assert np.array_equal(fy, rfy), f"fy: {np.abs(fy - rfy).max() / np.abs(rfy).max()}"
    assert np.array_equal(qg, rq)                       # the corner blocks carry the dir = 1 view, everything else untouched
    assert not np.array_equal(qg, q)


@pytest.mark.parametrize("nord", [0, 1, 2])
@pytest.mark.parametrize("tracer", [False, True])
def test_fv_tp_2d_damping(case_factory, nord, tracer):
    """deln_flux through fv_tp_2d: the mass-weighted form of the tracer branch (tp_core.F90:229-234) and the plain form of the
    delp / vorticity branch (:243-248); without `mass` the tracer branch must NOT damp."""
    case = case_factory(24, 16, 9, "float64")
    nlev = 2
    g = case.metrics()
    d6u, d6v, da_min = cs.damping_metrics(case.grid)
    q = ghosted_fields(case, nlev)
    crx, cry, xfx, yfx, ra_x, ra_y = operator_inputs(case, nlev)
    mfx = np.ascontiguousarray(case.mfx[:, :nlev]) if tracer else None
    mfy = np.ascontiguousarray(case.mfy[:, :nlev]) if tracer else None
    import oracle_binding as ob
    dst, src = ob.halo_offsets(case.n)
    mass = np.ascontiguousarray(case.dp1[:, :nlev]).copy()
    for l in range(nlev):
        flat = np.ascontiguousarray(mass[:, l]).reshape(-1)
        flat[dst] = flat[src]
        mass[:, l] = flat.reshape(mass[:, l].shape)
    damp_c = 0.12
    kw = dict(mfx=mfx, mfy=mfy, nord=nord, damp_c=damp_c)
    rfx, rfy, _ = oracle_planes(q, crx, cry, 8, xfx, yfx, ra_x, ra_y, g, d6u, d6v, da_min, mass=mass if tracer else None, **kw)
    pfx, pfy, _ = oracle_planes(q, crx, cry, 8, xfx, yfx, ra_x, ra_y, g, mfx=mfx, mfy=mfy)
    assert np.abs(rfx - pfx).max() > 1e-6 * np.abs(pfx).max()    # the damping does something
    ctx = TracerContext(case.n + 1, case.npz, case.nq, g, dtype=case.dtype)
    ctx.set_damping(d6u, d6v, da_min)
    fx, fy = ctx.fv_tp_2d(q.copy(), crx, cry, 8, xfx, yfx, ra_x, ra_y, mass=mass if tracer else None, **kw)
    assert np.array_equal(fx, rfx), f"fx: {np.abs(fx - rfx).max() / np.abs(rfx).max()}"
    assert np.array_equal(fy, rfy), f"fy: {np.abs(fy - rfy).max() / np.abs(rfy).max()}"
    if tracer:   # mass absent -> no damping in the tracer branch
        fx, fy = ctx.fv_tp_2d(q.copy(), crx, cry, 8, xfx, yfx, ra_x, ra_y, **kw)
        assert np.array_equal(fx, pfx) and np.array_equal(fy, pfy)
    ctx.close()


def test_fv_tp_2d_flux_form_update_matches_tracer_2d(oracle, case_factory):
    """The fluxes of the operator entry, put through the flux-form update of fv_tracer2d.F90:533-541, reproduce what tracer_2d itself
    does with the same level (one sub-step)."""
    case = case_factory(24, 16, 9, "float64")
    ref = oracle.tracer_2d(case, hord=8)
    assert ref["nsplt"] == 1
    g = case.metrics()
    n, nd = case.n, case.n + 6
    import oracle_binding as ob
    dst, src = ob.halo_offsets(n)
    k, iq = 5, 3
    stack = np.ascontiguousarray(case.q[:, iq, k]).reshape(-1).copy()
    stack[dst] = stack[src]
    q = stack.reshape(6, 1, nd, nd)
    ctx = TracerContext(n + 1, case.npz, case.nq, g, dtype=case.dtype)
    area = np.nan_to_num(g["area"], nan=1.0)
    # the operands tracer_2d itself builds for this level (fv_tracer2d.F90:387-405, 449-462)
    crx, cry = case.cx[:, k:k + 1], case.cy[:, k:k + 1]
    ss = np.nan_to_num(g["sin_sg"], nan=1.0)[:, None]         # [6, 1, 5, nd, nd]
    dxa, dya = np.nan_to_num(g["dxa"], nan=1.0)[:, None], np.nan_to_num(g["dya"], nan=1.0)[:, None]
    dx, dy = np.nan_to_num(g["dx"], nan=1.0)[:, None], np.nan_to_num(g["dy"], nan=1.0)[:, None]
    dyf = dy[..., 3:n + 4]
    xfx = np.where(crx > 0, crx * dxa[..., 2:n + 3] * dyf * ss[:, :, 2, :, 2:n + 3], crx * dxa[..., 3:n + 4] * dyf * ss[:, :, 0, :, 3:n + 4])
    dxf = dx[..., 3:n + 4, :]
    yfx = np.where(cry > 0, cry * dya[..., 2:n + 3, :] * dxf * ss[:, :, 3, 2:n + 3, :], cry * dya[..., 3:n + 4, :] * dxf * ss[:, :, 1, 3:n + 4, :])
    ra_x = area[:, None, :, 3:n + 3] + xfx[..., :-1] - xfx[..., 1:]
    ra_y = area[:, None, 3:n + 3, :] + yfx[..., :-1, :] - yfx[..., 1:, :]
    mfx, mfy = case.mfx[:, k:k + 1], case.mfy[:, k:k + 1]
    fx, fy = ctx.fv_tp_2d(q.copy(), crx, cry, 8, xfx, yfx, ra_x, ra_y, mfx=mfx, mfy=mfy)
    ctx.close()
    sl = slice(3, -3)
    dp1 = case.dp1[:, k, sl, sl]
    rarea = g["rarea"][:, sl, sl]
    dp2 = dp1 + (mfx[:, 0, :, :-1] - mfx[:, 0, :, 1:] + mfy[:, 0, :-1, :] - mfy[:, 0, 1:, :]) * rarea
    qn = (q[:, 0, sl, sl] * dp1 + (fx[:, 0, :, :-1] - fx[:, 0, :, 1:] + fy[:, 0, :-1, :] - fy[:, 0, 1:, :]) * rarea) / dp2
    assert np.array_equal(qn, ref["q"][:, iq, k, sl, sl])
