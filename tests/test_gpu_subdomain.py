"""Sub-tile contexts (fv3t_dims.sub_layout; SURVEY.md section 8 rows a18 and e): one global problem decomposed into L x L square
sub-domains per tile, each resident in a sub-tile context with its own edge / corner flags, halos (diagonal blocks included)
exchanged by gather lists.  The assembled result must equal the whole-tile run BIT FOR BIT -- the per-cell arithmetic does not
depend on the decomposition -- for a fast scheme (hord 8) and for exact-arithmetic ones (hord 10, -5), with sub-stepping, and
through the tracer remap."""
import numpy as np
import pytest

from fv3atm_b200.subdomain import SubMosaic, SubMosaicStep
from fv3atm_b200.tracer import TracerContext

pytestmark = pytest.mark.gpu
NG = 3


def whole_tile_run(case, hord, kord):
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "cx", "cy", "mfx", "mfy", "pe"):
        ctx.upload(f, getattr(case, f), nq=case.nq)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    nsplt = ctx.tracer_2d_resident(case.nq, hord)
    q1 = ctx.download("q", np.zeros_like(case.q), nq=case.nq)
    dp1 = ctx.download("dp1", np.zeros_like(case.dp1))
    ctx.remap_tracers_resident(case.nq, kord, True)
    q2 = ctx.download("q", np.zeros_like(case.q), nq=case.nq)
    delp = ctx.download("delp", np.zeros_like(case.dp1))
    ctx.close()
    return nsplt, q1, dp1, q2, delp


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 10, -5])
@pytest.mark.parametrize("L", [2, 3])
def test_submosaic_equals_whole_tiles(oracle, case_factory, L, hord, dtype):
    case = case_factory(24, 16, 9, dtype, courant=1.8)
    nsplt, q1, dp1, q2, delp = whole_tile_run(case, hord, 9)
    assert nsplt >= 2
    mo = SubMosaic(case.n, L)
    run = SubMosaicStep(mo, 0, 1, 0, case.npz, case.nq, case.dtype, case.metrics())
    assert len(run.ctxs) == (4 if L == 2 else 9)
    run.upload_case(case)
    assert run.tracer_2d(hord) == nsplt
    sl = slice(NG, -NG)
    g1 = run.download("q", np.zeros_like(case.q))
    gd = run.download("dp1", np.zeros_like(case.dp1))
    assert np.array_equal(g1[..., sl, sl], q1[..., sl, sl]), np.abs(g1[..., sl, sl] - q1[..., sl, sl]).max()
    assert np.array_equal(gd[..., sl, sl], dp1[..., sl, sl])
    run.remap(9)
    g2 = run.download("q", np.zeros_like(case.q))
    gp = run.download("delp", np.zeros_like(case.dp1))
    assert np.array_equal(g2[..., sl, sl], q2[..., sl, sl])
    assert np.array_equal(gp[..., sl, sl], delp[..., sl, sl])
    run.close()
    if hord != 8:   # exact-arithmetic schemes: also the oracle, bit for bit
        ref = oracle.tracer_2d(case, hord=hord)
        assert np.array_equal(g1[..., sl, sl], ref["q"][..., sl, sl])


def test_subtile_context_few_tracers_and_errors(case_factory):
    """Sub-tile contexts take any tracer count (whole-tile contexts with fewer than four tracers use k_advect4, which knows no
    sub-domains) and refuse the entries that need whole tiles."""
    from fv3atm_b200.lib import Fv3tError
    case = case_factory(24, 16, 9, "float64")
    nq = 2
    mo = SubMosaic(case.n, 2)
    run = SubMosaicStep(mo, 0, 1, 0, case.npz, nq, case.dtype, case.metrics())
    small = type("C", (), {})()
    for f in ("dp1", "cx", "cy", "mfx", "mfy", "pe", "ak", "bk", "ptop"):
        setattr(small, f, getattr(case, f))
    small.q = np.ascontiguousarray(case.q[:, :nq])
    run.upload_case(small)
    run.tracer_2d(10)
    got = run.download("q", np.zeros_like(small.q))
    ctx = TracerContext(case.n + 1, case.npz, nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "cx", "cy", "mfx", "mfy"):
        ctx.upload(f, getattr(small, f), nq=nq)
    ctx.tracer_2d_resident(nq, 10)
    want = ctx.download("q", np.zeros_like(small.q), nq=nq)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(got[..., sl, sl], want[..., sl, sl])
    with pytest.raises(Fv3tError, match="whole"):
        run.ctxs[0].tracer_2d_resident(nq, 8)
    with pytest.raises(Fv3tError, match="whole-tile"):
        run.ctxs[0].halo_pack(1, 0, 0, 0)
    run.close()
