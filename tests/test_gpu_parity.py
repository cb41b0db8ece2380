"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): max normalised difference <= 1e-12 per step in fp64, <= 1e-5 in the 32-bit
build.  The library is built FMA-free with IEEE division, the oracle likewise, so these tests additionally
assert the stronger property that the results are bit-identical wherever that is expected."""
import numpy as np
import pytest

from fv3atm_b200.tracer import TracerContext

pytestmark = pytest.mark.gpu
NG = 3


def norm_diff(a, b):
    """max |a-b| / max|b| per tracer (axis 1), compute domain only; arrays [6, nq, npz, n+6, n+6]"""
    sl = slice(NG, -NG)
    d = np.abs(a[..., sl, sl].astype(np.float64) - b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    s = np.abs(b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    return d / np.maximum(s, 1e-300)


def run_gpu_tracer_2d(case, hord, q_split=0, lim_fac=1.0):
    g = case.metrics()
    ctx = TracerContext(case.n + 1, case.npz, case.nq, g, dtype=case.dtype)
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    nsplt, ksplt = ctx.tracer_2d(out["q"], out["dp1"], out["mfx"], out["mfy"], out["cx"], out["cy"], hord, q_split=q_split,
                                 lim_fac=lim_fac)
    out["nsplt"], out["ksplt"] = nsplt, ksplt
    ctx.close()
    return out


TOL = {np.dtype("float64"): 1e-12, np.dtype("float32"): 1e-5}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 10, 9, 13, 12, 7, 11, -5, 5, 6, 1, 2, 3, 4])
def test_tracer_2d_parity_c24(oracle, case_factory, hord, dtype):
    case = case_factory(24, 16, 9, dtype)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_gpu_tracer_2d(case, hord)
    assert got["nsplt"] == ref["nsplt"]
    assert np.array_equal(got["ksplt"], ref["ksplt"])
    nd = norm_diff(got["q"], ref["q"])
    assert nd.max() <= TOL[case.dtype], f"hord={hord}: normalised diff {nd}"
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl]), f"hord={hord}: not bit-identical, nd={nd}"


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("courant", [1.8, 3.3])
def test_tracer_2d_subcycling(oracle, case_factory, courant, dtype):
    """nsplt > 1 with level-dependent ksplt(k): q, the advanced dp1 and the scaled cx/cy/mfx/mfy post-state."""
    case = case_factory(24, 16, 9, dtype, courant=courant)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    assert ref["nsplt"] >= 2 and len(set(ref["ksplt"].tolist())) > 1
    assert got["nsplt"] == ref["nsplt"] and np.array_equal(got["ksplt"], ref["ksplt"])
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl]), norm_diff(got["q"], ref["q"])
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("n,npz", [(48, 64), (40, 8)])
def test_tracer_2d_parity_config1(oracle, case_factory, n, npz):
    """BASELINE config 1 (C48 L64, 9 tracers, fp64, hord 8) and a ragged size (partial 32x16 blocks)."""
    case = case_factory(n, npz, 9, "float64")
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl]), norm_diff(got["q"], ref["q"])


@pytest.mark.parametrize("hord", [8, 10, -5, 7])
@pytest.mark.parametrize("nthreads", [32, 64, 96])
def test_tracer_2d_strip_decomposition(oracle, case_factory, monkeypatch, nthreads, hord):
    """The marching kernel splits a tile into strips of (threads - 6) columns: 1, 2 and 3 strips with ragged last
    strips must all reproduce the oracle bit-for-bit (C40: 32 -> 26+14, 64 -> 40, 96 -> 40)."""
    monkeypatch.setenv("FV3T_ADV_NT", str(nthreads))
    case = case_factory(40, 8, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_gpu_tracer_2d(case, hord)
    sl = slice(NG, -NG)
    assert got["nsplt"] == ref["nsplt"] >= 2
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl]), norm_diff(got["q"], ref["q"])
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kord", [9, 8, 10, 11, 12, 13, 14, 15, 16, 17])
def test_remap_parity(oracle, case_factory, kord, dtype):
    case = case_factory(24, 32, 9, dtype)
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    q = np.array(case.q, copy=True)
    delp = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, q, delp, kord, fill=True)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    nd = norm_diff(q, qref)
    assert nd.max() <= TOL[case.dtype], nd
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl]), nd


@pytest.mark.parametrize("kord", [9, 10, 7, 4, 6])
def test_remap_few_tracers_map1_q2(oracle, case_factory, kord):
    """nq <= 5 takes the map1_q2 (+ppm_profile for kord <= 7) branch of Lagrangian_to_Eulerian."""
    case = case_factory(24, 32, 5, "float64")
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    q = np.array(case.q, copy=True)
    delp = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, q, delp, kord, fill=True)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl]), norm_diff(q, qref)


def test_step_resident_matches_host_path(oracle, case_factory):
    """Device-resident advect + remap (the benchmarked path) against oracle advect + oracle remap."""
    case = case_factory(24, 16, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=8)
    kord = np.array([9] * 9, dtype=np.int32)
    qref, dref = oracle.remap_tracers(ref["q"], case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        ctx.upload(f, getattr(case, f), case.nq)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    nsplt = ctx.tracer_2d_resident(case.nq, 8)
    assert nsplt == ref["nsplt"] and nsplt >= 2
    ctx.remap_tracers_resident(case.nq, kord, fill=True)
    q = np.empty_like(case.q)
    delp = np.empty_like(case.dp1)
    ctx.download("q", q, case.nq)
    ctx.download("delp", delp, case.nq)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl]), norm_diff(q, qref)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
