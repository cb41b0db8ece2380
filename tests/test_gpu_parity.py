"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): max normalised difference <= 1e-12 per step in fp64, <= 1e-5 in the 32-bit
build.  Two kernel sets are checked (DESIGN.md "Strict and fast kernels"):
  strict (FV3T_STRICT=1): FMA-free, IEEE division, the reference's operation order -> additionally BIT-IDENTICAL
                          to the (FMA-free) oracle;
  fast   (the default):   FMA contraction, shared reciprocals -> within the bar (observed ~1e-15 in fp64).  Schemes
                          whose limiter is discontinuous on exact zeros (hord 1, 3-7, -5) always run the strict kernels."""
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
FAST_HORD = {8, 11, 2}   # fv3t::fast_hord_ok (every other scheme runs an exact-arithmetic kernel: bit-identical to the oracle)


@pytest.fixture(params=["strict", "fast"])
def mode(request, monkeypatch):
    """Kernel set used by the contexts a test creates (read by fv3t_*_create)."""
    monkeypatch.setenv("FV3T_STRICT", "1" if request.param == "strict" else "0")
    return request.param


def check_q(got, ref, mode, dtype, bit_exact_expected=True, what=""):
    nd = norm_diff(got, ref)
    assert nd.max() <= TOL[np.dtype(dtype)], f"{what}: normalised diff {nd}"
    if mode == "strict" or bit_exact_expected:
        sl = slice(NG, -NG)
        assert np.array_equal(got[..., sl, sl], ref[..., sl, sl]), f"{what}: not bit-identical, nd={nd}"


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 10, 9, 13, 12, 7, 11, -5, 5, 6, 1, 2, 3, 4])
def test_tracer_2d_parity_c24(oracle, case_factory, hord, dtype, mode):
    case = case_factory(24, 16, 9, dtype)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_gpu_tracer_2d(case, hord)
    assert got["nsplt"] == ref["nsplt"]
    assert np.array_equal(got["ksplt"], ref["ksplt"])
    check_q(got["q"], ref["q"], mode, dtype, bit_exact_expected=hord not in FAST_HORD, what=f"hord={hord}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("courant", [1.8, 3.3])
def test_tracer_2d_subcycling(oracle, case_factory, courant, dtype, mode):
    """nsplt > 1 with level-dependent ksplt(k): q, the advanced dp1 and the scaled cx/cy/mfx/mfy post-state."""
    case = case_factory(24, 16, 9, dtype, courant=courant)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    assert ref["nsplt"] >= 2 and len(set(ref["ksplt"].tolist())) > 1
    assert got["nsplt"] == ref["nsplt"] and np.array_equal(got["ksplt"], ref["ksplt"])
    sl = slice(NG, -NG)
    check_q(got["q"], ref["q"], mode, dtype, bit_exact_expected=False, what=f"courant={courant}")
    # the caller-visible post-state of dp1, cx, cy, mfx, mfy is bit-identical in both modes
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("n,npz", [(48, 64), (40, 8)])
def test_tracer_2d_parity_config1(oracle, case_factory, n, npz, mode):
    """BASELINE config 1 (C48 L64, 9 tracers, fp64, hord 8) and a ragged size (partial strips)."""
    case = case_factory(n, npz, 9, "float64")
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=False, what=f"C{n} L{npz}")


@pytest.mark.parametrize("hord", [8, 10, -5, 7])
@pytest.mark.parametrize("nthreads", [32, 64, 96])
def test_tracer_2d_strip_decomposition(oracle, case_factory, monkeypatch, nthreads, hord, mode):
    """The marching kernels split a tile into strips of (threads - 6) columns: 1, 2 and 3 strips with ragged last
    strips must all reproduce the oracle (C40: 32 -> 26+14, 64 -> 40, 96 -> 40)."""
    monkeypatch.setenv("FV3T_ADV_NT", str(nthreads))
    case = case_factory(40, 8, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_gpu_tracer_2d(case, hord)
    sl = slice(NG, -NG)
    assert got["nsplt"] == ref["nsplt"] >= 2
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=hord not in FAST_HORD, what=f"hord={hord} NT={nthreads}")
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])


FAST_KORD = {8, 9, 12, 13, 14, 17}   # fv3t::fast_kord_ok


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kord", [9, 8, 10, 11, 12, 13, 14, 15, 16, 17])
def test_remap_parity(oracle, case_factory, kord, dtype, mode):
    case = case_factory(24, 32, 9, dtype)
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    q = np.array(case.q, copy=True)
    delp = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, q, delp, kord, fill=True)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    check_q(q, qref, mode, dtype, bit_exact_expected=kord not in FAST_KORD, what=f"kord={kord}")


def test_remap_mixed_kord_uses_strict_kernel(oracle, case_factory, mode):
    """Per-tracer kord (cld_amt is forced to 9, fv_dynamics.F90:741-762): tracer sets with different limiters always run
    the strict kernel, bit-identical in both modes."""
    case = case_factory(24, 32, 9, "float64")
    kord = np.array([9, 10, 11, 12, 13, 14, 15, 16, 8], dtype=np.int32)
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    q = np.array(case.q, copy=True)
    delp = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, q, delp, kord, fill=True)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl])
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])


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


def test_step_resident_matches_host_path(oracle, case_factory, mode):
    """Device-resident advect + remap (the benchmarked path).  strict: bit-identical to oracle advect + oracle remap.
    fast: the advected q is within the bar, and the remap -- whose kord-9 limiter switches on the sign of cell-mean
    differences, so that O(1e-16) differences in its INPUT can flip it around the slotted cylinder -- is compared with
    the oracle remap of the SAME advected field."""
    case = case_factory(24, 16, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=8)
    kord = np.array([9] * 9, dtype=np.int32)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        ctx.upload(f, getattr(case, f), case.nq)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    nsplt = ctx.tracer_2d_resident(case.nq, 8)
    assert nsplt == ref["nsplt"] and nsplt >= 2
    qadv = np.empty_like(case.q)
    ctx.download("q", qadv, case.nq)
    check_q(qadv, ref["q"], mode, "float64", bit_exact_expected=False, what="advect")
    ctx.remap_tracers_resident(case.nq, kord, fill=True)
    q = np.empty_like(case.q)
    delp = np.empty_like(case.dp1)
    ctx.download("q", q, case.nq)
    ctx.download("delp", delp, case.nq)
    ctx.close()
    qref, dref = oracle.remap_tracers(qadv, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    sl = slice(NG, -NG)
    check_q(q, qref, mode, "float64", bit_exact_expected=False, what="remap of the advected field")
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])


def test_multi_step_mass_conservation_and_positivity(case_factory, mode):
    """BASELINE config 2 in miniature (C96, hord_tr 8 / kord_tr 9, device-resident multi-step run): flux form conserves the
    global tracer mass sum(q*dp*area) through tracer_2d, the remap conserves the mass it is handed (measured in the
    Lagrangian layers of pe) -- both to <= 1e-13 relative per step -- and the positive tracers stay non-negative.
    Size-independent properties: no oracle involved."""
    case = case_factory(96, 8, 9, "float64", courant=0.7)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        ctx.upload(f, getattr(case, f), case.nq)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    kord = np.full(case.nq, 9, dtype=np.int32)
    sl = slice(NG, -NG)
    area = case.metrics()["area"][:, None, None, sl, sl]
    rarea = case.metrics()["rarea"][:, None, sl, sl]
    dp1 = case.dp1[:, None, :, sl, sl]
    dp2 = (case.dp1[..., sl, sl] + (case.mfx[..., :, :-1] - case.mfx[..., :, 1:] + case.mfy[..., :-1, :] - case.mfy[..., 1:, :]) * rarea)[:, None]
    dp_lag = np.diff(case.pe, axis=2).transpose(0, 2, 1, 3)[:, None, :, 1:-1, 1:-1]
    positive = case.q[..., sl, sl].min(axis=(0, 2, 3, 4)) >= 0
    assert positive.sum() >= 6
    q = np.empty_like(case.q)
    delp = np.empty_like(case.dp1)

    def mass(thick):
        return (q[..., sl, sl] * thick * area).sum(axis=(0, 2, 3, 4))

    for step in range(4):
        ctx.download("q", q, case.nq)
        m0 = mass(dp1)
        assert ctx.tracer_2d_resident(case.nq, 8) == 1
        ctx.download("q", q, case.nq)
        m1 = mass(dp2)
        assert (np.abs(m1 - m0) <= 1e-13 * np.abs(m0)).all(), (step, (m1 - m0) / m0)
        assert (q[..., sl, sl].min(axis=(0, 2, 3, 4))[positive] >= 0).all()
        ml = mass(dp_lag)
        ctx.remap_tracers_resident(case.nq, kord, fill=True)
        ctx.download("q", q, case.nq)
        ctx.download("delp", delp, case.nq)
        m2 = mass(delp[:, None, :, sl, sl])
        assert (np.abs(m2 - ml)[positive] <= 1e-13 * np.abs(ml)[positive]).all(), (step, ((m2 - ml) / ml)[positive])
        assert (q[..., sl, sl].min(axis=(0, 2, 3, 4))[positive] >= 0).all()
    ctx.close()


# ---- edge cases and BASELINE configurations -----------------------------------------------------------------------------
def _remap_gpu(case, q0, kord, fill=True):
    ctx = TracerContext(case.n + 1, case.npz, q0.shape[1], case.metrics(), dtype=case.dtype)
    q = np.array(q0, copy=True)
    delp = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, q, delp, kord, fill=fill)
    ctx.close()
    return q, delp


def test_config1_full_step_c48_l64(oracle, case_factory, mode):
    """BASELINE config 1: C48 L64, 9 tracers, fp64, hord_tr 8 / kord_tr 9: one tracer_2d + one tracer remap (mapn_tracer)."""
    case = case_factory(48, 64, 9, "float64")
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=False, what="C48 L64 advect")
    qref, dref = oracle.remap_tracers(got["q"], case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    q, delp = _remap_gpu(case, got["q"], 9)
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    check_q(q, qref, mode, "float64", bit_exact_expected=False, what="C48 L64 remap")


@pytest.mark.parametrize("npz", [6, 127, 128])
def test_level_count_extremes(oracle, case_factory, npz, mode):
    """The smallest (6) and largest (128; operational 127) level counts the build supports."""
    case = case_factory(12, npz, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    assert got["nsplt"] == ref["nsplt"] and np.array_equal(got["ksplt"], ref["ksplt"])
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=False, what=f"npz={npz} advect")
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    q, delp = _remap_gpu(case, case.q, 9)
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    check_q(q, qref, mode, "float64", bit_exact_expected=False, what=f"npz={npz} remap")


@pytest.mark.parametrize("nq", [1, 6, 30])
def test_tracer_count_extremes(oracle, case_factory, nq, mode):
    """One tracer (map1_q2 branch of the remap), the smallest mapn_tracer set (6) and the 30-tracer aerosol-suite size."""
    base = case_factory(12, 16, 9, "float64")
    reps = -(-nq // 9)
    q0 = np.ascontiguousarray(np.concatenate([base.q * (1.0 + 0.25 * r) for r in range(reps)], axis=1)[:, :nq])
    import copy
    case = copy.copy(base)
    case.q, case.nq = q0, nq
    ref = oracle.tracer_2d(case, hord=8)
    got = run_gpu_tracer_2d(case, 8)
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=False, what=f"nq={nq} advect")
    qref, _ = oracle.remap_tracers(q0, case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    q, _ = _remap_gpu(case, q0, 9)
    check_q(q, qref, mode, "float64", bit_exact_expected=(nq <= 5), what=f"nq={nq} remap")


def test_q_split_and_no_fill(oracle, case_factory, mode):
    """q_split /= 0 fixes nsplt (fv_tracer2d.F90:437-441); fill = .false. skips fillz."""
    case = case_factory(24, 16, 9, "float64", courant=0.7)
    ref = oracle.tracer_2d(case, hord=8, q_split=3)
    got = run_gpu_tracer_2d(case, 8, q_split=3)
    assert got["nsplt"] == ref["nsplt"] == 3
    check_q(got["q"], ref["q"], mode, "float64", bit_exact_expected=False, what="q_split=3")
    sl = slice(NG, -NG)
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    qref, _ = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, 9, fill=False)
    q, _ = _remap_gpu(case, case.q, 9, fill=False)
    check_q(q, qref, mode, "float64", bit_exact_expected=False, what="fill=False")


def test_row_granular_mapn_tracer_entry(oracle, case_factory, mode):
    """fv3t_*_mapn_tracer with the reference's own argument list (fv_mapz.F90:1386-1402), one tile, row by row as the
    reference's j loop calls it: equals the batched remap of that tile."""
    case = case_factory(12, 16, 9, "float64")
    qref, _ = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    n, km, nq = case.n, case.npz, case.nq
    g = {k: v[:1] for k, v in case.metrics().items()}
    ctx = TracerContext(n + 1, km, nq, g, dtype=case.dtype, tiles=(1,))
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    q1 = np.ascontiguousarray(case.q[0])                      # (nq, km, n+6, n+6) = q1(isd:ied, jsd:jed, km, nq)
    kord = np.full(nq, 9, dtype=np.int32)
    for j in range(1, n + 1):
        pe1 = np.ascontiguousarray(case.pe[0, j, :, 1:-1])    # pe(is:ie, 1:km+1, j) -> (km+1, n)
        ps = pe1[-1]
        pe2 = case.ak[:, None] + case.bk[:, None] * ps[None, :]
        pe2[0], pe2[-1] = case.ptop, ps
        dp2 = np.ascontiguousarray(np.diff(pe2, axis=0))
        ctx.mapn_tracer(nq, km, pe1, np.ascontiguousarray(pe2), q1, dp2, kord, j, 1, n, -2, n + 3, -2, n + 3, 0.0, True)
    ctx.close()
    sl = slice(NG, -NG)
    assert np.array_equal(q1[..., sl, sl], qref[0][..., sl, sl])


def test_remap_prepare_overlap_is_transparent(oracle, case_factory, mode):
    """fv3t_*_remap_prepare (coefficients + delp on a side stream while tracer_2d runs) must not change any result, and a pe
    upload after it must invalidate it."""
    case = case_factory(24, 16, 9, "float64", courant=0.7)
    kord = np.full(9, 9, dtype=np.int32)
    outs = []
    for prepare in (False, True, "stale"):
        ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
        for f in ("q", "dp1", "mfx", "mfy", "cx", "cy"):
            ctx.upload(f, getattr(case, f), case.nq)
        ctx.set_vertical(case.ak, case.bk, case.ptop)
        if prepare == "stale":
            ctx.upload("pe", np.ascontiguousarray(case.pe[::-1]), case.nq)   # wrong pe ...
            ctx.remap_prepare()
        ctx.upload("pe", case.pe, case.nq)                                    # ... replaced: the prepared coefficients are dropped
        if prepare is True:
            ctx.remap_prepare()
        ctx.tracer_2d_resident(case.nq, 8)
        ctx.remap_tracers_resident(case.nq, kord, fill=True)
        q = np.empty_like(case.q)
        delp = np.empty_like(case.dp1)
        ctx.download("q", q, case.nq)
        ctx.download("delp", delp, case.nq)
        ctx.close()
        outs.append((q, delp))
    sl = slice(NG, -NG)
    for q, delp in outs[1:]:
        assert np.array_equal(q[..., sl, sl], outs[0][0][..., sl, sl])
        assert np.array_equal(delp[..., sl, sl], outs[0][1][..., sl, sl])


@pytest.mark.parametrize("courant", [0.7, 1.8])
def test_tracer_step_equals_the_two_separate_calls(oracle, case_factory, courant, mode):
    """fv3t_*_tracer_step (tracer_2d + remap on host arrays, pipelined per tracer when nsplt == 1) leaves exactly the
    post-state of fv3t_*_tracer_2d followed by fv3t_*_remap_tracers."""
    case = case_factory(24, 16, 9, "float64", courant=courant)
    kord = np.full(9, 9, dtype=np.int32)
    a = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    nsplt_a, _ = ctx.tracer_2d(a["q"], a["dp1"], a["mfx"], a["mfy"], a["cx"], a["cy"], 8)
    delp_a = np.zeros_like(case.dp1)
    ctx.remap_tracers(case.pe, case.ak, case.bk, case.ptop, a["q"], delp_a, kord, fill=True)
    ctx.close()
    b = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    delp_b = np.zeros_like(case.dp1)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    nsplt_b = ctx.tracer_step(b["q"], b["dp1"], b["mfx"], b["mfy"], b["cx"], b["cy"], case.pe, case.ak, case.bk, case.ptop, delp_b,
                              8, kord, fill=True)
    launches = ctx.kernel_launches()
    ctx.close()
    assert nsplt_a == nsplt_b and (nsplt_b == 1) == (courant < 1)
    sl = slice(NG, -NG)
    assert np.array_equal(a["q"][..., sl, sl], b["q"][..., sl, sl])
    assert np.array_equal(delp_a[..., sl, sl], delp_b[..., sl, sl])
    assert np.array_equal(a["dp1"][..., sl, sl], b["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(a[k], b[k]), k
    assert launches > 0


# ---- boundary: tracer_2d_1L and the row-granular mapn_tracer with the caller's own target grid ----------------------------------
@pytest.mark.parametrize("hord", [8, 10])
def test_tracer_2d_1l_entry(oracle, case_factory, hord, mode):
    """fv3t_*_tracer_2d_1L (fv_tracer2d.F90:92-321, z_tracer): q as tracer_2d; cx, cy, mfx, mfy and -- unlike tracer_2d -- the dp1
    post-state of a level advanced only between its own sub-steps (:305), all bit-identical to the oracle's tracer_2d_1L."""
    case = case_factory(24, 16, 9, "float64", courant=3.3)
    ref = oracle.tracer_2d_1l(case, hord=hord)
    std = oracle.tracer_2d(case, hord=hord)
    assert len(set(ref["ksplt"].tolist())) > 1 and not np.array_equal(ref["dp1"], std["dp1"])   # the two routines do differ here
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    nsplt, ksplt = ctx.tracer_2d_1L(out["q"], out["dp1"], out["mfx"], out["mfy"], out["cx"], out["cy"], hord)
    ctx.close()
    assert nsplt == ref["nsplt"] and np.array_equal(ksplt, ref["ksplt"])
    check_q(out["q"], ref["q"], mode, "float64", bit_exact_expected=hord not in FAST_HORD, what=f"1L hord={hord}")
    sl = slice(NG, -NG)
    assert np.array_equal(out["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(out[k], ref[k]), k


@pytest.mark.parametrize("nq", [3, 9])
def test_mapn_tracer_consumes_the_callers_target_grid(oracle, case_factory, nq, mode):
    """fv3t_*_mapn_tracer with a pe2 / dp2 that is NOT ak + bk*ps, and with nq <= 5 (where Lagrangian_to_Eulerian itself would
    have dispatched to map1_q2): the entry is mapn_tracer -- scalar_profile always, the caller's target grid as given -- and
    must equal the oracle's mapn_tracer column by column, bit for bit (strict kernel)."""
    import oracle_binding as ob
    case = case_factory(12, 16, 9, "float64")
    n, km = case.n, case.npz
    g = {k: v[:1] for k, v in case.metrics().items()}
    ctx = TracerContext(n + 1, km, nq, g, dtype=case.dtype, tiles=(1,))
    q1 = np.ascontiguousarray(case.q[0, :nq])
    q_in = q1.copy()
    kord = np.full(nq, 9, dtype=np.int32)
    rng = np.random.default_rng(7)
    rows = (1, 5, n)
    pe2_rows = {}
    for j in rows:
        pe1 = np.ascontiguousarray(case.pe[0, j, :, 1:-1])    # (km+1, n)
        w = rng.uniform(-0.3, 0.3, size=pe1.shape)
        pe2 = pe1 + w * np.gradient(pe1, axis=0)               # an arbitrary monotone target grid with the same ends
        pe2[0], pe2[-1] = pe1[0], pe1[-1]
        pe2 = np.ascontiguousarray(np.sort(pe2, axis=0))
        dp2 = np.ascontiguousarray(np.diff(pe2, axis=0))
        pe2_rows[j] = (pe1, pe2)
        ctx.mapn_tracer(nq, km, pe1, pe2, q1, dp2, kord, j, 1, n, -2, n + 3, -2, n + 3, 0.0, True)
    ctx.close()
    for j in rows:
        pe1, pe2 = pe2_rows[j]
        for i in range(n):
            col = np.ascontiguousarray(q_in[:, :, j + 2, i + 3])      # [nq, km]
            want = ob.map_col(0, pe1[:, i], pe2[:, i], col, kord, q_min=0.0, fill=True)
            assert np.array_equal(q1[:, :, j + 2, i + 3], want), (j, i)
    untouched = [j for j in range(1, n + 1) if j not in rows]
    assert np.array_equal(q1[:, :, [j + 2 for j in untouched]], q_in[:, :, [j + 2 for j in untouched]])


# ---- cs_profile on the GPU: map1_ppm / map_scalar (the rest of the vertical remap reaches cs_profile through them) ----------------
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("use_cs", [True, False])
@pytest.mark.parametrize("kord", [9, 11, 15, 10, 12, 13, 14, 16, 8, 17, 4, 6, 7])
def test_map1_ppm_and_map_scalar_against_the_oracle(oracle, case_factory, kord, use_cs, dtype):
    """fv3t_*_map1_ppm (cs_profile, fv_mapz.F90:2098-2498 -- including its own kord 9 / 11 / 15 branches and the
    6*a1 - 3*(a2+a3) rounding of :2324) and fv3t_*_map_scalar (scalar_profile with q_min) for iv = 0, 1, -1 and -2 (bottom value
    given): bit-identical to the oracle's restatement, column by column."""
    import oracle_binding as ob
    case = case_factory(12, 32, 9, dtype)
    n, km = case.n, case.npz
    ctx = TracerContext(n + 1, km, 1, case.metrics(), dtype=case.dtype)
    ctx.upload("pe", case.pe)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    rng = np.random.default_rng(kord)
    fields = {0: np.ascontiguousarray(case.q[:, 4]),                                           # sparse random positive (iv = 0)
              1: np.ascontiguousarray(case.q[:, 5] * 30 + 280),                                # temperature-like (iv = 1)
              -1: np.ascontiguousarray(case.q[:, 5] * 40),                                     # signed, wind-like (iv = -1)
              -2: np.ascontiguousarray(case.q[:, 1] * 1e3 + 1)}                                # with a bottom boundary value
    ak = np.asarray(case.ak, dtype=case.dtype)
    bk = np.asarray(case.bk, dtype=case.dtype)
    for iv, f in fields.items():
        got = f.copy()
        qs = np.ascontiguousarray(f[:, -1] * case.dtype.type(1.05)) if iv == -2 else None
        q_min = 0.0 if iv != 0 else 1e-6
        ctx.map_field(got, iv, kord, q_min=q_min, qs=qs, use_cs=use_cs)
        for t in (0, 3, 5):
            for (i, j) in [(1, 1), (n, n), (5, 7), (n, 2), (3, n)]:
                pe1 = np.ascontiguousarray(case.pe[t, j, :, i])
                ps = pe1[-1]
                pe2 = (ak + bk * ps).astype(case.dtype)
                pe2[0], pe2[-1] = case.dtype.type(case.ptop), ps
                col = np.ascontiguousarray(f[t, :, j + 2, i + 2])
                want = ob.map_field_col(use_cs, pe1, pe2, col, iv, kord, q_min=q_min, qs=float(qs[t, j + 2, i + 2]) if qs is not None else 0.0)
                assert np.array_equal(got[t, :, j + 2, i + 2], want), (iv, t, i, j, np.abs(got[t, :, j + 2, i + 2] - want).max())
    ctx.close()


# ---- tracer damping: deln_flux on the first sub-step -------------------------------------------------------------------------
@pytest.mark.parametrize("nord", [0, 1, 2])
@pytest.mark.parametrize("courant", [0.7, 1.8])
def test_tracer_damping_deln_flux(oracle, case_factory, nord, courant, mode):
    """tracer_2d with trdm2 > 1e-4 (fv_tracer2d.F90:487-494, 527-532 -> deln_flux, tp_core.F90:1239-1387): del-2 / del-4 / del-6,
    with and without sub-cycling.  The library applies the damping fluxes as a correction of the advected field (same terms, other
    summation order), so the bar is the north-star tolerance, not bit equality; the dp1 / cx / cy / mfx / mfy post-state and the
    sub-step counts are unchanged by damping."""
    import oracle_binding as ob
    from fv3atm_b200 import cubed_sphere as cs
    case = case_factory(24, 16, 9, "float64", courant=courant)
    d6u, d6v, da_min = cs.damping_metrics(case.grid)
    trdm = 0.15
    ref = ob.tracer_2d_damp(case, 8, nord, trdm, d6u, d6v, da_min)
    plain = oracle.tracer_2d(case, hord=8)
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    ctx.set_damping(d6u, d6v, da_min)
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    nsplt, ksplt = ctx.tracer_2d(out["q"], out["dp1"], out["mfx"], out["mfy"], out["cx"], out["cy"], 8, nord_tr=nord, trdm=trdm)
    ctx.close()
    assert nsplt == ref["nsplt"] and np.array_equal(ksplt, ref["ksplt"])
    nd = norm_diff(out["q"], ref["q"])
    assert nd.max() <= 1e-12, f"nord={nord}: {nd}"
    assert norm_diff(ref["q"], plain["q"]).max() > 1e-4   # the damping does something in this case
    sl = slice(NG, -NG)
    assert np.array_equal(out["dp1"][..., sl, sl], plain["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(out[k], plain[k]), k


def test_resident_calls_reject_a_tracer_count_other_than_the_uploaded_one(case_factory):
    """nq is the tile stride of the resident q: tracer_2d_resident / remap_tracers_resident with another count than the upload's
    would address other planes -- an error, not a silent wrong answer."""
    from fv3atm_b200.lib import Fv3tError
    case = case_factory(24, 16, 9, "float64")
    ctx = TracerContext(case.n + 1, case.npz, case.nq, case.metrics(), dtype=case.dtype)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        ctx.upload(f, getattr(case, f), case.nq)
    ctx.set_vertical(case.ak, case.bk, case.ptop)
    with pytest.raises(Fv3tError, match="uploaded as 9"):
        ctx.tracer_2d_resident(5, 8)
    with pytest.raises(Fv3tError, match="uploaded as 9"):
        ctx.remap_tracers_resident(6, 9, True)
    assert ctx.tracer_2d_resident(case.nq, 8) >= 1
    ctx.close()
