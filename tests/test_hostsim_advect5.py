"""CPU check of the product's multi-tracer advection path (fv3atm_b200/csrc/fv3t_advect5.cuh): the four phase functions of a
marching tracer group, the staged-box layout the TMA producer fills, the tracer-independent preparation (k_prep5) and the
sub-step bookkeeping are compiled for the host (tests/hostsim/, test infrastructure) and executed thread by thread; the result
must agree with the oracle to the north-star bar (max normalised difference <= 1e-12 in fp64, <= 1e-5 in fp32), and the
caller-visible post-state (dp1, cx, cy, mfx, mfy) bit for bit.  The GPU build of the same functions is checked by
tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, "hostsim")
NG = 3
TOL = {np.dtype("float64"): 1e-12, np.dtype("float32"): 1e-5}


@pytest.fixture(scope="module")
def sim5():
    so = os.path.join(SIM, "libhostsim_advect5.so")
    src = os.path.join(SIM, "advect5_hostsim.cu")
    csrc = os.path.join(HERE, "..", "fv3atm_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fv3t_advect5.cuh", "fv3t_advect4.cuh", "fv3t_advect3.cuh", "fv3t_advect2.cuh", "fv3t_advect.cuh",
                                                     "fv3t_ppm.cuh", "fv3t_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["nvcc", "-x", "cu", "-O1", "-std=c++17", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC,-fno-fast-math", "-shared", "-o", so, src], check=True, cwd=SIM)
    return C.CDLL(so)


def run_sim(sim, case, hord, ref, lim_fac=1.0, exact=False):
    sfx, ct = ("f64", C.c_double) if case.dtype == np.float64 else ("f32", C.c_float)
    g = case.metrics()
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    dst, src = ob.halo_offsets(case.n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ksplt = np.ascontiguousarray(ref["ksplt"], dtype=np.int32)
    rc = getattr(sim, f"hostsim5_tracer_2d_{sfx}")(
        case.n, case.npz, case.nq, p(out["q"]), p(out["dp1"]), p(out["mfx"]), p(out["mfy"]), p(out["cx"]), p(out["cy"]),
        p(g["area"]), p(g["rarea"]), p(g["dx"]), p(g["dy"]), p(g["dxa"]), p(g["dya"]), p(g["sin_sg"]), p(dst), p(src),
        C.c_int64(dst.size), int(hord), ct(lim_fac), int(ref["nsplt"]), p(ksplt), int(bool(exact)))
    assert rc == 0
    return out


def norm_diff(a, b):
    sl = slice(NG, -NG)
    d = np.abs(a[..., sl, sl].astype(np.float64) - b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    s = np.abs(b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    return d / np.maximum(s, 1e-300)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 11, 2])
def test_advect5_matches_oracle(sim5, oracle, case_factory, hord, dtype):
    """The schemes with a fast instantiation (fv3t::fast_hord_ok): shared reciprocals, within the north-star bar."""
    case = case_factory(12, 8, 9, dtype)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_sim(sim5, case, hord, ref)
    nd = norm_diff(got["q"], ref["q"])
    assert nd.max() <= TOL[case.dtype], f"hord={hord}: {nd}"


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 10, 9, 13, 12, 7, 11, -5, 5, 6, 1, 2, 3, 4])
def test_advect5_exact_instantiation_is_bit_identical(sim5, oracle, case_factory, hord, dtype):
    """Every scheme in the reference's own operation order (EX = true): bit-identical to the FMA-free oracle, including the
    caller-visible post-state, with sub-stepping (levels dropping out, lazily advanced dp1)."""
    case = case_factory(12, 8, 9, dtype, courant=1.8)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_sim(sim5, case, hord, ref, exact=True)
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl]), f"hord={hord}: {norm_diff(got['q'], ref['q'])}"
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("hord", [8, 11])
def test_advect5_substeps_and_post_state(sim5, oracle, case_factory, hord):
    """nsplt > 1: the lazily advanced dp1, the 1/ksplt scaling applied at the end, levels dropping out of the sub-step loop."""
    case = case_factory(12, 8, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=hord)
    assert ref["nsplt"] >= 2
    got = run_sim(sim5, case, hord, ref)
    assert norm_diff(got["q"], ref["q"]).max() <= 1e-12
    sl = slice(NG, -NG)
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_advect5_interior_strips_and_blocks(sim5, oracle, case_factory, dtype):
    """C128: three strips (the middle one runs the edge-free x code), 30 interior row blocks.  fp32: the scalar TMA boxes of
    strips 1 and 2 start at a column that is not a multiple of four and are read at a shift (A5Stage::shift)."""
    case = case_factory(128, 2, 3, dtype)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_sim(sim5, case, 8, ref)
    assert norm_diff(got["q"], ref["q"]).max() <= TOL[case.dtype]
    got = run_sim(sim5, case, 8, ref, exact=True)
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl])


# ---- sub-tile decomposition (layout L x L per tile): edge formulas on true tile edges only, corner views at true cube corners only,
#      halos (incl. the diagonal blocks) from the neighbouring sub-domains -----------------------------------------------------------
def run_sim_sub(sim, case, hord, ref, L, exact, lim_fac=1.0):
    """The same tracer_2d through the host-simulated kernels, one L x L sub-domain mosaic: returns q on the whole tiles."""
    from fv3atm_b200.subdomain import SubMosaic
    sfx, ct = ("f64", C.c_double) if case.dtype == np.float64 else ("f32", C.c_float)
    mo = SubMosaic(case.n, L)
    ns, m = len(mo), mo.m
    g = case.metrics()
    st = lambda f, a: np.ascontiguousarray(np.stack([f(a, s) for s in range(ns)]))
    q = st(mo.cells, case.q)
    dp1 = st(mo.cells, case.dp1)
    cx, cy = st(mo.xface, case.cx), st(mo.yface, case.cy)
    mfx, mfy = st(mo.mfx, case.mfx), st(mo.mfy, case.mfy)
    gm = {k: st(mo.cells, g[k]) for k in ("area", "rarea", "dxa", "dya", "sin_sg")}
    gm["dx"], gm["dy"] = st(mo.dx, g["dx"]), st(mo.dy, g["dy"])
    ds, do, ss, so = mo.halo_table()
    plane = (m + 6) * (m + 6)
    dst = np.ascontiguousarray(ds * plane + do)
    src = np.ascontiguousarray(ss * plane + so)
    sub = np.ascontiguousarray(np.array([mo.flags(s) for s in range(ns)], dtype=np.int32))
    ksplt = np.ascontiguousarray(ref["ksplt"], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = getattr(sim, f"hostsim5_tracer_2d_sub_{sfx}")(
        ns, m, case.npz, case.nq, p(q), p(dp1), p(mfx), p(mfy), p(cx), p(cy), p(gm["area"]), p(gm["rarea"]), p(gm["dx"]), p(gm["dy"]),
        p(gm["dxa"]), p(gm["dya"]), p(gm["sin_sg"]), p(dst), p(src), C.c_int64(dst.size), int(hord), ct(lim_fac), int(ref["nsplt"]),
        p(ksplt), int(bool(exact)), p(sub))
    assert rc == 0
    qw = np.array(case.q, copy=True)
    dw = np.array(case.dp1, copy=True)
    for s in range(ns):
        mo.put_interior(qw, s, q[s])
        mo.put_interior(dw, s, dp1[s])
    return {"q": qw, "dp1": dw}


def test_submosaic_halo_table_reproduces_the_whole_tile_exchange(case_factory):
    """Scattering a field to sub-domains, exchanging with the gather lists and reading the halos back gives exactly what the
    whole-tile exchange gives -- on every halo cell that is not part of a true cube-corner block, diagonal blocks included."""
    from fv3atm_b200 import cubed_sphere as cs
    from fv3atm_b200.subdomain import SubMosaic
    n = 24
    rng = np.random.default_rng(3)
    a = rng.standard_normal((6, 2, n + 6, n + 6))
    a[:, :, :3, :] = a[:, :, -3:, :] = a[:, :, :, :3] = a[:, :, :, -3:] = np.nan
    whole = cs.fill_edge_halos(a.copy(), n)
    for L in (2, 3):
        mo = SubMosaic(n, L)
        m = mo.m
        stack = np.stack([mo.cells(a, s) for s in range(len(mo))])
        interior = np.zeros((m + 6, m + 6), bool)
        interior[3:-3, 3:-3] = True
        stack[:, :, ~interior] = np.nan
        mo.fill_halos(stack)
        for s in range(len(mo)):
            want = mo.cells(whole, s)
            ok = ~np.isnan(want)
            assert np.array_equal(stack[s][ok], want[ok])
            assert np.isnan(stack[s][~ok]).all()            # true corner blocks stay untouched
            w, e, so_, no, cmask = mo.flags(s)
            assert (~ok[0]).sum() == 9 * bin(cmask).count("1")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord,exact", [(8, False), (10, True), (5, True), (13, True), (-5, True)])
@pytest.mark.parametrize("L", [2, 3])
def test_advect5_submosaic_is_bit_identical_to_whole_tiles(sim5, oracle, case_factory, hord, exact, L, dtype):
    """layout L x L: every sub-domain runs the same kernels with its own edge / corner flags and exchanged halos; the assembled
    field must equal the whole-tile run bit for bit (the per-cell arithmetic does not depend on the decomposition), with
    sub-stepping (two exchanges)."""
    case = case_factory(24, 4, 5, dtype, courant=1.8)
    ref = oracle.tracer_2d(case, hord=hord)
    assert ref["nsplt"] >= 2
    whole = run_sim(sim5, case, hord, ref, exact=exact)
    got = run_sim_sub(sim5, case, hord, ref, L, exact)
    sl = slice(NG, -NG)
    assert np.array_equal(got["q"][..., sl, sl], whole["q"][..., sl, sl]), norm_diff(got["q"], whole["q"])
    assert np.array_equal(got["dp1"][..., sl, sl], whole["dp1"][..., sl, sl])
    if exact:
        assert np.array_equal(got["q"][..., sl, sl], ref["q"][..., sl, sl])
