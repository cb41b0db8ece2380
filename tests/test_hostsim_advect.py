"""CPU check of the product's fast advection path: the phase functions of the marching CTA, the tracer-independent
preparation (k_prep3) and the sub-step bookkeeping of fv3atm_b200/csrc/fv3t_advect3.cuh are compiled for the host
(tests/hostsim/, test infrastructure) and executed thread by thread; the result must agree with the oracle to the
north-star bar (max normalised difference <= 1e-12 in fp64, <= 1e-5 in fp32).  The fast path multiplies by shared
reciprocals where the reference divides, so it is not expected to be bit-identical (the strict kernels are: see
tests/test_gpu_parity.py).  The GPU build of the same functions is checked by tests/test_gpu_parity.py."""
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
def sim():
    so = os.path.join(SIM, "libhostsim_advect3.so")
    src = os.path.join(SIM, "advect3_hostsim.cu")
    csrc = os.path.join(HERE, "..", "fv3atm_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fv3t_advect4.cuh", "fv3t_advect3.cuh", "fv3t_advect2.cuh", "fv3t_advect.cuh", "fv3t_ppm.cuh",
                                                     "fv3t_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["nvcc", "-x", "cu", "-O2", "-std=c++17", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC,-fno-fast-math", "-shared", "-o", so, src], check=True, cwd=SIM)
    return C.CDLL(so)


def run_sim(sim, case, hord, ref, NT, lim_fac=1.0, group=1):
    sim.hostsim_set_group(int(group))
    sfx, ct = ("f64", C.c_double) if case.dtype == np.float64 else ("f32", C.c_float)
    g = case.metrics()
    out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
    dst, src = ob.halo_offsets(case.n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ksplt = np.ascontiguousarray(ref["ksplt"], dtype=np.int32)
    rc = getattr(sim, f"hostsim_tracer_2d_{sfx}")(
        case.n, case.npz, case.nq, p(out["q"]), p(out["dp1"]), p(out["mfx"]), p(out["mfy"]), p(out["cx"]), p(out["cy"]),
        p(g["area"]), p(g["rarea"]), p(g["dx"]), p(g["dy"]), p(g["dxa"]), p(g["dya"]), p(g["sin_sg"]), p(dst), p(src),
        C.c_int64(dst.size), int(hord), ct(lim_fac), int(ref["nsplt"]), p(ksplt), int(NT))
    assert rc == 0
    return out


def norm_diff(a, b):
    sl = slice(NG, -NG)
    d = np.abs(a[..., sl, sl].astype(np.float64) - b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    s = np.abs(b[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    return d / np.maximum(s, 1e-300)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("hord", [8, 10, 9, 13, 12, 11, 2])
def test_fast_advection_matches_oracle(sim, oracle, case_factory, hord, dtype):
    """The schemes the product runs on the fast path (fv3t::fast_hord_ok)."""
    case = case_factory(12, 8, 9, dtype)
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_sim(sim, case, hord, ref, NT=32)
    nd = norm_diff(got["q"], ref["q"])
    assert nd.max() <= TOL[case.dtype], f"hord={hord}: {nd}"


@pytest.mark.parametrize("hord", [7, -5, 5, 6, 1, 3, 4])
def test_discontinuous_schemes_stay_on_the_strict_path(sim, oracle, case_factory, hord):
    """hord 1, 3-7, -5 switch on `bl*br < 0`-type tests that the reference evaluates on EXACT zeros for piecewise-
    constant data (tp_core.F90:421-519): re-associated arithmetic flips them, so fast_hord_ok() excludes these schemes
    and the product runs its bit-exact strict kernel for them.  This test documents why: the fast arithmetic agrees
    on every smooth tracer but is off by O(1e-4..1e-2) on the slotted cylinder (tracer 2)."""
    case = case_factory(12, 8, 9, "float64")
    ref = oracle.tracer_2d(case, hord=hord)
    got = run_sim(sim, case, hord, ref, NT=32)
    nd = norm_diff(got["q"], ref["q"])
    smooth = [i for i in range(9) if i != 2]
    assert nd[smooth].max() <= 1e-12, nd
    assert nd[2] > 1e-9, nd


def test_async_ring_variant_fp32(sim, oracle, case_factory):
    case = case_factory(12, 8, 9, "float32", courant=1.8)
    ref = oracle.tracer_2d(case, hord=8)
    ring = run_sim(sim, case, 8, ref, NT=32, group=4)
    regs = run_sim(sim, case, 8, ref, NT=32, group=1)
    assert np.array_equal(ring["q"], regs["q"])
    assert norm_diff(ring["q"], ref["q"]).max() <= 1e-5


@pytest.mark.parametrize("hord", [8, 10, 13])
@pytest.mark.parametrize("courant", [0.7, 1.8])
def test_async_ring_variant_equals_register_prefetch_variant(sim, oracle, case_factory, hord, courant):
    """k_advect4 (inputs staged through the cp.async ring; the product default) performs the same arithmetic as
    k_advect3: identical bits in the host simulation, and within the bar of the oracle."""
    case = case_factory(20, 8, 9, "float64", courant=courant)
    ref = oracle.tracer_2d(case, hord=hord)
    ring = run_sim(sim, case, hord, ref, NT=32, group=4)
    regs = run_sim(sim, case, hord, ref, NT=32, group=1)
    assert np.array_equal(ring["q"], regs["q"])
    if hord != 10:
        assert norm_diff(ring["q"], ref["q"]).max() <= 1e-12


@pytest.mark.parametrize("group", [2, 3])
def test_tracer_groups_per_thread(sim, oracle, case_factory, group):
    """G tracers marching in one thread (9 = 4*2+1 and 3*3: the padded last group must not store)."""
    case = case_factory(12, 8, 9, "float64", courant=1.8)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_sim(sim, case, 8, ref, NT=32, group=group)
    one = run_sim(sim, case, 8, ref, NT=32, group=1)
    assert norm_diff(got["q"], ref["q"]).max() <= 1e-12
    assert np.array_equal(got["q"], one["q"])


@pytest.mark.parametrize("NT", [32, 64])
@pytest.mark.parametrize("courant", [1.8, 3.3])
def test_fast_advection_subcycling(sim, oracle, case_factory, courant, NT):
    """nsplt > 1 with level-dependent ksplt: q within tolerance; dp1 and the scaled cx, cy, mfx, mfy post-state are
    bit-identical (uncontracted arithmetic in k_prep3 / k_cab3 / k_scale3)."""
    case = case_factory(20, 8, 9, "float64", courant=courant)
    ref = oracle.tracer_2d(case, hord=8)
    assert ref["nsplt"] >= 2 and len(set(ref["ksplt"].tolist())) > 1
    got = run_sim(sim, case, 8, ref, NT=NT)
    nd = norm_diff(got["q"], ref["q"])
    assert nd.max() <= 1e-12, nd
    sl = slice(NG, -NG)
    assert np.array_equal(got["dp1"][..., sl, sl], ref["dp1"][..., sl, sl])
    for k in ("cx", "cy", "mfx", "mfy"):
        assert np.array_equal(got[k], ref[k]), k


def test_fast_advection_conserves_mass_and_free_stream(sim, oracle, case_factory):
    """Flux form: global tracer mass sum(q*dp*area) is conserved to rounding; q == 1 stays 1 under non-divergent winds."""
    case = case_factory(12, 8, 9, "float64", courant=0.7, divergent=0.0)
    ref = oracle.tracer_2d(case, hord=8)
    got = run_sim(sim, case, 8, ref, NT=32)
    sl = slice(NG, -NG)
    area = case.metrics()["area"][:, None, None, sl, sl]
    m0 = (case.q[..., sl, sl] * case.dp1[:, None, :, sl, sl] * area).sum(axis=(0, 2, 3, 4))
    # dp2 of the (single) sub-step
    n = case.n
    mfx, mfy = case.mfx, case.mfy
    rarea = case.metrics()["rarea"][:, None, sl, sl]
    dp2 = case.dp1[..., sl, sl] + (mfx[..., :, :-1] - mfx[..., :, 1:] + mfy[..., :-1, :] - mfy[..., 1:, :]) * rarea
    m1 = (got["q"][..., sl, sl] * dp2[:, None] * area).sum(axis=(0, 2, 3, 4))
    assert (np.abs(m1 - m0) <= 1e-13 * np.abs(m0).max()).all(), (m1 - m0) / m0
    assert ref["nsplt"] == 1
