"""CPU check of the product's re-scheduled remap arithmetic: the __host__ __device__ column routine of
fv3atm_b200/csrc/fv3t_remap2.cuh is compiled for the host (tests/hostsim/, test infrastructure) and must reproduce the
oracle bit-for-bit.  The GPU build of the same template is checked by tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, "hostsim")
NG = 3


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM, "libhostsim.so")
    src = os.path.join(SIM, "remap_hostsim.cu")
    deps = [src] + [os.path.join(HERE, "..", "fv3atm_b200", "csrc", f) for f in ("fv3t_remap2.cuh", "fv3t_remap.cuh", "fv3t_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["nvcc", "-x", "cu", "-O2", "-std=c++17", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                        "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-o", so, src],
                       check=True, cwd=SIM)
    return C.CDLL(so)


def run_sim(sim, G, q, pe, ak, bk, ptop, kord, fill):
    nt, nq, km, nd, _ = q.shape
    n = nd - 6
    sfx, ct = ("f64", C.c_double) if q.dtype == np.float64 else ("f32", C.c_float)
    qs = np.ascontiguousarray(q)
    qd = np.array(q, copy=True)
    delp = np.zeros((nt, km, nd, nd), dtype=q.dtype)
    kord = np.ascontiguousarray(np.broadcast_to(np.asarray(kord, dtype=np.int32), (nq,)))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pe = np.ascontiguousarray(pe, dtype=q.dtype)
    ak = np.ascontiguousarray(ak, dtype=q.dtype)
    bk = np.ascontiguousarray(bk, dtype=q.dtype)
    getattr(sim, f"hostsim_remap_{sfx}")(G, nt, n, km, nq, p(pe), p(ak), p(bk), ct(ptop), p(qs), p(qd), p(delp), p(kord), int(fill))
    return qd, delp


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kord", [9, 8, 10, 11, 12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("G", [3, 1])
def test_streaming_remap_is_bit_identical_to_oracle(sim, oracle, case_factory, kord, dtype, G):
    case = case_factory(12, 32, 9, dtype)
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    q, delp = run_sim(sim, G, case.q, case.pe, case.ak, case.bk, case.ptop, kord, True)
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl]), np.abs(q[..., sl, sl] - qref[..., sl, sl]).max()


@pytest.mark.parametrize("fill", [False, True])
@pytest.mark.parametrize("nq,G", [(7, 3), (8, 3), (6, 2), (4, 3), (5, 2)])
def test_ragged_tracer_groups_and_map1_q2(sim, oracle, case_factory, nq, G, fill):
    """nq not a multiple of the group size; nq <= 5 takes the map1_q2 form of the overlap integrals."""
    case = case_factory(12, 32, 9, "float64")
    q0 = np.ascontiguousarray(case.q[:, :nq])
    kord = [9, 10, 11, 12, 13, 14, 15, 16, 8][:nq]
    qref, dref = oracle.remap_tracers(q0, case.pe, case.ak, case.bk, case.ptop, kord, fill=fill)
    q, delp = run_sim(sim, G, q0, case.pe, case.ak, case.bk, case.ptop, kord, fill)
    sl = slice(NG, -NG)
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl])


def test_fillz_paths_are_exercised(sim, oracle, case_factory):
    """A signed tracer forces negative mapped values: top/interior/bottom borrowing and the non-local rescale."""
    case = case_factory(12, 32, 9, "float64")
    q0 = np.array(case.q, copy=True)
    rng = np.random.default_rng(5)
    q0[:, 0] = rng.standard_normal(q0[:, 0].shape) * 1e-3 + 2e-4       # many negatives
    q0[:, 1, 0] = -np.abs(q0[:, 1, 0]) - 1e-6                          # negative top layer
    q0[:, 2, -1] = -np.abs(q0[:, 2, -1]) - 1e-6                        # negative bottom layer
    qref, _ = oracle.remap_tracers(q0, case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    q, _ = run_sim(sim, 3, q0, case.pe, case.ak, case.bk, case.ptop, 9, True)
    sl = slice(NG, -NG)
    assert (qref[:, 0, :, sl, sl] != q0[:, 0, :, sl, sl]).any()
    assert np.array_equal(q[..., sl, sl], qref[..., sl, sl])


# ---- the fast remap (fv3t_remap3.cuh): shared spline coefficients and reciprocals -> within the north-star bar ----------
@pytest.fixture(scope="module")
def sim3():
    so = os.path.join(SIM, "libhostsim_remap3.so")
    src = os.path.join(SIM, "remap3_hostsim.cu")
    csrc = os.path.join(HERE, "..", "fv3atm_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fv3t_remap5.cuh", "fv3t_remap3.cuh", "fv3t_remap2.cuh", "fv3t_remap.cuh", "fv3t_common.cuh",
                                                     "fv3t_advect3.cuh", "fv3t_advect4.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["nvcc", "-x", "cu", "-O2", "-std=c++17", "--extended-lambda", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC,-fno-fast-math", "-shared", "-o", so, src], check=True, cwd=SIM)
    return C.CDLL(so)


@pytest.fixture(params=[3, 5], ids=["three-walk", "two-walk"])
def variant(request, sim3):
    """k_remap3's column routine / k_remap5's (bottom-up elimination, fv3t_remap5.cuh)"""
    sim3.hostsim_remap_variant(request.param)
    yield request.param
    sim3.hostsim_remap_variant(3)


def run_sim3(sim3, q, pe, ak, bk, ptop, akord, fill):
    nt, nq, km, nd, _ = q.shape
    n = nd - 6
    sfx, ct = ("f64", C.c_double) if q.dtype == np.float64 else ("f32", C.c_float)
    qs = np.ascontiguousarray(q)
    qd = np.array(q, copy=True)
    delp = np.zeros((nt, km, nd, nd), dtype=q.dtype)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pe = np.ascontiguousarray(pe, dtype=q.dtype)
    ak = np.ascontiguousarray(ak, dtype=q.dtype)
    bk = np.ascontiguousarray(bk, dtype=q.dtype)
    rc = getattr(sim3, f"hostsim_remap3_{sfx}")(nt, n, km, nq, p(pe), p(ak), p(bk), ct(ptop), p(qs), p(qd), p(delp), int(akord),
                                                int(fill))
    assert rc == 0
    return qd, delp


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kord", [9, 8, 12, 13, 14, 17])
def test_fast_remap_matches_oracle(sim3, oracle, case_factory, kord, dtype, variant):
    """The limiters the product runs on the fast path (fv3t::fast_kord_ok).  kord 10, 11, 15, 16 compare COMPUTED interface
    values (ext5 / ext6, fv_mapz.F90:1820-1846) that are exactly equal on flat data, so any re-association flips them: the
    product keeps those on its bit-exact strict kernel."""
    case = case_factory(12, 32, 9, dtype)
    qref, dref = oracle.remap_tracers(case.q, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
    q, delp = run_sim3(sim3, case.q, case.pe, case.ak, case.bk, case.ptop, kord, True)
    sl = slice(NG, -NG)
    assert np.array_equal(delp[..., sl, sl], dref[..., sl, sl])
    d = np.abs(q[..., sl, sl].astype(np.float64) - qref[..., sl, sl]).max(axis=(0, 2, 3, 4))
    s = np.abs(qref[..., sl, sl].astype(np.float64)).max(axis=(0, 2, 3, 4))
    tol = 1e-12 if dtype == "float64" else 1e-5
    assert (d / np.maximum(s, 1e-300)).max() <= tol, d / s


def test_fast_remap_fillz_and_conservation(sim3, oracle, case_factory, variant):
    """Signed tracer: every fillz branch; and the column integral sum(q*dp) is conserved by the remap without fill."""
    case = case_factory(12, 32, 9, "float64")
    q0 = np.array(case.q, copy=True)
    rng = np.random.default_rng(5)
    q0[:, 0] = rng.standard_normal(q0[:, 0].shape) * 1e-3 + 2e-4
    q0[:, 1, 0] = -np.abs(q0[:, 1, 0]) - 1e-6
    q0[:, 2, -1] = -np.abs(q0[:, 2, -1]) - 1e-6
    qref, _ = oracle.remap_tracers(q0, case.pe, case.ak, case.bk, case.ptop, 9, fill=True)
    q, _ = run_sim3(sim3, q0, case.pe, case.ak, case.bk, case.ptop, 9, True)
    sl = slice(NG, -NG)
    d = np.abs(q[..., sl, sl] - qref[..., sl, sl]).max(axis=(0, 2, 3, 4)) / np.abs(qref[..., sl, sl]).max(axis=(0, 2, 3, 4))
    assert d.max() <= 1e-12, d
    # conservation without fill: sum_k q*dp1 (Lagrangian) == sum_k q2*dp2 (Eulerian)
    q2, delp = run_sim3(sim3, case.q, case.pe, case.ak, case.bk, case.ptop, 9, False)
    dp1 = np.diff(case.pe, axis=2).transpose(0, 2, 1, 3)[:, :, 1:-1, 1:-1]          # [6, km, n, n]
    m1 = (case.q[..., sl, sl] * dp1[:, None]).sum(axis=2)
    m2 = (q2[..., sl, sl] * delp[:, None, :, sl, sl]).sum(axis=2)
    assert np.abs(m2 - m1).max() <= 1e-13 * np.abs(m1).max()
