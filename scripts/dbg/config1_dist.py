import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_binding as ob
from fv3atm_b200 import synthetic as sy
from fv3atm_b200.tracer import TracerContext
case = sy.make_case(48, 64, 9, dtype="float64", courant=0.7)
kord = np.full(9, 9, dtype=np.int32)
ref = ob.tracer_2d(case, hord=8)
qref, dref = ob.remap_tracers(ref["q"], case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
ctx = TracerContext(49, 64, 9, case.metrics(), dtype=case.dtype)
out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
delp = np.zeros_like(case.dp1)
ctx.tracer_step(out["q"], out["dp1"], out["mfx"], out["mfy"], out["cx"], out["cy"], case.pe, case.ak, case.bk, case.ptop, delp, 8, kord, fill=True)
ctx.close()
sl = slice(3, -3)
for iq in range(9):
    s = np.abs(qref[:, iq][..., sl, sl]).max()
    d = np.abs(out["q"][:, iq][..., sl, sl] - qref[:, iq][..., sl, sl]) / s
    print(iq, "max", d.max(), "frac>1e-12", (d > 1e-12).mean(), ">1e-10", (d > 1e-10).mean(), ">1e-8", (d > 1e-8).mean(), ">1e-6", (d > 1e-6).mean(), ">1e-4", (d > 1e-4).mean(), ">1e-3", (d > 1e-3).mean())
# the same with the oracle remap fed by the oracle's advected field perturbed by 1 ulp (how sensitive is the REFERENCE algorithm?)
qp = ref["q"] * (1 + 2.2e-16 * np.sign(np.random.default_rng(0).standard_normal(ref["q"].shape)))
qref2, _ = ob.remap_tracers(qp, case.pe, case.ak, case.bk, case.ptop, kord, fill=True)
for iq in (2, 4):
    s = np.abs(qref[:, iq][..., sl, sl]).max()
    d = np.abs(qref2[:, iq][..., sl, sl] - qref[:, iq][..., sl, sl]) / s
    print("oracle vs 1-ulp-perturbed oracle input", iq, "max", d.max(), "frac>1e-12", (d > 1e-12).mean(), ">1e-6", (d > 1e-6).mean(), ">1e-3", (d > 1e-3).mean())
