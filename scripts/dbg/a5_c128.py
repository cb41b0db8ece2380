import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_binding as ob
from fv3atm_b200 import synthetic as sy
from fv3atm_b200.tracer import TracerContext
dt = sys.argv[1] if len(sys.argv) > 1 else "float32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
case = sy.make_case(n, 6, 9, dtype=dt, courant=0.7)
ctx = TracerContext(n + 1, 6, 9, case.metrics(), dtype=case.dtype)
out = {k: np.array(getattr(case, k), copy=True, order="C") for k in ("q", "dp1", "mfx", "mfy", "cx", "cy")}
nsplt, ksplt = ctx.tracer_2d(out["q"], out["dp1"], out["mfx"], out["mfy"], out["cx"], out["cy"], 8)
ctx.close()
ref = ob.tracer_2d(case, hord=8)
sl = slice(3, -3)
d = np.abs(out["q"][..., sl, sl].astype(np.float64) - ref["q"][..., sl, sl]).max(axis=(0, 2, 3, 4)) / np.abs(ref["q"][..., sl, sl]).max(axis=(0, 2, 3, 4))
print(dt, n, "nsplt", nsplt, "nd", d)
