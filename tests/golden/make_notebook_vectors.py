#!/usr/bin/env python
"""Generate golden vectors for the 1-D PPM operator from the REFERENCE's own executable documentation.

Source: /root/reference/atmos_cubed_sphere/docs/examples/tp_core.ipynb -- a self-contained numpy restatement of
xppm (atmos_cubed_sphere/model/tp_core.F90:332-704) on a periodic 1-D domain for hord 5, 6, 8, 10 and the
positive-definite option, written by the reference's authors.  It is the only runnable artefact of this path the
reference ships (its tests/ directory holds one unrelated namelist test).

This script does NOT copy the notebook: it loads the .ipynb at generation time, takes the code cells that define the
constants / periodic index arrays / initial conditions and the body of the integration loop (everything between
`qprev = q` and `time += dt`, i.e. the numerical part without plotting), executes them headless with the option
variables set per case, and records for every step the inputs (q, c) and the outputs (upwind flux value before it is
multiplied by c, and the updated q).  The vectors are committed as tests/golden/notebook_xppm.npz so that the tests
run where /root/reference does not exist.

    python tests/golden/make_notebook_vectors.py [path/to/tp_core.ipynb]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_NB = "/root/reference/atmos_cubed_sphere/docs/examples/tp_core.ipynb"


def load_cells(path):
    nb = json.load(open(path))
    return ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]


def find(cells, marker):
    hit = [c for c in cells if marker in c]
    assert len(hit) >= 1, marker
    return hit[0]


def loop_body(cells):
    """Numerical body of the `while time < tend:` loop: from `qprev = q` up to (excluding) `time += dt`."""
    src = find(cells, "#begin xppm").splitlines()
    i0 = next(i for i, l in enumerate(src) if l.strip() == "qprev = q")
    i1 = next(i for i, l in enumerate(src) if l.strip() == "time += dt")
    body = src[i0:i1]
    indent = len(body[0]) - len(body[0].lstrip())
    return "\n".join(l[indent:] if l.strip() else "" for l in body)


def run_case(cells, body, ord_, PD, tracer_type, courant, nsteps, nx=40, lim_fac=1.0):
    ns = {"np": np}
    # grid options of cell #2 (only the lines the numerics need)
    ns.update(dict(ord=ord_, PD=PD, dt=1, courant=courant, tracer_type=tracer_type, nx=nx, dx=1.0, lim_fac=lim_fac))
    exec("L = nx*dx\ndxa = dx*np.ones(nx)\nxi = np.concatenate((np.array([0]), np.cumsum(dxa)))\n"
         "xc = 0.5*(xi[1:]+xi[:-1])\nc0 = courant*dx/dt\n", ns)
    exec(find(cells, "#Constants"), ns)                      # c1..c3, p1, p2, r12, r3
    exec(find(cells, "#Define indices with periodicity"), ns)  # ix, ixp1, ...
    exec(find(cells, "#Define analytic profiles"), ns)       # tracer_* functions
    init = {0: "tracer_gaussian", 1: "tracer_tophat", 2: "tracer_2dx_tophat"}[tracer_type]
    ns["q"] = ns[init](ns["xc"])
    ns["c"] = ns["c0"] * np.ones(nx + 1)
    code = compile(body, "tp_core.ipynb:loop", "exec")
    qin, flux, qout = [], [], []
    for _ in range(nsteps):
        qin.append(np.array(ns["q"], dtype=np.float64))
        exec(code, ns)
        # the notebook's `flux` is (upwind value) * c at the end of the body
        flux.append(np.array(ns["flux"], dtype=np.float64))
        qout.append(np.array(ns["q"], dtype=np.float64))
    return np.stack(qin), np.stack(flux), np.stack(qout)


CASES = [
    # (name, ord, PD, tracer_type, courant, steps)
    ("ord8_gauss", 8, False, 0, 0.8, 12),
    ("ord8_tophat", 8, False, 1, 0.8, 12),
    ("ord8_2dx", 8, False, 2, 0.35, 12),
    ("ord8_gauss_neg", 8, False, 0, -0.6, 12),
    ("ord10_gauss", 10, False, 0, 0.8, 12),
    ("ord10_tophat", 10, False, 1, 0.8, 12),
    ("ord10_2dx", 10, False, 2, 0.35, 12),
    ("ord10_tophat_neg", 10, False, 1, -0.45, 12),
    ("ord5_gauss", 5, False, 0, 0.8, 12),
    ("ord5_tophat", 5, False, 1, 0.8, 12),
    ("ord5_2dx", 5, False, 2, 0.35, 12),
    ("ord6_gauss", 6, False, 0, 0.8, 12),
    ("ord6_tophat", 6, False, 1, 0.5, 12),
    ("ord5pd_tophat", 5, True, 1, 0.8, 12),
    ("ord5pd_2dx", 5, True, 2, 0.35, 12),
]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_NB
    cells = load_cells(path)
    body = loop_body(cells)
    out = {}
    for name, ord_, PD, tt, cour, steps in CASES:
        qin, flux, qout = run_case(cells, body, ord_, PD, tt, cour, steps)
        out[f"{name}__qin"] = qin
        out[f"{name}__flux_times_c"] = flux
        out[f"{name}__qout"] = qout
        out[f"{name}__meta"] = np.array([ord_, int(PD), tt, cour, steps], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "notebook_xppm.npz"), **out)
    print(f"wrote {len(CASES)} cases to tests/golden/notebook_xppm.npz")


if __name__ == "__main__":
    main()
