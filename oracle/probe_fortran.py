#!/usr/bin/env python
"""Probe for a Fortran compiler (test infrastructure).  The reference implementation of the path is Fortran
(atmos_cubed_sphere/model/tp_core.F90, fv_tracer2d.F90, fv_mapz.F90, fv_fill.F90); with a compiler it could be built into
oracle/_ref against stub modules and pin the C++ restatement.  This image has none, so the oracle stays
"parity unpinned" (DESIGN.md section 5); the probe is run by __graft_entry__.build() so that the fact is re-established on every
box instead of assumed.  Prints one JSON line; exit status 0 either way."""
import json
import os
import shutil
import subprocess
import sys

CANDIDATES = ["gfortran", "flang", "flang-new", "nvfortran", "pgfortran", "ifort", "ifx", "lfortran", "f95", "f77"]


def probe():
    found = {}
    for c in CANDIDATES:
        p = shutil.which(c)
        if p:
            try:
                v = subprocess.run([p, "--version"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:1]
            except Exception:
                v = []
            found[c] = {"path": p, "version": v[0] if v else ""}
    # gcc may carry the Fortran front end without a driver on PATH
    try:
        out = subprocess.run(["gcc", "-print-prog-name=f951"], capture_output=True, text=True, timeout=20).stdout.strip()
        if out and os.path.isabs(out) and os.path.exists(out):
            found["f951"] = {"path": out, "version": "gcc Fortran front end"}
    except Exception:
        pass
    return {"fortran_compilers": found, "reference_sources_present": os.path.isdir("/root/reference/atmos_cubed_sphere/model"),
            "oracle_ref_buildable": bool(found) and os.path.isdir("/root/reference/atmos_cubed_sphere/model")}


if __name__ == "__main__":
    print(json.dumps(probe()))
    sys.exit(0)
