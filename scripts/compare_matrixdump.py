#!/usr/bin/env python
"""Compare two PETSc binary `dump_matrix` files.

    python scripts/compare_matrixdump.py Velocity_1 ours_Velocity_1 [--rtol 1e-12] [--int64]

Each file holds MatView(A), VecView(b), VecView(x0) (femtools/Petsc_Tools.F90:1487-1504). One side
is the dump of a real Fluidity run (`solver/diagnostics/dump_matrix` in the flml,
femtools/Solvers.F90:1265-1272), the other is written by `formats.write_petsc_binary` from the
blocks `cgasm_momentum` / `cgasm_advdiff` return (`formats.blocks_to_petsc`, `formats.csr_to_petsc`).
Metric: max-norm relative error per matrix and per row, and per vector (1e-12 by default). Note that
the reference dump is taken at solve time: it contains the surface integrals and boundary conditions
added after the element loop, so a clean comparison needs a set-up without them (periodic, or all
boundaries unconstrained).
Exit status 0 = within tolerance."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidity_b200 import formats as fmt  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("reference")
    ap.add_argument("ours")
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--int64", action="store_true", help="files written by a 64-bit-index PETSc")
    a = ap.parse_args(argv)
    ref = fmt.read_petsc_binary(a.reference, a.int64)
    our = fmt.read_petsc_binary(a.ours, a.int64)
    report, ok = [], len(ref) == len(our)
    for k, (r, o) in enumerate(zip(ref, our)):
        if isinstance(r, fmt.PetscMat) != isinstance(o, fmt.PetscMat):
            report.append({"object": k, "ok": False, "reason": "matrix on one side, vector on the other"})
            ok = False
        elif isinstance(r, fmt.PetscMat):
            c = fmt.compare_petsc_mats(o, r, a.rtol)
            c.update(object=k, kind="matrix", rows=r.rows, nnz_reference=int(len(r.val)), nnz_ours=int(len(o.val)))
            report.append(c)
            ok &= c["ok"]
        else:
            scale = np.abs(r).max() if r.size else 0.0
            err = float(np.abs(o - r).max() / scale) if r.shape == o.shape and scale > 0 else float("inf") if r.shape != o.shape else 0.0
            report.append({"object": k, "kind": "vector", "n": int(r.size), "rel": err, "ok": bool(err <= a.rtol)})
            ok &= err <= a.rtol
    print(json.dumps({"ok": bool(ok), "objects": report}, indent=1))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
