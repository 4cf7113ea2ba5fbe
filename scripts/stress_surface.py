"""Stress loop for the rare failure of the surface tests on cube-parallel + GATHER: fresh handles, the test's calls, results
compared with the oracle and with the first iteration. usage: python scripts/stress_surface.py [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
from oracle import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mesh = load_golden_mesh("cube-parallel"); dim = 3
fs = syn.standard_fields(mesh)
sn, fe = syn.boundary_faces(mesh)
rng = np.random.default_rng(4); nf = len(fe)
vt = np.zeros((nf, dim), dtype=np.int32); vt[rng.random(nf) < 0.5] = abi.VBC_FREE_SURFACE
o = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=0.6, lump_mass=1, integrate_advection_by_parts=1)
ref = first = None
bad = 0
keep = []
for it in range(n):
    asm = cgasm.Assembler(mesh, tables.p1_tables(3)); asm.build_sparsity(); asm.set_fields(fs)
    asm.set_scatter(getattr(abi, 'SCATTER_' + os.environ.get('STRESS_SCATTER', 'GATHER'))); asm.set_surface(sn, fe, tables.p1_face_tables(3))
    if ref is None:
        findrm, colm, _ = asm.get_sparsity()
        ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
        orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, vt, np.zeros((nf, dim, dim)), ref["big_m"], ref["rhs"], masslump=ref["masslump"])
    asm.momentum_dev(o)
    before = asm.momentum_fetch()
    asm.momentum_surface_dev(o, vt)
    got = asm.momentum_fetch()
    errs = [rel_err(got["big_m"][d], ref["big_m"][d]) for d in range(dim)] + [rel_err(got["rhs"], ref["rhs"])]
    if first is None:
        first = before
    same_before = all(np.array_equal(before[k], first[k]) for k in ("big_m", "rhs", "masslump"))
    if max(errs) > 1e-12 or not same_before:
        bad += 1
        d = int(np.argmax(errs[:dim]))
        diff = np.abs(got["big_m"][d] - ref["big_m"][d]); w = np.nonzero(diff > 1e-10 * np.abs(ref["big_m"][d]).max())[0]
        rows = np.searchsorted(findrm - 1, w, side="right") - 1
        print("iteration %d: errs %s  element-loop result %s the first iteration's; block %d: %d entries off in rows %s" % (
            it, ["%.2e" % e for e in errs], "==" if same_before else "!=", d, len(w), np.unique(rows)[:12]), flush=True)
    if it % 3 == 0:
        keep.append(asm)          # leave some handles open, as the tests do
    else:
        asm.close()
print("%d of %d iterations off (scatter %s)" % (bad, n, os.environ.get("STRESS_SCATTER", "GATHER")))
