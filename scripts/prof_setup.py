"""Wall-clock of the library set-up phases (create, sparsity, fields, scatter plans) on a box mesh, the same mesh randomly
renumbered, and a Delaunay mesh. usage: python scripts/prof_setup.py [cells [delaunay_points]]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 0
box = syn.box_mesh((cells,) * 3)
cases = [("box", box), ("renumbered", syn.shuffled(box))]
if npts:
    cases.append(("delaunay", syn.delaunay_mesh(npts)))
for name, mesh in cases:
    fs = syn.standard_fields(mesh)
    t = [time.perf_counter()]
    asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0); t.append(time.perf_counter())
    asm.build_sparsity(); t.append(time.perf_counter())
    asm.set_fields(fs); t.append(time.perf_counter())
    asm.set_scatter(abi.SCATTER_STRIP); t.append(time.perf_counter())
    asm.momentum_dev(abi.common_momentum_opts()); asm.synchronize(); t.append(time.perf_counter())
    d = [t[i + 1] - t[i] for i in range(len(t) - 1)]
    print("%-11s %9d tets: create %.2f  sparsity %.2f  fields %.2f  scatter plans %.2f  first assembly %.2f  = %.2f s" % (
        name, mesh.n_elements, *d, sum(d)), flush=True)
    asm.close()
