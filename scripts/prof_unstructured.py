"""One momentum + tracer assembly on an unstructured (Delaunay) or renumbered box mesh, for ncu.
usage: python scripts/prof_unstructured.py delaunay 300000 | shuffled 64 | box 64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

kind, n = sys.argv[1], int(sys.argv[2])
mesh = syn.delaunay_mesh(n) if kind == "delaunay" else syn.box_mesh((n,) * 3)
if kind == "shuffled":
    mesh = syn.shuffled(mesh)
asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0)
asm.build_sparsity()
asm.set_fields(syn.standard_fields(mesh))
asm.set_scatter(abi.SCATTER_STRIP)
om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
for i in range(3):
    asm.momentum_dev(om); m = asm.last_kernel_ms()
    asm.advdiff_dev(oa); a = asm.last_kernel_ms()
print(kind, n, "elements", mesh.n_elements, "momentum %.4f tracer %.4f ms" % (m, a), "G el/s %.2f" % (mesh.n_elements / (m + a) / 1e6), asm.plan_stats())
