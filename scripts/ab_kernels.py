"""A/B of library builds on one GPU: kernel times (CUDA events, cgasm_last_kernel_ms) of the S3 option set.
usage: CGASM_LIB=ab/libcgasm_X.so python scripts/ab_kernels.py [cells]  -> one line"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "s3"   # s3 | 2d (cells^2 triangles) | abs (tracer with absorption + source)
mesh = syn.box_mesh((cells,) * (2 if mode == "2d" else 3))
asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim), device=0)
asm.build_sparsity()
asm.set_fields(syn.standard_fields(mesh))
asm.set_scatter(abi.SCATTER_STRIP)
om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
if mode == "abs":
    oa = abi.common_advdiff_opts(have_absorption=1, have_source=1)
mom, adv, fus = [], [], []
for i in range(9):
    asm.momentum_dev(om); m = asm.last_kernel_ms()
    asm.advdiff_dev(oa); a = asm.last_kernel_ms()
    asm.momentum_advdiff_dev(om, oa); f = asm.last_kernel_ms()
    if i >= 2:
        mom.append(m); adv.append(a); fus.append(f)
chk = float(np.abs(asm.momentum_fetch()["rhs"]).sum()) if hasattr(asm, "momentum_fetch") else 0.0
print("%-28s %s cells %d  momentum %.4f  tracer %.4f  fused %.4f ms   checksum %.12e" % (
    os.path.basename(os.environ.get("CGASM_LIB", "libcgasm.so")), mode, cells, statistics.median(mom), statistics.median(adv),
    statistics.median(fus), chk))
