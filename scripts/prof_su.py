"""S3 option set with streamline-upwind stabilisation in both loops on an N^3 box, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = syn.box_mesh((n,) * 3)
asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0)
asm.build_sparsity()
asm.set_fields(syn.standard_fields(mesh))
asm.set_scatter(abi.SCATTER_STRIP)
om = abi.common_momentum_opts(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND)
oa = abi.common_advdiff_opts(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND)
for i in range(3):
    asm.momentum_dev(om); m = asm.last_kernel_ms()
    asm.advdiff_dev(oa); a = asm.last_kernel_ms()
import os
print("%-24s su %d  momentum %.4f  tracer %.4f ms  paths %s" % (os.path.basename(os.environ.get("CGASM_LIB", "libcgasm.so")), n, m, a, asm.last_path()))
