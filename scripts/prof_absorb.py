"""backward_facing_step_3d option set (constant density, nodal vector absorption) on an N^3 box, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mesh = syn.box_mesh((n,) * 3)
asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0)
asm.build_sparsity()
asm.set_fields(syn.standard_fields(mesh))
asm.set_field(abi.F_DENSITY, np.ones(1), abi.FIELD_CONSTANT)
asm.set_scatter(abi.SCATTER_STRIP)
o = abi.common_momentum_opts(have_absorption=1, have_gravity=0)
for i in range(3):
    asm.momentum_dev(o)
print("absorb", n, asm.last_kernel_ms(), asm.last_path())
