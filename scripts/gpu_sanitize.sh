#!/bin/bash
# compute-sanitizer over the STRIP kernels (staged and per-entry), 2-D and 3-D, small meshes.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
for shape in ((9, 7, 6), (13, 11)):
    mesh = syn.shuffled(syn.box_mesh(shape), seed=3) if len(shape) == 3 else syn.box_mesh(shape)
    fs = syn.standard_fields(mesh)
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim), device=0)
    asm.build_sparsity(); asm.set_fields(fs); asm.set_scatter(abi.SCATTER_STRIP)
    for om in (abi.common_momentum_opts(), abi.common_momentum_opts(viscosity_shape=abi.TENSOR_FULL, have_gravity=0)):
        m = asm.momentum(om)
    for oa in (abi.common_advdiff_opts(), abi.common_advdiff_opts(lump_mass=1, diffusivity_shape=abi.TENSOR_FULL)):
        a = asm.advdiff(oa)
    print(shape, float(np.abs(m['big_m']).sum()), float(np.abs(a['matrix']).sum()))
    asm.close()
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool staged"; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -4
  echo "== $tool per-entry"; CGASM_STRIP_GLOBAL=1 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -4
done > gpurun_out/sanitize_strip.log 2>&1
cat gpurun_out/sanitize_strip.log
