#!/bin/bash
# Round 2, GPU call 15 (1 GPU): which case of the surface test broke with the degree-sorted rows; pipelined CMC expansion kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_surface_gpu.py tests/test_cmc.py tests/test_coo_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2_15_pytest.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2_15_debug.txt
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_surface_gpu as T
from fluidity_b200 import _abi as abi, synthetic as syn
from oracle import oracle as orc
mesh = T.meshes()["cube-parallel"]
fs = syn.standard_fields(mesh)
for scatter in (abi.SCATTER_ATOMIC, abi.SCATTER_GATHER, abi.SCATTER_STRIP):
    asm, sn, fe = T.make(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    for name, o in (("common", abi.common_advdiff_opts()), ("byparts_beta", abi.common_advdiff_opts(integrate_advection_by_parts=1, beta=0.25)),
                    ("byparts_nodiff", abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0)),
                    ("byparts_nodiff_theta0", abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0, theta=0.0))):
        ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
        for rep in range(3):
            asm.advdiff_dev(o)
            got = asm.advdiff_fetch()
            e = np.abs(got["matrix"] - ref["matrix"]).max() / np.abs(ref["matrix"]).max()
            print(scatter, name, rep, asm.last_path(), "err %.2e" % e, "bad entries", int((np.abs(got["matrix"] - ref["matrix"]) > 1e-9).sum()))
PY
timeout 600 python scripts/bench_cmc.py 2>&1 | tail -5 | tee gpurun_out/r2_15_cmc.txt
