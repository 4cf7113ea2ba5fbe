"""A/B on one GPU of the occupancy classes of the staged STRIP kernels on an unstructured (Delaunay) mesh: the same
handle with CGASM_STRIP_NOCLASSES set and unset. usage: python scripts/ab_classes.py [points [repetitions]]"""
import os, sys, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8   # 3 under compute-sanitizer
t0 = time.perf_counter()
mesh = syn.delaunay_mesh(npts)
print("Delaunay mesh: %d nodes, %d tets, %.1f s" % (mesh.n_nodes, mesh.n_elements, time.perf_counter() - t0), flush=True)
asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0)
asm.build_sparsity()
fs = syn.standard_fields(mesh)
asm.set_fields(fs)
asm.set_scatter(abi.SCATTER_STRIP)
print("plan", asm.plan_stats())
sets = {"S3": (abi.common_momentum_opts(), abi.common_advdiff_opts()),
        "S3 + tracer absorption/source": (abi.common_momentum_opts(), abi.common_advdiff_opts(have_absorption=1, have_source=1))}
ref = {}
for rep in range(2):
    for off in (True, False):
        if off:
            os.environ["CGASM_STRIP_NOCLASSES"] = "1"
        else:
            os.environ.pop("CGASM_STRIP_NOCLASSES", None)
        for name, (om, oa) in sets.items():
            mom, adv = [], []
            for i in range(reps):
                asm.momentum_dev(om); m = asm.last_kernel_ms()
                asm.advdiff_dev(oa); a = asm.last_kernel_ms()
                if i >= min(2, reps - 1):
                    mom.append(m); adv.append(a)
            got = asm.momentum_fetch(); ga = asm.advdiff_fetch()
            chk = (float(np.abs(got["big_m"]).sum()), float(np.abs(got["rhs"]).sum()), float(np.abs(ga["matrix"]).sum()), float(np.abs(ga["rhs"]).sum()))
            if name in ref:
                assert chk == ref[name], (chk, ref[name])  # the classes change where a block runs, not what it computes
            ref[name] = chk
            print("%-32s classes %-3s momentum %.4f  tracer %.4f ms  = %.2f G el/s   launches %d" % (
                name, "off" if off else "on", statistics.median(mom), statistics.median(adv),
                mesh.n_elements / (statistics.median(mom) + statistics.median(adv)) / 1e6, asm.launch_count()), flush=True)
