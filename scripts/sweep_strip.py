"""Times the STRIP kernels under several tuning switches in ONE process: the mesh and the plans are built
once, the switches are environment variables libcgasm reads at every launch.

    python scripts/sweep_strip.py --cells 256 --reps 10 [--configs name=K1=V1,K2=V2 ...]

Prints the median device time (cgasm_last_kernel_ms, CUDA events on the handle's stream) of the momentum and
tracer kernels per configuration and writes gpurun_out/sweep_strip_<cells>.json. Tuning aid, not a bench.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = [
    "default=",
    "m_mb4_pf0=CGASM_STRIP_MINB=4,CGASM_STRIP_PF=0",
    "m_mb3_pf1=CGASM_STRIP_MINB=3,CGASM_STRIP_PF=1",
    "a_mb4_pf0=CGASM_STRIP_MINB_ADV=4,CGASM_STRIP_PF_ADV=0",
    "a_mb3_pf1=CGASM_STRIP_MINB_ADV=3,CGASM_STRIP_PF_ADV=1",
    "global=CGASM_STRIP_GLOBAL=1",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--configs", nargs="*", default=DEFAULT)
    ap.add_argument("--shuffle", action="store_true", help="random node and element numbering (synthetic.shuffled)")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables, partition as part

    c = args.cells
    import time
    lp = part.slab_partition((c, c, c), 1, 0)
    mesh = lp.mesh
    gnode = lp.global_node
    if args.shuffle:
        mesh = syn.shuffled(mesh)
        gnode = np.arange(mesh.n_nodes)
    F = part.global_nodal_fields(3, mesh.X, gnode)
    t0 = time.perf_counter()
    asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=0)
    asm.build_sparsity()
    g = np.zeros((1, 3))
    g[0, 2] = -1.0
    asm.set_field(abi.F_GRAVITY, g, abi.FIELD_CONSTANT)
    asm.set_field(abi.F_VISCOSITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_T_DIFFUSIVITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    for slot, a in [(abi.F_NU, F["nu"]), (abi.F_OLDU, F["oldu"]), (abi.F_DENSITY, F["density"]),
                    (abi.F_BUOYANCY, F["buoyancy"]), (abi.F_T, F["t"])]:
        asm.set_field(slot, a)
    asm.set_scatter(abi.SCATTER_STRIP)
    print("library set-up (create + sparsity + fields + strip plans): %.2f s" % (time.perf_counter() - t0), flush=True)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    n_el = mesh.n_elements
    touched = set()
    out = []
    for spec in args.configs:
        name, _, kv = spec.partition("=")
        for k in touched:
            os.environ.pop(k, None)
        for item in filter(None, kv.split(",")):
            k, _, v = item.partition("=")
            os.environ[k] = v
            touched.add(k)
        mom, adv = [], []
        for i in range(args.reps + 2):
            asm.momentum_dev(om)
            m = asm.last_kernel_ms()
            asm.advdiff_dev(oa)
            a = asm.last_kernel_ms()
            if i >= 2:
                mom.append(m)
                adv.append(a)
        rec = dict(name=name, env=kv, mom_ms=statistics.median(mom), adv_ms=statistics.median(adv),
                   mom_min=min(mom), adv_min=min(adv))
        rec["gel_s"] = n_el / ((rec["mom_ms"] + rec["adv_ms"]) * 1e-3) / 1e9
        out.append(rec)
        print("%-22s mom %.3f (min %.3f)  tracer %.3f (min %.3f)  %.2f G el/s" %
              (name, rec["mom_ms"], rec["mom_min"], rec["adv_ms"], rec["adv_min"], rec["gel_s"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_strip_%d%s.json" % (c, args.tag)), "w") as f:
        json.dump(out, f, indent=1)
    asm.close()


if __name__ == "__main__":
    main()
