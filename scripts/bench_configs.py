"""Device time of the two element loops for the option sets of the reference's example configs
(SURVEY.md section 0 / 8(a) "Branch taken by each config"), on synthetic meshes, STRIP vs GATHER variant.

    python scripts/bench_configs.py [--cells3 128] [--cells2 2048]

Not the headline bench (bench.py): the example meshes are not checked in, so these are the examples' OPTION
SETS on the S3 / S2 meshes. Writes gpurun_out/bench_configs.json."""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells3", type=int, default=128)
    ap.add_argument("--cells2", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

    cm, ca = abi.common_momentum_opts, abi.common_advdiff_opts
    cases = [
        # name, dim, momentum opts, tracer opts, field tweaks
        ("driven_cavity (2-D, nodal density, constant isotropic viscosity, no gravity)", 2, cm(have_gravity=0), ca(), None),
        ("lock_exchange (2-D, Boussinesq, nodal buoyancy, gravity)", 2, cm(), ca(), "const_density"),
        ("backward_facing_step_3d (3-D, constant density, nodal vector absorption)", 3, cm(have_absorption=1, have_gravity=0),
         ca(), "const_density"),
        ("flow_past_sphere_Re100 (3-D, constant anisotropic viscosity)", 3, cm(viscosity_shape=abi.TENSOR_FULL, have_gravity=0),
         ca(), "aniso"),
        ("S3 / S2 common option set", 3, cm(), ca(), None),
    ]
    meshes = {}
    out = []
    for name, dim, om, oa, tweak in cases:
        if dim not in meshes:
            c = args.cells3 if dim == 3 else args.cells2
            mesh = syn.box_mesh((c,) * dim)
            meshes[dim] = (mesh, syn.standard_fields(mesh))
        mesh, fs0 = meshes[dim]
        asm = cgasm.Assembler(mesh, tables.p1_tables(dim), device=0)
        asm.build_sparsity()
        asm.set_fields(fs0)
        if tweak == "const_density":
            asm.set_field(abi.F_DENSITY, np.ones(1), abi.FIELD_CONSTANT)
        if tweak == "aniso":
            asm.set_field(abi.F_VISCOSITY, syn.aniso_tensor(dim), abi.FIELD_CONSTANT)
        rec = dict(config=name, dim=dim, elements=mesh.n_elements)
        for label, variant in (("strip", abi.SCATTER_STRIP), ("gather", abi.SCATTER_GATHER)):
            asm.set_scatter(variant)
            mom, adv = [], []
            for i in range(args.reps + 2):
                asm.momentum_dev(om)
                m = asm.last_kernel_ms()
                asm.advdiff_dev(oa)
                a = asm.last_kernel_ms()
                if i >= 2:
                    mom.append(m)
                    adv.append(a)
            rec[label] = dict(momentum_ms=statistics.median(mom), tracer_ms=statistics.median(adv),
                              gel_s=mesh.n_elements / ((statistics.median(mom) + statistics.median(adv)) * 1e-3) / 1e9)
        out.append(rec)
        print("%-88s %9d el  strip %.3f + %.3f ms = %.2f G el/s   gather %.3f + %.3f ms = %.2f G el/s" %
              (name, mesh.n_elements, rec["strip"]["momentum_ms"], rec["strip"]["tracer_ms"], rec["strip"]["gel_s"],
               rec["gather"]["momentum_ms"], rec["gather"]["tracer_ms"], rec["gather"]["gel_s"]), flush=True)
        asm.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
