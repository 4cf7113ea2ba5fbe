#!/usr/bin/env python
"""Times the kernels next to the hot path on one B200 (not the headline bench): the lumped-mass pressure matrix
C M_L^-1 C^T (cgasm_cmc_dev) and the surface loops, on an N^3 x 6 Kuhn box. CUDA events on the handle's stream
(cgasm_last_kernel_ms) for CMC, wall clock around a synchronize for the O(boundary) surface calls.
    python scripts/bench_cmc.py [N=96] > gpurun_out/bench_cmc.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    mesh = syn.box_mesh((n, n, n))
    fs = syn.standard_fields(mesh)
    asm = cgasm.Assembler(mesh, tables.p1_tables(3))
    nnz = asm.build_sparsity()
    asm.set_fields(fs)
    t0 = time.time()
    nnz2 = asm.cmc_build_sparsity()
    t_pattern = time.time() - t0
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1)
    asm.momentum_dev(o)
    ms = []
    for _ in range(5):
        asm.cmc_dev()
        asm.synchronize()
        ms.append(asm.last_kernel_ms())
    cmc_ms = float(np.median(ms[1:]))
    # compulsory traffic: first-order pattern + dim value blocks + inverse mass read once, second-order colm read and
    # values written once
    bytes_alg = 4 * (mesh.n_nodes + 1) * 2 + 4 * nnz + 8 * 3 * nnz + 8 * 3 * mesh.n_nodes + (4 + 8) * nnz2
    sn, fe = syn.boundary_faces(mesh)
    asm.set_surface(sn, fe, tables.p1_face_tables(3))
    oa = abi.common_advdiff_opts(integrate_advection_by_parts=1)
    asm.advdiff_dev(oa)
    asm.synchronize()
    bt = np.zeros(len(fe), dtype=np.int32)
    t0 = time.time()
    for _ in range(5):
        asm.advdiff_surface_dev(oa, bt)
    asm.synchronize()
    surf_ms = (time.time() - t0) / 5 * 1e3
    print(json.dumps({"mesh": "%d^3 x 6 Kuhn tets" % n, "n_nodes": mesh.n_nodes, "n_elements": mesh.n_elements, "nnz": nnz,
                      "nnz_second_order": nnz2, "second_order_pattern_host_s": round(t_pattern, 3), "cmc_kernel_ms": cmc_ms,
                      "cmc_entries_per_s": nnz2 / (cmc_ms * 1e-3), "cmc_algorithmic_GBs": bytes_alg / (cmc_ms * 1e-3) / 1e9,
                      "n_boundary_faces": int(len(fe)), "advdiff_surface_call_ms_incl_upload": surf_ms}))


if __name__ == "__main__":
    main()
