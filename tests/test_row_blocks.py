"""CPU checks of the row blocks the row-owner kernels work on (host_mesh.cpp: morton_order, form_row_blocks)
through the host-only diagnostics entry point cgasm_row_blocks_host."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden_mesh
from fluidity_b200 import synthetic as syn, cgasm

BR = 128


def row_blocks(mesh, block_rows=BR):
    lib = cgasm.load()
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    X = np.ascontiguousarray(mesh.X, dtype=np.float64)
    cap = mesh.n_nodes  # never more blocks than nodes
    rows = np.zeros(cap * block_rows, dtype=np.int32)
    nb = C.c_int(0)
    scale = np.zeros(3)
    st = lib.cgasm_row_blocks_host(C.c_int(mesh.dim), C.c_int(mesh.n_nodes), C.c_int(mesh.n_elements),
                                   nd.ctypes.data_as(C.POINTER(C.c_int)), X.ctypes.data_as(C.POINTER(C.c_double)),
                                   C.c_int(block_rows), rows.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(cap), C.byref(nb),
                                   scale.ctypes.data_as(C.POINTER(C.c_double)))
    assert st == 0, lib.cgasm_last_error()
    return rows[:nb.value * block_rows].reshape(nb.value, block_rows), scale[:mesh.dim]


@pytest.mark.parametrize("name", ["box3", "box2", "shuffled3", "cube-parallel", "2d_square", "cell3"])
def test_blocks_partition_the_nodes(name):
    mesh = {"box3": lambda: syn.box_mesh((9, 7, 6)), "box2": lambda: syn.box_mesh((23, 17)),
            "shuffled3": lambda: syn.shuffled(syn.box_mesh((8, 8, 5)), seed=4),
            "cube-parallel": lambda: load_golden_mesh("cube-parallel"), "2d_square": lambda: load_golden_mesh("2d_square"),
            "cell3": lambda: syn.box_mesh((1, 1, 1))}[name]()
    rows, _ = row_blocks(mesh)
    real = rows[rows > 0]
    assert sorted(real.tolist()) == list(range(1, mesh.n_nodes + 1))  # every node in exactly one block
    for b in rows:  # padding only at the end of a block
        n = int((b > 0).sum())
        assert (b[:n] > 0).all() and (b[n:] == 0).all()
    assert rows.shape[0] <= 1.15 * (-(-mesh.n_nodes // BR)) + 1


def test_lattice_matches_a_jittered_box_mesh():
    """The lattice spacing comes from a sample of element edges. A strided sample whose stride shares a factor
    with the 6 Kuhn tets per cube saw ONE tet shape and produced an anisotropic lattice (139 x 108 x 139
    instead of 128^3 on S3/8): blocks were not bricks. The sample is hashed now; the sampling only kicks in
    above 4 M elements, so this pins the small-mesh path and the brick structure."""
    c = 16
    mesh = syn.box_mesh((c, c, c), jitter=0.1)
    rows, scale = row_blocks(mesh)
    assert np.allclose(scale, c)  # cells per unit length on the unit cube
    # full blocks are 8 x 4 x 4 bricks of lattice points: extents of their nodes' lattice coordinates
    q = np.rint(mesh.X * c).astype(int)
    n_bricks = 0
    for b in rows:
        if (b > 0).sum() != BR:
            continue
        ext = np.sort(q[b - 1].max(axis=0) - q[b - 1].min(axis=0) + 1)
        n_bricks += ext.tolist() == [4, 4, 8]  # (merged boundary slabs are full blocks too, but not bricks)
    assert n_bricks == (c // 8) * (c // 4) * (c // 4)  # every completely filled brick is one block


def test_merged_boundary_blocks_stay_compact():
    """Partly filled bricks on the boundary are merged only while the block's distinct neighbour set stays
    near a full brick's (what the staged kernels keep in shared memory)."""
    c = 24
    mesh = syn.box_mesh((c, c, c), jitter=0.1)
    rows, _ = row_blocks(mesh)
    nd = mesh.ndglno
    nbrs = [set() for _ in range(mesh.n_nodes + 1)]
    for e in nd:
        for v in e:
            nbrs[v].update(int(x) for x in e)
    touched = []
    for b in rows:
        s = set()
        for v in b[b > 0]:
            s |= nbrs[int(v)]
        touched.append(len(s))
    full = [t for t, b in zip(touched, rows) if (b > 0).sum() == BR]
    assert max(touched) <= 1.15 * np.median(full) + 1


def test_lattice_from_the_hashed_edge_sample():
    """Above 4 M elements the spacing is estimated from a hashed sample of the elements: it must still see all
    six Kuhn shapes (the strided sample it replaces did not) and recover the lattice exactly."""
    c = 90  # 6 * 90^3 = 4.37 M tets
    mesh = syn.box_mesh((c, c, c), jitter=0.1)
    rows, scale = row_blocks(mesh)
    assert np.allclose(scale, c)
    assert rows.shape[0] <= 1.15 * (-(-mesh.n_nodes // BR)) + 1


def plan_stats(mesh):
    lib = cgasm.load()
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    X = np.ascontiguousarray(mesh.X, dtype=np.float64)
    st = (C.c_double * 5)()
    rc = lib.cgasm_plan_host_stats(C.c_int(mesh.dim), C.c_int(mesh.n_nodes), C.c_int(mesh.n_elements),
                                   nd.ctypes.data_as(C.POINTER(C.c_int)), X.ctypes.data_as(C.POINTER(C.c_double)), st)
    assert rc == 0, lib.cgasm_last_error()
    return dict(zip(("entries_per_pair", "max_nodes_per_block", "lds128_wavefronts", "walk_ratio", "compute_ratio"), st))


def test_renumbering_does_not_change_the_plan_shape():
    """The staged STRIP plan is built on geometric keys (rows of a block, the block's node list, the labels of the strip
    builder): a randomly renumbered structured mesh gets the plan of the lexicographically numbered one -- same strip
    lengths, the same (near conflict-free) shared-memory access pattern, no extra divergence. With id-ordered labels the
    renumbered mesh read its staged records with 2.5 wavefronts per quarter-warp (2.1x measured on the B200)."""
    box = syn.box_mesh((40, 40, 40))
    a, b = plan_stats(box), plan_stats(syn.shuffled(box, seed=3))
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-3 * max(1.0, abs(a[k])), (k, a[k], b[k])  # (ties between equal keys go by id)
    assert a["lds128_wavefronts"] < 1.4 and a["compute_ratio"] < 1.1


def test_unstructured_plan_stats_are_sane():
    st = plan_stats(syn.delaunay_mesh(6000, seed=2))
    assert 1.0 <= st["entries_per_pair"] < 1.6 and st["max_nodes_per_block"] <= 1024
    assert 1.0 <= st["walk_ratio"] < 1.6 and 1.0 <= st["compute_ratio"] < 2.0 and 1.0 <= st["lds128_wavefronts"] <= 8.0


def test_bank_groups_are_searched_on_unstructured_meshes_only(monkeypatch):
    """Delaunay mesh: no two links are congruent, the 8 lanes of a quarter-warp read 8 unrelated staged records, 2.5
    wavefronts per LDS.128 with the block's nodes in any fixed order. assign_banks (strip_plan.cpp) moves nodes between
    the bank groups (local index mod 8): fewer wavefronts, the same strips. A structured mesh keeps its geometric order
    (near conflict-free, contiguous staging copies)."""
    mesh = syn.delaunay_mesh(20000, seed=4)
    on = plan_stats(mesh)
    monkeypatch.setenv("CGASM_STRIP_BANKS", "0")
    off = plan_stats(mesh)
    monkeypatch.delenv("CGASM_STRIP_BANKS")
    assert on["lds128_wavefronts"] < 0.9 * off["lds128_wavefronts"], (on, off)
    for k in ("entries_per_pair", "walk_ratio", "compute_ratio"):
        assert on[k] == off[k]
    box = syn.box_mesh((24, 24, 24))
    a = plan_stats(box)
    monkeypatch.setenv("CGASM_STRIP_BANKS", "0")
    b = plan_stats(box)
    assert a == b
