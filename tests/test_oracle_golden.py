"""Pins the CPU oracle against every known answer the reference's own unit tests hold for
pieces of the hot path. CPU only."""
import json
import math
import os
import numpy as np
import pytest

from conftest import load_golden_mesh, GOLDEN
from fluidity_b200 import synthetic as syn


@pytest.mark.parametrize("dim", [2, 3])
def test_quadrature_integrates_monomials(orc, dim):
    # femtools/tests/test_quadrature.F90: sum_g w_g l(g,1)^p == p!/(p+dim)! for p <= degree,
    # compared with the reference's .fne. (|a-b| < max(eps,|a|eps), eps = 100*epsilon,
    # femtools/Unittest_tools.F90:184-203)
    l, w = orc.quadrature(dim)
    assert len(w) == (5 if dim == 3 else 4)
    for power in range(0, 4):
        got = float(np.sum(w * l[:, 0] ** power))
        want = math.factorial(power) / math.factorial(power + dim)
        tol = 100 * np.finfo(float).eps
        assert abs(got - want) < max(tol, abs(got) * tol), (dim, power, got - want)
    # the rule is symmetric: same for every barycentric coordinate
    for j in range(dim + 1):
        assert abs(np.sum(w * l[:, j] ** 3) - 6 / math.factorial(3 + dim)) < 1e-15
    assert abs(l.sum(axis=1) - 1).max() < 1e-15


@pytest.mark.parametrize("dim", [2, 3])
def test_shape_functions_integrate(orc, dim):
    # femtools/tests/test_shape_functions.F90 for degree 1: with f_i = x(node i)^p,
    # sum_i sum_g w_g f_i n(i,g) == p!/(p+dim)! and sum_i sum_g w_g f_i dn(i,g,1) ==
    # p*(p-1)!/(p-1+dim)!  (eps 1e-12). local_coords(i) of P1 node i = unit vector e_i.
    n, dn, w = orc.tables(dim)
    loc, ngi = dim + 1, len(w)
    N = n.reshape(ngi, loc).T              # N[i, g]
    DN = dn.reshape(dim, ngi, loc).transpose(2, 1, 0)  # DN[i, g, k]
    coords1 = np.array([1.0 if i == 0 else 0.0 for i in range(loc)])  # l(1) at node i
    for power in range(0, 2):
        f = coords1 ** power if power else np.ones(loc)
        got = sum(w[g] * f[i] * N[i, g] for i in range(loc) for g in range(ngi))
        want = math.factorial(power) / math.factorial(power + dim)
        assert abs(got - want) < 1e-12
        gotd = sum(w[g] * f[i] * DN[i, g, 0] for i in range(loc) for g in range(ngi))
        wantd = power * (math.factorial(power - 1) / math.factorial(power - 1 + dim)) if power else 0.0
        assert abs(gotd - wantd) < 1e-12
    # partition of unity / derivative sum
    assert np.abs(N.sum(axis=0) - 1).max() < 1e-15
    assert np.abs(DN.sum(axis=0)).max() == 0.0


def test_jacobian_embedded_triangle_area(orc):
    # femtools/tests/test_jacobian.F90: triangle (0,0,0),(2,0,0),(0,1,1) has sum(detwei) =
    # sqrt(2), tol 1e-10. The hot path only sees dim x dim Jacobians, so the same triangle is
    # expressed in its own plane: edges 2 and sqrt(2), right angle.
    X = np.array([[0.0, 0.0], [2.0, 0.0], [0.0, math.sqrt(2.0)]])
    ds, detwei, _ = orc.transform_to_physical(2, X)
    assert abs(detwei.sum() - math.sqrt(2.0)) < 1e-10
    # and a rotated copy: area is invariant
    c, s = math.cos(0.7), math.sin(0.7)
    R = np.array([[c, -s], [s, c]])
    ds2, detwei2, _ = orc.transform_to_physical(2, X @ R.T + 3.0)
    assert abs(detwei2.sum() - math.sqrt(2.0)) < 1e-10


def test_colouring_reference_mesh(orc):
    # femtools/tests/test_colouring.F90 on data/square-cavity-2d: no more colours than
    # max degree + 1, valid colouring, colour sets partition 1..n_elements (sum check).
    mesh = load_golden_mesh("square-cavity-2d")
    col, nc = orc.colour_elements(mesh)
    nd = mesh.ndglno.astype(np.int64) - 1
    # adjacency: elements sharing a node
    from collections import defaultdict
    n2e = defaultdict(list)
    for e, row in enumerate(nd):
        for v in row:
            n2e[v].append(e)
    maxdeg = 0
    for e, row in enumerate(nd):
        neigh = set()
        for v in row:
            neigh.update(n2e[v])
        maxdeg = max(maxdeg, len(neigh))  # row length incl. diagonal
        for e2 in neigh:
            if e2 != e:
                assert col[e2] != col[e]
    assert nc <= maxdeg + 1
    ptr, els = orc.colour_sets(col, nc)
    assert int(els.astype(np.int64).sum()) == mesh.n_elements * (mesh.n_elements + 1) // 2
    assert ptr[0] == 1 and ptr[-1] == mesh.n_elements + 1
    for c in range(nc):
        seg = els[ptr[c] - 1:ptr[c + 1] - 1]
        assert (np.diff(seg) > 0).all() and (col[seg - 1] == c + 1).all()


def test_block_addto_known_answer(orc):
    # femtools/tests/test_petsc_csr_matrix.F90: addto(A,1,1,(1..4),(1..4),vals) gives values
    # 1..16 and colm = 1..4 per row; adding again doubles them (1e-12).
    findrm = np.array([1, 5, 9, 13, 17], dtype=np.int32)
    colm = np.tile(np.arange(1, 5, dtype=np.int32), 4)
    vals = np.arange(1.0, 17.0).reshape(4, 4)  # row-major 1..16 == transpose(reshape(...))
    val = np.zeros(16)
    orc.block_addto(findrm, colm, [1, 2, 3, 4], [1, 2, 3, 4], vals, val)
    assert np.abs(val - np.arange(1.0, 17.0)).max() < 1e-12
    orc.block_addto(findrm, colm, [1, 2, 3, 4], [1, 2, 3, 4], vals, val)
    assert np.abs(val - 2 * np.arange(1.0, 17.0)).max() < 1e-12


@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "2d_square"])
def test_make_sparsity_reference_meshes(orc, name):
    # femtools/tests/test_make_sparsity.F90 is a smoke test; here the pattern is checked
    # against its definition: row i = ascending unique nodes sharing an element with i.
    mesh = load_golden_mesh(name)
    findrm, colm, centrm = orc.make_sparsity(mesh)
    nd = mesh.ndglno.astype(np.int64)
    rows = [set() for _ in range(mesh.n_nodes)]
    for e in nd:
        for i in e:
            rows[i - 1].update(e.tolist())
    assert findrm[0] == 1
    for i in range(mesh.n_nodes):
        got = colm[findrm[i] - 1:findrm[i + 1] - 1]
        assert got.tolist() == sorted(rows[i])
        assert colm[centrm[i] - 1] == i + 1
    assert findrm[-1] - 1 == sum(len(r) for r in rows)


def test_halo_fixture_trailing_receives(orc):
    # tests/meshconv_test/src/prectangle_{0,1}.halo: a real 2-rank decomposition. Checks the
    # conventions the halo update relies on: receives are numbered
    # after the n_private_nodes owned nodes, sends are owned nodes, and what rank p sends to
    # q has the length q expects to receive from p. Then drives orc.halo_copy with them.
    with open(os.path.join(GOLDEN, "prectangle_halos.json")) as f:
        H = json.load(f)
    for lvl in ("1", "2"):
        h0, h1 = H["0"]["levels"][lvl], H["1"]["levels"][lvl]
        assert len(h0["sends"]["1"]) == len(h1["receives"]["0"])
        assert len(h1["sends"]["0"]) == len(h0["receives"]["1"])
        for h, other in ((h0, "1"), (h1, "0")):
            assert min(h["receives"][other]) > h["n_private_nodes"]
            assert max(h["sends"][other]) <= h["n_private_nodes"]
    h0, h1 = H["0"]["levels"]["2"], H["1"]["levels"]["2"]
    n0 = max(h0["receives"]["1"])
    n1 = max(h1["receives"]["0"])
    f0 = np.arange(1.0, 2 * n0 + 1).reshape(n0, 2)
    f1 = -np.arange(1.0, 2 * n1 + 1).reshape(n1, 2)
    g0, g1 = f0.copy(), f1.copy()
    orc.halo_copy(2, f0, h0["sends"]["1"], g1, h1["receives"]["0"])
    orc.halo_copy(2, f1, h1["sends"]["0"], g0, h0["receives"]["1"])
    for k, (s, r) in enumerate(zip(h0["sends"]["1"], h1["receives"]["0"])):
        assert (g1[r - 1] == f0[s - 1]).all()
    for s, r in zip(h1["sends"]["0"], h0["receives"]["1"]):
        assert (g0[r - 1] == f1[s - 1]).all()
    assert (g0[:h0["n_private_nodes"]] == f0[:h0["n_private_nodes"]]).all()
