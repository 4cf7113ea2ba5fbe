"""P1-P1 pressure stabilisation (assemble_kmk_matrix, assemble/Momentum_CG.F90:2707-2766).

CPU: the oracle's restatement of simplex_tensor / edge_length_from_eigenvalue / the kt element loop /
mult_div_invscalar_div_T is pinned on (i) the reference's own known answers (error_measures/tests/
test_simplex_tensor.F90: the metric of data/triangle.1; test_simplex_tensor_edgelens.F90: the square root of the
metric maps every edge to unit length), (ii) numpy.linalg -- the LAPACK routines the reference itself calls (DGESV for
the metric's system, a symmetric eigensolver for its power) -- and (iii) an independent scipy assembly.
GPU: cgasm_kmk_dev against the oracle through the C ABI."""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden_mesh
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

TOL = 1e-12


def _meshes():
    return {"box3": syn.box_mesh((4, 3, 3), seed=2), "box2": syn.box_mesh((6, 5), seed=2),
            "cube-parallel": load_golden_mesh("cube-parallel"), "cavity": load_golden_mesh("square-cavity-2d"),
            "delaunay3": syn.delaunay_mesh(300, seed=5)}


def _metric_numpy(P):
    """The reference's system (Metric_tools.F90:875-896) solved by LAPACK through numpy."""
    loc, dim = P.shape
    pairs = [(k, l) for k in range(dim) for l in range(k, dim)]
    A = np.array([[(1.0 if k == l else 2.0) * (P[j] - P[i])[k] * (P[j] - P[i])[l] for k, l in pairs]
                  for i, j in itertools.combinations(range(loc), 2)])
    x = np.linalg.solve(A, np.ones(len(pairs)))
    M = np.zeros((dim, dim))
    for v, (k, l) in zip(x, pairs):
        M[k, l] = M[l, k] = v
    return M


def test_simplex_tensor_reference_known_answer(orc):
    # error_measures/tests/test_simplex_tensor.F90 on tests/data/triangle.1.msh: nodes (0, 0.5), (0, -0.5), (1, 0)
    m = orc.simplex_tensor(np.array([[0.0, 0.5], [0.0, -0.5], [1.0, 0.0]]))
    assert np.abs(m - np.array([[0.75, 0.0], [0.0, 1.0]])).max() < 1e-15


@pytest.mark.parametrize("name", ["box3", "box2", "cube-parallel", "cavity", "delaunay3"])
def test_metric_and_edge_lengths(orc, name):
    mesh = _meshes()[name]
    rng = np.random.default_rng(0)
    for e in rng.choice(mesh.n_elements, size=min(60, mesh.n_elements), replace=False):
        P = mesh.X[mesh.ndglno[e] - 1]
        M = orc.simplex_tensor(P)
        Mn = _metric_numpy(P)
        assert np.abs(M - Mn).max() <= 1e-12 * np.abs(Mn).max()
        # test_simplex_tensor_edgelens.F90: the metric to the power 1/2 maps every edge to length 1
        w, V = np.linalg.eigh(Mn)
        root = (V * np.sqrt(w)) @ V.T
        for i, j in itertools.combinations(range(mesh.loc), 2):
            assert abs(np.linalg.norm(root @ (P[i] - P[j])) - 1.0) < 1e-11
        H = orc.edge_length_from_metric(M)
        Hn = (V / np.sqrt(np.abs(w))) @ V.T
        assert np.abs(H - Hn).max() <= 1e-12 * np.abs(Hn).max()


def _kt_numpy(mesh):
    n, dim = mesh.n_nodes, mesh.dim
    rows, cols, vals = [], [], []
    ml = np.zeros(n)
    for nd in mesh.ndglno - 1:
        P = mesh.X[nd]
        E = (P[1:] - P[0]).T                      # columns = edges
        G = np.zeros((dim + 1, dim))
        G[1:] = np.linalg.inv(E)                  # rows = gradients of lambda_1..dim
        G[0] = -G[1:].sum(axis=0)
        vol = abs(np.linalg.det(E)) / (2.0 if dim == 2 else 6.0)
        w, V = np.linalg.eigh(_metric_numpy(P))
        H = (V / np.sqrt(np.abs(w))) @ V.T
        K = 0.5 * vol * G @ H @ G.T
        rows += list(np.repeat(nd, dim + 1))
        cols += list(np.tile(nd, dim + 1))
        vals += list(K.ravel())
        ml[nd] += vol / (dim + 1)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n)), ml


@pytest.mark.parametrize("name", ["box3", "box2", "cube-parallel", "cavity"])
def test_oracle_kmk_against_scipy(orc, name):
    mesh = _meshes()[name]
    findrm, colm, _ = orc.make_sparsity(mesh)
    findrm2, colm2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    theta = 0.7
    kmk, kt, ml = orc.assemble_kmk(mesh, findrm, colm, findrm2, colm2, theta)
    K, mln = _kt_numpy(mesh)
    ktn = np.asarray(K[np.repeat(np.arange(mesh.n_nodes), np.diff(findrm)), colm - 1]).ravel()
    assert np.abs(kt - ktn).max() <= TOL * np.abs(ktn).max()
    assert np.abs(ml - mln).max() <= TOL * mln.max()
    assert abs(ml.sum() - mln.sum()) <= TOL * mln.sum()
    ref = (K @ sp.diags(1.0 / (theta * mln)) @ K.T).tocsr()
    refv = np.asarray(ref[np.repeat(np.arange(mesh.n_nodes), np.diff(findrm2)), colm2 - 1]).ravel()
    assert np.abs(kmk - refv).max() <= TOL * np.abs(refv).max()
    # a stiffness matrix: constants are in the kernel of kt, hence of kmk
    assert np.abs(K @ np.ones(mesh.n_nodes)).max() <= 1e-12 * np.abs(ktn).max()


@pytest.mark.gpu
@pytest.mark.parametrize("theta", [1.0, 0.55])
@pytest.mark.parametrize("name", ["box3", "box2", "cube-parallel", "cavity", "delaunay3"])
def test_kmk_device_matches_the_oracle(orc, name, theta):
    mesh = _meshes()[name]
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim))
    asm.build_sparsity()
    findrm, colm, _ = asm.get_sparsity()
    asm.cmc_build_sparsity()
    findrm2, colm2 = asm.cmc_get_sparsity()
    ref, ref_kt, ref_ml = orc.assemble_kmk(mesh, findrm, colm, findrm2, colm2, theta)
    got, kt, ml = asm.kmk(theta, want_parts=True)
    assert np.abs(kt - ref_kt).max() <= TOL * np.abs(ref_kt).max()
    assert np.abs(ml - ref_ml).max() <= TOL * ref_ml.max()
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max()
    # adopted (caller's) second-order sparsity gives the same values
    asm.cmc_set_sparsity(findrm2, colm2)
    assert np.abs(asm.kmk(theta) - ref).max() <= TOL * np.abs(ref).max()
    asm.close()


@pytest.mark.gpu
def test_kmk_needs_the_second_order_sparsity():
    mesh = syn.box_mesh((2, 2, 2))
    asm = cgasm.Assembler(mesh, tables.p1_tables(3))
    asm.build_sparsity()
    with pytest.raises(cgasm.CgasmError) as e:
        asm.kmk(1.0)
    assert e.value.code == abi.ESTATE
    asm.close()
