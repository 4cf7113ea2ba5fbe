"""Device hand-off to PETSc: cgasm_coo_pattern_dev / cgasm_coo_values_dev against the matrix the
reference's petsc_csr_matrix insertion would have built (formats.blocks_to_petsc: Sparse_Tools_Petsc.F90:848-879 with the
serial numbering of Petsc_Tools.F90:184-199), and against the masking of non-owned rows (:220-227)."""
import numpy as np
import pytest
import scipy.sparse as sp

from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables, formats as fmt, partition as part

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _asm(mesh, fs):
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim))
    asm.build_sparsity()
    asm.set_fields(fs)
    asm.set_scatter(abi.SCATTER_STRIP)
    return asm


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("group_size", [1, "dim"])
def test_coo_triplets_build_the_reference_petsc_matrix(orc, dim, group_size):
    mesh = syn.box_mesh((5, 4, 3)[:dim], seed=41)
    fs = syn.standard_fields(mesh)
    gs = dim if group_size == "dim" else 1
    asm = _asm(mesh, fs)
    om, oa = abi.common_momentum_opts(have_absorption=1), abi.common_advdiff_opts()
    findrm, colm, _ = asm.get_sparsity()
    nn = mesh.n_nodes
    # momentum: dim diagonal blocks, absorption makes them differ
    num = fmt.petsc_row_numbering(nn, dim, gs)
    ncoo = asm.coo_pattern(0, num)
    assert ncoo == dim * len(colm)
    asm.momentum_dev(om)
    i, j, v = asm.coo_fetch(0, ncoo)
    A = sp.coo_matrix((v, (i, j)), shape=(nn * dim, nn * dim)).tocsr()
    ref = orc.assemble_momentum(mesh, fs, om, findrm, colm)
    want = fmt.blocks_to_petsc(findrm, colm, ref["big_m"], nn, group_size=gs)
    got = fmt.PetscMat(nn * dim, nn * dim, A.indptr, A.indices, A.data)
    rep = fmt.compare_petsc_mats(got, want, rtol=TOL)
    assert rep["ok"], rep
    # a second assembly with other options: same pattern, new values, no new pattern call
    asm.momentum_dev(abi.common_momentum_opts())
    _, _, v2 = asm.coo_fetch(0, ncoo)
    got2 = asm.momentum_fetch()
    assert (v2 == got2["big_m"].ravel()).all()       # uncompacted values ARE the result buffer, [block][entry]
    # tracer
    n1 = asm.coo_pattern(1, fmt.petsc_row_numbering(nn, 1))
    asm.advdiff_dev(oa)
    i, j, v = asm.coo_fetch(1, n1)
    A = sp.coo_matrix((v, (i, j)), shape=(nn, nn)).tocsr()
    want = fmt.csr_to_petsc(findrm, colm, orc.assemble_advdiff(mesh, fs, oa, findrm, colm)["matrix"], nn)
    rep = fmt.compare_petsc_mats(fmt.PetscMat(nn, nn, A.indptr, A.indices, A.data), want, rtol=TOL)
    assert rep["ok"], rep


def test_coo_drops_the_rows_of_nodes_the_process_does_not_own(orc):
    """A partition's matrix: rows of halo nodes are masked with -1 in the row numbering (Sparse_Tools_Petsc.F90:220-227).
    Uncompacted they stay in the list with a negative row (PETSc ignores them); compacted they are gone. The owned rows
    of the two partitions together are the global matrix."""
    cells = (4, 4, 6)
    whole = part.block_partition(cells, (1, 1, 1), 0)
    gm = whole.mesh
    GF = part.global_nodal_fields(3, gm.X, whole.global_node)
    nn_glob = gm.n_nodes

    def fields(mesh, F):
        fs = syn.standard_fields(mesh)
        for slot, name in ((abi.F_NU, "nu"), (abi.F_OLDU, "oldu"), (abi.F_DENSITY, "density"), (abi.F_BUOYANCY, "buoyancy"),
                           (abi.F_T, "t")):
            fs.set(slot, F[name])
        return fs

    oa = abi.common_advdiff_opts()
    gf, gc, _ = orc.make_sparsity(gm)
    ref = fmt.csr_to_petsc(gf, gc, orc.assemble_advdiff(gm, fields(gm, GF), oa, gf, gc)["matrix"], nn_glob)
    R = sp.csr_matrix((ref.val, ref.colm, ref.findrm), shape=(nn_glob, nn_glob))
    total = sp.csr_matrix((nn_glob, nn_glob))
    for rank in range(2):
        lp = part.block_partition(cells, (1, 1, 2), rank)
        F = part.global_nodal_fields(3, lp.mesh.X, lp.global_node)
        asm = _asm(lp.mesh, fields(lp.mesh, F))
        # universal numbers = global node ids here; rows of non-owned nodes masked
        rown = lp.global_node.astype(np.int32).copy()
        rown[lp.n_owned:] = -1
        coln = lp.global_node.astype(np.int32)
        nnz = asm.nnz
        n_all = asm.coo_pattern(1, rown[:, None], coln[:, None], compact=False)
        assert n_all == nnz
        asm.advdiff_dev(oa)
        i, j, v = asm.coo_fetch(1, n_all)
        assert (i < 0).sum() > 0 and (j >= 0).all()
        n_kept = asm.coo_pattern(1, rown[:, None], coln[:, None], compact=True)
        ic, jc, vc = asm.coo_fetch(1, n_kept)
        keep = i >= 0
        assert n_kept == keep.sum() and (ic == i[keep]).all() and (jc == j[keep]).all() and (vc == v[keep]).all()
        total = total + sp.coo_matrix((vc, (ic, jc)), shape=(nn_glob, nn_glob)).tocsr()
        asm.close()
    D = abs(total - R)
    assert D.max() <= TOL * abs(R).max()
