"""Host logic of the multi-GPU path, CPU only: the node-based decomposition,
its Fluidity-convention halos, and the property that makes the path shard without a data-path
collective -- every rank assembling ALL its local elements reproduces the global rows of its
owned nodes exactly (same contributions, same per-row element order up to renumbering)."""
import json
import os
import numpy as np
import pytest

from conftest import GOLDEN, load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, partition as part


def _exchange(orc, parts, arrays, block):
    """halo_update of per-rank arrays using the oracle's owner->ghost copy."""
    for p, lp in enumerate(parts):
        for q, lq in enumerate(parts):
            if p != q and len(lp.sends[q]):
                orc.halo_copy(block, arrays[p], lp.sends[q], arrays[q], lq.recvs[p])


def _owners_geometric(mesh, nprocs, axis=0):
    order = np.argsort(mesh.X[:, axis], kind="stable")
    owner = np.zeros(mesh.n_nodes, dtype=np.int64)
    for r, chunk in enumerate(np.array_split(order, nprocs)):
        owner[chunk] = r
    return owner


@pytest.mark.parametrize("name,nprocs", [("cube-parallel", 2), ("cube-parallel", 3), ("2d_square", 4)])
def test_partition_by_owner_conventions(orc, name, nprocs):
    mesh = load_golden_mesh(name)
    owner = _owners_geometric(mesh, nprocs)
    parts = part.partition_by_owner(mesh, owner, nprocs)
    assert sum(lp.n_owned for lp in parts) == mesh.n_nodes
    for r, lp in enumerate(parts):
        # trailing receives: owned first, every receive id > n_owned, every send id <= n_owned
        assert (owner[lp.global_node[:lp.n_owned]] == r).all()
        assert (owner[lp.global_node[lp.n_owned:]] != r).all()
        for p in range(nprocs):
            assert len(lp.sends[p]) == len(parts[p].recvs[r])
            if len(lp.recvs[p]):
                assert lp.recvs[p].min() > lp.n_owned
                assert (owner[lp.global_node[lp.recvs[p] - 1]] == p).all()
            if len(lp.sends[p]):
                assert lp.sends[p].max() <= lp.n_owned
                # same global node on both sides, in the same order
                assert (lp.global_node[lp.sends[p] - 1] == parts[p].global_node[parts[p].recvs[r] - 1]).all()
        allrecv = np.concatenate([lp.recvs[p] for p in range(nprocs)])
        assert sorted(allrecv.tolist()) == list(range(lp.n_owned + 1, lp.mesh.n_nodes + 1))
        # every element touching an owned node is local (so owned rows are complete)
        nd0 = mesh.ndglno.astype(np.int64) - 1
        touching = np.flatnonzero((owner[nd0] == r).any(axis=1))
        assert np.isin(touching, lp.global_element).all()


def test_real_halo_fixture_has_the_same_shape():
    # the reference's own 2-rank decomposition (prectangle): L1 receives precede L2-only ones
    with open(os.path.join(GOLDEN, "prectangle_halos.json")) as f:
        H = json.load(f)
    for r, other in (("0", "1"), ("1", "0")):
        l1 = H[r]["levels"]["1"]["receives"][other]
        l2 = H[r]["levels"]["2"]["receives"][other]
        assert set(l1) <= set(l2)
        assert max(l1) < min(set(l2) - set(l1))


@pytest.mark.parametrize("nprocs", [2, 3])
def test_owned_rows_equal_global_assembly(orc, nprocs):
    mesh = syn.box_mesh((5, 4, 9), seed=3)
    fs = syn.standard_fields(mesh)
    owner = _owners_geometric(mesh, nprocs, axis=2)
    parts = part.partition_by_owner(mesh, owner, nprocs)
    findrm, colm, _ = orc.make_sparsity(mesh)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    gm = orc.assemble_momentum(mesh, fs, om, findrm, colm)
    ga = orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    # per-rank field arrays: owned values only, halo zeroed, then halo_update
    slots = [abi.F_NU, abi.F_OLDU, abi.F_DENSITY, abi.F_BUOYANCY, abi.F_T]
    local_fs = []
    for lp in parts:
        lf = syn.standard_fields(lp.mesh)
        for s in slots:
            v = fs.get(s)[0][lp.global_node].copy()
            v[lp.n_owned:] = 0.0
            lf.set(s, v)
        local_fs.append(lf)
    for s in slots:
        arrays = [lf.get(s)[0] for lf in local_fs]
        _exchange(orc, parts, arrays, arrays[0].shape[1] if arrays[0].ndim > 1 else 1)
    for lp, lf in zip(parts, local_fs):
        for s in slots:
            assert (lf.get(s)[0] == fs.get(s)[0][lp.global_node]).all()
        lfind, lcolm, _ = orc.make_sparsity(lp.mesh)
        lm = orc.assemble_momentum(lp.mesh, lf, om, lfind, lcolm)
        la = orc.assemble_advdiff(lp.mesh, lf, oa, lfind, lcolm)
        for i in range(lp.n_owned):
            gi = lp.global_node[i]
            lrow = slice(lfind[i] - 1, lfind[i + 1] - 1)
            grow = slice(findrm[gi] - 1, findrm[gi + 1] - 1)
            # same columns (as global ids) ...
            lcols = lp.global_node[lcolm[lrow] - 1]
            gcols = colm[grow] - 1
            perm = np.argsort(lcols)
            assert (lcols[perm] == gcols).all()
            # ... and the same values to rounding (element order inside a row may differ)
            for d in range(3):
                assert np.abs(lm["big_m"][d][lrow][perm] - gm["big_m"][d][grow]).max() <= 1e-12 * np.abs(gm["big_m"][d]).max()
            assert np.abs(la["matrix"][lrow][perm] - ga["matrix"][grow]).max() <= 1e-12 * np.abs(ga["matrix"]).max()
        own = lp.global_node[:lp.n_owned]
        assert rel_err(lm["rhs"][:lp.n_owned], gm["rhs"][own]) < 1e-12
        assert rel_err(lm["masslump"][:lp.n_owned], gm["masslump"][own]) < 1e-12
        assert rel_err(la["rhs"][:lp.n_owned], ga["rhs"][own]) < 1e-12


@pytest.mark.parametrize("nprocs", [1, 2, 4])
def test_slab_partition_matches_generic_partition(nprocs):
    cells = (4, 3, 11)
    # global mesh with the same hash jitter = the nprocs=1 slab
    g = part.slab_partition(cells, 1, 0)
    assert g.n_owned == g.mesh.n_nodes and (g.global_node == np.arange(g.mesh.n_nodes)).all()
    Xe = g.mesh.X[g.mesh.ndglno - 1]
    assert np.linalg.det(Xe[:, :3] - Xe[:, 3:]).min() > 0
    npts = (5, 4, 12)
    L = part.slab_layers(npts[2], nprocs)
    layer = np.arange(g.mesh.n_nodes) // (npts[0] * npts[1])
    owner = np.searchsorted(np.array(L[1:]), layer, side="right")
    ref = part.partition_by_owner(g.mesh, owner, nprocs)
    for r in range(nprocs):
        lp = part.slab_partition(cells, nprocs, r)
        assert lp.n_owned == ref[r].n_owned
        assert (lp.global_node[:lp.n_owned] == ref[r].global_node[:lp.n_owned]).all()
        assert sorted(lp.global_node.tolist()) == sorted(ref[r].global_node.tolist())
        assert lp.n_l1 == ref[r].n_l1
        assert sorted(lp.global_node[lp.n_owned:lp.n_owned + lp.n_l1].tolist()) == \
            sorted(ref[r].global_node[ref[r].n_owned:ref[r].n_owned + ref[r].n_l1].tolist())
        # coordinates agree with the global mesh on every local node (hash jitter is global)
        assert (lp.mesh.X == g.mesh.X[lp.global_node]).all()
        # same element set, same connectivity in global numbering
        assert sorted(lp.global_element.tolist()) == sorted(ref[r].global_element.tolist())
        gl = lp.global_node[lp.mesh.ndglno.astype(np.int64) - 1]
        assert (gl == g.mesh.ndglno[lp.global_element].astype(np.int64) - 1).all()
        for p in range(nprocs):
            assert sorted(lp.global_node[lp.sends[p] - 1].tolist()) == sorted(ref[r].global_node[ref[r].sends[p] - 1].tolist())
            assert sorted(lp.global_node[lp.recvs[p] - 1].tolist()) == sorted(ref[r].global_node[ref[r].recvs[p] - 1].tolist())
    # send order on p == receive order on q
    lps = [part.slab_partition(cells, nprocs, r) for r in range(nprocs)]
    for p in range(nprocs):
        for q in range(nprocs):
            assert (lps[p].global_node[lps[p].sends[q] - 1] == lps[q].global_node[lps[q].recvs[p] - 1]).all()


def test_global_fields_agree_on_shared_nodes():
    cells = (3, 3, 8)
    lps = [part.slab_partition(cells, 2, r) for r in range(2)]
    F = [part.global_nodal_fields(3, lp.mesh.X, lp.global_node) for lp in lps]
    common, i0, i1 = np.intersect1d(lps[0].global_node, lps[1].global_node, return_indices=True)
    assert len(common) > 0
    for k in F[0]:
        assert (F[0][k][i0] == F[1][k][i1]).all()


@pytest.mark.parametrize("name,nprocs", [("cube-parallel", 5), ("2d_square", 8), ("square-cavity-2d", 3)])
def test_rcb_partition(orc, name, nprocs):
    """Geometric stand-in for Zoltan/METIS: balanced, deterministic, every rank a
    box of the bisection tree, and valid input for partition_by_owner."""
    mesh = load_golden_mesh(name)
    owner = part.rcb_owner(mesh.X, nprocs)
    assert (owner == part.rcb_owner(mesh.X, nprocs)).all()
    counts = np.bincount(owner, minlength=nprocs)
    assert counts.min() >= mesh.n_nodes // nprocs - 1 and counts.max() <= -(-mesh.n_nodes // nprocs) + 1
    parts = part.partition_by_owner(mesh, owner, nprocs)
    q = part.partition_quality(mesh, parts)
    assert q["owned_imbalance"] < 1.01
    # compact parts: far less halo than a random owner map of the same balance
    rnd = np.random.default_rng(0).permutation(owner)
    q_rnd = part.partition_quality(mesh, part.partition_by_owner(mesh, rnd, nprocs))
    assert q["halo_nodes_sent_total"] < 0.5 * q_rnd["halo_nodes_sent_total"]
    assert q["element_redundancy"] < q_rnd["element_redundancy"]
    for r, lp in enumerate(parts):
        for p in range(nprocs):
            assert len(lp.sends[p]) == len(parts[p].recvs[r])
    # the bounding boxes of two parts overlap at most in a thin layer along one axis
    lo = np.array([mesh.X[owner == r].min(axis=0) for r in range(nprocs)])
    hi = np.array([mesh.X[owner == r].max(axis=0) for r in range(nprocs)])
    for a in range(nprocs):
        for b in range(a + 1, nprocs):
            overlap = np.minimum(hi[a], hi[b]) - np.maximum(lo[a], lo[b])
            assert (overlap <= 1e-12).any() or overlap.min() < 0.05 * (mesh.X.max() - mesh.X.min())


@pytest.mark.parametrize("name,nprocs", [("cube-parallel", 3), ("square-cavity-2d", 2)])
def test_surface_loops_shard_like_the_element_loops(orc, name, nprocs):
    """Every rank runs the face loop over the boundary faces it holds (partition_by_owner(..., sndgln): the faces
    fldecomp writes into its mesh file); the rows of owned nodes equal the global element + face assembly."""
    mesh = load_golden_mesh(name)
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    sn, fe = syn.boundary_faces(mesh)
    rng = np.random.default_rng(3)
    bt = rng.choice([0, 1, 4], size=len(fe)).astype(np.int32)
    t_bc, t_bc_2 = rng.uniform(size=(len(fe), dim)), rng.uniform(0.5, 2.0, size=(len(fe), dim))
    vt = rng.choice([0, 1, 5], size=(len(fe), dim)).astype(np.int32)
    vbc = rng.uniform(-1, 1, size=(len(fe), dim, dim))
    findrm, colm, _ = orc.make_sparsity(mesh)
    oa = abi.common_advdiff_opts(integrate_advection_by_parts=1, beta=0.2)
    om = abi.common_momentum_opts(integrate_advection_by_parts=1)
    ga = orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    orc.assemble_advdiff_surface(mesh, fs, oa, findrm, colm, sn, fe, bt, t_bc, t_bc_2, ga["matrix"], ga["rhs"])
    gm = orc.assemble_momentum(mesh, fs, om, findrm, colm)
    orc.assemble_momentum_surface(mesh, fs, om, findrm, colm, sn, fe, vt, vbc, gm["big_m"], gm["rhs"])
    parts = part.partition_by_owner(mesh, part.rcb_owner(mesh.X, nprocs), nprocs, sn, np.arange(len(sn)))
    slots = [abi.F_NU, abi.F_OLDU, abi.F_DENSITY, abi.F_BUOYANCY, abi.F_T]
    for lp in parts:
        lf = syn.standard_fields(lp.mesh)
        for s in slots:
            lf.set(s, fs.get(s)[0][lp.global_node])
        lfind, lcolm, _ = orc.make_sparsity(lp.mesh)
        # owning element of every local face: the local element that contains all its nodes
        ele_of = {}
        for e, nd in enumerate(lp.mesh.ndglno):
            for k in range(mesh.loc):
                ele_of.setdefault(tuple(sorted(np.delete(nd, k))), e + 1)
        lfe = np.array([ele_of[tuple(sorted(f))] for f in lp.sndgln], dtype=np.int32)
        gf = lp.global_face
        # the same face may list its nodes in a different local order than the global file: boundary values follow the nodes
        order = np.array([[list(sn[g]).index(n) for n in lp.global_node[f - 1] + 1] for f, g in zip(lp.sndgln, gf)])
        take = lambda a: np.take_along_axis(a[gf], order, axis=1)
        la = orc.assemble_advdiff(lp.mesh, lf, oa, lfind, lcolm)
        orc.assemble_advdiff_surface(lp.mesh, lf, oa, lfind, lcolm, lp.sndgln, lfe, bt[gf], take(t_bc), take(t_bc_2),
                                     la["matrix"], la["rhs"])
        lm = orc.assemble_momentum(lp.mesh, lf, om, lfind, lcolm)
        orc.assemble_momentum_surface(lp.mesh, lf, om, lfind, lcolm, lp.sndgln, lfe, vt[gf],
                                      np.take_along_axis(vbc[gf], order[:, :, None], axis=1), lm["big_m"], lm["rhs"])
        for i in range(lp.n_owned):
            gi = lp.global_node[i]
            lrow, grow = slice(lfind[i] - 1, lfind[i + 1] - 1), slice(findrm[gi] - 1, findrm[gi + 1] - 1)
            perm = np.argsort(lp.global_node[lcolm[lrow] - 1])
            assert np.abs(la["matrix"][lrow][perm] - ga["matrix"][grow]).max() <= 1e-12 * np.abs(ga["matrix"]).max()
            for d in range(dim):
                assert np.abs(lm["big_m"][d][lrow][perm] - gm["big_m"][d][grow]).max() <= 1e-12 * np.abs(gm["big_m"][d]).max()
        own = lp.global_node[:lp.n_owned]
        assert rel_err(la["rhs"][:lp.n_owned], ga["rhs"][own]) < 1e-12
        assert rel_err(lm["rhs"][:lp.n_owned], gm["rhs"][own]) < 1e-12


# ---- block partition (strong scaling of ONE box: bench.py --gpus N) ----------------------------------------------
def _block_owner(shape, pgrid, gid):
    npts = [c + 1 for c in shape]
    L = [part.slab_layers(npts[k], pgrid[k]) for k in range(3)]
    idx = [gid % npts[0], (gid // npts[0]) % npts[1], gid // (npts[0] * npts[1])]
    owner = np.zeros(len(gid), dtype=np.int64)
    for k in (2, 1, 0):
        owner = owner * pgrid[k] + (np.searchsorted(np.asarray(L[k]), idx[k], side="right") - 1)
    return owner


@pytest.mark.parametrize("shape,pgrid", [((6, 5, 7), (2, 2, 2)), ((5, 5, 9), (1, 2, 2)), ((4, 4, 8), (1, 1, 2)),
                                         ((7, 6, 5), (2, 1, 3))])
def test_block_partition_is_the_fldecomp_partition_of_the_block_owner_map(shape, pgrid):
    """block_partition never builds the global mesh; on small boxes it must equal partition_by_owner (= the reference's
    fldgmsh.cpp writer, tests/test_formats.py) list for list: numbering, elements, coordinates, both halo lists."""
    whole = part.block_partition(shape, (1, 1, 1), 0)
    nprocs = int(np.prod(pgrid))
    ref = part.partition_by_owner(whole.mesh, _block_owner(shape, pgrid, whole.global_node), nprocs)
    n_owned = 0
    for r in range(nprocs):
        lp, a = part.block_partition(shape, pgrid, r), ref[r]
        assert lp.n_owned == a.n_owned and lp.n_l1 == a.n_l1
        assert (lp.global_node == a.global_node).all() and (lp.global_element == a.global_element).all()
        assert (lp.mesh.ndglno == a.mesh.ndglno).all() and (lp.mesh.X == a.mesh.X).all()
        for p in range(nprocs):
            assert (lp.recvs[p] == a.recvs[p]).all() and (lp.sends[p] == a.sends[p]).all()
        n_owned += lp.n_owned
    assert n_owned == whole.mesh.n_nodes
    # the one-block case is the slab generator's whole box (same coordinates bit for bit: bench N=1 is unchanged)
    w2 = part.slab_partition(shape, 1, 0)
    assert (whole.mesh.X == w2.mesh.X).all() and (whole.mesh.ndglno == w2.mesh.ndglno).all()


def test_block_grid():
    assert [part.block_grid(n) for n in (1, 2, 4, 8)] == [(1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
    assert int(np.prod(part.block_grid(6))) == 6 and int(np.prod(part.block_grid(3))) == 3
