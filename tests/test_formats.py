"""On-disk formats either side of the path: gmsh 2.x,
.halo XML, `_<rank>` decompositions, PETSc binary dumps. CPU only.

Pins: (1) the reference's own fixture files, read where they lie when /root/reference exists
(this container) and compared with the independent line parser of tests/golden/make_golden.py;
(2) the committed conversions of the reference's real 2-rank decomposition `prectangle`, which
must drop into the multi-rank path unchanged: halo lists pair up node for node and the owned
rows assembled per rank equal the rows of the re-joined global mesh."""
import json
import os
import numpy as np
import pytest

from conftest import GOLDEN, load_golden_mesh, rel_err
from fluidity_b200 import formats as fmt, partition as part, synthetic as syn, _abi as abi

REF = os.environ.get("FLUIDITY_REFERENCE", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests", "data")),
                                     reason="reference fixtures are only present in the build container")


# ---- gmsh ------------------------------------------------------------------------------------------------
@needs_reference
@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "square-cavity-2d", "2d_square"])
def test_gmsh_reader_against_reference_fixtures(name):
    g = fmt.read_gmsh(os.path.join(REF, "tests", "data", name + ".msh"))
    ref = load_golden_mesh(name)  # converted by make_golden.read_gmsh_ascii
    assert g.mesh.dim == ref.dim and not g.binary
    assert (g.mesh.ndglno == ref.ndglno).all()
    assert (g.mesh.X == ref.X).all()
    sloc = g.mesh.dim
    assert g.sndgln.shape[1] == sloc
    if len(g.sndgln):
        # every face is a face of some element
        faces = {tuple(sorted(f)) for e in g.mesh.ndglno for f in
                 [tuple(np.delete(e, k)) for k in range(g.mesh.loc)]}
        assert all(tuple(sorted(f)) in faces for f in g.sndgln)


@needs_reference
def test_gmsh_binary_reader_reads_the_reference_decomposition():
    src = os.path.join(REF, "tests", "meshconv_test", "src")
    for r in (0, 1):
        g = fmt.read_gmsh(os.path.join(src, "prectangle_%d.msh" % r))
        z = np.load(os.path.join(GOLDEN, "prectangle_%d.npz" % r))
        assert g.binary and g.version == (2, 1)
        assert (g.mesh.ndglno == z["ndglno"]).all() and (g.mesh.X == z["X"]).all()
        assert g.element_owner is not None  # 4 face tags: surface id, 0, 0, owning element
        own = g.mesh.ndglno[g.element_owner - 1]
        assert all(set(f) <= set(e) for f, e in zip(g.sndgln, own))
        h = fmt.read_halo(os.path.join(src, "prectangle_%d.halo" % r))
        with open(os.path.join(GOLDEN, "prectangle_halos.json")) as f:
            ref = json.load(f)[str(r)]
        for lv, hl in h.levels.items():
            assert hl.n_private_nodes == ref["levels"][str(lv)]["n_private_nodes"]
            for p in range(h.nprocs):
                assert hl.sends[p].tolist() == ref["levels"][str(lv)]["sends"][str(p)]
                assert hl.receives[p].tolist() == ref["levels"][str(lv)]["receives"][str(p)]


@needs_reference
def test_serial_dummy_halo_with_comment_and_legacy_tag():
    # tests/data/cube-parallel_0.halo: a comment in front of the declaration, `tag=` instead of `level=`
    h = fmt.read_halo(os.path.join(REF, "tests", "data", "cube-parallel_0.halo"))
    assert (h.process, h.nprocs) == (0, 1) and sorted(h.levels) == [1, 2]
    assert h.levels[1].n_private_nodes == 665
    assert h.levels[2].receives[0].min() == 666  # "665 private nodes, receives 666..."


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("name", ["cube.1", "square-cavity-2d"])
def test_gmsh_round_trip(tmp_path, name, binary):
    m = load_golden_mesh(name)
    rng = np.random.default_rng(4)
    sloc = m.dim
    faces = np.array([np.delete(m.ndglno[e], e % m.loc) for e in range(min(7, m.n_elements))], dtype=np.int32)
    g = fmt.GmshMesh(mesh=m, sndgln=faces, boundary_ids=rng.integers(1, 9, len(faces)).astype(np.int32),
                     element_owner=np.arange(1, len(faces) + 1, dtype=np.int32),
                     region_ids=rng.integers(1, 4, m.n_elements).astype(np.int32))
    p = str(tmp_path / "m.msh")
    fmt.write_gmsh(p, g, binary=binary)
    b = fmt.read_gmsh(p)
    assert b.binary == binary and b.mesh.dim == m.dim and b.sndgln.shape == (len(faces), sloc)
    assert (b.mesh.ndglno == m.ndglno).all() and (b.mesh.X == m.X).all()  # exact, also in ASCII
    assert (b.sndgln == faces).all() and (b.boundary_ids == g.boundary_ids).all()
    assert (b.element_owner == g.element_owner).all() and (b.region_ids == g.region_ids).all()


def test_gmsh_reader_refuses_what_the_reference_refuses(tmp_path):
    def write(text):
        p = tmp_path / "bad.msh"
        p.write_text(text)
        return str(p)
    with pytest.raises(fmt.FormatError, match="version"):
        fmt.read_gmsh(write("$MeshFormat\n3.0 0 8\n$EndMeshFormat\n"))
    with pytest.raises(fmt.FormatError, match="data size"):
        fmt.read_gmsh(write("$MeshFormat\n2.2 0 4\n$EndMeshFormat\n"))
    with pytest.raises(fmt.FormatError, match="nodes field < 2"):
        fmt.read_gmsh(write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n"))
    body = "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n2 1 0 0\n3 0 1 0\n$EndNodes\n$Elements\n"
    with pytest.raises(fmt.FormatError, match="Unsupported element type"):
        fmt.read_gmsh(write(body + "1\n1 9 2 0 0 1 2 3 1 2 3\n$EndElements\n"))
    with pytest.raises(fmt.FormatError, match="Inconsistent number of element tags"):
        fmt.read_gmsh(write(body + "2\n1 2 2 0 0 1 2 3\n2 2 3 0 0 0 1 2 3\n$EndElements\n"))
    with pytest.raises(fmt.FormatError, match=r"\$EndElements"):
        fmt.read_gmsh(write(body + "1\n1 2 2 0 0 1 2 3\n"))
    g = fmt.read_gmsh(write(body + "1\n1 2 2 5 0 1 2 3\n$EndElements\n"))
    assert g.mesh.n_elements == 1 and g.region_ids.tolist() == [5] and g.boundary_ids is None


# ---- .halo -----------------------------------------------------------------------------------------------
def _golden_halos(r):
    with open(os.path.join(GOLDEN, "prectangle_halos.json")) as f:
        H = json.load(f)[str(r)]
    out = fmt.Halos(process=H["process"], nprocs=H["nprocs"])
    for lv, e in H["levels"].items():
        out.levels[int(lv)] = fmt.HaloLevel(e["n_private_nodes"],
                                            [np.array(e["sends"][str(p)], dtype=np.int32) for p in range(H["nprocs"])],
                                            [np.array(e["receives"][str(p)], dtype=np.int32) for p in range(H["nprocs"])])
    return out


def test_halo_round_trip_and_validity_rules(tmp_path):
    h = _golden_halos(0)
    p = str(tmp_path / "x_0.halo")
    fmt.write_halo(p, h)
    b = fmt.read_halo(p)
    assert (b.process, b.nprocs) == (0, 2) and sorted(b.levels) == [1, 2]
    for lv in (1, 2):
        assert b.levels[lv].n_private_nodes == h.levels[lv].n_private_nodes
        for q in range(2):
            assert (b.levels[lv].sends[q] == h.levels[lv].sends[q]).all()
            assert (b.levels[lv].receives[q] == h.levels[lv].receives[q]).all()
        assert fmt.trailing_receives_consistent(b.levels[lv])
    text = open(p).read()
    if os.path.isdir(os.path.join(REF, "tests", "meshconv_test", "src")):
        # byte-identical with the file WriteHalos produced (TinyXML layout: 4-blank indent, "id " lists)
        assert text == open(os.path.join(REF, "tests", "meshconv_test", "src", "prectangle_0.halo")).read()
    for bad in (text.replace('process="0" nprocs="2"', 'process="2" nprocs="2"'),       # process >= nprocs
                text.replace('n_private_nodes="16"', 'n_private_nodes="-1"', 1),        # negative
                text.replace('<halo_data process="1">', '<halo_data process="0">', 1),  # repeated process
                text.replace('<halo_data process="1">', '<halo_data process="5">', 1),  # out of range
                text.replace(' level="1"', "", 1),                                       # neither level nor tag
                text[:200]):                                                             # truncated XML
        q = tmp_path / "bad_0.halo"
        q.write_text(bad)
        with pytest.raises(fmt.FormatError):
            fmt.read_halo(str(q))
    with pytest.raises(fmt.FormatError):
        fmt.read_halo(str(tmp_path / "absent_0.halo"))


# ---- decompositions --------------------------------------------------------------------------------------
def _owned_rows_match(orc, parts, gmesh, l2g):
    """Every rank assembles all its local elements; rows of owned nodes == global rows."""
    dim = gmesh.dim
    fs = syn.standard_fields(gmesh)
    findrm, colm, _ = orc.make_sparsity(gmesh)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    gm = orc.assemble_momentum(gmesh, fs, om, findrm, colm)
    ga = orc.assemble_advdiff(gmesh, fs, oa, findrm, colm)
    slots = [abi.F_NU, abi.F_OLDU, abi.F_DENSITY, abi.F_BUOYANCY, abi.F_T]
    local_fs = []
    for lp, g in zip(parts, l2g):
        lf = syn.standard_fields(lp.mesh)
        for s in slots:
            v = fs.get(s)[0][g].copy()
            v[lp.n_owned:] = 0.0  # stale halo values, to be filled by the exchange
            lf.set(s, v)
        local_fs.append(lf)
    for s in slots:
        arrays = [lf.get(s)[0] for lf in local_fs]
        block = arrays[0].shape[1] if arrays[0].ndim > 1 else 1
        for p, lp in enumerate(parts):
            for q, lq in enumerate(parts):
                if p != q and len(lp.sends[q]):
                    orc.halo_copy(block, arrays[p], lp.sends[q], arrays[q], lq.recvs[p])
    for lp, lf, g in zip(parts, local_fs, l2g):
        for s in slots:
            assert (lf.get(s)[0] == fs.get(s)[0][g]).all()  # the halo lists reach every non-owned node
        lfind, lcolm, _ = orc.make_sparsity(lp.mesh)
        lm = orc.assemble_momentum(lp.mesh, lf, om, lfind, lcolm)
        la = orc.assemble_advdiff(lp.mesh, lf, oa, lfind, lcolm)
        for i in range(lp.n_owned):
            lrow = slice(lfind[i] - 1, lfind[i + 1] - 1)
            grow = slice(findrm[g[i]] - 1, findrm[g[i] + 1] - 1)
            lcols = g[lcolm[lrow] - 1]
            perm = np.argsort(lcols)
            assert (lcols[perm] == colm[grow] - 1).all()
            for d in range(dim):
                assert np.abs(lm["big_m"][d][lrow][perm] - gm["big_m"][d][grow]).max() <= 1e-12 * np.abs(gm["big_m"][d]).max()
            assert np.abs(la["matrix"][lrow][perm] - ga["matrix"][grow]).max() <= 1e-12 * np.abs(ga["matrix"]).max()
        own = g[:lp.n_owned]
        assert rel_err(lm["rhs"][:lp.n_owned], gm["rhs"][own]) < 1e-12
        assert rel_err(la["rhs"][:lp.n_owned], ga["rhs"][own]) < 1e-12


def test_reference_decomposition_drops_into_the_multirank_path(orc, tmp_path):
    """prectangle_{0,1}: flredecomp output of the reference (binary gmsh + L1/L2 .halo). Written back
    with our writers, read with read_decomposition, joined by coordinates into the global mesh."""
    base = str(tmp_path / "prectangle")
    for r in (0, 1):
        z = np.load(os.path.join(GOLDEN, "prectangle_%d.npz" % r))
        g = fmt.GmshMesh(mesh=syn.Mesh(dim=int(z["dim"]), ndglno=z["ndglno"], X=z["X"]), sndgln=z["sndgln"],
                         boundary_ids=z["boundary_ids"], element_owner=z["element_owner"], region_ids=z["region_ids"])
        fmt.write_gmsh(fmt.parallel_filename(base, r, ".msh"), g, binary=True)
        fmt.write_halo(fmt.parallel_filename(base, r, ".halo"), _golden_halos(r))
    parts = [fmt.read_decomposition(base, r)[0] for r in (0, 1)]
    assert [lp.n_owned for lp in parts] == [16, 7] and [lp.n_l1 for lp in parts] == [6, 7]
    # both sides of every list are the same points, in the same order
    for r, p in ((0, 1), (1, 0)):
        assert len(parts[r].sends[p]) == len(parts[p].recvs[r]) > 0
        assert (parts[r].mesh.X[parts[r].sends[p] - 1] == parts[p].mesh.X[parts[p].recvs[r] - 1]).all()
    # global numbering: owned nodes of rank 0, then of rank 1; halo nodes found by their coordinates
    owned_X = np.concatenate([lp.mesh.X[:lp.n_owned] for lp in parts])
    assert len(np.unique(owned_X, axis=0)) == 23
    key = {tuple(x): i for i, x in enumerate(owned_X)}
    l2g = [np.array([key[tuple(x)] for x in lp.mesh.X]) for lp in parts]
    eles = {}
    for lp, g in zip(parts, l2g):
        for e in g[lp.mesh.ndglno - 1]:
            eles.setdefault(tuple(sorted(e)), e)
    gmesh = syn.Mesh(dim=2, ndglno=(np.array(list(eles.values())) + 1).astype(np.int32), X=owned_X)
    # every element around an owned node is present locally
    nd0 = gmesh.ndglno - 1
    for r, (lp, g) in enumerate(zip(parts, l2g)):
        mine = np.isin(nd0, g[:lp.n_owned]).any(axis=1).sum()
        local = np.isin(g[lp.mesh.ndglno - 1], g[:lp.n_owned]).any(axis=1).sum()
        assert mine == local
    _owned_rows_match(orc, parts, gmesh, l2g)


@pytest.mark.parametrize("binary", [False, True])
def test_our_decomposition_round_trips_through_the_reference_formats(orc, tmp_path, binary):
    mesh = load_golden_mesh("cube-parallel")
    nprocs = 3
    order = np.argsort(mesh.X[:, 0], kind="stable")
    owner = np.zeros(mesh.n_nodes, dtype=np.int64)
    for r, chunk in enumerate(np.array_split(order, nprocs)):
        owner[chunk] = r
    parts = part.partition_by_owner(mesh, owner, nprocs)
    base = str(tmp_path / "cube")
    fmt.write_decomposition(base, parts, binary=binary)
    for r, lp in enumerate(parts):
        assert os.path.exists("%s_%d.msh" % (base, r)) and os.path.exists("%s_%d.halo" % (base, r))
        back, gm, hs = fmt.read_decomposition(base, r)
        assert back.n_owned == lp.n_owned and back.n_l1 == lp.n_l1
        assert (back.mesh.ndglno == lp.mesh.ndglno).all() and (back.mesh.X == lp.mesh.X).all()
        for p in range(nprocs):
            assert (back.sends[p] == lp.sends[p]).all() and (back.recvs[p] == lp.recvs[p]).all()
        # level 1 is the level-2 list restricted to the first n_l1 receive nodes, pairwise consistent
        l1 = hs.levels[1]
        assert fmt.trailing_receives_consistent(l1) and fmt.trailing_receives_consistent(hs.levels[2])
        allr = np.concatenate(l1.receives)
        assert sorted(allr.tolist()) == list(range(lp.n_owned + 1, lp.n_owned + lp.n_l1 + 1))
        for p in range(nprocs):
            other = fmt.read_halo("%s_%d.halo" % (base, p)).levels[1]
            assert len(l1.sends[p]) == len(other.receives[r])
            assert (lp.global_node[l1.sends[p] - 1] == parts[p].global_node[other.receives[r] - 1]).all()
    with pytest.raises(fmt.FormatError, match="process number"):
        os.replace("%s_1.halo" % base, "%s_7.halo" % base)
        os.replace("%s_1.msh" % base, "%s_7.msh" % base)
        fmt.read_decomposition(base, 7)


# ---- PETSc binary dumps ----------------------------------------------------------------------------------
def test_petsc_binary_layout_is_the_published_one(tmp_path):
    # 2 x 3 matrix [[1, 0, 2], [0, 3, 0]] and a vector, written by hand byte for byte
    import struct
    raw = struct.pack(">4i", 1211216, 2, 3, 3) + struct.pack(">2i", 2, 1) + struct.pack(">3i", 0, 2, 1) + \
        struct.pack(">3d", 1.0, 2.0, 3.0) + struct.pack(">2i", 1211214, 2) + struct.pack(">2d", 0.5, -4.0)
    p = tmp_path / "matrixdump"
    p.write_bytes(raw)
    A, b = fmt.read_petsc_binary(str(p))
    assert (A.rows, A.cols) == (2, 3) and A.findrm.tolist() == [0, 2, 3] and A.colm.tolist() == [0, 2, 1]
    assert A.val.tolist() == [1.0, 2.0, 3.0] and b.tolist() == [0.5, -4.0]
    q = tmp_path / "again"
    fmt.write_petsc_binary(str(q), [A, b])
    assert q.read_bytes() == raw
    fmt.write_petsc_binary(str(q), [A, b], int64=True)
    A8, b8 = fmt.read_petsc_binary(str(q), int64=True)
    assert A8.colm.tolist() == [0, 2, 1] and b8.tolist() == [0.5, -4.0]
    (tmp_path / "cut").write_bytes(raw[:40])
    with pytest.raises(fmt.FormatError, match="truncated"):
        fmt.read_petsc_binary(str(tmp_path / "cut"))
    (tmp_path / "junk").write_bytes(struct.pack(">2i", 77, 1))
    with pytest.raises(fmt.FormatError, match="class id"):
        fmt.read_petsc_binary(str(tmp_path / "junk"))
    assert fmt.dump_name_parts("/x/Velocity_12") == ("Velocity", 12) and fmt.dump_name_parts("matrixdump") == ("matrixdump", None)


def test_petsc_numbering_and_block_expansion(orc):
    # femtools/Petsc_Tools.F90:184-199: field-major without groups, node-major inside a group
    n = fmt.petsc_row_numbering(4, 3)
    assert n[:, 0].tolist() == [0, 1, 2, 3] and n[:, 2].tolist() == [8, 9, 10, 11]
    g = fmt.petsc_row_numbering(4, 3, group_size=3)
    assert g[0].tolist() == [0, 1, 2] and g[3].tolist() == [9, 10, 11]
    # femtools/tests/test_petsc_csr_matrix.F90: 2 x 2 blocks of a 2-node dense sparsity, entry
    # values 1..16 in (block, node) order land at gnn2unn rows/columns
    findrm, colm = np.array([1, 3, 5]), np.array([1, 2, 1, 2])
    blocks = np.arange(1.0, 17.0).reshape(2, 2, 4)
    M = fmt.blocks_to_petsc(findrm, colm, blocks, 2, diagonal=False)
    import scipy.sparse as sp
    D = sp.csr_matrix((M.val, M.colm, M.findrm), shape=(4, 4)).toarray()
    for bi in range(2):
        for bj in range(2):
            assert (D[2 * bi:2 * bi + 2, 2 * bj:2 * bj + 2].ravel() == blocks[bi, bj]).all()
    # an assembled momentum matrix survives dump -> read -> compare, and a 1e-9 perturbation does not
    mesh = load_golden_mesh("cube.1")
    fs = syn.standard_fields(mesh)
    fr, cm, _ = orc.make_sparsity(mesh)
    res = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(), fr, cm)
    A = fmt.blocks_to_petsc(fr, cm, np.array(res["big_m"]), mesh.n_nodes)
    assert A.rows == 3 * mesh.n_nodes and (np.diff(A.colm)[np.diff(np.repeat(np.arange(A.rows), np.diff(A.findrm))) == 0] > 0).all()
    same = fmt.compare_petsc_mats(A, A)
    assert same["ok"] and same["block_rel"] == 0.0
    # PETSc drops nothing, but a dump made with MAT_IGNORE_ZERO_ENTRIES has fewer stored entries
    B = fmt.blocks_to_petsc(fr, cm, np.array(res["big_m"]), mesh.n_nodes, keep_zeros=False)
    assert fmt.compare_petsc_mats(A, B)["ok"]
    C = fmt.PetscMat(A.rows, A.cols, A.findrm, A.colm, A.val * (1 + 1e-9))
    assert not fmt.compare_petsc_mats(C, A)["ok"]


def test_vector_fields_follow_the_matrix_numbering(orc):
    mesh = load_golden_mesh("cube.1")
    fs = syn.standard_fields(mesh)
    fr, cm, _ = orc.make_sparsity(mesh)
    res = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(), fr, cm)
    x = np.random.default_rng(0).uniform(size=(mesh.n_nodes, 3))
    import scipy.sparse as sp
    want = np.stack([sp.csr_matrix((res["big_m"][d], cm - 1, fr - 1), shape=(mesh.n_nodes,) * 2) @ x[:, d] for d in range(3)], axis=1)
    for gs in (1, 3):
        A = fmt.blocks_to_petsc(fr, cm, np.array(res["big_m"]), mesh.n_nodes, group_size=gs)
        M = sp.csr_matrix((A.val, A.colm, A.findrm), shape=(A.rows, A.cols))
        assert np.abs(M @ fmt.field_to_petsc(x, gs) - fmt.field_to_petsc(want, gs)).max() < 1e-14
    assert (fmt.field_to_petsc(x[:, 0]) == x[:, 0]).all()


def test_compare_matrixdump_tool(orc, tmp_path):
    import subprocess
    import sys
    mesh = load_golden_mesh("cube.1")
    fs = syn.standard_fields(mesh)
    fr, cm, _ = orc.make_sparsity(mesh)
    res = orc.assemble_advdiff(mesh, fs, abi.common_advdiff_opts(), fr, cm)
    A = fmt.csr_to_petsc(fr, cm, res["matrix"], mesh.n_nodes)
    x0 = np.zeros(mesh.n_nodes)
    fmt.write_petsc_binary(str(tmp_path / "T_1"), [A, res["rhs"], x0])
    fmt.write_petsc_binary(str(tmp_path / "T_ours"), [A, res["rhs"] * (1 + 1e-14), x0])
    fmt.write_petsc_binary(str(tmp_path / "T_off"), [A, res["rhs"] * (1 + 1e-8), x0])
    tool = os.path.join(os.path.dirname(GOLDEN), "..", "scripts", "compare_matrixdump.py")
    good = subprocess.run([sys.executable, tool, str(tmp_path / "T_1"), str(tmp_path / "T_ours")], capture_output=True, text=True)
    assert good.returncode == 0 and json.loads(good.stdout)["ok"]
    bad = subprocess.run([sys.executable, tool, str(tmp_path / "T_1"), str(tmp_path / "T_off")], capture_output=True, text=True)
    assert bad.returncode == 1 and not json.loads(bad.stdout)["objects"][1]["ok"]


# ---- against the reference's own compiled halo reader / writer (oracle/_ref) ------------------------------
def _ref_halos():
    from oracle import ref_halos
    if not ref_halos.available():
        pytest.skip("oracle/_ref/libref_halos_io.so is built from /root/reference (build container only)")
    return ref_halos


def test_our_halo_files_are_read_by_the_reference_reader(tmp_path):
    """Halos_IO.cpp, compiled unmodified from the reference tree (`make -C oracle ref`), reads the files
    formats.write_decomposition produces and returns the lists we wrote."""
    rh = _ref_halos()
    mesh = load_golden_mesh("cube-parallel")
    nprocs = 3
    parts = part.partition_by_owner(mesh, part.rcb_owner(mesh.X, nprocs), nprocs)
    base = str(tmp_path / "cube")
    fmt.write_decomposition(base, parts, binary=True)
    for r, lp in enumerate(parts):
        got = rh.read(base, r, nprocs)
        ours = fmt.read_halo(fmt.parallel_filename(base, r, ".halo"))
        for lv in (1, 2):
            npn, sends, recvs = got[lv]
            assert npn == lp.n_owned == ours.levels[lv].n_private_nodes
            for p in range(nprocs):
                assert (sends[p] == ours.levels[lv].sends[p]).all() and (recvs[p] == ours.levels[lv].receives[p]).all()
        assert all((got[2][1][p] == lp.sends[p]).all() and (got[2][2][p] == lp.recvs[p]).all() for p in range(nprocs))
    # a file for the wrong process and a corrupt file are refused by the reference reader as by ours
    os.replace("%s_1.halo" % base, "%s_5.halo" % base)
    os.replace("%s_1.msh" % base, "%s_5.msh" % base)
    with pytest.raises(ValueError):
        rh.read(base, 5, 6)
    with pytest.raises(fmt.FormatError):
        fmt.read_decomposition(base, 5)


def test_reference_writer_and_ours_produce_the_same_bytes(tmp_path):
    rh = _ref_halos()
    mesh = load_golden_mesh("2d_square")
    nprocs = 4
    parts = part.partition_by_owner(mesh, part.rcb_owner(mesh.X, nprocs), nprocs)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "theirs")
    fmt.write_decomposition(ours, parts)
    for r in range(nprocs):
        h = fmt.read_halo(fmt.parallel_filename(ours, r, ".halo"))
        rh.write(theirs, r, nprocs, {lv: (hl.n_private_nodes, hl.sends, hl.receives) for lv, hl in h.levels.items()})
        a = open(fmt.parallel_filename(ours, r, ".halo"), "rb").read()
        b = open(fmt.parallel_filename(theirs, r, ".halo"), "rb").read()
        assert a == b
        back = fmt.read_halo(fmt.parallel_filename(theirs, r, ".halo"))
        assert all((back.levels[2].sends[p] == parts[r].sends[p]).all() for p in range(nprocs))


@needs_reference
def test_reference_reader_and_ours_agree_on_the_reference_fixtures():
    rh = _ref_halos()
    src = os.path.join(REF, "tests", "meshconv_test", "src", "prectangle")
    for r in (0, 1):
        got = rh.read(src, r, 2)
        ours = fmt.read_halo(fmt.parallel_filename(src, r, ".halo"))
        for lv in (1, 2):
            assert got[lv][0] == ours.levels[lv].n_private_nodes
            for p in range(2):
                assert (got[lv][1][p] == ours.levels[lv].sends[p]).all() and (got[lv][2][p] == ours.levels[lv].receives[p]).all()
    serial = rh.read(os.path.join(REF, "tests", "data", "cube-parallel"), 0, 1)   # comment before the declaration, `tag=`
    ours = fmt.read_halo(os.path.join(REF, "tests", "data", "cube-parallel_0.halo"))
    assert serial[1][0] == 665 and (serial[2][2][0] == ours.levels[2].receives[0]).all()


@pytest.mark.parametrize("name,nparts", [("cube-parallel", 3), ("2d_square", 4), ("square-cavity-2d", 2), ("cube.1", 2)])
def test_partition_by_owner_is_the_reference_decomposition_writer(tmp_path, name, nparts):
    """fldecomp/fldgmsh.cpp write_partitions_gmsh -- the reference's own code that turns a node -> partition map
    into per-rank meshes and halos -- compiled unmodified into oracle/_ref and fed OUR owner map: our LocalParts
    and the files we write for them are identical, byte for byte (.msh: nodes, faces, elements; .halo: both levels)."""
    from oracle import ref_fldecomp as rf
    if not rf.available():
        pytest.skip("oracle/_ref/libref_fldecomp.so is built from /root/reference (build container only)")
    mesh = load_golden_mesh(name)
    owner = part.rcb_owner(mesh.X, nparts)
    sn, fe = syn.boundary_faces(mesh)
    bid = (np.arange(len(sn)) % 5 + 1).astype(np.int32)
    rid = (np.arange(mesh.n_elements) % 3 + 7).astype(np.int32)
    theirs, ours = str(tmp_path / "theirs"), str(tmp_path / "ours")
    rf.write_partitions(theirs, mesh, owner, nparts, sn, bid, rid)
    parts = part.partition_by_owner(mesh, owner, nparts, sn, bid)
    fmt.write_decomposition(ours, parts, style="fldecomp", region_ids=rid)
    for r, mine in enumerate(parts):
        lp, gm, hs = fmt.read_decomposition(theirs, r)
        assert lp.n_owned == mine.n_owned and lp.n_l1 == mine.n_l1
        assert (lp.mesh.X == mine.mesh.X).all() and (lp.mesh.ndglno == mine.mesh.ndglno).all()
        assert (gm.sndgln == mine.sndgln).all() and (gm.boundary_ids == mine.boundary_ids).all()
        assert (gm.region_ids == rid[mine.global_element]).all()
        for p in range(nparts):
            assert (lp.recvs[p] == mine.recvs[p]).all() and (lp.sends[p] == mine.sends[p]).all()
        for ext in (".msh", ".halo"):
            a = open(fmt.parallel_filename(theirs, r, ext), "rb").read()
            b = open(fmt.parallel_filename(ours, r, ext), "rb").read()
            assert a == b, (r, ext)
        # every local face lies in one local element, and every global boundary face of an owned node is held
        if len(mine.sndgln):
            ele_sets = {tuple(sorted(np.delete(e, k))) for e in mine.mesh.ndglno for k in range(mesh.loc)}
            assert all(tuple(sorted(f)) in ele_sets for f in mine.sndgln)
        touching = np.flatnonzero((owner[sn - 1] == r).any(axis=1))
        assert np.isin(touching, mine.global_face).all()


@needs_reference
@pytest.mark.parametrize("rel", ["tests/data/cube-parallel.msh", "tests/data/square-cavity-2d.msh", "tests/data/cube.1.msh",
                                 "tests/meshconv_test/src/prectangle_0.msh", "tests/meshconv_test/src/prectangle_1.msh"])
def test_gmsh_reader_against_the_reference_python_reader(rel):
    """python/fluidity/diagnostics/gmshtools.py ReadMsh (the reference's own Python gmsh reader, imported
    unmodified) on ASCII and binary fixtures. Its package imports vtk at module level for unrelated VTU helpers;
    vtk is absent here, so a MagicMock stands in for it -- the gmsh parsing code never touches it."""
    import sys
    from unittest import mock
    stub_vtk = "vtk" not in sys.modules
    if stub_vtk:
        sys.modules["vtk"] = mock.MagicMock()
    sys.path.insert(0, os.path.join(REF, "python"))
    try:
        import fluidity.diagnostics.gmshtools as gmshtools
        path = os.path.join(REF, rel)
        cwd = os.getcwd()
        os.chdir(os.path.dirname(path))  # ReadMsh looks for a .halo file next to the mesh
        try:
            ref = gmshtools.ReadMsh(path)
        finally:
            os.chdir(cwd)
    finally:
        sys.path.remove(os.path.join(REF, "python"))
        if stub_vtk:
            del sys.modules["vtk"]
    g = fmt.read_gmsh(path)
    m = g.mesh
    assert ref.NodeCount() == m.n_nodes and ref.VolumeElementCount() == m.n_elements and ref.SurfaceElementCount() == len(g.sndgln)
    assert (np.array([ref.GetNodeCoord(i)[:m.dim] for i in range(m.n_nodes)]) == m.X).all()
    assert (np.array([ref.GetVolumeElement(e).GetNodes() for e in range(m.n_elements)]) + 1 == m.ndglno).all()
    if g.region_ids is not None:
        assert [ref.GetVolumeElement(e).GetIds()[0] for e in range(m.n_elements)] == g.region_ids.tolist()
    if len(g.sndgln):
        assert (np.array([ref.GetSurfaceElement(f).GetNodes() for f in range(len(g.sndgln))]) + 1 == g.sndgln).all()
        if g.boundary_ids is not None:
            assert [ref.GetSurfaceElement(f).GetIds()[0] for f in range(len(g.sndgln))] == g.boundary_ids.tolist()
        else:
            assert all(len(ref.GetSurfaceElement(f).GetIds()) == 0 for f in range(len(g.sndgln)))
        if g.element_owner is not None:
            assert [ref.GetSurfaceElement(f).GetIds()[3] for f in range(len(g.sndgln))] == g.element_owner.tolist()


@pytest.mark.parametrize("kind", ["random", "stripes"])
def test_reference_decomposition_writer_with_pathological_owner_maps(tmp_path, kind):
    """Scattered ownership (every rank neighbours every other, halos cover most of the mesh): still the same bytes as
    fldecomp's writer."""
    from oracle import ref_fldecomp as rf
    if not rf.available():
        pytest.skip("oracle/_ref/libref_fldecomp.so is built from /root/reference (build container only)")
    mesh = syn.shuffled(syn.box_mesh((5, 4, 4)), seed=3) if kind == "random" else syn.box_mesh((7, 6), seed=2)
    nparts = 5 if kind == "random" else 3
    owner = np.random.default_rng(1).integers(0, nparts, mesh.n_nodes) if kind == "random" else np.arange(mesh.n_nodes) % nparts
    sn, _ = syn.boundary_faces(mesh)
    bid = np.ones(len(sn), dtype=np.int32)
    theirs, ours = str(tmp_path / "theirs"), str(tmp_path / "ours")
    rf.write_partitions(theirs, mesh, owner, nparts, sn, bid)
    fmt.write_decomposition(ours, part.partition_by_owner(mesh, owner, nparts, sn, bid), style="fldecomp")
    for r in range(nparts):
        for ext in (".msh", ".halo"):
            assert open(fmt.parallel_filename(theirs, r, ext), "rb").read() == open(fmt.parallel_filename(ours, r, ext), "rb").read()


# ---- .stat files ---------------------------------------------------------------------------------------------------
def _same_tree(mine, ref, path=""):
    n = 0
    for k, v in ref.items():
        if isinstance(v, dict):
            n += _same_tree(mine[k], v, path + "/" + k)
        else:
            assert np.array_equal(np.asarray(mine[k]), np.asarray(v)), path + "/" + k
            n += 1
    return n


@needs_reference
@pytest.mark.parametrize("rel", ["tests/Instability/Reference.stat",
                                 "tests/backward_facing_step_2d_zoltan_sam/sam_output/backward_facing_step_2d_sam.stat"])
def test_read_stat_equals_the_reference_stat_parser(rel):
    """formats.read_stat against the reference's own python/fluidity_tools.py stat_parser (imported unmodified; its
    module imports vtk for unrelated helpers, absent here: a MagicMock stands in) on the reference's .stat fixtures."""
    import sys
    from unittest import mock
    path = os.path.join(REF, rel)
    stub_vtk = "vtk" not in sys.modules
    if stub_vtk:
        sys.modules["vtk"] = mock.MagicMock()
    sys.path.insert(0, os.path.join(REF, "python"))
    try:
        import fluidity_tools
        ref = fluidity_tools.stat_parser(path)
    finally:
        sys.path.remove(os.path.join(REF, "python"))
        if stub_vtk:
            del sys.modules["vtk"]
    mine = fmt.read_stat(path)
    assert _same_tree(mine, ref) >= 20
    sub = fmt.read_stat(path, subsample=3)
    assert np.array_equal(sub["ElapsedTime"]["value"], mine["ElapsedTime"]["value"][::3])


def test_read_stat_plain_and_binary(tmp_path):
    header = ('<header>\n<constant name="FluidityVersion" type="string" value="x" />\n%s'
              '<field column="1" name="ElapsedTime" statistic="value"/>\n'
              '<field column="2" name="Velocity" statistic="max" material_phase="Water" components="3"/>\n'
              '<field column="5" name="Tracer" statistic="l2norm" material_phase="Water"/>\n</header>\n')
    data = np.arange(20, dtype=np.float64).reshape(4, 5) * 0.25
    p = tmp_path / "a.stat"
    p.write_text(header % "" + "".join(" ".join("%.17g" % v for v in row) + "\n" for row in data))
    s = fmt.read_stat(str(p))
    assert np.array_equal(s["ElapsedTime"]["value"], data[:, 0])
    assert np.array_equal(s["Water"]["Velocity"]["max"], data[:, 1:4].T) and np.array_equal(s["Water"]["Tracer"]["l2norm"], data[:, 4])
    assert s["__constants__"]["FluidityVersion"] == ("string", "x")
    b = tmp_path / "b.stat"
    b.write_text(header % ('<constant name="format" type="string" value="binary" />\n'
                           '<constant name="real_size" type="integer" value="8" />\n'
                           '<constant name="integer_size" type="integer" value="4" />\n'))
    np.concatenate([data.ravel(), [1.0, 2.0]]).tofile(str(b) + ".dat")  # + an incomplete last line, ignored
    sb = fmt.read_stat(str(b))
    assert np.array_equal(sb["Water"]["Velocity"]["max"], data[:, 1:4].T) and len(sb["ElapsedTime"]["value"]) == 4
    bad = tmp_path / "c.stat"
    bad.write_text(header % "" + "1 2 3\n")
    with pytest.raises(fmt.FormatError):
        fmt.read_stat(str(bad))
