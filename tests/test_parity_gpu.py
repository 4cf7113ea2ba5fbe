"""Parity of the CUDA path (through the C ABI) against the CPU oracle. Needs a B200.

Bar (north_star): findrm/colm/centrm and colouring bit-exact; element matrices, assembled
CSR values, rhs and lumped mass within 1e-12 relative (norm-relative per block and per row,
SURVEY.md 8(c)(ii)) because summation order differs.
"""
import numpy as np
import pytest

from conftest import load_golden_mesh, rel_err, row_rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
from variants import momentum_variants, advdiff_variants, boussinesq_variants, fields_for

pytestmark = pytest.mark.gpu
TOL = 1e-12

ALL_SCATTERS = [pytest.param(abi.SCATTER_ATOMIC, id="atomic"), pytest.param(abi.SCATTER_COLOURED, id="coloured"),
                pytest.param(abi.SCATTER_WARPAGG, id="warpagg"), pytest.param(abi.SCATTER_TILED, id="tiled"),
                pytest.param(abi.SCATTER_GATHER, id="gather"), pytest.param(abi.SCATTER_STRIP, id="strip")]


def make_asm(mesh, fields=None, scatter=None):
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim))
    asm.build_sparsity()
    if fields is not None:
        asm.set_fields(fields)
    if scatter is not None:
        asm.set_scatter(scatter)
    return asm


def check_momentum(got, ref, findrm, dim):
    for d in range(dim):
        assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
        assert row_rel_err(got["big_m"][d], ref["big_m"][d], findrm) < TOL
        assert rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL
        if ref.get("masslump") is not None and got.get("masslump") is not None:
            assert rel_err(got["masslump"][:, d], ref["masslump"][:, d]) < TOL
        if ref.get("ct_m") is not None and got.get("ct_m") is not None:
            assert rel_err(got["ct_m"][d], ref["ct_m"][d]) < TOL


# ---- sparsity / colouring: bit exact --------------------------------------------------------
@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "square-cavity-2d", "2d_square"])
def test_sparsity_bit_exact_reference_meshes(orc, name):
    mesh = load_golden_mesh(name)
    asm = make_asm(mesh)
    f, c, ce = asm.get_sparsity()
    of, oc, oce = orc.make_sparsity(mesh)
    assert (f == of).all() and (c == oc).all() and (ce == oce).all()


@pytest.mark.parametrize("shape", [(5, 4, 3), (7, 6), (1, 1, 1), (1, 1)])
def test_sparsity_bit_exact_box_and_shuffled(orc, shape):
    mesh = syn.box_mesh(shape)
    for m in (mesh, syn.shuffled(mesh, seed=3)):
        asm = make_asm(m)
        f, c, ce = asm.get_sparsity()
        of, oc, oce = orc.make_sparsity(m)
        assert (f == of).all() and (c == oc).all() and (ce == oce).all()


@pytest.mark.parametrize("name", ["square-cavity-2d", "cube-parallel"])
def test_colouring_bit_exact(orc, name):
    mesh = load_golden_mesh(name)
    asm = make_asm(mesh)
    nc = asm.build_colouring()
    ptr, els = asm.get_colouring(nc)
    col, onc = orc.colour_elements(mesh)
    optr, oels = orc.colour_sets(col, onc)
    assert nc == onc and (ptr == optr).all() and (els == oels).all()


def test_reference_sparsity_and_colouring_can_be_adopted(orc):
    mesh = load_golden_mesh("cube-parallel")
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    col, nc = orc.colour_elements(mesh)
    asm = cgasm.Assembler(mesh, tables.p1_tables(3))
    asm.set_sparsity(findrm, colm)
    asm.set_colouring(*orc.colour_sets(col, nc))
    asm.set_fields(fs)
    asm.set_scatter(abi.SCATTER_COLOURED)
    o = abi.common_momentum_opts()
    check_momentum(asm.momentum(o), orc.assemble_momentum(mesh, fs, o, findrm, colm), findrm, 3)
    # an invalid colouring (everything in one colour) must be refused, not raced
    with pytest.raises(cgasm.CgasmError):
        asm.set_colouring(np.array([1, mesh.n_elements + 1]), np.arange(1, mesh.n_elements + 1))
    # a pattern with unsorted rows must be refused
    bad = colm.copy()
    bad[0], bad[1] = bad[1], bad[0]
    with pytest.raises(cgasm.CgasmError):
        asm.set_sparsity(findrm, bad)


# ---- element matrices ---------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(momentum_variants().keys()))
def test_momentum_element_matrices(orc, dim, variant):
    mesh = syn.box_mesh((4, 3, 5)[:dim], seed=21)
    o = momentum_variants()[variant]
    fs = fields_for(mesh, variant)
    asm = make_asm(mesh, fs)
    rng = np.random.default_rng(0)
    for ele in [1, mesh.n_elements] + rng.integers(1, mesh.n_elements + 1, size=6).tolist():
        T, r, ml, gp = asm.momentum_element(o, ele)
        oT, orr, oml, ogp = orc.momentum_element(mesh, fs, o, ele)
        scale = max(np.abs(oT).max(), 1e-300)
        assert np.abs(T - oT).max() <= TOL * scale, (variant, ele)
        assert rel_err(r, orr) < TOL
        assert rel_err(ml, oml) < TOL
        if o.assemble_ct_matrix_here:
            assert rel_err(gp, ogp) < TOL


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(advdiff_variants().keys()))
def test_advdiff_element_matrices(orc, dim, variant):
    mesh = syn.box_mesh((4, 3, 5)[:dim], seed=22)
    o = advdiff_variants()[variant]
    fs = fields_for(mesh, variant)
    asm = make_asm(mesh, fs)
    rng = np.random.default_rng(1)
    for ele in [1, mesh.n_elements] + rng.integers(1, mesh.n_elements + 1, size=6).tolist():
        A, r = asm.advdiff_element(o, ele)
        oA, orr = orc.advdiff_element(mesh, fs, o, ele)
        assert rel_err(A, oA) < TOL, (variant, ele)
        assert rel_err(r, orr) < TOL


# ---- assembled values, every scatter variant ----------------------------------------------
@pytest.mark.parametrize("scatter", ALL_SCATTERS)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(momentum_variants().keys()))
def test_momentum_assembly(orc, scatter, dim, variant):
    mesh = syn.box_mesh((6, 5, 4)[:dim], seed=31)
    o = momentum_variants()[variant]
    if o.stabilisation_scheme and scatter not in (abi.SCATTER_ATOMIC, abi.SCATTER_GATHER, abi.SCATTER_STRIP):
        asm = make_asm(mesh, fields_for(mesh, variant), scatter)
        with pytest.raises(cgasm.CgasmError) as ei:
            asm.momentum(o)
        assert ei.value.code == abi.EUNSUPPORTED  # the caller keeps the Fortran loop (or picks GATHER)
        return
    fs = fields_for(mesh, variant)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    got = asm.momentum(o)
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=bool(o.assemble_ct_matrix_here))
    check_momentum(got, ref, findrm, dim)


@pytest.mark.parametrize("scatter", ALL_SCATTERS)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(advdiff_variants().keys()))
def test_advdiff_assembly(orc, scatter, dim, variant):
    mesh = syn.box_mesh((6, 5, 4)[:dim], seed=32)
    o = advdiff_variants()[variant]
    if o.stabilisation_scheme and scatter not in (abi.SCATTER_ATOMIC, abi.SCATTER_GATHER, abi.SCATTER_STRIP):
        asm = make_asm(mesh, fields_for(mesh, variant), scatter)
        with pytest.raises(cgasm.CgasmError) as ei:
            asm.advdiff(o)
        assert ei.value.code == abi.EUNSUPPORTED
        return
    fs = fields_for(mesh, variant)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    got = asm.advdiff(o)
    ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    assert rel_err(got["matrix"], ref["matrix"]) < TOL
    assert row_rel_err(got["matrix"], ref["matrix"], findrm) < TOL
    assert rel_err(got["rhs"], ref["rhs"]) < TOL


@pytest.mark.parametrize("scatter", ALL_SCATTERS)
@pytest.mark.parametrize("name", ["cube-parallel", "2d_square", "cube.1"])
def test_reference_fixture_meshes(orc, scatter, name):
    # unstructured gmsh meshes from the reference's tests/data
    mesh = load_golden_mesh(name)
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    o = abi.common_momentum_opts(have_absorption=1)
    check_momentum(asm.momentum(o), orc.assemble_momentum(mesh, fs, o, findrm, colm), findrm, mesh.dim)
    oa = abi.common_advdiff_opts(have_source=1)
    got, ref = asm.advdiff(oa), orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    assert rel_err(got["matrix"], ref["matrix"]) < TOL and rel_err(got["rhs"], ref["rhs"]) < TOL


@pytest.mark.parametrize("scatter", ALL_SCATTERS)
def test_shuffled_numbering_s3_small(orc, scatter):
    # S3-small (32^3 x 6 = 196 608 tets) with nodes and elements randomly renumbered: no
    # structure for the kernels to lean on
    mesh = syn.shuffled(syn.box_mesh((32, 32, 32)), seed=5)
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    o = abi.common_momentum_opts()
    check_momentum(asm.momentum(o), orc.assemble_momentum(mesh, fs, o, findrm, colm), findrm, 3)
    oa = abi.common_advdiff_opts()
    got, ref = asm.advdiff(oa), orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    assert rel_err(got["matrix"], ref["matrix"]) < TOL and rel_err(got["rhs"], ref["rhs"]) < TOL


@pytest.mark.parametrize("scatter", ALL_SCATTERS)
@pytest.mark.parametrize("dim", [2, 3])
def test_delaunay_unstructured_mesh(orc, scatter, dim):
    """A Delaunay triangulation of graded random points (what examples/flow_past_sphere_Re100 looks like to the
    kernels: node degrees 8-58, row lengths and strip lengths all different, both local orientations), with the
    anisotropic-viscosity option set of that example, the common set and the constant-density absorption set."""
    mesh = syn.delaunay_mesh(4000 if dim == 3 else 6000, dim=dim, seed=3)
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    of, oc, _ = orc.make_sparsity(mesh)
    assert (findrm == of).all() and (colm == oc).all()
    oa = abi.common_advdiff_opts()
    got, ref = asm.advdiff(oa), orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    assert rel_err(got["matrix"], ref["matrix"]) < TOL and rel_err(got["rhs"], ref["rhs"]) < TOL
    o = abi.common_momentum_opts()
    check_momentum(asm.momentum(o), orc.assemble_momentum(mesh, fs, o, findrm, colm), findrm, dim)
    fs.set(abi.F_VISCOSITY, syn.aniso_tensor(dim), abi.FIELD_CONSTANT)
    fs.set(abi.F_DENSITY, np.array([1.1]), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_VISCOSITY, syn.aniso_tensor(dim), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_DENSITY, np.array([1.1]), abi.FIELD_CONSTANT)
    for o in (abi.common_momentum_opts(viscosity_shape=abi.TENSOR_FULL, have_gravity=0),
              abi.common_momentum_opts(have_absorption=1, have_gravity=0)):
        check_momentum(asm.momentum(o), orc.assemble_momentum(mesh, fs, o, findrm, colm), findrm, dim)


def test_occupancy_classes_change_where_a_block_runs_not_what_it_computes(orc, monkeypatch):
    """On an unstructured mesh the most crowded row block sets the chunk stride and the accumulator length of the staged
    STRIP kernels; the blocks that fit smaller ones are launched as a class of their own at more blocks per SM
    (strip_staged.cuh, staged_classes). Same node lists, same entries: the results are bitwise those of the single launch
    (CGASM_STRIP_NOCLASSES), and they match the oracle."""
    mesh = syn.delaunay_mesh(30000, dim=3, seed=5)
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, abi.SCATTER_STRIP)
    findrm, colm, _ = asm.get_sparsity()
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts(have_absorption=1, have_source=1)
    monkeypatch.setenv("CGASM_STRIP_NOCLASSES", "1")
    l0 = asm.launch_count()
    m_one, a_one = asm.momentum(om), asm.advdiff(oa)
    one = asm.launch_count() - l0
    monkeypatch.delenv("CGASM_STRIP_NOCLASSES")
    l0 = asm.launch_count()
    m_cls, a_cls = asm.momentum(om), asm.advdiff(oa)
    cls = asm.launch_count() - l0
    assert asm.last_path() == ("strip_staged", "strip_staged")
    assert cls > one, "this mesh is meant to split into occupancy classes (plan: %s)" % asm.plan_stats()
    for k in ("big_m", "rhs", "masslump"):
        assert np.array_equal(m_one[k], m_cls[k])
    for k in ("matrix", "rhs"):
        assert np.array_equal(a_one[k], a_cls[k])
    check_momentum(m_cls, orc.assemble_momentum(mesh, fs, om, findrm, colm), findrm, 3)
    ref = orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    assert rel_err(a_cls["matrix"], ref["matrix"]) < TOL and rel_err(a_cls["rhs"], ref["rhs"]) < TOL


# ---- golden vectors from the reference's own Python element machinery ------------------------------
@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "square-cavity-2d", "prectangle_0"])
def test_element_matrices_match_reference_python(name):
    """tests/golden/pyref_*.npz were computed by python/fluidity/state_types.py of the reference
    (imported unmodified, tests/golden/make_pyref_golden.py): mass, lumped mass, tracer mass and
    grad_p_u_mat of every element, here against cgasm_momentum_element / cgasm_advdiff_element."""
    import pyref_checks as pc
    mesh, fs, z = pc.load(name)
    asm = make_asm(mesh, fs)
    worst = pc.check_elements(mesh, fs, z, asm.momentum_element, asm.advdiff_element, zero_tol=TOL)
    assert worst < TOL
    # the common option set (+ absorption + sources) from reference-computed ingredients
    assert pc.check_composed(mesh, fs, z, asm.momentum_element, asm.advdiff_element) < TOL


@pytest.mark.parametrize("scatter", ALL_SCATTERS)
@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "square-cavity-2d", "prectangle_0"])
def test_assembled_mass_and_ct_match_reference_python(scatter, name):
    """Assembled lumped mass, ct_m blocks and the tracer mass matrix of every scatter variant against
    what the reference's Python accumulated with its own Field.addto."""
    import pyref_checks as pc
    mesh, fs, z = pc.load(name)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    pc.check_assembled(mesh, fs, z, findrm, colm, asm.momentum, asm.advdiff, zero_tol=TOL)
    # assembled CSR values and rhs of the common option set (+ absorption + sources)
    pc.check_assembled_composed(mesh, fs, z, findrm, colm, asm.momentum, asm.advdiff)


# ---- size-independent properties at a size the oracle would not finish quickly ---------------
@pytest.mark.parametrize("scatter", [pytest.param(abi.SCATTER_ATOMIC, id="atomic"), pytest.param(abi.SCATTER_TILED, id="tiled"),
                                     pytest.param(abi.SCATTER_GATHER, id="gather"), pytest.param(abi.SCATTER_STRIP, id="strip")])
def test_large_mesh_properties(scatter):
    mesh = syn.box_mesh((96, 96, 96), jitter=0.1)
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_T, np.ones(mesh.n_nodes))
    fs.set(abi.F_DENSITY, np.ones(1), abi.FIELD_CONSTANT)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, centrm = asm.get_sparsity()
    # (1) constants are in the kernel of advection(beta=0)+diffusion: tracer rhs == 0 for T == 1
    out = asm.advdiff(abi.common_advdiff_opts())
    assert np.abs(out["rhs"]).max() < 1e-12 * np.abs(out["matrix"]).max()
    # (2) the consistent mass matrix sums to the volume of the unit cube
    out = asm.advdiff(abi.common_advdiff_opts(have_advection=0, have_diffusivity=0))
    assert abs(out["matrix"].sum() - 1.0) < 1e-10
    # (3) lumped mass (rho = 1) sums to the volume too, identically for every component, and
    #     equals the big_m diagonal when nothing else is assembled
    o = abi.common_momentum_opts(exclude_advection=1, have_viscosity=0, have_gravity=0)
    out = asm.momentum(o)
    for d in range(3):
        assert abs(out["masslump"][:, d].sum() - 1.0) < 1e-10
        assert rel_err(out["big_m"][d][centrm - 1], out["masslump"][:, d]) < TOL
        assert np.abs(out["rhs"]).max() == 0.0
    # (4) linearity in oldu: rhs(2*oldu) - rhs(oldu) == rhs(oldu) - rhs(0)  (no mass in rhs)
    o = abi.common_momentum_opts(have_gravity=0)
    oldu, _ = fs.get(abi.F_OLDU)
    r1 = asm.momentum(o)["rhs"].copy()
    asm.set_field(abi.F_OLDU, 2 * oldu)
    r2 = asm.momentum(o)["rhs"].copy()
    assert rel_err(r2, 2 * r1) < 1e-11
    # (5) idempotence: same call twice gives identical sums for the deterministic variant
    a = asm.momentum(o)["big_m"].copy()
    b = asm.momentum(o)["big_m"]
    if scatter in (abi.SCATTER_TILED, abi.SCATTER_GATHER, abi.SCATTER_STRIP):
        assert (a == b).all()
    else:
        assert rel_err(a, b) < TOL


def test_strip_matches_gather_at_size():
    """STRIP and GATHER kernels are independent formulations of the same sums: at 64^3 (1.57 M tets,
    jittered, every boundary shape) they must agree to 1e-12, and STRIP must be bitwise reproducible."""
    mesh = syn.box_mesh((64, 64, 64), jitter=0.1)
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, abi.SCATTER_GATHER)
    findrm, _, _ = asm.get_sparsity()
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    gm = {k: v.copy() for k, v in asm.momentum(om).items() if v is not None}
    ga = {k: v.copy() for k, v in asm.advdiff(oa).items()}
    asm.set_scatter(abi.SCATTER_STRIP)
    sm = {k: v.copy() for k, v in asm.momentum(om).items() if v is not None}
    sa = {k: v.copy() for k, v in asm.advdiff(oa).items()}
    check_momentum(sm, gm, findrm, 3)
    assert rel_err(sa["matrix"], ga["matrix"]) < TOL and row_rel_err(sa["matrix"], ga["matrix"], findrm) < TOL
    assert rel_err(sa["rhs"], ga["rhs"]) < TOL
    sm2, sa2 = asm.momentum(om), asm.advdiff(oa)
    assert (sm2["big_m"] == sm["big_m"]).all() and (sm2["rhs"] == sm["rhs"]).all()
    assert (sa2["matrix"] == sa["matrix"]).all() and (sa2["rhs"] == sa["rhs"]).all()


def test_async_host_flavour_matches_blocking_calls():
    """cgasm_set_async: uploads and downloads are queued on two streams; after cgasm_synchronize the host
    buffers hold exactly what the blocking calls return, also when a loop is re-run while the previous
    download may still be in flight."""
    mesh = syn.box_mesh((12, 10, 8))
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, abi.SCATTER_STRIP)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    nu, _ = fs.get(abi.F_NU)
    T, _ = fs.get(abi.F_T)
    ref = []
    for scale in (1.0, 2.0):
        asm.set_field(abi.F_NU, scale * nu)
        asm.set_field(abi.F_T, scale * T)
        m = asm.momentum(om)
        a = asm.advdiff(oa)
        ref.append(({k: v.copy() for k, v in m.items() if v is not None}, {k: v.copy() for k, v in a.items()}))
    asm.set_async(True)
    nn, nnz = mesh.n_nodes, asm.nnz
    for rep in range(2):
        for scale, (rm, ra) in zip((1.0, 2.0), ref):
            out_m = dict(big_m=np.empty((1, nnz)), rhs=np.empty((nn, 3)), masslump=np.empty((nn, 3)))
            out_a = dict(matrix=np.empty(nnz), rhs=np.empty(nn))
            up_nu, up_T = scale * nu, scale * T  # must stay alive until synchronize
            asm.set_field(abi.F_NU, up_nu)
            asm.set_field(abi.F_T, up_T)
            asm.advdiff_dev(oa)
            asm.advdiff_fetch_into(out_a)
            assert asm.momentum_host(om, out_m) == 1
            asm.advdiff_dev(oa)  # re-run while the first tracer download may still be in flight
            asm.synchronize()
            assert (out_a["matrix"] == ra["matrix"]).all() and (out_a["rhs"] == ra["rhs"]).all()
            assert (out_m["big_m"][0] == rm["big_m"][0]).all() and (out_m["rhs"] == rm["rhs"]).all()
            assert (out_m["masslump"] == rm["masslump"]).all()
    asm.set_async(False)
    m = asm.momentum(om)
    assert (m["big_m"] == ref[1][0]["big_m"]).all()


def test_identical_blocks_partial_fetch(orc):
    mesh = syn.box_mesh((5, 4, 3))
    fs = syn.standard_fields(mesh)
    asm = make_asm(mesh, fs, abi.SCATTER_GATHER)
    findrm, colm, _ = asm.get_sparsity()
    o = abi.common_momentum_opts()
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    out = dict(big_m=np.empty((1, asm.nnz)), rhs=np.empty((mesh.n_nodes, 3)), masslump=np.empty((mesh.n_nodes, 3)))
    assert asm.momentum_host(o, out) == 1
    for d in range(3):  # one fetched block stands for all three
        assert rel_err(out["big_m"][0], ref["big_m"][d]) < TOL
    assert rel_err(out["rhs"], ref["rhs"]) < TOL and rel_err(out["masslump"], ref["masslump"]) < TOL
    o2 = abi.common_momentum_opts(have_absorption=1)
    out3 = dict(big_m=np.empty((3, asm.nnz)), rhs=np.empty((mesh.n_nodes, 3)), masslump=np.empty((mesh.n_nodes, 3)))
    assert asm.momentum_host(o2, out3) == 3
    ref2 = orc.assemble_momentum(mesh, fs, o2, findrm, colm)
    for d in range(3):
        assert rel_err(out3["big_m"][d], ref2["big_m"][d]) < TOL
    with pytest.raises(cgasm.CgasmError):
        asm.momentum_fetch_blocks(2, 2, np.empty((2, asm.nnz)))


def test_unsupported_options_are_refused():
    mesh = syn.box_mesh((2, 2, 2))
    asm = make_asm(mesh, syn.standard_fields(mesh))
    for kw in (dict(have_les=1), dict(stress_form=1), dict(have_coriolis=1), dict(move_mesh=1)):
        with pytest.raises(cgasm.CgasmError) as ei:
            asm.momentum(abi.common_momentum_opts(**kw))
        assert ei.value.code == abi.EUNSUPPORTED
    with pytest.raises(cgasm.CgasmError) as ei:
        asm.advdiff(abi.common_advdiff_opts(multiphase=1))
    assert ei.value.code == abi.EUNSUPPORTED


def test_missing_field_is_state_error():
    mesh = syn.box_mesh((2, 2, 2))
    asm = make_asm(mesh)
    with pytest.raises(cgasm.CgasmError) as ei:
        asm.momentum(abi.common_momentum_opts())
    assert ei.value.code == abi.ESTATE


@pytest.mark.parametrize("name", ["cube.1", "cube-parallel", "square-cavity-2d", "prectangle_0"])
def test_every_non_stabilised_variant_matches_reference_python(name):
    """The CUDA element routines (cgasm_momentum_element / cgasm_advdiff_element) against the `v_*` goldens: every
    non-stabilised option variant of tests/variants.py, contractions run by the reference's own Python loops."""
    import pyref_checks as pc
    mesh, fs, z = pc.load(name)
    asms = []

    def elements_for(f):
        asm = make_asm(mesh, f)
        asms.append(asm)
        return asm.momentum_element, asm.advdiff_element

    assert pc.check_variants(mesh, z, elements_for) < TOL


# ---- STRIP: absorption, sources, reference profile with a constant density -----------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("in_loop", [True, False])
@pytest.mark.parametrize("variant", list(boussinesq_variants().keys()))
def test_strip_additive_pass_constant_density(orc, dim, variant, in_loop, monkeypatch):
    """Constant density (every example config is Boussinesq): the STRIP variant assembles these option sets in its own
    kernels and must agree with the oracle like any other. The full absorption matrix is carried by the common kernel's
    own loop (strip_absorb.cu; in_loop) or, where that does not fit, by one more strip pass per component
    (strip_extra.cu; forced here with CGASM_STRIP_NO_ABSORB); the per-row quantities (lumped absorption, sources,
    reference profile) by strip_extra.cu's per-row pass."""
    mesh = syn.box_mesh((6, 5, 4)[:dim], seed=33)
    o = boussinesq_variants()[variant]
    full = bool(o.have_absorption and not o.lump_absorption)
    light = bool((o.have_absorption and o.lump_absorption) or o.have_source or (o.have_gravity and o.subtract_out_reference_profile))
    if not in_loop:
        if not full:
            pytest.skip("no full absorption matrix: same kernels as in_loop")
        monkeypatch.setenv("CGASM_STRIP_NO_ABSORB", "1")
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.array([1.3]), abi.FIELD_CONSTANT)
    if variant == "everything":
        fs.set(abi.F_VISCOSITY, syn.aniso_tensor(dim), abi.FIELD_CONSTANT)
    asm = make_asm(mesh, fs, abi.SCATTER_STRIP)
    findrm, colm, _ = asm.get_sparsity()
    l0 = asm.launch_count()
    got = asm.momentum(o)
    # not the two-pass GATHER staging path (element kernel + row kernel)
    assert asm.launch_count() - l0 == 1 + int(full and not in_loop) + int(light), "the option set did not take the STRIP kernels"
    assert asm.last_path()[0] == "strip_staged"
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_masslump=bool(o.assemble_inverse_masslump))
    check_momentum(got, ref, findrm, dim)
    # a second assembly overwrites, it does not accumulate on top of the first
    got2 = asm.momentum(o)
    for k in ("big_m", "rhs"):
        assert (got2[k] == got[k]).all()


# ---- both loops in one call ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("case", ["common", "excluded_mass_lumped_tracer", "no_gravity_no_ml", "fallback_absorption",
                                  "fallback_tensor"])
def test_fused_momentum_tracer_call_equals_the_two_calls(orc, dim, case):
    """cgasm_momentum_advdiff_dev: one fused STRIP kernel for the common option sets (same device functions, same
    operand order as the separate kernels: the results must be BITWISE those of the two calls), the two loops one
    after the other for anything else."""
    mesh = syn.box_mesh((7, 5, 4)[:dim], seed=35)
    fs = syn.standard_fields(mesh)
    cm, ca = abi.common_momentum_opts, abi.common_advdiff_opts
    om, oa, fused = {
        "common": (cm(), ca(), True),
        "excluded_mass_lumped_tracer": (cm(exclude_mass=1), ca(lump_mass=1), True),
        "no_gravity_no_ml": (cm(have_gravity=0, assemble_inverse_masslump=0), ca(have_diffusivity=0), True),
        "fallback_absorption": (cm(), ca(have_absorption=1), False),
        "fallback_tensor": (cm(viscosity_shape=abi.TENSOR_FULL), ca(), False),
    }[case]
    asm = make_asm(mesh, fs, abi.SCATTER_STRIP)
    findrm, colm, _ = asm.get_sparsity()
    asm.momentum_dev(om)
    asm.advdiff_dev(oa)
    want_ml = bool(om.assemble_inverse_masslump)
    sep_m, sep_a = asm.momentum_fetch(want_masslump=want_ml), asm.advdiff_fetch()
    l0 = asm.launch_count()
    asm.momentum_advdiff_dev(om, oa)
    assert asm.launch_count() - l0 == (1 if fused else 2)
    got_m, got_a = asm.momentum_fetch(want_masslump=want_ml), asm.advdiff_fetch()
    for k in ("big_m", "rhs") + (("masslump",) if want_ml else ()):
        assert (got_m[k] == sep_m[k]).all(), k
    for k in ("matrix", "rhs"):
        assert (got_a[k] == sep_a[k]).all(), k
    ref_m = orc.assemble_momentum(mesh, fs, om, findrm, colm, want_masslump=want_ml)
    ref_a = orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    check_momentum(got_m, ref_m, findrm, dim)
    assert rel_err(got_a["matrix"], ref_a["matrix"]) < TOL and rel_err(got_a["rhs"], ref_a["rhs"]) < TOL


# ---- the `mass` matrix (assemble_mass_matrix) and continuity by parts ------------------------------------------------
@pytest.mark.parametrize("scatter", [pytest.param(abi.SCATTER_ATOMIC, id="atomic"), pytest.param(abi.SCATTER_STRIP, id="strip")])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("case", ["plain", "lumped_excluded", "pressure_corrected_absorption"])
def test_assemble_mass_matrix(orc, scatter, dim, case):
    """Momentum_CG.F90:1567-1571 / :2073-2078: the consistent density-weighted mass matrix on every diagonal block,
    independent of what lump_mass / exclude_mass do to big_m, plus dt*theta*absorption_mat when pressure-corrected."""
    mesh = syn.box_mesh((5, 4, 3)[:dim], seed=36)
    fs = syn.standard_fields(mesh)
    c = abi.common_momentum_opts
    o = {"plain": c(assemble_mass_matrix=1, lump_mass=0),
         "lumped_excluded": c(assemble_mass_matrix=1, exclude_mass=1),
         "pressure_corrected_absorption": c(assemble_mass_matrix=1, have_absorption=1, lump_absorption=1,
                                            pressure_corrected_absorption=1)}[case]
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    got = asm.momentum(o)
    mass = asm.momentum_mass_fetch()
    ref_mass = orc.assemble_momentum_mass(mesh, fs, o, findrm, colm)
    for d in range(dim):
        assert rel_err(mass[d], ref_mass[d]) < TOL and row_rel_err(mass[d], ref_mass[d], findrm) < TOL
    # big_m itself is what it is without the flag
    o2 = abi.MomentumOpts.from_buffer_copy(o)
    o2.assemble_mass_matrix = 0
    check_momentum(got, orc.assemble_momentum(mesh, fs, o2, findrm, colm), findrm, dim)
    # and the matrix is not there when it was not asked for
    asm.momentum_dev(o2)
    with pytest.raises(cgasm.CgasmError):
        asm.momentum_mass_fetch()


@pytest.mark.parametrize("scatter", [pytest.param(abi.SCATTER_ATOMIC, id="atomic"), pytest.param(abi.SCATTER_GATHER, id="gather"),
                                     pytest.param(abi.SCATTER_STRIP, id="strip")])
@pytest.mark.parametrize("dim", [2, 3])
def test_continuity_by_parts_volume_form(orc, scatter, dim):
    """integrate_continuity_by_parts: ct_m's volume form is -dshape_shape (Momentum_CG.F90:1379-1383)."""
    mesh = syn.box_mesh((5, 4, 3)[:dim], seed=37)
    fs = syn.standard_fields(mesh)
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1, integrate_continuity_by_parts=1)
    asm = make_asm(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    got = asm.momentum(o)
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    check_momentum(got, ref, findrm, dim)
    for ele in (1, mesh.n_elements // 2, mesh.n_elements):
        _, _, _, gp = asm.momentum_element(o, ele)
        _, _, _, ogp = orc.momentum_element(mesh, fs, o, ele)
        assert rel_err(gp, ogp) < TOL
