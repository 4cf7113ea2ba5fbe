"""Surface-element loops and strong Dirichlet conditions on the device, through the C ABI, against the
oracle. Needs a B200. The per-face arithmetic is additionally checked on the CPU by
tests/test_surface_math.py; here the kernels, the CSR position search and the atomics around it."""
import numpy as np
import pytest

from conftest import load_golden_mesh, rel_err, row_rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make(mesh, fs, scatter=None):
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim))
    asm.build_sparsity()
    asm.set_fields(fs)
    if scatter is not None:
        asm.set_scatter(scatter)
    sn, fe = syn.boundary_faces(mesh)
    asm.set_surface(sn, fe, tables.p1_face_tables(mesh.dim))
    return asm, sn, fe


def meshes():
    return {"box2": syn.box_mesh((9, 7), seed=3), "box3": syn.box_mesh((5, 4, 6), seed=5),
            "cube-parallel": load_golden_mesh("cube-parallel"), "cavity": load_golden_mesh("square-cavity-2d")}


SCATTERS = [pytest.param(abi.SCATTER_ATOMIC, id="atomic"), pytest.param(abi.SCATTER_GATHER, id="gather"),
            pytest.param(abi.SCATTER_STRIP, id="strip")]


@pytest.mark.parametrize("scatter", SCATTERS)
@pytest.mark.parametrize("name", ["box2", "box3", "cube-parallel", "cavity"])
def test_tracer_surface_loop_and_dirichlet(orc, scatter, name):
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    asm, sn, fe = make(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    rng = np.random.default_rng(7)
    nf, sloc = len(fe), mesh.dim
    t_bc, t_bc_2 = rng.uniform(size=(nf, sloc)), rng.uniform(0.5, 2.0, size=(nf, sloc))
    nodes = np.unique(sn[: nf // 3].ravel())
    vals = rng.uniform(size=len(nodes))
    cases = [(abi.common_advdiff_opts(), rng.choice([0, 1, 3, 4], size=nf)),                                   # Neumann / Robin
             (abi.common_advdiff_opts(integrate_advection_by_parts=1, beta=0.25), rng.choice([0, 1, 3, 4], size=nf)),
             (abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0), rng.choice([0, 2, 3], size=nf)),
             (abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0, theta=0.0), rng.choice([0, 2], size=nf))]
    for o, bt in cases:
        ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
        orc.assemble_advdiff_surface(mesh, fs, o, findrm, colm, sn, fe, bt, t_bc, t_bc_2, ref["matrix"], ref["rhs"])
        orc.apply_dirichlet_scalar(nodes, vals, fs.get(abi.F_T)[0], o.dt, ref["rhs"])
        asm.advdiff_dev(o)
        asm.advdiff_surface_dev(o, bt, t_bc, t_bc_2)
        asm.advdiff_dirichlet_dev(nodes, vals, o.dt)
        got = asm.advdiff_fetch()
        assert rel_err(got["matrix"], ref["matrix"]) < TOL and row_rel_err(got["matrix"], ref["matrix"], findrm) < TOL
        assert rel_err(got["rhs"], ref["rhs"]) < TOL
    # without a time step the Dirichlet rows receive the boundary value itself
    asm.advdiff_dev(cases[0][0])
    asm.advdiff_dirichlet_dev(nodes, vals)
    assert (asm.advdiff_fetch()["rhs"][nodes - 1] == vals).all()


@pytest.mark.parametrize("name", ["box2", "box3"])
def test_by_parts_identity_on_the_device(name):
    """by-parts volume loop + face loop (no boundary condition) == plain volume loop: ties the face kernel's
    normals, measures and face mass matrices to the element kernels without the oracle."""
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.array([1.3]), abi.FIELD_CONSTANT)
    asm, sn, fe = make(mesh, fs)
    plain = asm.advdiff(abi.common_advdiff_opts(beta=0.3))
    o = abi.common_advdiff_opts(beta=0.3, integrate_advection_by_parts=1)
    asm.advdiff_dev(o)
    asm.advdiff_surface_dev(o, np.zeros(len(fe), dtype=np.int32))
    got = asm.advdiff_fetch()
    assert rel_err(got["matrix"], plain["matrix"]) < TOL and rel_err(got["rhs"], plain["rhs"]) < TOL
    pm = asm.momentum(abi.common_momentum_opts(beta=0.4))
    om = abi.common_momentum_opts(beta=0.4, integrate_advection_by_parts=1)
    asm.momentum_dev(om)
    asm.momentum_surface_dev(om, np.zeros((len(fe), mesh.dim), dtype=np.int32))
    gm = asm.momentum_fetch()
    for d in range(mesh.dim):
        assert rel_err(gm["big_m"][d], pm["big_m"][d]) < TOL and rel_err(gm["rhs"][:, d], pm["rhs"][:, d]) < TOL
    assert asm.momentum_identical_blocks()


@pytest.mark.parametrize("scatter", SCATTERS)
@pytest.mark.parametrize("name", ["box2", "box3", "cube-parallel"])
def test_momentum_surface_loop(orc, scatter, name):
    mesh = meshes()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    asm, sn, fe = make(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    rng = np.random.default_rng(9)
    nf = len(fe)
    vbc = rng.uniform(-1, 1, size=(nf, dim, dim))
    vt = rng.choice([0, 1, 2, 3, 4, 5], size=(nf, dim))
    pt = rng.choice([0, 1], size=nf)
    for o in (abi.common_momentum_opts(integrate_advection_by_parts=1, beta=0.5), abi.common_momentum_opts(),
              abi.common_momentum_opts(integrate_advection_by_parts=1, have_absorption=1)):
        for ptype in (None, pt):
            ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
            orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, vt, vbc, ref["big_m"], ref["rhs"], ptype)
            asm.momentum_dev(o)
            asm.momentum_surface_dev(o, vt, vbc, ptype)
            got = asm.momentum_fetch()
            for d in range(dim):
                assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
                assert row_rel_err(got["big_m"][d], ref["big_m"][d], findrm) < TOL
                assert rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL
                assert rel_err(got["masslump"][:, d], ref["masslump"][:, d]) < TOL
    # weak Dirichlet on some components only: the diagonal blocks are no longer identical
    o = abi.common_momentum_opts(integrate_advection_by_parts=1)
    asm.momentum_dev(o)
    assert asm.momentum_identical_blocks()
    asm.momentum_surface_dev(o, vt, vbc)
    assert not asm.momentum_identical_blocks()


@pytest.mark.parametrize("scatter", SCATTERS)
@pytest.mark.parametrize("name", ["box2", "box3", "cube-parallel"])
@pytest.mark.parametrize("case", ["lumped", "lumped_pressure_corrected", "consistent"])
def test_free_surface_stabilisation(orc, scatter, name, case):
    """Momentum_CG.F90:1108-1176: faces of type FREE_SURFACE with have_fs_stab add dt g fs_sf (n . k) k mass terms to
    big_m, rhs and -- pressure-corrected absorption on a lumped mass -- masslump."""
    mesh = meshes()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    asm, sn, fe = make(mesh, fs, scatter)
    findrm, colm, _ = asm.get_sparsity()
    rng = np.random.default_rng(4)
    nf = len(fe)
    vt = np.zeros((nf, dim), dtype=np.int32)
    top = rng.random(nf) < 0.5
    vt[top] = abi.VBC_FREE_SURFACE
    pc = int(case == "lumped_pressure_corrected")
    o = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=0.6, lump_mass=int(case != "consistent"),
                                 have_absorption=pc, lump_absorption=pc, pressure_corrected_absorption=pc,
                                 integrate_advection_by_parts=1)
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    before = ref["masslump"].copy()
    orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, vt, np.zeros((nf, dim, dim)), ref["big_m"], ref["rhs"],
                                  masslump=ref["masslump"])
    assert (np.abs(ref["masslump"] - before).max() > 0) == bool(pc)
    asm.momentum_dev(o)
    asm.momentum_surface_dev(o, vt)
    got = asm.momentum_fetch()
    for d in range(dim):
        assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
        assert row_rel_err(got["big_m"][d], ref["big_m"][d], findrm) < TOL
        assert rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL
        assert rel_err(got["masslump"][:, d], ref["masslump"][:, d]) < TOL
    assert not asm.momentum_identical_blocks()
    # the reference exits on a consistent mass with pressure-corrected absorption (:1161-1163): so does the library
    bad = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=0.6, lump_mass=0, have_absorption=1,
                                   pressure_corrected_absorption=1)
    asm.momentum_dev(bad)
    with pytest.raises(cgasm.CgasmError) as e:
        asm.momentum_surface_dev(bad, vt)
    assert e.value.code == abi.EUNSUPPORTED


def test_surface_calls_refuse_what_they_must():
    mesh = syn.box_mesh((3, 3), seed=1)
    fs = syn.standard_fields(mesh)
    asm = cgasm.Assembler(mesh, tables.p1_tables(2))
    asm.build_sparsity()
    asm.set_fields(fs)
    sn, fe = syn.boundary_faces(mesh)
    o = abi.common_advdiff_opts()

    def code(fn, *a):
        with pytest.raises(cgasm.CgasmError) as ei:
            fn(*a)
        return ei.value.code

    asm.advdiff_dev(o)
    assert code(asm.advdiff_surface_dev, o, np.zeros(len(fe))) == abi.ESTATE           # no surface yet
    bad = sn.copy()
    bad[0, 0] = (set(range(1, mesh.n_nodes + 1)) - set(mesh.ndglno[fe[0] - 1].tolist())).pop()
    assert code(asm.set_surface, bad, fe, tables.p1_face_tables(2)) == abi.EARG        # face node outside its element
    assert code(asm.set_surface, sn, fe, tables.p1_tables(2)) == abi.EUNSUPPORTED      # wrong face element
    asm.set_surface(sn, fe, tables.p1_face_tables(2))
    assert code(asm.advdiff_surface_dev, o, np.full(len(fe), abi.TBC_WEAKDIRICHLET), np.zeros((len(fe), 2))) == abi.EUNSUPPORTED
    assert code(asm.advdiff_surface_dev, o, np.full(len(fe), 9)) == abi.EARG
    assert code(asm.advdiff_surface_dev, o, np.full(len(fe), abi.TBC_NEUMANN)) == abi.EARG   # values missing
    assert code(asm.advdiff_dirichlet_dev, [mesh.n_nodes + 1], [0.0], 0.1) == abi.EARG
    assert code(asm.advdiff_dirichlet_dev, [1], [0.0], 0.0) == abi.EARG
    om = abi.common_momentum_opts()
    assert code(asm.momentum_surface_dev, om, np.zeros((len(fe), 2))) == abi.ESTATE    # no momentum result yet
    asm.momentum_dev(om)
    assert code(asm.momentum_surface_dev, om, np.full((len(fe), 2), abi.VBC_FLUX)) == abi.EARG   # values missing
    assert code(asm.momentum_surface_dev, om, np.full((len(fe), 2), 6)) == abi.EARG
    # nothing was added by the refused calls
    ref = asm.advdiff(o)
    asm.advdiff_dev(o)
    asm.advdiff_surface_dev(abi.common_advdiff_opts(have_diffusivity=0), np.full(len(fe), abi.TBC_NEUMANN))  # loop does not run
    got = asm.advdiff_fetch()
    # (two runs of the default ATOMIC scatter variant: equal up to the order of the atomic additions)
    assert rel_err(got["matrix"], ref["matrix"]) < 1e-14 and rel_err(got["rhs"], ref["rhs"]) < 1e-14


@pytest.mark.parametrize("name", ["box2", "box3", "cube-parallel"])
def test_continuity_by_parts_boundary_blocks(orc, name):
    """integrate_continuity_by_parts: the element loop makes -dshape_shape, the surface loop adds
    shape_shape_vector(p_shape, u_shape, detwei_bdy, normal_bdy) on every face that is neither no-normal-flow nor
    free-surface (Momentum_CG.F90:1073-1088). By the divergence theorem the two together equal the plain form on rows
    whose faces all take part; here simply against the oracle."""
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    asm, sn, fe = make(mesh, fs, abi.SCATTER_STRIP)
    findrm, colm, _ = asm.get_sparsity()
    dim, nf = mesh.dim, len(fe)
    rng = np.random.default_rng(11)
    bt = np.zeros((nf, dim), dtype=np.int32)
    kind = rng.choice([0, 1, 2, 4], size=nf)          # none / weak Dirichlet / no normal flow / free surface
    bt[:, 0] = kind
    bt[kind == 1, 1:] = 1
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1, integrate_continuity_by_parts=1, integrate_advection_by_parts=1)
    vbc = rng.uniform(size=(nf, dim, dim))
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, bt, vbc, ref["big_m"], ref["rhs"])
    orc.assemble_ct_surface(mesh, fs, o, findrm, colm, sn, fe, bt, ref["ct_m"])
    asm.momentum_dev(o)
    asm.momentum_surface_dev(o, bt, vbc)
    got = asm.momentum_fetch(want_masslump=True, want_ct=True)
    for d in range(dim):
        assert rel_err(got["ct_m"][d], ref["ct_m"][d]) < TOL
        assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
    assert rel_err(got["rhs"], ref["rhs"]) < TOL
    # pressure conditions with by-parts continuity stay with the Fortran loop
    with pytest.raises(cgasm.CgasmError) as ei:
        asm.momentum_surface_dev(o, bt, vbc, np.ones(nf, dtype=np.int32))
    assert ei.value.code == abi.EUNSUPPORTED


@pytest.mark.parametrize("name", ["box2", "box3", "cube-parallel"])
def test_vector_dirichlet_lifting_and_velocity_correction(orc, name):
    """Strong Dirichlet conditions on big_m (lift_boundary_conditions: MatZeroRowsColumns + fix_scaling) and
    correct_masslumped_velocity, both on the device-resident results, against the C restatements in the oracle."""
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    asm, sn, fe = make(mesh, fs, abi.SCATTER_STRIP)
    findrm, colm, _ = asm.get_sparsity()
    dim, nn = mesh.dim, mesh.n_nodes
    rng = np.random.default_rng(13)
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1, have_absorption=1)
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    # conditions: all components on a third of the boundary nodes, the first component only on another third,
    # a few pairs listed twice (the later value wins)
    bnodes = np.unique(sn.ravel())
    a, b = bnodes[: len(bnodes) // 3], bnodes[len(bnodes) // 3: 2 * len(bnodes) // 3]
    nodes = np.concatenate([np.repeat(a, dim), b, a[:3]])
    comps = np.concatenate([np.tile(np.arange(1, dim + 1), len(a)), np.ones(len(b), dtype=np.int64), np.ones(3, dtype=np.int64)])
    vals = rng.uniform(-1, 1, size=len(nodes))
    want_rhs = ref["rhs"].copy()
    want_rhs[nodes - 1, comps - 1] = vals          # numpy fancy assignment: later entries win, like the reference's set()
    want_bigm = ref["big_m"].copy()
    orc.lift_boundary_conditions(findrm, colm, want_bigm, want_rhs, nodes, comps)
    asm.momentum_dev(o)
    asm.momentum_dirichlet_dev(nodes, comps, vals)
    got = asm.momentum_fetch(want_masslump=True, want_ct=True)
    for d in range(dim):
        assert rel_err(got["big_m"][d], want_bigm[d]) < TOL
        lifted = np.zeros(nn, dtype=bool)
        lifted[nodes[comps == d + 1] - 1] = True
        rows = np.repeat(np.arange(nn), np.diff(findrm))
        offdiag = (colm - 1) != rows
        assert (got["big_m"][d][lifted[rows] & offdiag] == 0.0).all() and (got["big_m"][d][lifted[colm - 1] & offdiag] == 0.0).all()
    assert rel_err(got["rhs"], want_rhs) < TOL
    assert not asm.momentum_identical_blocks()
    # velocity correction with the resident ct_m and with a host copy
    iml = 1.0 / ref["masslump"]
    dp = rng.uniform(-1, 1, size=nn)
    u0 = fs.get(abi.F_NU)[0].copy()
    want_u = orc.correct_masslumped_velocity(findrm, colm, ref["ct_m"], iml, dp, u0.copy())
    got_u = asm.correct_masslumped_velocity(iml, dp, u0.copy())
    assert rel_err(got_u, want_u) < TOL
    got_u2 = asm.correct_masslumped_velocity(iml, dp, u0.copy(), ct_m=ref["ct_m"])
    assert (got_u2 == want_u).all()      # same inputs, same summation order, explicit multiplies and adds: bitwise


def test_many_handles_in_one_process_give_the_same_bits(orc):
    """Regression test of the upload race found at the end of round 2: plan arrays went up with plain cudaMemcpy from
    pageable memory, which returns once the data is STAGED -- its DMA is ordered in the legacy stream only, the handles'
    streams are non-blocking, so a kernel launched right after could read a destination whose tail had not landed. It took
    ~100 handles in one process to show (wrong rows at the high node numbers of this mesh on the two-pass GATHER path,
    sometimes an illegal address in gather_pairs_kernel). scripts/stress_surface.py is the long version."""
    mesh = load_golden_mesh("cube-parallel")
    fs = syn.standard_fields(mesh)
    o = abi.common_momentum_opts(lump_mass=1, integrate_advection_by_parts=1)   # outside the STRIP kernels: pair lists + two passes
    first = None
    keep = []
    for it in range(160):
        asm = cgasm.Assembler(mesh, tables.p1_tables(3))
        asm.build_sparsity()
        asm.set_fields(fs)
        asm.set_scatter(abi.SCATTER_GATHER)
        asm.momentum_dev(o)
        got = asm.momentum_fetch()
        if first is None:
            first = got
            findrm, colm, _ = asm.get_sparsity()
            ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
            for d in range(3):
                assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
        else:
            for k in ("big_m", "rhs", "masslump"):
                assert np.array_equal(got[k], first[k]), "handle %d: %s differs from the first handle's" % (it, k)
        if it % 3 == 0:
            keep.append(asm)   # some handles stay open, as in the rest of the suite
        else:
            asm.close()
    for asm in keep:
        asm.close()
