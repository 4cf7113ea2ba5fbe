"""CPU checks of the STRIP variant's host side (fluidity_b200/csrc/strip_plan.cpp) and of the
algebra its kernels use (tests/strip_emulation.py follows strip.cu line by line in numpy):
every (row, element) pair of the reference's addto loop (femtools/Sparse_Tools.F90:2680-2703) is
computed exactly once, slots point at the right columns, and the emulated assembly matches the
oracle within the north-star tolerance."""
import numpy as np
import pytest

from conftest import load_golden_mesh, rel_err, row_rel_err
from fluidity_b200 import synthetic as syn, _abi as abi
import strip_emulation as se

TOL = 1e-12


def meshes():
    return {
        "box3": syn.box_mesh((3, 3, 2)),
        "box3_shuffled": syn.shuffled(syn.box_mesh((3, 2, 2)), seed=5),
        "box2": syn.box_mesh((5, 4)),
        "box2_shuffled": syn.shuffled(syn.box_mesh((4, 3)), seed=2),
        "cell3": syn.box_mesh((1, 1, 1)),
        "cell2": syn.box_mesh((1, 1)),
        "cube.1": load_golden_mesh("cube.1"),
        "2d_square": load_golden_mesh("2d_square"),
    }


@pytest.mark.parametrize("name", list(meshes().keys()) + ["cube-parallel"])
def test_strip_covers_every_pair_once(orc, name):
    mesh = load_golden_mesh(name) if name == "cube-parallel" else meshes()[name]
    findrm, colm, _ = orc.make_sparsity(mesh)
    row_ptr, ent = se.strip_plan(mesh)
    nd = mesh.ndglno
    incident = [[] for _ in range(mesh.n_nodes)]
    for e in range(mesh.n_elements):
        for v in nd[e]:
            incident[v - 1].append(frozenset(int(x) for x in nd[e]))
    w = mesh.dim
    for r in range(mesh.n_nodes):
        seen = []
        fifo = []
        for node1, meta in ent[row_ptr[r]:row_ptr[r + 1]]:
            assert colm[findrm[r] - 1 + (meta & 0xff)] == node1
            fifo = (fifo + [int(node1)])[-w:]
            if meta & se.COMPUTE:
                assert len(set(fifo)) == w and (r + 1) not in fifo
                seen.append(frozenset(fifo + [r + 1]))
        assert sorted(map(sorted, seen)) == sorted(map(sorted, incident[r]))


def test_strip_length_on_kuhn_meshes():
    # interior node of a Kuhn mesh: 24 elements in 30 pushes (three bands of eight triangles: the bounded search of
    # strip_plan.cpp finds them on structured meshes; the greedy alone needs 33, 29 is infeasible); a 2-D interior node: 6 in 7
    m3 = syn.box_mesh((4, 4, 4))
    rp, _ = se.strip_plan(m3)
    deg = np.bincount(m3.ndglno.ravel() - 1, minlength=m3.n_nodes)
    assert (np.diff(rp)[deg == 24] == 30).all()
    assert (np.diff(rp) >= deg + 2).all()
    m2 = syn.box_mesh((5, 5))
    rp, _ = se.strip_plan(m2)
    deg = np.bincount(m2.ndglno.ravel() - 1, minlength=m2.n_nodes)
    assert (np.diff(rp)[deg == 6] == 7).all()
    assert (np.diff(rp) <= deg + 2).all()  # 2-D links are paths or cycles: at most one restart


def test_strip_search_can_be_switched_off(monkeypatch):
    # CGASM_STRIP_SEARCH=0: the greedy strips (what unstructured meshes get, whose every link is a class of its own)
    monkeypatch.setenv("CGASM_STRIP_SEARCH", "0")
    m3 = syn.box_mesh((4, 4, 4))
    rp, _ = se.strip_plan(m3)
    deg = np.bincount(m3.ndglno.ravel() - 1, minlength=m3.n_nodes)
    assert (np.diff(rp)[deg == 24] == 33).all()


def test_strip_pattern_is_translation_invariant():
    # rows with the same local topology get the same compute pattern (lanes of a warp stay aligned)
    m3 = syn.box_mesh((5, 5, 5))
    rp, ent = se.strip_plan(m3)
    deg = np.bincount(m3.ndglno.ravel() - 1, minlength=m3.n_nodes)
    interior = np.nonzero(deg == 24)[0]
    pats = {tuple((ent[rp[r]:rp[r + 1], 1] & se.COMPUTE).tolist()) for r in interior}
    assert len(pats) == 1


@pytest.mark.parametrize("name", list(meshes().keys()))
def test_emulated_momentum_matches_oracle(orc, name):
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    o = abi.common_momentum_opts()
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    got = se.emulate_momentum(mesh, fs, o, findrm, colm)
    for d in range(mesh.dim):
        assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
        assert row_rel_err(got["big_m"][d], ref["big_m"][d], findrm) < TOL
        assert rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL
        assert rel_err(got["masslump"][:, d], ref["masslump"][:, d]) < TOL


@pytest.mark.parametrize("theta", [0.5, 0.0])
@pytest.mark.parametrize("name", list(meshes().keys()))
def test_emulated_advdiff_matches_oracle(orc, name, theta):
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    o = abi.common_advdiff_opts(theta=theta)
    ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    got = se.emulate_advdiff(mesh, fs, o, findrm, colm)
    assert rel_err(got["matrix"], ref["matrix"]) < TOL
    assert row_rel_err(got["matrix"], ref["matrix"], findrm) < TOL
    assert rel_err(got["rhs"], ref["rhs"]) < TOL


@pytest.mark.parametrize("name", ["box3", "box2_shuffled", "cube.1"])
def test_emulated_momentum_option_variants_match_oracle(orc, name):
    # momentum closed forms planned for the strip kernels in round 2: nodal source, subtract_out_reference_profile,
    # consistent mass, the beta term of the plain advection form, the by-parts volume form, lumped source, lumped
    # (and pressure-corrected) absorption
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    for o in (abi.common_momentum_opts(have_source=1), abi.common_momentum_opts(subtract_out_reference_profile=1),
              abi.common_momentum_opts(have_source=1, subtract_out_reference_profile=1, theta=1.0),
              abi.common_momentum_opts(lump_mass=0), abi.common_momentum_opts(beta=1.0), abi.common_momentum_opts(beta=0.3, lump_mass=0),
              abi.common_momentum_opts(exclude_mass=1), abi.common_momentum_opts(integrate_advection_by_parts=1),
              abi.common_momentum_opts(integrate_advection_by_parts=1, beta=0.3, lump_mass=0),
              abi.common_momentum_opts(have_source=1, lump_source=1),
              abi.common_momentum_opts(have_absorption=1, lump_absorption=1),
              abi.common_momentum_opts(have_absorption=1, lump_absorption=1, pressure_corrected_absorption=1)):
        ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
        got = se.emulate_momentum(mesh, fs, o, findrm, colm)
        for d in range(mesh.dim):
            assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL
            assert rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL
            assert rel_err(got["masslump"][:, d], ref["masslump"][:, d]) < TOL


@pytest.mark.parametrize("name", ["box3", "box2_shuffled", "cube.1"])
def test_emulated_momentum_tensor_viscosity_matches_oracle(orc, name):
    # nodal full-tensor viscosity (element average x Wsum), constant anisotropic and diagonal tensors
    mesh = meshes()[name]
    findrm, colm, _ = orc.make_sparsity(mesh)
    cases = [(syn.standard_fields(mesh, nodal_viscosity=True), abi.common_momentum_opts(viscosity_shape=abi.TENSOR_FULL))]
    for shape in (abi.TENSOR_FULL, abi.TENSOR_DIAGONAL):
        fs = syn.standard_fields(mesh)
        fs.set(abi.F_VISCOSITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
        cases.append((fs, abi.common_momentum_opts(viscosity_shape=shape)))
    for fs, o in cases:
        ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
        got = se.emulate_momentum(mesh, fs, o, findrm, colm)
        for d in range(mesh.dim):
            assert rel_err(got["big_m"][d], ref["big_m"][d]) < TOL and rel_err(got["rhs"][:, d], ref["rhs"][:, d]) < TOL


@pytest.mark.parametrize("name", ["box3", "box2_shuffled", "cube.1", "2d_square"])
def test_emulated_advdiff_with_absorption_and_source_matches_oracle(orc, name):
    # the tracer closed forms planned for the strip kernels in round 2 (absorption, nodal source)
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    for o in (abi.common_advdiff_opts(have_absorption=1), abi.common_advdiff_opts(have_source=1),
              abi.common_advdiff_opts(have_absorption=1, have_source=1, theta=1.0), abi.common_advdiff_opts(beta=1.0),
              abi.common_advdiff_opts(beta=0.4, have_absorption=1), abi.common_advdiff_opts(integrate_advection_by_parts=1),
              abi.common_advdiff_opts(integrate_advection_by_parts=1, beta=1.0),
              abi.common_advdiff_opts(integrate_advection_by_parts=1, beta=0.25, have_source=1)):
        ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
        got = se.emulate_advdiff(mesh, fs, o, findrm, colm)
        assert rel_err(got["matrix"], ref["matrix"]) < TOL and row_rel_err(got["matrix"], ref["matrix"], findrm) < TOL
        assert rel_err(got["rhs"], ref["rhs"]) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_absorption_row_closed_form_for_the_strip_kernels(orc, dim):
    """The row-owner closed form planned for absorption in the strip kernels: with
    constant density, row 0 of Ab^d = shape_shape_vector(test, u, detwei*rho, sigma) (Momentum_CG.F90:2038) is
    rho |J| [Qa sigma_0 + Qaab S] on the diagonal and rho |J| [Qd (sigma_0 + sigma_k) + Qabc S] off it, S = sum of
    sigma_d over the element -- the density-weighted mass row with sigma_d in the place of rho. Checked per element
    against the oracle (difference of the element matrices with and without absorption)."""
    mesh = syn.box_mesh((3, 2, 2)[:dim], seed=9)
    fs = syn.standard_fields(mesh)
    rho = 1.3
    fs.set(abi.F_DENSITY, np.array([rho]), abi.FIELD_CONSTANT)
    m = se.moments(dim)
    Qa, Qd = m["Qaaa"] - m["Qaab"], m["Qaab"] - m["Qabc"]
    sig = fs.get(abi.F_ABSORPTION)[0]
    o0, o1 = abi.common_momentum_opts(), abi.common_momentum_opts(have_absorption=1)
    dtt = o0.dt * o0.theta
    for ele in range(1, mesh.n_elements + 1):
        nd = mesh.ndglno[ele - 1] - 1
        T0 = orc.momentum_element(mesh, fs, o0, ele)[0]
        T1 = orc.momentum_element(mesh, fs, o1, ele)[0]
        X = mesh.X[nd]
        adet = abs(np.linalg.det(X[1:] - X[0]))
        for d in range(dim):
            Ab = (T1[d, d] - T0[d, d]) / dtt
            s = sig[nd, d]
            S = s.sum()
            want = np.array([[rho * adet * ((Qa * s[i] + m["Qaab"] * S) if i == k else (Qd * (s[i] + s[k]) + m["Qabc"] * S))
                              for k in range(dim + 1)] for i in range(dim + 1)])
            assert np.abs(Ab - want).max() <= 1e-9 * np.abs(want).max()   # (difference of two O(1) numbers / dt theta)


@pytest.mark.parametrize("name", ["box3", "box2_shuffled", "cube.1"])
def test_emulated_absorption_pass_matches_oracle(orc, name):
    """common strip result + the planned absorption pass == the oracle's assembly with have_absorption
    (the backward_facing_step_3d option set: constant density, nodal vector absorption)."""
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.full(mesh.n_nodes, 1.25))  # uniform, stored nodally (the emulation indexes it by node)
    findrm, colm, _ = orc.make_sparsity(mesh)
    o = abi.common_momentum_opts(have_absorption=1)
    ref = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    base = se.emulate_momentum(mesh, fs, abi.common_momentum_opts(), findrm, colm)
    add = se.emulate_absorption_pass(mesh, fs, o, findrm, colm)
    for d in range(mesh.dim):
        got = base["big_m"][d] + add["big_m"][d]
        assert rel_err(got, ref["big_m"][d]) < TOL and row_rel_err(got, ref["big_m"][d], findrm) < TOL
        assert rel_err(base["rhs"][:, d] + add["rhs"][:, d], ref["rhs"][:, d]) < TOL
    assert rel_err(base["masslump"], ref["masslump"]) < TOL


def test_strip_search_leaves_unstructured_meshes_to_the_greedy(monkeypatch):
    # strip_search_prescan: a Delaunay mesh's links are all different (no congruence classes to amortise a search over),
    # so the default plan is the greedy's, entry for entry
    mesh = syn.delaunay_mesh(3000, dim=3, seed=9)
    rp_default, ent_default = se.strip_plan(mesh)
    monkeypatch.setenv("CGASM_STRIP_SEARCH", "0")
    rp_greedy, ent_greedy = se.strip_plan(mesh)
    assert np.array_equal(rp_default, rp_greedy) and np.array_equal(ent_default, ent_greedy)
