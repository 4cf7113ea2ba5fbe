"""The oracle against golden vectors computed by the reference's own Python implementation of
transform_to_physical / shape_shape / shape_dshape (python/fluidity/state_types.py, imported
unmodified by tests/golden/make_pyref_golden.py). This is what pins VALUES of the path to reference
code rather than to a restatement: geometry, the mass and lumped-mass terms, the tracer mass matrix
and grad_p_u_mat / ct_m. (Advection, viscosity, absorption, sources and buoyancy have no reference
implementation outside the Fortran; they stay pinned by closed forms and the independent numpy
evaluation, tests/test_oracle_closed_forms.py.) CPU only; tests/test_parity_gpu.py runs the same
checks on the CUDA path."""
import pytest

import pyref_checks as pc
from conftest import rel_err


@pytest.mark.parametrize("name", pc.CASES)
def test_transform_to_physical_matches_reference_python(orc, name):
    mesh, fs, z = pc.load(name)
    for e in range(mesh.n_elements):
        Xv = mesh.X[mesh.ndglno[e] - 1]
        ds, detwei, _ = orc.transform_to_physical(mesh.dim, Xv)
        assert rel_err(detwei, z["detwei"][e]) < pc.TOL
        assert rel_err(ds, z["dshape"][e]) < pc.TOL
    # the fixture is not degenerate: positive volumes, Sum detwei = mesh volume of the kept elements
    assert (z["detwei"].sum(axis=1) > 0).all()


@pytest.mark.parametrize("name", pc.CASES)
def test_element_matrices_match_reference_python(orc, name):
    mesh, fs, z = pc.load(name)
    worst = pc.check_elements(mesh, fs, z, lambda o, ele: orc.momentum_element(mesh, fs, o, ele),
                              lambda o, ele: orc.advdiff_element(mesh, fs, o, ele))
    assert worst < pc.TOL


@pytest.mark.parametrize("name", pc.CASES)
def test_assembled_mass_lumped_mass_and_ct_match_reference_python(orc, name):
    mesh, fs, z = pc.load(name)
    findrm, colm, _ = orc.make_sparsity(mesh)
    pc.check_assembled(mesh, fs, z, findrm, colm,
                       lambda o: orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True),
                       lambda o: orc.assemble_advdiff(mesh, fs, o, findrm, colm))
    pc.check_assembled_composed(mesh, fs, z, findrm, colm,
                                lambda o: orc.assemble_momentum(mesh, fs, o, findrm, colm),
                                lambda o: orc.assemble_advdiff(mesh, fs, o, findrm, colm))


@pytest.mark.parametrize("name", pc.CASES)
def test_common_option_set_from_reference_python_ingredients(orc, name):
    mesh, fs, z = pc.load(name)
    worst = pc.check_composed(mesh, fs, z, lambda o, ele: orc.momentum_element(mesh, fs, o, ele),
                              lambda o, ele: orc.advdiff_element(mesh, fs, o, ele))
    assert worst < pc.TOL


@pytest.mark.parametrize("name", pc.CASES)
def test_every_non_stabilised_variant_from_reference_python_loops(orc, name):
    """Tensor / diagonal viscosity and diffusivity, by-parts advection, the beta term, consistent / lumped / excluded mass,
    full / lumped / pressure-corrected absorption, sources, the reference profile, constant density: the oracle against
    goldens whose every quadrature contraction ran in the reference's own Python loops."""
    mesh, fs, z = pc.load(name)
    worst = pc.check_variants(mesh, z, lambda f: (lambda o, ele: orc.momentum_element(mesh, f, o, ele),
                                                  lambda o, ele: orc.advdiff_element(mesh, f, o, ele)))
    assert worst < pc.TOL


def test_quadrature_point_gather_matches_reference_python(orc):
    # Field.ele_val_at_quad (state_types.py:113-117) == ele_val . shape%n (Fields_Base.F90:2256-2310)
    mesh, fs, z = pc.load("cube-parallel")
    n, _, w = orc.tables(mesh.dim)
    N = n.reshape(len(w), mesh.loc)  # n[i + loc*g]
    rho = z["density"]
    for e in range(mesh.n_elements):
        assert rel_err(N @ rho[mesh.ndglno[e] - 1], z["rho_q"][e]) < 1e-15
