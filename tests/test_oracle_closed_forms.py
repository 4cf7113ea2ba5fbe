"""Closed forms that follow from the reference code and an independent
numpy evaluation of the element formulae, both against the C oracle. CPU only."""
import numpy as np
import pytest

from conftest import load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi
import np_formulas as npf

TOL = 1e-12


def _variants_momentum():
    c = abi.common_momentum_opts
    return {
        "common": c(),
        "consistent_mass": c(lump_mass=0),
        "by_parts_beta": c(integrate_advection_by_parts=1, beta=0.3),
        "beta1": c(beta=1.0),
        "absorption": c(have_absorption=1),
        "absorption_lumped_pc": c(have_absorption=1, lump_absorption=1, pressure_corrected_absorption=1),
        "source": c(have_source=1),
        "source_lumped": c(have_source=1, lump_source=1),
        "ref_profile": c(subtract_out_reference_profile=1),
        "aniso": c(viscosity_shape=abi.TENSOR_FULL),
        "diagvisc": c(viscosity_shape=abi.TENSOR_DIAGONAL),
        "no_adv_no_mass": c(exclude_advection=1, exclude_mass=1),
        "stokes": c(exclude_advection=1, have_gravity=0),
        "su_optimal": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, nu_bar_scheme=abi.NU_BAR_OPTIMAL),
        "su_unity_noviscosity": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, have_viscosity=0),
        "supg_critical": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_CRITICAL_RULE, nu_bar_scale=1.0,
                           lump_mass=0, have_absorption=1, have_source=1),
        "supg_asymptotic_by_parts": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_DOUBLY_ASYMPTOTIC,
                                      integrate_advection_by_parts=1, beta=0.5),
    }


def _variants_advdiff():
    c = abi.common_advdiff_opts
    return {
        "common": c(),
        "lumped": c(lump_mass=1),
        "by_parts": c(integrate_advection_by_parts=1, beta=0.25),
        "beta": c(beta=1.0),
        "absorb_source": c(have_absorption=1, have_source=1),
        "tensor_diff": c(diffusivity_shape=abi.TENSOR_FULL),
        "pure_diffusion": c(have_advection=0),
        "mass_only": c(have_advection=0, have_diffusivity=0),
        "theta0": c(theta=0.0),
        "su_optimal": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND),
        "su_unity_nodiff": c(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND, have_diffusivity=0),
        "supg_optimal_tensor": c(stabilisation_scheme=abi.STAB_SUPG, diffusivity_shape=abi.TENSOR_FULL, have_source=1,
                                 have_absorption=1, lump_mass=1),
        "supg_critical_by_parts": c(stabilisation_scheme=abi.STAB_SUPG, nu_bar_scheme=abi.NU_BAR_CRITICAL_RULE,
                                    integrate_advection_by_parts=1, beta=0.3),
    }


def _fields(mesh, variant):
    fs = syn.standard_fields(mesh)
    if variant in ("aniso", "diagvisc", "tensor_diff", "supg_optimal_tensor"):
        fs.set(abi.F_VISCOSITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
        fs.set(abi.F_T_DIFFUSIVITY, syn.aniso_tensor(mesh.dim), abi.FIELD_CONSTANT)
    return fs


@pytest.mark.parametrize("dim", [2, 3])
def test_mass_matrix_closed_form(orc, dim):
    # P1 mass matrix = vol/20 (1+delta_ij) on a tet, area/12 (1+delta_ij) on a triangle;
    # lumped = vol/4, area/3; sum_j K_ij = 0.
    mesh = syn.box_mesh((3,) * dim)
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.ones((1,)), abi.FIELD_CONSTANT)
    o = abi.common_momentum_opts(lump_mass=0, exclude_advection=1, have_viscosity=0, have_gravity=0)
    Xe = mesh.X[mesh.ndglno - 1]
    vol = np.abs(np.linalg.det(Xe[:, :dim] - Xe[:, dim:])) / (2 if dim == 2 else 6)
    fac = 12.0 if dim == 2 else 20.0
    for ele in (1, 7, mesh.n_elements):
        T, r, ml, gp = orc.momentum_element(mesh, fs, o, ele)
        want = vol[ele - 1] / fac * (1 + np.eye(dim + 1))
        for d in range(dim):
            assert rel_err(T[d, d], want) < TOL
            assert rel_err(ml[d], np.full(dim + 1, vol[ele - 1] / (dim + 1))) < TOL
        assert np.abs(r).max() == 0.0
        # grad_p_u_mat(d,i,j) = int N_i dN_j/dx_d: summing over i gives vol * dN_j/dx_d and
        # summing over j gives 0
        assert np.abs(gp.sum(axis=2)).max() < 1e-15
    o2 = abi.common_momentum_opts(exclude_mass=1, exclude_advection=1, have_gravity=0,
                                  assemble_inverse_masslump=0)
    T, r, ml, gp = orc.momentum_element(mesh, fs, o2, 5)
    assert np.abs(T[0, 0].sum(axis=1)).max() < 1e-18 + 1e-14 * np.abs(T[0, 0]).max()
    assert np.abs(T[0, 0] - T[0, 0].T).max() == 0.0 or rel_err(T[0, 0], T[0, 0].T) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_linear_field_gradient_exact(orc, dim):
    # transform: physical gradients reproduce the gradient of a linear field exactly
    mesh = syn.box_mesh((3,) * dim)
    g = np.array([0.3, -1.2, 2.0])[:dim]
    for ele in (1, 4, 11):
        Xv = mesh.X[mesh.ndglno[ele - 1] - 1]
        ds, detwei, J = orc.transform_to_physical(dim, Xv, want_J=True)
        f = Xv @ g + 0.7
        grad = np.einsum("i,igk->gk", f, ds)
        assert np.abs(grad - g[None, :]).max() < 1e-12
        # J(:,:,gi) = transpose(J_local_T): J[a,k] = dx_k/dxi_a
        want = (Xv[:dim] - Xv[dim:]).copy()
        assert np.abs(J[:, :, 0] - want).max() < 1e-15


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(_variants_momentum().keys()))
def test_momentum_oracle_vs_numpy(orc, dim, variant):
    mesh = syn.box_mesh((4, 3, 5)[:dim], seed=11)
    o = _variants_momentum()[variant]
    fs = _fields(mesh, variant)
    L, rhs, ml, gp = npf.momentum_local(orc, mesh, fs, o)
    for ele in (1, 2, 17, mesh.n_elements):
        T, r, m, g = orc.momentum_element(mesh, fs, o, ele)
        for d in range(dim):
            assert rel_err(T[d, d], L[ele - 1, d]) < TOL, (variant, ele, d)
            for d2 in range(dim):
                if d2 != d:
                    assert np.abs(T[d, d2]).max() == 0.0
        assert rel_err(r, rhs[ele - 1]) < TOL
        assert rel_err(m, ml[ele - 1]) < TOL
    findrm, colm, centrm = orc.make_sparsity(mesh)
    out = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    big_m, R, ML = npf.scatter_momentum(mesh, findrm, colm, L, rhs, ml)
    for d in range(dim):
        assert rel_err(out["big_m"][d], big_m[d]) < TOL
        assert rel_err(out["rhs"][:, d], R[:, d]) < TOL
        assert rel_err(out["masslump"][:, d], ML[:, d]) < TOL


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("variant", list(_variants_advdiff().keys()))
def test_advdiff_oracle_vs_numpy(orc, dim, variant):
    mesh = syn.box_mesh((4, 3, 5)[:dim], seed=12)
    o = _variants_advdiff()[variant]
    fs = _fields(mesh, variant)
    A, rhs = npf.advdiff_local(orc, mesh, fs, o)
    for ele in (1, 3, 19, mesh.n_elements):
        a, r = orc.advdiff_element(mesh, fs, o, ele)
        assert rel_err(a, A[ele - 1]) < TOL, (variant, ele)
        assert rel_err(r, rhs[ele - 1]) < TOL
    findrm, colm, centrm = orc.make_sparsity(mesh)
    out = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    val, R = npf.scatter_advdiff(mesh, findrm, colm, A, rhs)
    assert rel_err(out["matrix"], val) < TOL
    assert rel_err(out["rhs"], R) < TOL


def test_coloured_openmp_order_matches_serial(orc):
    # the reference's OpenMP path (colour by colour) only changes summation order
    mesh = load_golden_mesh("cube-parallel")
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    col, nc = orc.colour_elements(mesh)
    sets = orc.colour_sets(col, nc)
    o = abi.common_momentum_opts()
    a = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    b = orc.assemble_momentum(mesh, fs, o, findrm, colm, colouring=sets)
    for k in ("big_m", "rhs", "masslump"):
        assert rel_err(b[k], a[k]) < TOL
    oa = abi.common_advdiff_opts()
    a = orc.assemble_advdiff(mesh, fs, oa, findrm, colm)
    b = orc.assemble_advdiff(mesh, fs, oa, findrm, colm, colouring=sets)
    assert rel_err(b["matrix"], a["matrix"]) < TOL and rel_err(b["rhs"], a["rhs"]) < TOL


def test_assembled_row_sums(orc):
    # consistency: constant T is in the kernel of advection (beta=0) + diffusion, so with
    # T == 1 and no source the tracer rhs vanishes; K rows of big_m sum like the mass.
    mesh = syn.box_mesh((4, 4, 4))
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_T, np.ones(mesh.n_nodes))
    findrm, colm, _ = orc.make_sparsity(mesh)
    out = orc.assemble_advdiff(mesh, fs, abi.common_advdiff_opts(), findrm, colm)
    scale = np.abs(out["matrix"]).max()
    assert np.abs(out["rhs"]).max() < 1e-12 * scale
    # total mass: sum of all consistent-mass entries = volume of the unit cube
    o = abi.common_advdiff_opts(have_advection=0, have_diffusivity=0)
    out = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    assert abs(out["matrix"].sum() - 1.0) < 1e-12


def test_unsupported_option_is_refused(orc):
    mesh = syn.box_mesh((2, 2, 2))
    fs = syn.standard_fields(mesh)
    with pytest.raises(RuntimeError):
        orc.momentum_element(mesh, fs, abi.common_momentum_opts(have_les=1), 1)
