"""Checks shared by the CPU (oracle) and GPU (CUDA path through the C ABI) tests against the golden
vectors that the reference's own Python element machinery produced
(tests/golden/make_pyref_golden.py imports python/fluidity/state_types.py unmodified):
transform_to_physical (detwei, physical gradients), momentum mass / lumped mass, tracer mass,
grad_p_u_mat / ct_m. Tolerance 1e-12 relative per block; the reference
Python takes |det J| from an SVD and J^-1 from pinv, i.e. a different operation order."""
import os
import numpy as np

from conftest import GOLDEN, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi

CASES = ["cube.1", "cube-parallel", "square-cavity-2d", "prectangle_0"]
TOL = 1e-12


def load(name):
    z = np.load(os.path.join(GOLDEN, "pyref_%s.npz" % name))
    mesh = syn.Mesh(dim=int(z["dim"]), ndglno=np.ascontiguousarray(z["ndglno"], dtype=np.int32),
                    X=np.ascontiguousarray(z["X"]))
    fs = syn.standard_fields(mesh)
    assert (fs.get(abi.F_DENSITY)[0] == z["density"]).all()  # the generator used the same nodal density
    return mesh, fs, z


def mass_only_momentum(lump):
    return abi.common_momentum_opts(lump_mass=1 if lump else 0, exclude_advection=1, have_viscosity=0, have_gravity=0,
                                    assemble_ct_matrix_here=1)


def mass_only_tracer():
    return abi.common_advdiff_opts(have_advection=0, have_diffusivity=0)


def check_elements(mesh, fs, z, momentum_element, advdiff_element, elements=None, zero_tol=0.0):
    """momentum_element(opts, ele) -> (T, rhs, masslump, grad_p_u_mat); advdiff_element(opts, ele) -> (A, rhs)."""
    dim, loc = mesh.dim, mesh.loc
    if elements is None:
        elements = range(1, mesh.n_elements + 1)
    worst = 0.0
    ztol = zero_tol * np.abs(z["M_rho"]).max()  # 0: terms that are switched off contribute exact zeros
    for ele in elements:
        e = ele - 1
        M, Mr, G = z["M"][e], z["M_rho"][e], z["G"][e]
        T, r, ml, gp = momentum_element(mass_only_momentum(False), ele)
        for d1 in range(dim):
            for d2 in range(dim):
                ref = Mr if d1 == d2 else np.zeros_like(Mr)
                assert np.abs(T[d1, d2] - ref).max() <= TOL * np.abs(Mr).max(), (ele, d1, d2)
        assert np.abs(r).max() <= ztol  # the mass term has no rhs part on a static mesh (Momentum_CG.F90:1573)
        for d in range(dim):
            assert rel_err(ml[d], Mr.sum(1)) < TOL
            assert rel_err(gp[d], G[:, :, d]) < TOL, (ele, d)
        T, r, ml, gp = momentum_element(mass_only_momentum(True), ele)
        for d in range(dim):
            assert rel_err(T[d, d], np.diag(Mr.sum(1))) < TOL
            assert rel_err(ml[d], Mr.sum(1)) < TOL
        A, ra = advdiff_element(mass_only_tracer(), ele)
        assert rel_err(A, M) < TOL, ele
        assert np.abs(ra).max() <= ztol
        worst = max(worst, rel_err(A, M), rel_err(gp[0], G[:, :, 0]))
    return worst


def _dense_at(findrm, colm, dense):
    rows = np.repeat(np.arange(len(findrm) - 1), np.diff(findrm))
    return dense[rows, np.asarray(colm) - 1]


def check_assembled(mesh, fs, z, findrm, colm, assemble_momentum, assemble_advdiff, zero_tol=0.0):
    """assemble_momentum(opts) -> dict(big_m, rhs, masslump, ct_m); assemble_advdiff(opts) -> dict(matrix, rhs)."""
    dim = mesh.dim
    ztol = zero_tol * np.abs(z["masslump"]).max()
    # the sparsity holds exactly the node pairs the reference Python loop touched
    pattern = np.zeros_like(z["mass_dense"], dtype=bool)
    rows = np.repeat(np.arange(mesh.n_nodes), np.diff(findrm))
    pattern[rows, np.asarray(colm) - 1] = True
    assert (pattern == (z["mass_dense"] != 0.0)).all()
    got = assemble_momentum(mass_only_momentum(True))
    for d in range(dim):
        assert rel_err(got["masslump"][:, d], z["masslump"]) < TOL
        assert rel_err(got["ct_m"][d], _dense_at(findrm, colm, z["ct_dense"][d])) < TOL
        # lumped mass on the diagonal of big_m, not scaled by dt*theta (Momentum_CG.F90:1550)
        D = np.zeros((mesh.n_nodes, mesh.n_nodes))
        D[rows, np.asarray(colm) - 1] = got["big_m"][d]
        assert rel_err(np.diag(D), z["masslump"]) < TOL and np.abs(D - np.diag(np.diag(D))).max() <= ztol
    assert np.abs(got["rhs"]).max() <= ztol
    adv = assemble_advdiff(mass_only_tracer())
    assert rel_err(adv["matrix"], _dense_at(findrm, colm, z["mass_dense"])) < TOL
    assert np.abs(adv["rhs"]).max() <= ztol


def check_composed(mesh, fs, z, momentum_element, advdiff_element, elements=None):
    """The common option set (and + absorption + sources) against the `c_*` arrays: contractions
    restated in the generator, every ingredient (detwei, gradients, quadrature-point values, mass-type
    matrices) computed by the reference Python."""
    if elements is None:
        elements = range(1, mesh.n_elements + 1)
    mom = {"common": abi.common_momentum_opts(), "abs_src": abi.common_momentum_opts(have_absorption=1, have_source=1)}
    adv = {"common": abi.common_advdiff_opts(), "abs_src": abi.common_advdiff_opts(have_absorption=1, have_source=1)}
    worst = 0.0
    for ele in elements:
        e = ele - 1
        for tag, o in mom.items():
            T, r, ml, _ = momentum_element(o, ele)
            refT, refr = z["c_mom_T_" + tag][e], z["c_mom_rhs_" + tag][e]
            errs = (np.abs(T - refT).max() / np.abs(refT).max(), rel_err(r, refr), rel_err(ml[0], z["M_rho"][e].sum(1)))
            assert max(errs) < TOL, (ele, tag, errs)
            worst = max(worst, *errs)
        for tag, o in adv.items():
            A, r = advdiff_element(o, ele)
            errs = (rel_err(A, z["c_adv_A_" + tag][e]), rel_err(r, z["c_adv_rhs_" + tag][e]))
            assert max(errs) < TOL, (ele, tag, errs)
            worst = max(worst, *errs)
    return worst


def check_assembled_composed(mesh, fs, z, findrm, colm, assemble_momentum, assemble_advdiff, tags=("common", "abs_src")):
    """Assembled big_m / rhs / tracer matrix of the common option set against the plain sum of the
    `c_*` element matrices (ascending element order, like the serial reference loop)."""
    dim, n = mesh.dim, mesh.n_nodes
    nd0 = mesh.ndglno.astype(np.int64) - 1
    mom = {"common": abi.common_momentum_opts(), "abs_src": abi.common_momentum_opts(have_absorption=1, have_source=1)}
    adv = {"common": abi.common_advdiff_opts(), "abs_src": abi.common_advdiff_opts(have_absorption=1, have_source=1)}
    for tag in tags:
        big = np.zeros((dim, n, n))
        rhs = np.zeros((n, dim))
        mat = np.zeros((n, n))
        trhs = np.zeros(n)
        for e in range(mesh.n_elements):
            idx = nd0[e]
            for d in range(dim):
                big[d][np.ix_(idx, idx)] += z["c_mom_T_" + tag][e][d, d]
                np.add.at(rhs[:, d], idx, z["c_mom_rhs_" + tag][e][d])
            mat[np.ix_(idx, idx)] += z["c_adv_A_" + tag][e]
            np.add.at(trhs, idx, z["c_adv_rhs_" + tag][e])
        got = assemble_momentum(mom[tag])
        for d in range(dim):
            ref = _dense_at(findrm, colm, big[d])
            assert rel_err(got["big_m"][d], ref) < TOL, (tag, d)
            starts = np.asarray(findrm[:-1]) - 1
            num = np.maximum.reduceat(np.abs(got["big_m"][d] - ref), starts)
            den = np.maximum.reduceat(np.abs(ref), starts)
            assert (num <= TOL * den).all(), (tag, d)
            assert rel_err(got["rhs"][:, d], rhs[:, d]) < TOL, (tag, d)
            assert rel_err(got["masslump"][:, d], z["masslump"]) < TOL
        a = assemble_advdiff(adv[tag])
        assert rel_err(a["matrix"], _dense_at(findrm, colm, mat)) < TOL, tag
        assert rel_err(a["rhs"], trhs) < TOL, tag


def check_variants(mesh, z, elements_for):
    """Every non-stabilised option variant of tests/variants.py against its `v_*` golden (reference Python loops on
    reference-computed ingredients; only the assembly of the terms is restated in the generator).
    elements_for(fs) -> (momentum_element(opts, ele), advdiff_element(opts, ele)) for a field set."""
    import variants as V
    dim = mesh.dim
    nvar = int(z["n_variant_elements"])
    cache = {}
    worst = 0.0
    for tag, kind, o, ftag in V.variant_cases():
        key = V.variant_field_key(ftag)
        if key not in cache:
            cache[key] = elements_for(V.variant_fields(mesh, ftag))
        mom_el, adv_el = cache[key]
        for e in range(nvar):
            if kind == "mom":
                T, r, ml, _ = mom_el(o, e + 1)
                refT, refr, refml = z["v_mom_T_" + tag][e], z["v_mom_rhs_" + tag][e], z["v_mom_ml_" + tag][e]
                scale = np.abs(refT).max()
                errs = [rel_err(r, refr)]
                for d1 in range(dim):
                    for d2 in range(dim):
                        ref = refT[d1] if d1 == d2 else np.zeros_like(refT[0])
                        errs.append(np.abs(T[d1, d2] - ref).max() / scale)
                if o.assemble_inverse_masslump:
                    errs.append(rel_err(ml, refml))
            else:
                A, r = adv_el(o, e + 1)
                errs = [rel_err(A, z["v_adv_A_" + tag][e]), rel_err(r, z["v_adv_rhs_" + tag][e])]
            assert max(errs) < TOL, (tag, e, errs)
            worst = max(worst, *errs)
    return worst
