"""Surface-element loops and strong Dirichlet conditions of the oracle, CPU only.

The reference has no unit test at this level, so the restatement is pinned by identities that hold
for the reference's formulae: the divergence theorem for the facet transform, the integration-by-
parts identity that ties the face term to the two volume forms (every integrand is a polynomial the
degree-3 rules integrate exactly), and the exact simplex moments int prod lambda^a = |F| m! prod a! /
(m + sum a)! for the Neumann / Robin / weak-Dirichlet / flux terms."""
import itertools
import math
import numpy as np
import pytest

from conftest import load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, tables

TOL = 1e-12
NEUMANN, WEAKDIRICHLET, INTERNAL, ROBIN = 1, 2, 3, 4           # Advection_Diffusion_CG.F90:74-75
V_WEAK, V_NNF, V_INTERNAL, V_FREE, V_FLUX = 1, 2, 3, 4, 5     # Momentum_CG.F90:138-140


def meshes():
    return {"box2": syn.box_mesh((5, 4), seed=3), "box3": syn.box_mesh((3, 4, 2), seed=5),
            "cube.1": load_golden_mesh("cube.1"), "cavity": load_golden_mesh("square-cavity-2d")}


def face_measure(X):
    e = X[1:] - X[0]
    return np.linalg.norm(e[0]) if len(X) == 2 else 0.5 * np.linalg.norm(np.cross(e[0], e[1]))


def moment(m, *idx):
    """int over an m-simplex of measure 1 of prod_k lambda_{idx_k}"""
    a = np.bincount(idx) if idx else np.zeros(1, dtype=int)
    return math.factorial(m) * np.prod([math.factorial(int(x)) for x in a]) / math.factorial(m + len(idx))


def outward_normal(Xf, Xv):
    c = Xf.mean(0) - Xv.mean(0)
    if Xf.shape[1] == 2:
        t = Xf[1] - Xf[0]
        n = np.array([-t[1], t[0]])
    else:
        n = np.cross(Xf[1] - Xf[0], Xf[2] - Xf[0])
    n = n / np.linalg.norm(n)
    return n if n @ c > 0 else -n


@pytest.mark.parametrize("dim", [2, 3])
def test_face_tables_equal_the_restated_rule(orc, dim):
    n, dn, w = orc.face_tables(dim)
    pn, pdn, pw = tables.p1_face_tables(dim)
    assert (n == pn).all() and (dn == pdn).all() and (w == pw).all()
    sloc, sngi = dim, len(w)
    N = n.reshape(sngi, sloc)
    assert np.abs(N.sum(1) - 1).max() < 1e-15
    # the rule integrates every monomial of degree <= 3 of the face's barycentric coordinates
    vol = 1.0 / math.factorial(dim - 1)
    for deg in range(4):
        for idx in itertools.combinations_with_replacement(range(sloc), deg):
            q = (w * np.prod([N[:, i] for i in idx], axis=0)).sum() if idx else w.sum()
            assert abs(q - vol * moment(dim - 1, *idx)) < 4e-15, idx


@pytest.mark.parametrize("name", ["box2", "box3", "cube.1", "cavity"])
def test_facet_transform_divergence_theorem(orc, name):
    mesh = meshes()[name]
    dim = mesh.dim
    sn, fe = syn.boundary_faces(mesh)
    area, nint, xn = 0.0, np.zeros(dim), 0.0
    for f in range(len(fe)):
        Xf, Xv = mesh.X[sn[f] - 1], mesh.X[mesh.ndglno[fe[f] - 1] - 1]
        dw, nrm = orc.transform_facet_to_physical(dim, Xf, Xv)
        assert np.abs(nrm - outward_normal(Xf, Xv)).max() < 1e-14 and (nrm == nrm[0]).all()
        assert abs(dw.sum() - face_measure(Xf)) < 1e-14 * (1 + face_measure(Xf))
        area += dw.sum()
        nint += dw.sum() * nrm[0]
        xn += dw.sum() * (Xf.mean(0) @ nrm[0])
    vol = sum(abs(np.linalg.det(mesh.X[e[1:] - 1] - mesh.X[e[0] - 1])) for e in mesh.ndglno) / math.factorial(dim)
    assert np.abs(nint).max() < 1e-13 * area and abs(xn / dim - vol) < 1e-13 * vol


@pytest.mark.parametrize("beta", [0.0, 0.3, 1.0])
@pytest.mark.parametrize("name", ["box2", "box3", "cavity"])
def test_tracer_by_parts_plus_faces_equals_the_plain_form(orc, name, beta):
    """int N_i u.grad N_j = -int (u.grad N_i) N_j - int N_i N_j div u + oint N_i N_j u.n: the by-parts volume
    loop followed by the face loop (no boundary condition anywhere) reproduces the plain assembly."""
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    sn, fe = syn.boundary_faces(mesh)
    plain = orc.assemble_advdiff(mesh, fs, abi.common_advdiff_opts(beta=beta), findrm, colm)
    o = abi.common_advdiff_opts(beta=beta, integrate_advection_by_parts=1)
    got = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    before = got["matrix"].copy()
    orc.assemble_advdiff_surface(mesh, fs, o, findrm, colm, sn, fe, np.zeros(len(fe), dtype=np.int32), None, None,
                                 got["matrix"], got["rhs"])
    assert rel_err(before, plain["matrix"]) > 1e-6  # the face term is not negligible
    assert rel_err(got["matrix"], plain["matrix"]) < TOL and rel_err(got["rhs"], plain["rhs"]) < TOL
    # internal faces are skipped, and without by-parts advection or diffusivity the loop does nothing
    again = {k: v.copy() for k, v in plain.items()}
    orc.assemble_advdiff_surface(mesh, fs, abi.common_advdiff_opts(have_diffusivity=0), findrm, colm, sn, fe,
                                 np.full(len(fe), NEUMANN, dtype=np.int32), np.ones((len(fe), mesh.dim)), None,
                                 again["matrix"], again["rhs"])
    assert (again["matrix"] == plain["matrix"]).all() and (again["rhs"] == plain["rhs"]).all()
    orc.assemble_advdiff_surface(mesh, fs, o, findrm, colm, sn, fe, np.full(len(fe), INTERNAL, dtype=np.int32), None, None,
                                 again["matrix"], again["rhs"])
    assert (again["matrix"] == plain["matrix"]).all()


@pytest.mark.parametrize("name", ["box2", "box3"])
def test_momentum_by_parts_plus_faces_equals_the_plain_form(orc, name):
    mesh = meshes()[name]
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.array([1.3]), abi.FIELD_CONSTANT)  # Boussinesq: keeps the face integrand cubic
    findrm, colm, _ = orc.make_sparsity(mesh)
    sn, fe = syn.boundary_faces(mesh)
    for beta in (0.0, 0.4):
        plain = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(beta=beta), findrm, colm)
        o = abi.common_momentum_opts(beta=beta, integrate_advection_by_parts=1)
        got = orc.assemble_momentum(mesh, fs, o, findrm, colm)
        bt = np.zeros((len(fe), mesh.dim), dtype=np.int32)
        orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, bt, np.zeros((len(fe), mesh.dim, mesh.dim)),
                                      got["big_m"], got["rhs"])
        for d in range(mesh.dim):
            assert rel_err(got["big_m"][d], plain["big_m"][d]) < TOL
            assert rel_err(got["rhs"][:, d], plain["rhs"][:, d]) < TOL
    # faces whose only condition is no-normal-flow are skipped (Momentum_CG.F90:799-803) ...
    o = abi.common_momentum_opts(integrate_advection_by_parts=1)
    base = orc.assemble_momentum(mesh, fs, o, findrm, colm)
    skip = {k: (v.copy() if v is not None else None) for k, v in base.items()}
    bt = np.zeros((len(fe), mesh.dim), dtype=np.int32)
    bt[:, 0] = V_NNF
    orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, bt, np.zeros((len(fe), mesh.dim, mesh.dim)),
                                  skip["big_m"], skip["rhs"])
    assert (skip["big_m"] == base["big_m"]).all() and (skip["rhs"] == base["rhs"]).all()
    # ... unless they carry a pressure condition; then the advection term is still off (type(1) = no normal flow)
    orc.assemble_momentum_surface(mesh, fs, o, findrm, colm, sn, fe, bt, np.zeros((len(fe), mesh.dim, mesh.dim)),
                                  skip["big_m"], skip["rhs"], pressure_bc_type=np.ones(len(fe), dtype=np.int32))
    assert (skip["big_m"] == base["big_m"]).all() and (skip["rhs"] == base["rhs"]).all()


def _face_mats(mesh, sn, fe, f):
    """|F|, outward n, P1 face mass matrix and the third-moment tensor int N_i N_j N_k"""
    dim = mesh.dim
    Xf, Xv = mesh.X[sn[f] - 1], mesh.X[mesh.ndglno[fe[f] - 1] - 1]
    F = face_measure(Xf)
    sloc = dim
    M = np.array([[F * moment(dim - 1, i, j) for j in range(sloc)] for i in range(sloc)])
    Q = np.array([[[F * moment(dim - 1, i, j, k) for k in range(sloc)] for j in range(sloc)] for i in range(sloc)])
    return F, outward_normal(Xf, Xv), M, Q


@pytest.mark.parametrize("name", ["box2", "box3", "cube.1"])
def test_tracer_face_terms_against_exact_moments(orc, name):
    mesh = meshes()[name]
    dim = sloc = mesh.dim
    fs = syn.standard_fields(mesh)
    sn, fe = syn.boundary_faces(mesh)
    rng = np.random.default_rng(12)
    T, U = fs.get(abi.F_T)[0], fs.get(abi.F_NU)[0]
    dt, theta = 0.01, 0.5
    for f in rng.choice(len(fe), size=min(12, len(fe)), replace=False):
        F, n, M, Q = _face_mats(mesh, sn, fe, f)
        bc, bc2 = rng.uniform(size=sloc), rng.uniform(0.5, 2.0, size=sloc)
        Tf = T[sn[f] - 1]
        # Neumann: rhs += int N_i g  (Advection_Diffusion_CG.F90:1359-1360)
        A, r = orc.advdiff_face(mesh, fs, abi.common_advdiff_opts(), sn, fe, f + 1, NEUMANN, bc)
        assert np.abs(A).max() == 0.0 and rel_err(r, M @ bc) < TOL
        # Robin: rhs += int N_i g - R T, matrix += dt theta R, R = int N_i N_j h  (:1361-1369)
        R = np.einsum("ijk,k->ij", Q, bc2)
        A, r = orc.advdiff_face(mesh, fs, abi.common_advdiff_opts(), sn, fe, f + 1, ROBIN, bc, bc2)
        assert rel_err(A, dt * theta * R) < TOL and rel_err(r, M @ bc - R @ Tf) < TOL
        # theta = 0: the matrix part is guarded by |dt theta| > epsilon (:1365)
        A, r = orc.advdiff_face(mesh, fs, abi.common_advdiff_opts(theta=0.0), sn, fe, f + 1, ROBIN, bc, bc2)
        assert np.abs(A).max() == 0.0 and rel_err(r, M @ bc - R @ Tf) < TOL
        # by-parts advection: adv = int N_i N_j u.n; weak Dirichlet keeps it out of the matrix (:1329-1337)
        un = U[sn[f] - 1] @ n
        Adv = np.einsum("ijk,k->ij", Q, un)
        o = abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0)
        A, r = orc.advdiff_face(mesh, fs, o, sn, fe, f + 1, 0)
        assert rel_err(A, dt * theta * Adv) < TOL and rel_err(r, -Adv @ Tf) < TOL
        A, r = orc.advdiff_face(mesh, fs, o, sn, fe, f + 1, WEAKDIRICHLET, bc)
        assert np.abs(A).max() == 0.0 and rel_err(r, -theta * Adv @ (bc - Tf) - Adv @ Tf) < TOL
    # weak Dirichlet together with diffusivity is refused like the reference's FLExit (:1375)
    with pytest.raises(RuntimeError):
        orc.advdiff_face(mesh, fs, abi.common_advdiff_opts(), sn, fe, 1, WEAKDIRICHLET, np.zeros(sloc))
    with pytest.raises(RuntimeError):
        orc.advdiff_face(mesh, fs, abi.common_advdiff_opts(), sn, fe, 1, 7)


@pytest.mark.parametrize("name", ["box2", "box3"])
def test_momentum_face_terms_against_exact_moments(orc, name):
    mesh = meshes()[name]
    dim = sloc = mesh.dim
    fs = syn.standard_fields(mesh)
    fs.set(abi.F_DENSITY, np.array([0.8]), abi.FIELD_CONSTANT)
    sn, fe = syn.boundary_faces(mesh)
    rng = np.random.default_rng(13)
    U, oldu = fs.get(abi.F_NU)[0], fs.get(abi.F_OLDU)[0]
    dt, theta = 0.01, 0.5
    o = abi.common_momentum_opts(integrate_advection_by_parts=1)
    for f in rng.choice(len(fe), size=min(10, len(fe)), replace=False):
        F, n, M, Q = _face_mats(mesh, sn, fe, f)
        Adv = 0.8 * np.einsum("ijk,k->ij", Q, U[sn[f] - 1] @ n)
        bc = rng.uniform(size=(sloc, dim))
        bt = np.zeros(dim, dtype=np.int32)
        bt[0], bt[dim - 1] = V_WEAK, V_FLUX
        B, r = orc.momentum_face(mesh, fs, o, sn, fe, f + 1, bt, bc)
        ou = oldu[sn[f] - 1]
        for d in range(dim):
            if bt[d] == V_WEAK:  # Momentum_CG.F90:1052-1056
                assert np.abs(B[d]).max() == 0.0 and rel_err(r[d], -Adv @ bc[:, d]) < TOL
            else:                # :1057-1064 (+ the flux condition :1180-1187)
                want = -Adv @ ou[:, d] + (M @ bc[:, d] if bt[d] == V_FLUX else 0.0)
                assert rel_err(B[d], dt * theta * Adv) < TOL and rel_err(r[d], want) < TOL
        # not by parts: only the flux term is left
        B, r = orc.momentum_face(mesh, fs, abi.common_momentum_opts(), sn, fe, f + 1, bt, bc)
        assert np.abs(B).max() == 0.0 and rel_err(r[dim - 1], M @ bc[:, dim - 1]) < TOL and np.abs(r[0]).max() == 0.0
    # nodal density enters at the quadrature points (degree-4 integrand: compared with the rule itself)
    fs2 = syn.standard_fields(mesh)
    n_f, _, w = orc.face_tables(dim)
    N = n_f.reshape(len(w), sloc)
    f = 0
    F, n, M, Q = _face_mats(mesh, sn, fe, f)
    rho_q = N @ fs2.get(abi.F_DENSITY)[0][sn[f] - 1]
    un_q = N @ (U[sn[f] - 1] @ n)
    Adv = np.einsum("gi,gj,g->ij", N, N, w * F * math.factorial(dim - 1) * un_q * rho_q)
    B, r = orc.momentum_face(mesh, fs2, o, sn, fe, f + 1, np.zeros(dim, dtype=np.int32))
    assert rel_err(B[0], dt * theta * Adv) < TOL


def test_strong_dirichlet_scalar(orc):
    # Boundary_Conditions.F90:1982-2024: rate-of-change form with dt, plain value without; rows flagged
    rng = np.random.default_rng(2)
    T, rhs = rng.uniform(size=20), rng.uniform(size=20)
    nodes, vals = np.array([3, 7, 20, 1]), np.array([1.0, -2.0, 0.5, 4.0])
    want = rhs.copy()
    want[nodes - 1] = (vals - T[nodes - 1]) / 0.25
    flags = np.zeros(20, dtype=np.int32)
    orc.apply_dirichlet_scalar(nodes, vals, T, 0.25, rhs, flags)
    assert (rhs == want).all() and flags.sum() == 4 and flags[[2, 6, 19, 0]].all()
    orc.apply_dirichlet_scalar(nodes, vals, T, None, rhs)
    assert (rhs[nodes - 1] == vals).all()


@pytest.mark.parametrize("name", ["box2", "box3", "cavity"])
def test_continuity_by_parts_plus_boundary_blocks_equals_the_plain_form(orc, name):
    """int N_i d_d N_j = -int d_d N_i N_j + oint N_i N_j n_d: the by-parts ct_m (Momentum_CG.F90:1377-1383) plus the
    boundary blocks of the surface loop (:1073-1088) is the plain ct_m (:1401). (Oracle only: the device path still
    returns CGASM_EUNSUPPORTED for integrate_continuity_by_parts.)"""
    mesh = meshes()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    sn, fe = syn.boundary_faces(mesh)
    plain = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(assemble_ct_matrix_here=1), findrm, colm, want_ct=True)
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1, integrate_continuity_by_parts=1)
    got = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    for k in ("big_m", "rhs", "masslump"):
        assert (got[k] == plain[k]).all()           # only ct_m depends on the option
    assert rel_err(got["ct_m"], plain["ct_m"]) > 1e-3
    bt = np.zeros((len(fe), dim), dtype=np.int32)
    orc.assemble_ct_surface(mesh, fs, o, findrm, colm, sn, fe, bt, got["ct_m"])
    for d in range(dim):
        assert rel_err(got["ct_m"][d], plain["ct_m"][d]) < TOL
    # no-normal-flow faces carry no boundary block: ct_m keeps the by-parts volume form there (the condition is natural)
    skip = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    vol = skip["ct_m"].copy()
    bt[:, 0] = V_NNF
    orc.assemble_ct_surface(mesh, fs, o, findrm, colm, sn, fe, bt, skip["ct_m"], pressure_bc_type=np.ones(len(fe), dtype=np.int32))
    assert (skip["ct_m"] == vol).all()
    # one face: blocks, weak Dirichlet moved to ct_rhs, pressure condition on the momentum rhs
    f = 0
    F, n, M, Q = _face_mats(mesh, sn, fe, f)
    bc = np.random.default_rng(5).uniform(size=(dim, dim))
    pbc = np.random.default_rng(6).uniform(size=dim)
    vt = np.zeros(dim, dtype=np.int32)
    vt[0] = V_WEAK
    Cb, cr, r = orc.momentum_face_ct(mesh, fs, o, sn, fe, f + 1, vt, bc, pressure_bc_type=1, pressure_bc=pbc,
                                     include_pressure_and_continuity_bcs=True)
    for d in range(dim):
        blk = M * n[d]
        assert rel_err(r[d], -(pbc @ blk)) < TOL
        if d == 0:
            assert np.abs(Cb[d]).max() == 0.0 and rel_err(cr, -(blk @ bc[:, 0])) < TOL
        else:
            assert rel_err(Cb[d], blk) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_free_surface_stabilisation_against_exact_moments(orc, dim):
    """Momentum_CG.F90:1108-1176 on the flat top of the unit box with constant density and vertical gravity: the face
    matrices sum to dt theta * rho dt g fs_sf * area in the vertical block and vanish in the others; side faces
    (normal orthogonal to gravity) contribute nothing; the lumped form is the row sums on the diagonal."""
    mesh = syn.box_mesh((3,) * dim, jitter=0.0)
    fs = syn.standard_fields(mesh)
    rho, g_mag, fs_sf = 1.7, 9.81, 0.35
    fs.set(abi.F_DENSITY, np.array([rho]), abi.FIELD_CONSTANT)
    grav = np.zeros((1, dim))
    grav[0, dim - 1] = -1.0
    fs.set(abi.F_GRAVITY, grav, abi.FIELD_CONSTANT)
    sn, fe = syn.boundary_faces(mesh)
    bt = np.array([abi.VBC_FREE_SURFACE] * dim, dtype=np.int32)
    o = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=fs_sf, lump_mass=0, gravity_magnitude=g_mag)
    ol = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=fs_sf, lump_mass=1, gravity_magnitude=g_mag)
    total = np.zeros(dim)
    for f in range(len(fe)):
        X = mesh.X[sn[f] - 1]
        B, r = orc.momentum_face(mesh, fs, o, sn, fe, f + 1, bt)
        Bl, rl = orc.momentum_face(mesh, fs, ol, sn, fe, f + 1, bt)
        on_top = np.allclose(X[:, dim - 1], 1.0)
        on_bottom = np.allclose(X[:, dim - 1], 0.0)
        if not (on_top or on_bottom):
            assert np.abs(B).max() == 0 and np.abs(Bl).max() == 0
            continue
        # n . up = +1 on the top, -1 on the bottom; the sign of the vertical block follows it
        for d in range(dim - 1):
            assert np.abs(B[d]).max() == 0
        assert np.allclose(np.diag(Bl[dim - 1]), B[dim - 1].sum(axis=1), rtol=1e-13, atol=0)
        assert np.abs(Bl[dim - 1] - np.diag(np.diag(Bl[dim - 1]))).max() == 0
        if on_top:
            total += [B[d].sum() for d in range(dim)]
            assert B[dim - 1].sum() > 0
        else:
            assert B[dim - 1].sum() < 0
    want = o.dt * o.theta * rho * o.dt * g_mag * fs_sf * 1.0
    assert abs(total[dim - 1] - want) <= 1e-13 * want and np.abs(total[:dim - 1]).max() == 0
