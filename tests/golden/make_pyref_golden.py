"""Golden vectors computed by the REFERENCE'S OWN Python implementation of pieces of this path.

The reference ships a Python restatement of its element machinery for python diagnostics:
`python/fluidity/state_types.py` -- `Transform` (Jacobian J = X.dn, |det J|, detwei, inverse;
:338-377 = femtools/Transform_elements.F90:807-887), `Transform.grad` (physical shape-function
gradients, :379-386), `Transform.shape_shape` (mass matrix with an optional quadrature-point
coefficient, :388-406 = femtools/FETools.F90:206-226), `Transform.shape_dshape` (:408-427 =
FETools.F90:332-362), `Field.ele_val_at_quad` (:113-117 = femtools/Fields_Base.F90:2256-2310),
`Field.addto` (:71-81). This script IMPORTS that module unmodified from /root/reference (build
container only; the GPU box has no reference, so the outputs are committed) and runs it on
reference fixture meshes, producing what the momentum / tracer element loops are built from:

  per element   detwei(ngi), du_t = dshape(loc,ngi,dim)       transform_to_physical
                M      = shape_shape(N, N, detwei)            tracer consistent mass (Advection_Diffusion_CG.F90:909)
                M_rho  = shape_shape(N, N, detwei*rho_g)      momentum mass (Momentum_CG.F90:1535), rho_g = ele_val_at_quad
                G      = shape_dshape(N, du_t)                grad_p_u_mat / ct_m entries (Momentum_CG.F90:1401)
  assembled     lumped mass sum_j M_rho_ij per node           masslump (Momentum_CG.F90:1541,1560-1565), via Field.addto as
                                                              the module's own test_shape_dshape does (:452-458)
                dense tracer mass matrix and dense ct blocks  (plain += of the element matrices)
  composed      the momentum / tracer element matrices and rhs of the common option set (+ absorption, sources), following
                Momentum_CG.F90:1535-1552,1675-1680,1737-1748,1770-1789,2038-2059,2304-2346 and
                Advection_Diffusion_CG.F90:909-920,1093-1125,1139,1156-1160,1192-1200. Every ingredient is
                reference-computed (detwei, du_t = Transform.grad, N, every field at the quadrature points through
                Field.ele_val_at_quad) and every quadrature contraction runs in the reference's own loops: mass,
                absorption and source matrices through Transform.shape_shape(coeff), advection and
                viscosity / diffusivity through Transform.shape_dshape with the coefficient folded into detwei (and, for
                the stiffness term, dN/dx_k in the place of N). What is restated here is only the assembly of those
                pieces (dt*theta scaling, sums, products with oldu / T, the buoyancy and tracer-source vectors). Kept
                apart in the file as `c_*`.

  variants      `v_mom_T_<tag>` (diagonal blocks), `v_mom_rhs_<tag>`, `v_mom_ml_<tag>`, `v_adv_A_<tag>`, `v_adv_rhs_<tag>` for EVERY
                non-stabilised entry of tests/variants.py (momentum_variants, advdiff_variants, boussinesq_variants = the
                four example option sets and their neighbours) on the first VARIANT_ELEMENTS elements: consistent / lumped /
                excluded mass, plain / by-parts advection with the beta term (Momentum_CG.F90:1646-1680: by parts =
                minus the TRANSPOSE of the reference's shape_dshape contraction; div(nu) from Transform.grad), isotropic /
                diagonal / full tensor viscosity and diffusivity (dshape_tensor_dshape, FETools.F90:551-698 = the reference's
                shape_dshape loop with dN/dx_a in the place of N and V_ab folded into detwei), full / lumped /
                pressure-corrected absorption (:2036-2073), consistent / lumped sources (:1717-1751),
                subtract_out_reference_profile (:1767-1771), constant or nodal density. What stays Fortran-only (no
                reference implementation outside the Fortran): SU / SUPG and their nu_bar schemes.

The element tables n / dn / weights handed to `Element` / `Quadrature` are the degree-3 P1 tables
of fluidity_b200/tables.py (at run time the reference fills these objects from its Fortran
element_type; the tables are pinned separately against femtools/tests/test_quadrature.F90 and
test_shape_functions.F90 by tests/test_oracle_golden.py).

    python tests/golden/make_pyref_golden.py        # writes tests/golden/pyref_<mesh>.npz
"""
import copy
import os
import sys
import numpy as np

REF = os.environ.get("FLUIDITY_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "python"))

sys.path.insert(0, os.path.join(ROOT, "tests"))

from fluidity import state_types as st  # noqa: E402  (the reference module, unmodified)
from fluidity_b200 import tables, synthetic  # noqa: E402
import variants  # noqa: E402  (tests/variants.py: the option tables the parity tests run)

VARIANT_ELEMENTS = 40
EPS = 2.220446049250313e-16

# (fixture, elements taken): whole small meshes, the first elements of the unstructured ones
CASES = [("cube.1", None), ("cube-parallel", 160), ("square-cavity-2d", 200), ("prectangle_0", None)]


def reference_shape(dim):
    n, dn, w = tables.p1_tables(dim)
    loc, ngi = dim + 1, len(w)
    n = n.reshape(ngi, loc).T.copy()            # n[i, g]
    dn = dn.reshape(dim, ngi, loc).transpose(2, 1, 0).copy()  # dn[i, g, k]
    el = st.Element(dim, loc, ngi, 1, n, dn, np.zeros((loc, dim)), 0, 0, 0, 0, "lagrangian", "simplex")
    l, _ = tables.quadrature_degree3(dim)
    el.set_quadrature(st.Quadrature(w, l, dim, 3, loc, ngi))
    return el


def density_of(mesh):
    return synthetic.standard_fields(mesh).get(synthetic.abi.F_DENSITY)[0]


DT, THETA, GMAG = 0.01, 0.5, 10.0  # _abi.common_momentum_opts / common_advdiff_opts


def scalar_components(name, val, mesh):
    """A (n_nodes, k) array as k reference ScalarFields (ele_val_at_quad is scalar-only, state_types.py:113-117)."""
    val = np.asarray(val)
    if val.ndim == 1:
        val = val[:, None]
    out = []
    for k in range(val.shape[1]):
        f = st.ScalarField("%s%d" % (name, k), np.ascontiguousarray(val[:, k]), 0, "")
        f.set_mesh(mesh)
        out.append(f)
    return out


def composed(e, t, du_t, shape, F, dim, loc):
    """Element matrices of the common option set with absorption and sources switched on or off."""
    A = synthetic.abi
    N = np.asarray(shape.n)                       # [i, g]
    dN = np.asarray(du_t.dn)                      # [i, g, k]
    dw = np.asarray(t.detwei)
    q = lambda comps: np.array([c.ele_val_at_quad(e) for c in comps])  # [k, g]
    nodes = F["rho"][0].ele_nodes(e)
    nv = lambda comps: np.array([[c.node_val(n) for n in nodes] for c in comps])  # [k, i]
    rho_g, u_g, b_g = q(F["rho"])[0], q(F["nu"]), q(F["buoy"])[0]
    sig_g, tsig_g, ts_g = q(F["absn"]), q(F["t_abs"])[0], q(F["t_src"])[0]
    oldu, src, T = nv(F["oldu"]), nv(F["src"]), nv(F["T"])[0]
    Mr = t.shape_shape(shape, shape, rho_g)
    M = t.shape_shape(shape, shape)
    m = Mr.sum(1)
    def with_detwei(weights, fn):
        """Runs a reference contraction with a quadrature-point coefficient folded into detwei, the way the Fortran
        passes detwei*coefficient to FETools (e.g. Momentum_CG.F90:1675-1680)."""
        keep = t.detwei
        t.detwei = np.asarray(weights)
        try:
            return np.asarray(fn())
        finally:
            t.detwei = keep

    def advection(weight_g):
        # shape_vector_dot_dshape (FETools.F90:749-772) = sum_k shape_dshape(N, du_t, detwei*u_k)[:, :, k], through
        # the reference's Transform.shape_dshape
        return sum(with_detwei(weight_g * u_g[k], lambda: t.shape_dshape(shape, du_t))[:, :, k] for k in range(dim))

    def stiffness(weight_g):
        # dshape_dot_dshape (FETools.F90:391-453): the same reference loop with dN_i/dx_k in the place of N_i
        out = np.zeros((loc, loc))
        for k in range(dim):
            grad_k = copy.copy(shape)
            grad_k.n = dN[:, :, k]
            out += with_detwei(weight_g, lambda: t.shape_dshape(grad_k, du_t))[:, :, k]
        return out

    Adv = advection(rho_g * dw)
    K = stiffness(F["mu"] * dw)
    out = {}
    for tag, have_abs, have_src in (("common", 0, 0), ("abs_src", 1, 1)):
        Tm = np.zeros((dim, dim, loc, loc))
        rhs = np.zeros((dim, loc))
        for d in range(dim):
            Ab = t.shape_shape(shape, shape, rho_g * sig_g[d]) if have_abs else np.zeros((loc, loc))
            L = Adv + K + Ab
            Tm[d, d] = np.diag(m) + DT * THETA * L
            rhs[d] = -L @ oldu[d] + N @ (F["gdir"][d] * GMAG * b_g * dw)
            if have_src:
                rhs[d] += Mr @ src[d]
        out["c_mom_T_" + tag], out["c_mom_rhs_" + tag] = Tm, rhs
    Atr = advection(dw)
    D = stiffness(F["kappa"] * dw)
    for tag, on in (("common", 0), ("abs_src", 1)):
        Ab = t.shape_shape(shape, shape, tsig_g) if on else np.zeros((loc, loc))
        L = Atr + D + Ab
        out["c_adv_A_" + tag] = M + DT * THETA * L
        out["c_adv_rhs_" + tag] = -L @ T + (N @ (ts_g * dw) if on else 0.0)
    return out


class RefFields:
    """The fields of a variant as reference ScalarFields per component (CONSTANT fields become nodal fields holding the
    constant: ele_val_at_quad of a constant is the constant)."""
    def __init__(self, fs, mesh, n_nodes, dim):
        A = synthetic.abi
        self.dim = dim

        def comps(slot, shape):
            val, ftype = fs.get(slot)
            val = np.asarray(val, dtype=np.float64)
            if ftype == A.FIELD_CONSTANT:
                val = np.broadcast_to(val.reshape((1,) + shape), (n_nodes,) + shape)
            return scalar_components("f%d_" % slot, np.ascontiguousarray(val.reshape(n_nodes, -1)), mesh)

        self.rho = comps(A.F_DENSITY, ())
        self.nu, self.oldu = comps(A.F_NU, (dim,)), comps(A.F_OLDU, (dim,))
        self.buoy, self.hb = comps(A.F_BUOYANCY, ()), comps(A.F_HB_DENSITY, ())
        self.absn, self.src = comps(A.F_ABSORPTION, (dim,)), comps(A.F_SOURCE, (dim,))
        self.grav = comps(A.F_GRAVITY, (dim,))
        self.visc = comps(A.F_VISCOSITY, (dim, dim))       # component b*dim + a holds V(a, b) (synthetic.py layout)
        self.T = comps(A.F_T, ())
        self.t_abs, self.t_src = comps(A.F_T_ABSORPTION, ()), comps(A.F_T_SOURCE, ())
        self.kappa = comps(A.F_T_DIFFUSIVITY, (dim, dim))


class RefElement:
    """Reference-computed ingredients of one element and the reference's own contraction loops."""
    def __init__(self, e, t, du_t, shape, dim, loc):
        self.e, self.t, self.du_t, self.shape, self.dim, self.loc = e, t, du_t, shape, dim, loc
        self.N = np.asarray(shape.n)
        self.dN = np.asarray(du_t.dn)
        self.dw = np.asarray(t.detwei)

    def q(self, comps):
        return np.array([c.ele_val_at_quad(self.e) for c in comps])    # [k, g]

    def nv(self, comps):
        nodes = comps[0].ele_nodes(self.e)
        return np.array([[c.node_val(n) for n in nodes] for c in comps])  # [k, i]

    def with_detwei(self, weights, fn):
        keep = self.t.detwei
        self.t.detwei = np.asarray(weights)
        try:
            return np.asarray(fn())
        finally:
            self.t.detwei = keep

    def mass(self, coeff_g=None):
        return np.asarray(self.t.shape_shape(self.shape, self.shape, coeff_g) if coeff_g is not None
                          else self.t.shape_shape(self.shape, self.shape))

    def advection(self, weight_g, u_g):
        # shape_vector_dot_dshape (FETools.F90:749-772): sum_k shape_dshape(N, du_t, detwei*u_k)[:, :, k]
        return sum(self.with_detwei(weight_g * u_g[k], lambda: self.t.shape_dshape(self.shape, self.du_t))[:, :, k]
                   for k in range(self.dim))

    def tensor_stiffness(self, V_g, which):
        """dshape_dot_dshape / dshape_diagtensor_dshape / dshape_tensor_dshape (FETools.F90:391-453,551-698) through the
        reference's shape_dshape loop: dN/dx_a in the place of N, V_ab(g) * detwei as the weight, component b taken.
        V_g[a, b, g]; which: 0 isotropic (V_11 on every a = b), 1 diagonal, 2 full."""
        out = np.zeros((self.loc, self.loc))
        for a in range(self.dim):
            grad_a = copy.copy(self.shape)
            grad_a.n = self.dN[:, :, a]
            for b in range(self.dim):
                if which < 2 and a != b:
                    continue
                w = V_g[0, 0] if which == 0 else V_g[a, b]
                out += self.with_detwei(w * self.dw, lambda: self.t.shape_dshape(grad_a, self.du_t))[:, :, b]
        return out

    def div_at_quad(self, comps):
        # ele_div_at_quad (Fields_Base.F90): sum_k sum_i u_k(i) dN_i/dx_k, from the reference's Transform.grad
        u = self.nv(comps)                                               # [k, i]
        return np.einsum("ki,igk->g", u, self.dN)

    def tensor_at_quad(self, comps):
        v = self.q(comps)                                                # [b*dim + a, g]
        return v.reshape(self.dim, self.dim, -1).transpose(1, 0, 2)      # [a, b, g]


def momentum_variant_element(o, R, F):
    """construct_momentum_element_cg for the option set o (no stabilisation): every contraction in the reference's
    Python loops, only the assembly of the terms restated (Momentum_CG.F90:1492-1575 mass, :1602-1714 advection,
    :1717-1751 sources, :1753-1795 buoyancy, :2036-2073 absorption, :2286-2359 viscosity)."""
    A = synthetic.abi
    dim, loc = R.dim, R.loc
    dtt = o.dt * o.theta
    rho_g = R.q(F.rho)[0]
    oldu = R.nv(F.oldu)
    Mr = R.mass(rho_g)
    m = Mr.sum(1)
    T = np.zeros((dim, loc, loc))
    rhs = np.zeros((dim, loc))
    ml = np.zeros((dim, loc))
    if o.assemble_inverse_masslump:
        ml += m
    if not o.exclude_mass:
        T += np.diag(m) if o.lump_mass else Mr
    if not o.exclude_advection:
        u_g = R.q(F.nu)
        div_g = R.div_at_quad(F.nu)
        if o.integrate_advection_by_parts:
            adv = -R.advection(rho_g * R.dw, u_g).T - (1.0 - o.beta) * R.mass(div_g * rho_g)
        else:
            adv = R.advection(rho_g * R.dw, u_g) + o.beta * R.mass(div_g * rho_g)
        T += dtt * adv
        rhs -= oldu @ adv.T
    if o.have_source:
        src = R.nv(F.src)
        rhs += (m * src) if o.lump_source else src @ Mr.T
    if o.have_gravity:
        b_g = R.q(F.buoy)[0]
        if o.subtract_out_reference_profile:
            b_g = b_g - R.q(F.hb)[0]
        g_g = R.q(F.grav)
        for d in range(dim):
            rhs[d] += R.N @ (g_g[d] * o.gravity_magnitude * b_g * R.dw)
    if o.have_absorption:
        sig_g = R.q(F.absn)
        for d in range(dim):
            Ab = R.mass(rho_g * sig_g[d])
            if o.lump_absorption:
                al = Ab.sum(1)
                T[d] += dtt * np.diag(al)
                rhs[d] -= al * oldu[d]
                if o.pressure_corrected_absorption and o.assemble_inverse_masslump:
                    ml[d] += dtt * al
            else:
                T[d] += dtt * Ab
                rhs[d] -= Ab @ oldu[d]
    if o.have_viscosity:
        K = R.tensor_stiffness(R.tensor_at_quad(F.visc), o.viscosity_shape)
        T += dtt * K
        rhs -= oldu @ K.T
    return T, rhs, ml


def advdiff_variant_element(o, R, F):
    """assemble_advection_diffusion_element_cg (Advection_Diffusion_CG.F90:867-944 mass, :946-1127 advection,
    :1129-1162 source / absorption, :1164-1202 diffusivity), default equation type, no stabilisation."""
    dim, loc = R.dim, R.loc
    dtt = o.dt * o.theta
    implicit = abs(dtt) > EPS
    Tn = R.nv(F.T)[0]
    Amat = np.zeros((loc, loc))
    rhs = np.zeros(loc)
    if o.have_mass:
        M = R.mass()
        Amat += np.diag(M.sum(1)) if o.lump_mass else M
    terms = []
    if o.have_advection:
        u_g = R.q(F.nu)
        if o.integrate_advection_by_parts:
            adv = -R.advection(R.dw, u_g).T
            if abs(1.0 - o.beta) > EPS:
                adv = adv - (1.0 - o.beta) * R.mass(R.div_at_quad(F.nu))
        else:
            adv = R.advection(R.dw, u_g)
            if abs(o.beta) > EPS:
                adv = adv + o.beta * R.mass(R.div_at_quad(F.nu))
        terms.append(adv)
    if o.have_absorption:
        terms.append(R.mass(R.q(F.t_abs)[0]))
    if o.have_diffusivity:
        which = 0 if o.diffusivity_shape == synthetic.abi.TENSOR_ISOTROPIC else 2
        terms.append(R.tensor_stiffness(R.tensor_at_quad(F.kappa), which))
    for L in terms:
        if implicit:
            Amat += dtt * L
        rhs -= L @ Tn
    if o.have_source:
        rhs += R.N @ (R.q(F.t_src)[0] * R.dw)
    return Amat, rhs


def run(name, nele):
    z = np.load(os.path.join(OUT, name + ".npz"))
    dim = int(z["dim"])
    nd = np.ascontiguousarray(z["ndglno"], dtype=np.int32)
    if nele is not None:
        # keep the first `nele` elements and renumber their nodes compactly (ascending old id)
        nd = nd[:nele]
        used = np.unique(nd)
        remap = np.zeros(int(used.max()) + 1, dtype=np.int32)
        remap[used] = np.arange(1, len(used) + 1)
        nd = remap[nd]
        X = np.ascontiguousarray(z["X"][used - 1])
    else:
        X = np.ascontiguousarray(z["X"])
    loc = dim + 1
    n_nodes, n_ele = X.shape[0], nd.shape[0]
    mesh = st.Mesh(nd.ravel(), n_ele, n_nodes, 0, "CoordinateMesh", "", None)
    mesh.shape = reference_shape(dim)
    ngi = mesh.shape.ngi
    coord = st.VectorField("Coordinate", X, 0, "", dim)
    coord.set_mesh(mesh)
    rho_val = density_of(synthetic.Mesh(dim=dim, ndglno=nd, X=X))
    rho = st.ScalarField("Density", rho_val, 0, "")
    rho.set_mesh(mesh)
    lump = st.ScalarField("LumpMass", np.zeros(n_nodes), 0, "")
    lump.set_mesh(mesh)
    A = synthetic.abi
    fs = synthetic.standard_fields(synthetic.Mesh(dim=dim, ndglno=nd, X=X))
    F = dict(rho=[rho], nu=scalar_components("nu", fs.get(A.F_NU)[0], mesh), oldu=scalar_components("oldu", fs.get(A.F_OLDU)[0], mesh),
             buoy=scalar_components("b", fs.get(A.F_BUOYANCY)[0], mesh), absn=scalar_components("abs", fs.get(A.F_ABSORPTION)[0], mesh),
             src=scalar_components("src", fs.get(A.F_SOURCE)[0], mesh), T=scalar_components("T", fs.get(A.F_T)[0], mesh),
             t_abs=scalar_components("tabs", fs.get(A.F_T_ABSORPTION)[0], mesh), t_src=scalar_components("tsrc", fs.get(A.F_T_SOURCE)[0], mesh),
             mu=float(fs.get(A.F_VISCOSITY)[0][0, 0, 0]), kappa=float(fs.get(A.F_T_DIFFUSIVITY)[0][0, 0, 0]),
             gdir=fs.get(A.F_GRAVITY)[0][0])
    comp = {}

    detwei = np.zeros((n_ele, ngi))
    dshape = np.zeros((n_ele, loc, ngi, dim))
    M = np.zeros((n_ele, loc, loc))
    M_rho = np.zeros((n_ele, loc, loc))
    G = np.zeros((n_ele, loc, loc, dim))
    rho_q = np.zeros((n_ele, ngi))
    mass_dense = np.zeros((n_nodes, n_nodes))
    ct_dense = np.zeros((dim, n_nodes, n_nodes))
    for e in range(n_ele):
        t = st.Transform(e, coord)
        du_t = t.grad(mesh.shape)
        detwei[e] = t.detwei
        dshape[e] = du_t.dn
        rho_q[e] = rho.ele_val_at_quad(e)
        M[e] = t.shape_shape(mesh.shape, mesh.shape)
        M_rho[e] = t.shape_shape(mesh.shape, mesh.shape, rho_q[e])
        G[e] = np.asarray(t.shape_dshape(mesh.shape, du_t))
        for k, v in composed(e, t, du_t, mesh.shape, F, dim, loc).items():
            comp.setdefault(k, []).append(v)
        nodes = lump.ele_nodes(e)
        lump.addto(nodes, M_rho[e].sum(1))
        for i in range(loc):
            for j in range(loc):
                mass_dense[nodes[i], nodes[j]] += M[e][i, j]
                for d in range(dim):
                    ct_dense[d, nodes[i], nodes[j]] += G[e][i, j, d]
    # every non-stabilised option variant of the parity tests on the first elements
    smesh = synthetic.Mesh(dim=dim, ndglno=nd, X=X)
    nvar = min(n_ele, VARIANT_ELEMENTS)
    field_cache = {}
    for tag, kind, o, ftag in variants.variant_cases():
        fkey = variants.variant_field_key(ftag)
        if fkey not in field_cache:
            field_cache[fkey] = RefFields(variants.variant_fields(smesh, ftag), mesh, n_nodes, dim)
        RF = field_cache[fkey]
        for e in range(nvar):
            t = st.Transform(e, coord)
            R = RefElement(e, t, t.grad(mesh.shape), mesh.shape, dim, loc)
            if kind == "mom":
                Tm, r, mlv = momentum_variant_element(o, R, RF)
                comp.setdefault("v_mom_T_" + tag, []).append(Tm)
                comp.setdefault("v_mom_rhs_" + tag, []).append(r)
                comp.setdefault("v_mom_ml_" + tag, []).append(mlv)
            else:
                Am, r = advdiff_variant_element(o, R, RF)
                comp.setdefault("v_adv_A_" + tag, []).append(Am)
                comp.setdefault("v_adv_rhs_" + tag, []).append(r)
    out = os.path.join(OUT, "pyref_%s.npz" % name)
    np.savez_compressed(out, dim=dim, n_variant_elements=nvar, ndglno=nd, X=X, density=rho_val, rho_q=rho_q, detwei=detwei, dshape=dshape,
                        M=M, M_rho=M_rho, G=G, masslump=lump.val, mass_dense=mass_dense, ct_dense=ct_dense,
                        **{k: np.array(v) for k, v in comp.items()})
    print(name, "dim", dim, "elements", n_ele, "nodes", n_nodes, "-> %s (%d bytes)" % (out, os.path.getsize(out)))


if __name__ == "__main__":
    for name, nele in CASES:
        run(name, nele)
