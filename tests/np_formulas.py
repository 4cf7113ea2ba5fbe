"""Vectorised numpy evaluation of the element formulae of SURVEY.md 8(a) ("config-
specialised element formulae"), written independently of oracle/cg_oracle.c (einsum over
all elements, different summation order). Test infrastructure: cross-checks the C oracle."""
import numpy as np
from fluidity_b200 import _abi as abi


def _tables(orc, dim):
    n, dn, w = orc.tables(dim)
    loc, ngi = dim + 1, len(w)
    N = n.reshape(ngi, loc).T.copy()                       # N[i,g]
    DN = dn.reshape(dim, ngi, loc).transpose(2, 1, 0)[:, 0, :].copy()  # DN[i,k] (const in g)
    return N, DN, w


def geometry(orc, mesh):
    dim = mesh.dim
    N, DN, w = _tables(orc, dim)
    Xe = mesh.X[mesh.ndglno.astype(np.int64) - 1]          # (E, loc, dim)
    JT = np.einsum("eia,ik->eak", Xe, DN)                  # dx_a/dxi_k
    detJ = np.linalg.det(JT)
    invJT = np.linalg.inv(JT)                              # dxi_k/dx_a at [k,a]
    grad = np.einsum("eka,ik->eia", invJT, DN)             # dN_i/dx_a
    detwei = np.abs(detJ)[:, None] * w[None, :]
    return N, grad, detwei


def _gather(mesh, fields, slot, kind):
    val, ft = fields.get(slot)
    nd = mesh.ndglno.astype(np.int64) - 1
    if ft == abi.FIELD_CONSTANT:
        v = val.reshape((1,) + val.shape[1:]) if val.ndim > 1 else val.reshape(1)
        return np.broadcast_to(v[0], (mesh.n_elements, mesh.loc) + v.shape[1:])
    return val[nd]


def nu_bar(u_g, J, diff_g, scheme, scale):
    """nu_bar_scaled_q (assemble/Upwind_Stabilisation.F90:225-320), vectorised over (e, g).
    u_g (E,G,d); J (E,d,d) with J[e,a,k] = J(a,k); diff_g (E,G,d,d) [a,b] or None."""
    norm = np.einsum("ega,ega->eg", u_g, u_g)
    uJ = np.einsum("ega,eak->egk", u_g, J)
    if diff_g is None or scheme == abi.NU_BAR_UNITY:
        val = np.abs(uJ).sum(axis=2)
    else:
        inv = np.linalg.inv(diff_g)
        pe = 0.5 * np.einsum("ega,eab,egbk->egk", u_g, J, inv)
        if scheme == abi.NU_BAR_OPTIMAL:
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                xi = 1.0 / np.tanh(pe) - 1.0 / pe
                xi = np.where(pe > 11.859499013855018, 1.0 - 1.0 / pe, xi)
                xi = np.where(pe < -11.859499013855018, -1.0 - 1.0 / pe, xi)
                xi = np.where(np.abs(pe) < 1e-10, 0.0, xi)
        elif scheme == abi.NU_BAR_DOUBLY_ASYMPTOTIC:
            xi = np.where(np.abs(pe) <= 3.0, pe / 3.0, np.sign(pe))
        else:
            with np.errstate(divide="ignore"):
                xi = np.where(np.abs(pe) <= 1.0, 0.0, np.sign(pe) - 1.0 / pe)
        val = np.einsum("egk,egk->eg", xi, uJ)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(norm < 1e-10, 0.0, val / norm)
    return out * scale


def jacobian(orc, mesh):
    """J[e,a,k] = J(a,k,gi) = transpose(J_local_T) (Transform_elements.F90:878-882)."""
    dim = mesh.dim
    _, DN, _ = _tables(orc, dim)
    Xe = mesh.X[mesh.ndglno.astype(np.int64) - 1]
    JT = np.einsum("eia,ik->eak", Xe, DN)
    return JT.transpose(0, 2, 1)


def momentum_local(orc, mesh, fields, o):
    """Returns L[e,d,i,j] (diagonal blocks incl. lumped diagonal), rhs[e,d,i], ml[e,d,i]."""
    dim, loc = mesh.dim, mesh.loc
    N, grad, detwei = geometry(orc, mesh)
    E = mesh.n_elements
    rho = _gather(mesh, fields, abi.F_DENSITY, 0)           # (E,loc)
    rho_g = rho @ N                                        # (E,ngi)
    nu = _gather(mesh, fields, abi.F_NU, 1)                # (E,loc,dim)
    u_g = np.einsum("eid,ig->egd", nu, N)
    oldu = _gather(mesh, fields, abi.F_OLDU, 1)
    divu = np.einsum("eid,eid->e", nu, grad)
    L = np.zeros((E, dim, loc, loc))
    diag = np.zeros((E, dim, loc))
    rhs = np.zeros((E, dim, loc))
    ml = np.zeros((E, dim, loc))
    dtt = o.dt * o.theta
    Nt = np.broadcast_to(N[None], (E,) + N.shape)      # test function (E,i,g)
    stab_mat = None
    if o.stabilisation_scheme != abi.STAB_NONE:
        dq = None
        if o.have_viscosity:
            visc0 = _gather(mesh, fields, abi.F_VISCOSITY, 2)
            Vq = np.einsum("eiba,ig->egab", visc0, N)
            dq = np.zeros_like(Vq)
            for a in range(dim):
                dq[:, :, a, a] = Vq[:, :, a, a]
        nb = nu_bar(u_g, jacobian(orc, mesh), dq, o.nu_bar_scheme, o.nu_bar_scale)
        udn = np.einsum("ega,eia->eig", u_g, grad)
        if o.stabilisation_scheme == abi.STAB_SUPG:
            Nt = N[None] + nb[:, None, :] * udn
        else:
            stab_mat = np.einsum("eig,ejg,eg->eij", udn, udn, nb * detwei)
    M = np.einsum("eig,jg,eg->eij", Nt, N, rho_g * detwei)
    m = M.sum(axis=2)
    if not o.exclude_mass:
        if o.lump_mass:
            diag += m[:, None, :]
        else:
            L += M[:, None]
    if o.assemble_inverse_masslump:
        ml += m[:, None, :]
    if not o.exclude_advection:
        ugradN = np.einsum("egd,ejd->egj", u_g, grad)      # u_g . grad N_j
        NN = np.einsum("eig,jg,eg->eij", Nt, N, divu[:, None] * rho_g * detwei)
        if o.integrate_advection_by_parts:
            A = -np.einsum("egi,jg,eg->eij", ugradN, N, rho_g * detwei) - (1 - o.beta) * NN
        else:
            A = np.einsum("eig,egj,eg->eij", Nt, ugradN, rho_g * detwei) + o.beta * NN
        if stab_mat is not None:
            A = A + stab_mat
        L += dtt * A[:, None]
        rhs -= np.einsum("eij,edj->edi", A, oldu.transpose(0, 2, 1))
    if o.have_source:
        S = np.einsum("eig,jg,eg->eij", Nt, N, rho_g * detwei)
        src = _gather(mesh, fields, abi.F_SOURCE, 1)
        if o.lump_source:
            rhs += S.sum(axis=2)[:, None, :] * src.transpose(0, 2, 1)
        else:
            rhs += np.einsum("eij,ejd->edi", S, src)
    if o.have_gravity:
        b_g = _gather(mesh, fields, abi.F_BUOYANCY, 0) @ N
        if o.subtract_out_reference_profile:
            b_g = b_g - _gather(mesh, fields, abi.F_HB_DENSITY, 0) @ N
        g_g = np.einsum("eid,ig->egd", _gather(mesh, fields, abi.F_GRAVITY, 1), N)
        rhs += np.einsum("eig,egd,eg->edi", Nt, g_g, o.gravity_magnitude * b_g * detwei)
    if o.have_absorption:
        s_g = np.einsum("eid,ig->egd", _gather(mesh, fields, abi.F_ABSORPTION, 1), N)
        Ab = np.einsum("eig,jg,egd,eg->edij", Nt, N, s_g, rho_g * detwei)
        if o.lump_absorption:
            al = Ab.sum(axis=3)
            diag += dtt * al
            rhs -= al * oldu.transpose(0, 2, 1)
            if o.pressure_corrected_absorption and o.assemble_inverse_masslump:
                ml += dtt * al
        else:
            L += dtt * Ab
            rhs -= np.einsum("edij,edj->edi", Ab, oldu.transpose(0, 2, 1))
    if o.have_viscosity:
        visc = _gather(mesh, fields, abi.F_VISCOSITY, 2)    # (E,loc,b,a) = T(a,b)
        V_g = np.einsum("eiba,ig->egab", visc, N)
        if o.viscosity_shape == abi.TENSOR_ISOTROPIC:
            K = np.einsum("eia,eja,eg->eij", grad, grad, V_g[:, :, 0, 0] * detwei)
        elif o.viscosity_shape == abi.TENSOR_DIAGONAL:
            dg = np.einsum("egaa->ega", V_g)
            K = np.einsum("eia,ega,eja,eg->eij", grad, dg, grad, detwei)
        else:
            K = np.einsum("eia,egab,ejb,eg->eij", grad, V_g, grad, detwei)
        L += dtt * K[:, None]
        rhs -= np.einsum("eij,edj->edi", K, oldu.transpose(0, 2, 1))
    idx = np.arange(loc)
    L[:, :, idx, idx] += diag
    gp = np.einsum("ig,ejd,eg->edij", N, grad, detwei)
    return L, rhs, ml, gp


def advdiff_local(orc, mesh, fields, o):
    dim, loc = mesh.dim, mesh.loc
    N, grad, detwei = geometry(orc, mesh)
    E = mesh.n_elements
    T = _gather(mesh, fields, abi.F_T, 0)
    A_tot = np.zeros((E, loc, loc))
    rhs = np.zeros((E, loc))
    dtt = o.dt * o.theta
    Nt = np.broadcast_to(N[None], (E,) + N.shape)
    stab_mat = None
    if o.stabilisation_scheme != abi.STAB_NONE:
        uu = _gather(mesh, fields, abi.F_NU, 1)
        uq = np.einsum("eid,ig->egd", uu, N)
        dq = None
        if o.have_diffusivity:
            kap0 = _gather(mesh, fields, abi.F_T_DIFFUSIVITY, 2)
            dq = np.einsum("eiba,ig->egab", kap0, N)
        nb = nu_bar(uq, jacobian(orc, mesh), dq, o.nu_bar_scheme, o.nu_bar_scale)
        udn = np.einsum("ega,eia->eig", uq, grad)
        if o.stabilisation_scheme == abi.STAB_SUPG:
            Nt = N[None] + nb[:, None, :] * udn
        else:
            stab_mat = np.einsum("eig,ejg,eg->eij", udn, udn, nb * detwei)
    if o.have_mass:
        M = np.einsum("eig,jg,eg->eij", Nt, N, detwei)
        if o.lump_mass:
            idx = np.arange(loc)
            A_tot[:, idx, idx] += M.sum(axis=2)
        else:
            A_tot += M
    if o.have_advection:
        u = _gather(mesh, fields, abi.F_NU, 1)
        u_g = np.einsum("eid,ig->egd", u, N)
        divu = np.einsum("eid,eid->e", u, grad)
        ugradN = np.einsum("egd,ejd->egj", u_g, grad)
        NN = np.einsum("eig,jg,eg->eij", Nt, N, divu[:, None] * detwei)
        if o.integrate_advection_by_parts:
            A = -np.einsum("egi,jg,eg->eij", ugradN, N, detwei) - (1 - o.beta) * NN
        else:
            A = np.einsum("eig,egj,eg->eij", Nt, ugradN, detwei) + o.beta * NN
        if stab_mat is not None:
            A = A + stab_mat
        A_tot += dtt * A
        rhs -= np.einsum("eij,ej->ei", A, T)
    if o.have_absorption:
        s_g = _gather(mesh, fields, abi.F_T_ABSORPTION, 0) @ N
        Ab = np.einsum("eig,jg,eg->eij", Nt, N, s_g * detwei)
        A_tot += dtt * Ab
        rhs -= np.einsum("eij,ej->ei", Ab, T)
    if o.have_diffusivity:
        kap = _gather(mesh, fields, abi.F_T_DIFFUSIVITY, 2)
        K_g = np.einsum("eiba,ig->egab", kap, N)
        if o.diffusivity_shape == abi.TENSOR_ISOTROPIC:
            D = np.einsum("eia,eja,eg->eij", grad, grad, K_g[:, :, 0, 0] * detwei)
        else:
            D = np.einsum("eia,egab,ejb,eg->eij", grad, K_g, grad, detwei)
        A_tot += dtt * D
        rhs -= np.einsum("eij,ej->ei", D, T)
    if o.have_source:
        s_g = _gather(mesh, fields, abi.F_T_SOURCE, 0) @ N
        rhs += np.einsum("eig,eg->ei", Nt, s_g * detwei)
    return A_tot, rhs


def csr_positions(mesh, findrm, colm):
    """pos[e,i,j] 0-based position of (node_i, node_j) in colm."""
    nd = mesh.ndglno.astype(np.int64)
    E, loc = nd.shape
    pos = np.zeros((E, loc, loc), dtype=np.int64)
    fr = findrm.astype(np.int64)
    for i in range(loc):
        for j in range(loc):
            r = nd[:, i] - 1
            lo = fr[r] - 1
            hi = fr[r + 1] - 1
            # vectorised bisection on sorted rows
            target = nd[:, j]
            l, h = lo.copy(), hi.copy()
            while True:
                active = l < h
                if not active.any():
                    break
                mid = (l + h) // 2
                less = np.zeros(E, dtype=bool)
                less[active] = colm[mid[active]] < target[active]
                l = np.where(active & less, mid + 1, l)
                h = np.where(active & ~less, mid, h)
            assert (colm[l] == target).all()
            pos[:, i, j] = l
    return pos


def scatter_momentum(mesh, findrm, colm, L, rhs, ml):
    dim = mesh.dim
    nnz = len(colm)
    pos = csr_positions(mesh, findrm, colm)
    big_m = np.zeros((dim, nnz))
    for d in range(dim):
        np.add.at(big_m[d], pos.ravel(), L[:, d].ravel())
    nd = mesh.ndglno.astype(np.int64) - 1
    R = np.zeros((mesh.n_nodes, dim))
    ML = np.zeros((mesh.n_nodes, dim))
    for d in range(dim):
        np.add.at(R[:, d], nd.ravel(), rhs[:, d].ravel())
        np.add.at(ML[:, d], nd.ravel(), ml[:, d].ravel())
    return big_m, R, ML


def scatter_advdiff(mesh, findrm, colm, A, rhs):
    pos = csr_positions(mesh, findrm, colm)
    val = np.zeros(len(colm))
    np.add.at(val, pos.ravel(), A.ravel())
    nd = mesh.ndglno.astype(np.int64) - 1
    R = np.zeros(mesh.n_nodes)
    np.add.at(R, nd.ravel(), rhs.ravel())
    return val, R
