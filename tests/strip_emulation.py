"""numpy emulation of the STRIP row-owner kernels (fluidity_b200/csrc/strip.cu), run on the CPU
from the plan that cgasm_strip_plan_host builds: same FIFO windows, same closed forms, same
per-row epilogue. It is TEST infrastructure: it lets the CPU suite check the plan and the
algebra of the device code against the oracle without a GPU. Common option set only.
"""
import ctypes as C

import numpy as np

from fluidity_b200 import _abi as abi, cgasm, tables

COMPUTE = 0x100


def strip_plan(mesh):
    """row_ptr (n_nodes+1), entries (n, 2): node (1-based), meta."""
    lib = cgasm.load()
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    row_ptr = np.zeros(mesh.n_nodes + 1, dtype=np.int64)
    needed = C.c_longlong(0)
    args = [C.c_int(mesh.loc), C.c_int(mesh.n_nodes), C.c_int(mesh.n_elements), nd.ctypes.data_as(C.POINTER(C.c_int)),
            row_ptr.ctypes.data_as(C.POINTER(C.c_longlong))]
    st = lib.cgasm_strip_plan_host(*args, None, C.c_longlong(0), C.byref(needed))
    assert st == 0, lib.cgasm_last_error()
    ent = np.zeros((max(needed.value, 1), 2), dtype=np.int32)
    st = lib.cgasm_strip_plan_host(*args, ent.ctypes.data_as(C.POINTER(C.c_int)), C.c_longlong(needed.value),
                                   C.byref(needed))
    assert st == 0, lib.cgasm_last_error()
    return row_ptr, ent[:needed.value]


def moments(dim):
    """Pd, Po, Qaaa, Qaab, Qabc, Wsum of the degree-3 rule (element_math.cuh Tables)."""
    l, w = tables.quadrature_degree3(dim)
    N = l.T  # N[i, g]
    P = np.einsum("ig,kg,g->ik", N, N, w)
    Q = np.einsum("ig,kg,lg,g->ikl", N, N, N, w)
    return dict(Pd=P[0, 0], Po=P[0, 1], Qaaa=Q[0, 0, 0], Qaab=Q[0, 0, 1], Qabc=Q[0, 1, 2], Wsum=w.sum())


def _geometry(e):
    """cofactor vectors c[k] (gradN_k = c[k]/det) and det of the window edges e (dim, dim)."""
    dim = e.shape[0]
    if dim == 3:
        c = np.array([np.cross(e[1], e[2]), np.cross(e[2], e[0]), np.cross(e[0], e[1])])
    else:
        c = np.array([[e[1][1], -e[1][0]], [-e[0][1], e[0][0]]])
    return c, float(e[0] @ c[0])


def emulate_momentum(mesh, fs, o, findrm, colm):
    dim, nn = mesh.dim, mesh.n_nodes
    m = moments(dim)
    X = mesh.X
    nu, oldu = fs.get(abi.F_NU)[0], fs.get(abi.F_OLDU)[0]
    rho, bb = fs.get(abi.F_DENSITY)[0], fs.get(abi.F_BUOYANCY)[0]
    if o.have_gravity and o.subtract_out_reference_profile:
        bb = bb - fs.get(abi.F_HB_DENSITY)[0]       # Momentum_CG.F90:1781-1784: (buoyancy - hb_density) at the quadrature points
    src = fs.get(abi.F_SOURCE)[0] if (o.have_source and not o.lump_source) else None
    rsrc = np.zeros((nn, dim))
    lump_abs = bool(o.have_absorption and o.lump_absorption)
    sig = fs.get(abi.F_ABSORPTION)[0] if lump_abs else None
    l, wq = tables.quadrature_degree3(dim)
    Q3 = np.einsum("g,gl,gm->lm", wq * l[:, 0], l, l)      # Q_{0lm} = sum_g w_g N_0g N_lg N_mg (row node first)
    visc_field = fs.get(abi.F_VISCOSITY)[0]            # (1 | n_nodes, dim, dim), [node, b, a] = V(a, b)
    mu = visc_field.reshape(-1)[0]
    general_visc = o.have_viscosity and (o.viscosity_shape != abi.TENSOR_ISOTROPIC or visc_field.shape[0] > 1)
    g = fs.get(abi.F_GRAVITY)[0].reshape(-1)[:dim]
    row_ptr, ent = strip_plan(mesh)
    f0, c0 = findrm - 1, colm - 1
    nnz = len(colm)
    big_m = np.zeros((dim, nnz))
    rhs = np.zeros((nn, dim))
    ml = np.zeros((nn, dim))
    dtt = o.dt * o.theta
    Qa, Qd = m["Qaaa"] - m["Qaab"], m["Qaab"] - m["Qabc"]
    for r in range(nn):
        s0, s1 = f0[r], f0[r + 1]
        acc = np.zeros(s1 - s0)
        mass = np.zeros(s1 - s0)
        own = int(np.searchsorted(c0[s0:s1], r))
        fifo = []  # (node, slot)
        msum = nbsum = 0.0
        alump = np.zeros(dim)
        for node1, meta in ent[row_ptr[r]:row_ptr[r + 1]]:
            fifo.append((node1 - 1, meta & 0xff))
            fifo = fifo[-dim:]
            if not (meta & COMPUTE):
                continue
            nodes = [q for q, _ in fifo]
            e = X[nodes] - X[r]
            c, det = _geometry(e)
            rd = 1.0 / det
            if lump_abs:
                # lumped absorption (:2041-2056): sum_j Ab^d_0j = sum_g N_0g sigma_dg rho_g detwei = |J| sum_lm Q_0lm sigma_dl rho_m
                ids_a = [r] + nodes
                alump += abs(det) * np.einsum("lm,ld,m->d", Q3, sig[ids_a], rho[ids_a])
            sc = c.sum(axis=0)
            S = rho[r] + rho[nodes].sum()
            M0 = Qa * rho[r] + m["Qaab"] * S
            w = M0 * nu[r]
            for k, q in enumerate(nodes):
                w = w + (Qd * (rho[r] + rho[q]) + m["Qabc"] * S) * nu[q]
            if src is not None:
                # add_sources_element_cg (:1737-1748): rhs(d, 0) += sum_k M_0k src(d, k), M = the density-weighted mass row
                ws = M0 * src[r]
                for k, q in enumerate(nodes):
                    ws = ws + (Qd * (rho[r] + rho[q]) + m["Qabc"] * S) * src[q]
                rsrc[r] += abs(det) * ws
            if general_visc:
                # dshape_tensor_dshape / dshape_diagtensor_dshape (:2318-2339): K_0k = |J| sum_g w_g gradN_0^T V_g gradN_k; V is P1
                # and the rule symmetric, so sum_g w_g V_g = (Wsum / loc) sum_l V_l (constant fields: Wsum V)
                ids_v = ([r] + nodes) if visc_field.shape[0] > 1 else [0]
                Vbar = m["Wsum"] * visc_field[ids_v].mean(axis=0).T        # Vbar[a, b]
                if o.viscosity_shape == abi.TENSOR_DIAGONAL:
                    Vbar = np.diag(np.diag(Vbar))
                elif o.viscosity_shape == abi.TENSOR_ISOTROPIC:
                    Vbar = Vbar[0, 0] * np.eye(dim)
                u = np.sign(det) * (w - rd * (sc @ Vbar))
            else:
                u = np.sign(det) * (w - (mu * m["Wsum"] * rd) * sc)
            tot = 0.0
            ad = abs(det)
            if o.integrate_advection_by_parts and not o.exclude_advection:
                # by-parts volume form (:1646-1669): A_0j = -gradN_0 . W_j - (1 - beta) div(nu) M_0j with
                # W_j = sum_g u_g N_jg rho_g detwei = |J| sum_k M~_jk nu_k (M~ = density-weighted mass moments, every row j)
                ids = [r] + nodes
                Mt = np.array([[(Qa * rho[a] + m["Qaab"] * S) if a == b else (Qd * (rho[a] + rho[b]) + m["Qabc"] * S)
                                for b in ids] for a in ids])
                Wj = Mt @ nu[ids]                                           # (loc, dim), without |J|
                divu = (float(-nu[r] @ sc) + sum(float(nu[q] @ c[k]) for k, q in enumerate(nodes))) * rd
                visc = -np.sign(det) * (mu * m["Wsum"] * rd)                # K_0k = visc * sc . c[k]
                row = np.sign(det) * (Wj @ sc) - (1.0 - o.beta) * divu * ad * Mt[0]      # advective part, columns ids
                tot_k = 0.0
                for k, (q, slot) in enumerate(fifo):
                    kk = visc * float(sc @ c[k])
                    acc[slot] += row[k + 1] + kk
                    tot_k += kk
                    if not o.lump_mass and not o.exclude_mass:
                        mass[slot] += ad * Mt[0, k + 1]
                acc[own] += row[0] - tot_k
                if not o.lump_mass and not o.exclude_mass:
                    mass[own] += ad * M0
                msum += ad * ((m["Pd"] - m["Po"]) * rho[r] + m["Po"] * S)
                nbsum += ad * ((m["Pd"] - m["Po"]) * bb[r] + m["Po"] * (bb[r] + bb[nodes].sum()))
                continue
            # beta term of the plain form (:1675-1680): beta * div(nu) * density-weighted mass row; for P1 div(nu) is
            # constant on the element: sum_k nu_k . gradN_k with gradN_0 = -sc/det, gradN_k = c[k]/det
            divu = (float(-nu[r] @ sc) + sum(float(nu[q] @ c[k]) for k, q in enumerate(nodes))) * rd
            bm = o.beta * divu * ad
            for k, (q, slot) in enumerate(fifo):
                sk = float(u @ c[k])
                Mk = Qd * (rho[r] + rho[q]) + m["Qabc"] * S
                acc[slot] += sk + bm * Mk
                if not o.lump_mass and not o.exclude_mass:
                    mass[slot] += ad * Mk               # consistent mass block (:1554-1556), not scaled by dt*theta
                tot += sk
            acc[own] += bm * M0 - tot
            if not o.lump_mass and not o.exclude_mass:
                mass[own] += ad * M0
            msum += ad * ((m["Pd"] - m["Po"]) * rho[r] + m["Po"] * S)
            nbsum += ad * ((m["Pd"] - m["Po"]) * bb[r] + m["Po"] * (bb[r] + bb[nodes].sum()))
        cols = c0[s0:s1]
        if o.have_source and o.lump_source:
            rsrc[r] = msum * fs.get(abi.F_SOURCE)[0][r]     # (:1739-1742): lumped density-weighted mass times the nodal source
        rhs[r] = o.gravity_magnitude * g * nbsum - acc @ oldu[cols] + rsrc[r] - alump * oldu[r]
        vals = dtt * acc + mass
        if o.lump_mass and not o.exclude_mass:
            vals[own] += msum
        big_m[:, s0:s1] = vals
        big_m[:, s0 + own] += dtt * alump
        ml[r] = msum + (dtt * alump if (lump_abs and o.pressure_corrected_absorption) else 0.0)
    return dict(big_m=big_m, rhs=rhs, masslump=ml)


def emulate_advdiff(mesh, fs, o, findrm, colm):
    """have_absorption / have_source (not in the device strip kernels yet: the closed forms planned for round 2,
    DESIGN.md section 7 item 2): Ab_0k = |J| [Qa s_0 + Qaab S | Qd (s_0 + s_k) + Qabc S]
    (Advection_Diffusion_CG.F90:1156), source rhs_0 += |J| [(Pd - Po) q_0 + Po sum q] (:1139)."""
    dim, nn = mesh.dim, mesh.n_nodes
    m = moments(dim)
    X = mesh.X
    nu, T = fs.get(abi.F_NU)[0], fs.get(abi.F_T)[0]
    sig = fs.get(abi.F_T_ABSORPTION)[0] if o.have_absorption else np.zeros(nn)
    src = fs.get(abi.F_T_SOURCE)[0] if o.have_source else np.zeros(nn)
    Qa, Qd = m["Qaaa"] - m["Qaab"], m["Qaab"] - m["Qabc"]
    kappa = fs.get(abi.F_T_DIFFUSIVITY)[0].reshape(-1)[0]
    row_ptr, ent = strip_plan(mesh)
    f0, c0 = findrm - 1, colm - 1
    matrix = np.zeros(len(colm))
    rhs = np.zeros(nn)
    dtt = o.dt * o.theta
    dtt = dtt if abs(dtt) > 2.220446049250313e-16 else 0.0
    for r in range(nn):
        s0, s1 = f0[r], f0[r + 1]
        a = np.zeros(s1 - s0)
        vol = np.zeros(s1 - s0)
        own = int(np.searchsorted(c0[s0:s1], r))
        fifo = []
        rh = 0.0
        for node1, meta in ent[row_ptr[r]:row_ptr[r + 1]]:
            fifo.append((node1 - 1, meta & 0xff))
            fifo = fifo[-dim:]
            if not (meta & COMPUTE):
                continue
            nodes = [q for q, _ in fifo]
            e = X[nodes] - X[r]
            c, det = _geometry(e)
            sc = c.sum(axis=0)
            Su = nu[r] + nu[nodes].sum(axis=0)
            v = (m["Pd"] - m["Po"]) * nu[r] + m["Po"] * Su
            u = np.sign(det) * (v - (kappa * m["Wsum"] / det) * sc)
            tot = 0.0
            ad = abs(det)
            Ss = sig[r] + sig[nodes].sum()
            # beta term (:1093-1098): beta * div(nu) * mass row, div(nu) constant on a P1 element
            divu = (float(-nu[r] @ sc) + sum(float(nu[q] @ c[k]) for k, q in enumerate(nodes))) / det
            bm = o.beta * divu * ad if o.have_advection else 0.0
            if o.have_advection and o.integrate_advection_by_parts:
                # by-parts volume form (:1043-1049): A_0j = -gradN_0 . W_j - (1 - beta) div(nu) P_0j, W_j = |J| [(Pd - Po) nu_j
                # + Po sum nu]; the second term is skipped when |1 - beta| <= epsilon. Diffusion as in the plain form.
                ids = [r] + nodes
                Wj = (m["Pd"] - m["Po"]) * nu[ids] + m["Po"] * Su                 # (loc, dim), without |J|
                row = np.sign(det) * (Wj @ sc)
                if abs(1.0 - o.beta) > 2.220446049250313e-16:
                    row = row - (1.0 - o.beta) * divu * ad * np.array([m["Pd"]] + [m["Po"]] * dim)
                dif = -np.sign(det) * (kappa * m["Wsum"] / det)
                tot_k = 0.0
                for k, (q, slot) in enumerate(fifo):
                    kk = dif * float(sc @ c[k])
                    ab = ad * (Qd * (sig[r] + sig[q]) + m["Qabc"] * Ss)
                    a[slot] += row[k + 1] + kk + ab
                    vol[slot] += ad
                    rh -= (row[k + 1] + kk + ab) * T[q]
                    tot_k += kk
                ab0 = ad * (Qa * sig[r] + m["Qaab"] * Ss)
                a[own] += row[0] - tot_k + ab0
                vol[own] += ad
                rh -= (row[0] - tot_k + ab0) * T[r]
                rh += ad * ((m["Pd"] - m["Po"]) * src[r] + m["Po"] * (src[r] + src[nodes].sum()))
                continue
            for k, (q, slot) in enumerate(fifo):
                sk = float(u @ c[k]) + bm * m["Po"]
                ab = ad * (Qd * (sig[r] + sig[q]) + m["Qabc"] * Ss)
                a[slot] += sk + ab
                vol[slot] += ad
                rh -= (sk + ab) * T[q]
                tot += sk
            ab0 = ad * (Qa * sig[r] + m["Qaab"] * Ss)
            # diagonal of the advective + diffusive part: minus the off-diagonal row sum of the beta-free part, plus
            # the beta mass diagonal
            d0 = -(tot - dim * bm * m["Po"]) + bm * m["Pd"]
            a[own] += ab0 + d0
            vol[own] += ad
            rh -= (ab0 + d0) * T[r]
            rh += ad * ((m["Pd"] - m["Po"]) * src[r] + m["Po"] * (src[r] + src[nodes].sum()))
        vals = dtt * a + m["Po"] * vol
        vals[own] = dtt * a[own] + m["Pd"] * vol[own]
        matrix[s0:s1] = vals
        rhs[r] = rh
    return dict(matrix=matrix, rhs=rhs)


def emulate_absorption_pass(mesh, fs, o, findrm, colm):
    """The second strip pass planned for round 2, emulated: per row the same strip,
    per computed window only |J| (one triple product) and the closed-form absorption row
    rho |J| [Qa s_0 + Qaab S | Qd (s_0 + s_k) + Qabc S] for every component d. Constant density (Boussinesq).
    Returns what has to be ADDED to the common result: big_m (dim, nnz) and rhs (n_nodes, dim)."""
    dim, nn = mesh.dim, mesh.n_nodes
    m = moments(dim)
    Qa, Qd = m["Qaaa"] - m["Qaab"], m["Qaab"] - m["Qabc"]
    X = mesh.X
    rho = float(fs.get(abi.F_DENSITY)[0].reshape(-1)[0])
    sig, oldu = fs.get(abi.F_ABSORPTION)[0], fs.get(abi.F_OLDU)[0]
    row_ptr, ent = strip_plan(mesh)
    f0, c0 = findrm - 1, colm - 1
    big_m = np.zeros((dim, len(colm)))
    rhs = np.zeros((nn, dim))
    dtt = o.dt * o.theta
    for r in range(nn):
        s0, s1 = f0[r], f0[r + 1]
        acc = np.zeros((dim, s1 - s0))
        own = int(np.searchsorted(c0[s0:s1], r))
        fifo = []
        for node1, meta in ent[row_ptr[r]:row_ptr[r + 1]]:
            fifo.append((node1 - 1, meta & 0xff))
            fifo = fifo[-dim:]
            if not (meta & COMPUTE):
                continue
            nodes = [q for q, _ in fifo]
            ad = rho * abs(np.linalg.det(X[nodes] - X[r]))
            S = sig[r] + sig[nodes].sum(axis=0)                      # (dim,)
            acc[:, own] += ad * (Qa * sig[r] + m["Qaab"] * S)
            for q, slot in fifo:
                acc[:, slot] += ad * (Qd * (sig[r] + sig[q]) + m["Qabc"] * S)
        cols = c0[s0:s1]
        big_m[:, s0:s1] = dtt * acc
        rhs[r] = -np.einsum("ds,sd->d", acc, oldu[cols])
    return dict(big_m=big_m, rhs=rhs)
