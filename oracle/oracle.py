"""ctypes wrapper of oracle/liborc.so (cg_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of cg_oracle.c. Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the
product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from fluidity_b200 import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build():
    """Compiles liborc.so with the committed Makefile (gcc; seconds)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


_FLAVOUR = "strict"


def select(flavour):
    """"strict" (default): liborc.so, -ffp-contract=off, baseline x86-64-v2 -- the build every parity test uses.
    "fast": liborc_fast.so, -O3 -march=native with FMA contraction, compiled on THIS host on first use -- only for
    the timing legs of bench.py (cpu_baseline, --impl reference): the CPU arm at its best, not a parity oracle."""
    global _FLAVOUR, _LIB
    assert flavour in ("strict", "fast")
    if flavour != _FLAVOUR:
        _FLAVOUR, _LIB = flavour, None


def lib():
    global _LIB
    if _LIB is None:
        name = "liborc.so" if _FLAVOUR == "strict" else "liborc_fast.so"
        path = os.path.join(_HERE, name)
        src = os.path.join(_HERE, "cg_oracle.c")
        if _FLAVOUR == "fast":
            # -march=native: always rebuilt where it runs (a copy built on another host may not run here)
            subprocess.run(["make", "-s", "-B", "-C", _HERE, name], check=True)
        elif not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_make_sparsity.restype = C.c_int
        _LIB.orc_colour_elements.restype = C.c_int
    return _LIB


def set_threads(n):
    """OpenMP threads of the assembly loops (omp_set_num_threads in the library's own runtime)."""
    lib().orc_set_threads(C.c_int(int(n)))


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


class _Field(C.Structure):
    _fields_ = [("val", c_dp), ("field_type", C.c_int)]


class _Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("loc", C.c_int), ("ngi", C.c_int), ("n_nodes", C.c_int),
                ("n_elements", C.c_int), ("ndglno", c_ip), ("n", c_dp), ("dn", c_dp),
                ("weight", c_dp), ("X", c_dp)]


class _MomFields(C.Structure):
    _fields_ = [(k, _Field) for k in ("nu", "oldu", "density", "viscosity", "buoyancy",
                                     "hb_density", "gravity", "absorption", "source")]


class _AdvFields(C.Structure):
    _fields_ = [(k, _Field) for k in ("t", "velocity", "source", "absorption", "diffusivity")]


_MOM_SLOTS = {"nu": abi.F_NU, "oldu": abi.F_OLDU, "density": abi.F_DENSITY,
              "viscosity": abi.F_VISCOSITY, "buoyancy": abi.F_BUOYANCY,
              "hb_density": abi.F_HB_DENSITY, "gravity": abi.F_GRAVITY,
              "absorption": abi.F_ABSORPTION, "source": abi.F_SOURCE}
_ADV_SLOTS = {"t": abi.F_T, "velocity": abi.F_NU, "source": abi.F_T_SOURCE,
              "absorption": abi.F_T_ABSORPTION, "diffusivity": abi.F_T_DIFFUSIVITY}


def quadrature(dim):
    """(l (ngi, loc), weight (ngi)) of the degree-3 rule."""
    loc = dim + 1
    l = np.zeros((loc, 5), dtype=np.float64)  # column-major (ngi, loc) with ngi<=5
    w = np.zeros(5)
    ngi = 5 if dim == 3 else 4
    lbuf = np.zeros(ngi * loc)
    lib().orc_quadrature_degree3(C.c_int(dim), _dp(lbuf), _dp(w))
    l = lbuf.reshape(loc, ngi).T.copy()  # l[g, j]
    return l, w[:ngi].copy()


def tables(dim):
    """P1 tables in the raw column-major buffers the C ABI takes:
    n (loc*ngi), dn (loc*ngi*dim), weight (ngi)."""
    loc = dim + 1
    l, w = quadrature(dim)
    ngi = len(w)
    lbuf = np.ascontiguousarray(l.T).ravel()  # l[g + ngi*j]
    n = np.zeros(loc * ngi)
    dn = np.zeros(loc * ngi * dim)
    lib().orc_shape_p1(C.c_int(dim), C.c_int(ngi), _dp(lbuf), _dp(n), _dp(dn))
    return n, dn, w


def transform_to_physical(dim, X_val, want_J=False):
    """X_val (loc, dim) rows = node positions. Returns dshape (loc, ngi, dim), detwei, J."""
    loc = dim + 1
    n, dn, w = tables(dim)
    ngi = len(w)
    Xv = np.ascontiguousarray(X_val, dtype=np.float64)  # memory = X_val(dim, loc) col-major
    dshape = np.zeros(loc * ngi * dim)
    detwei = np.zeros(ngi)
    J = np.zeros(dim * dim * ngi) if want_J else None
    lib().orc_transform_to_physical(C.c_int(dim), C.c_int(ngi), _dp(Xv), _dp(dn), _dp(w),
                                    _dp(dshape), _dp(detwei), _dp(J))
    ds = dshape.reshape(dim, ngi, loc).transpose(2, 1, 0).copy()
    Jm = J.reshape(ngi, dim, dim).transpose(2, 1, 0).copy() if want_J else None  # J[a,k,g]
    return ds, detwei, Jm


def make_sparsity(mesh):
    """findrm (n+1), colm (nnz), centrm (n): 1-based, as lists2csr_sparsity builds them."""
    f, c, ce = c_ip(), c_ip(), c_ip()
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    nnz = lib().orc_make_sparsity(C.c_int(mesh.n_nodes), C.c_int(mesh.n_elements),
                                  C.c_int(mesh.loc), _ip(nd), C.byref(f), C.byref(c), C.byref(ce))
    findrm = np.ctypeslib.as_array(f, shape=(mesh.n_nodes + 1,)).copy()
    colm = np.ctypeslib.as_array(c, shape=(max(nnz, 1),)).copy()[:nnz]
    centrm = np.ctypeslib.as_array(ce, shape=(mesh.n_nodes,)).copy()
    for p in (f, c, ce):
        lib().orc_free(p)
    return findrm, colm, centrm


def colour_elements(mesh):
    """(colour_of (n_elements) 1-based colours, ncolours)."""
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    col = np.zeros(mesh.n_elements, dtype=np.int32)
    nc = lib().orc_colour_elements(C.c_int(mesh.n_nodes), C.c_int(mesh.n_elements),
                                   C.c_int(mesh.loc), _ip(nd), _ip(col))
    return col, nc


def colour_sets(colour_of, ncolours):
    """colour_sets (femtools/Colouring.F90:250-262): colour_ptr (ncolours+1, 1-based offsets),
    colour_elements (1-based ids ascending inside each colour)."""
    order = np.argsort(colour_of, kind="stable").astype(np.int32)
    counts = np.bincount(colour_of, minlength=ncolours + 1)[1:]
    ptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int32)
    return ptr, (order + 1).astype(np.int32)


class _Ctx:
    """Keeps numpy buffers alive while C structs point at them."""

    def __init__(self, mesh, fields):
        self.keep = []
        dim = mesh.dim
        n, dn, w = tables(dim)
        nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
        X = np.ascontiguousarray(mesh.X, dtype=np.float64)
        self.keep += [n, dn, w, nd, X]
        self.mesh = _Mesh(dim, dim + 1, len(w), mesh.n_nodes, mesh.n_elements, _ip(nd), _dp(n),
                          _dp(dn), _dp(w), _dp(X))
        self.fields = fields

    def _field(self, slot):
        if slot not in self.fields.data:
            return _Field(None, 0)
        val, ft = self.fields.get(slot)
        self.keep.append(val)
        return _Field(_dp(val), ft)

    def mom(self):
        return _MomFields(*[self._field(_MOM_SLOTS[k]) for k, _ in _MomFields._fields_])

    def adv(self):
        return _AdvFields(*[self._field(_ADV_SLOTS[k]) for k, _ in _AdvFields._fields_])


def momentum_element(mesh, fields, opts, ele):
    """Returns big_m_tensor_addto (dim,dim,loc,loc), rhs_addto (dim,loc), masslump_addto
    (dim,loc), grad_p_u_mat (dim,loc,loc) as numpy arrays indexed like the Fortran ones."""
    ctx = _Ctx(mesh, fields)
    dim, loc = mesh.dim, mesh.loc
    T = np.zeros(dim * dim * loc * loc)
    r = np.zeros(dim * loc)
    ml = np.zeros(dim * loc)
    gp = np.zeros(dim * loc * loc)
    mf = ctx.mom()
    st = lib().orc_momentum_element(C.byref(ctx.mesh), C.byref(mf), C.byref(opts), C.c_int(ele),
                                    _dp(T), _dp(r), _dp(ml), _dp(gp))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return (T.reshape(loc, loc, dim, dim).transpose(3, 2, 1, 0).copy(),
            r.reshape(loc, dim).T.copy(), ml.reshape(loc, dim).T.copy(),
            gp.reshape(loc, loc, dim).transpose(2, 1, 0).copy())


def advdiff_element(mesh, fields, opts, ele):
    ctx = _Ctx(mesh, fields)
    loc = mesh.loc
    A = np.zeros(loc * loc)
    r = np.zeros(loc)
    af = ctx.adv()
    st = lib().orc_advdiff_element(C.byref(ctx.mesh), C.byref(af), C.byref(opts), C.c_int(ele),
                                   _dp(A), _dp(r))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return A.reshape(loc, loc).T.copy(), r


def assemble_momentum(mesh, fields, opts, findrm, colm, colouring=None, want_masslump=True,
                      want_ct=False):
    """Serial reference order unless colouring=(colour_ptr, colour_elements) is given (then
    OpenMP over each colour like the reference). Returns dict of big_m (dim, nnz), rhs
    (n_nodes, dim), masslump (n_nodes, dim) | None, ct_m (dim, nnz) | None."""
    ctx = _Ctx(mesh, fields)
    dim = mesh.dim
    nnz = int(findrm[-1] - 1)
    big_m = np.zeros((dim, nnz))
    rhs = np.zeros((mesh.n_nodes, dim))
    ml = np.zeros((mesh.n_nodes, dim)) if want_masslump else None
    ct = np.zeros((dim, nnz)) if want_ct else None
    findrm = np.ascontiguousarray(findrm, dtype=np.int32)
    colm = np.ascontiguousarray(colm, dtype=np.int32)
    if colouring is not None:
        cptr = np.ascontiguousarray(colouring[0], dtype=np.int32)
        cel = np.ascontiguousarray(colouring[1], dtype=np.int32)
        nc = len(cptr) - 1
    else:
        cptr = cel = None
        nc = 0
    mf = ctx.mom()
    st = lib().orc_assemble_momentum(C.byref(ctx.mesh), C.byref(mf), C.byref(opts), _ip(findrm),
                                     _ip(colm), C.c_int(nc), _ip(cptr), _ip(cel), _dp(big_m),
                                     _dp(rhs), _dp(ml), _dp(ct))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return dict(big_m=big_m, rhs=rhs, masslump=ml, ct_m=ct)


def assemble_momentum_mass(mesh, fields, opts, findrm, colm):
    """The `mass` matrix of construct_momentum_cg (assemble_mass_matrix): (dim, nnz) diagonal blocks."""
    ctx = _Ctx(mesh, fields)
    nnz = int(findrm[-1] - 1)
    mass = np.zeros((mesh.dim, nnz))
    mf = ctx.mom()
    st = lib().orc_assemble_momentum_mass(C.byref(ctx.mesh), C.byref(mf), C.byref(opts),
                                          _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                          _ip(np.ascontiguousarray(colm, dtype=np.int32)), _dp(mass))
    if st:
        raise RuntimeError("orc_assemble_momentum_mass status %d" % st)
    return mass


def assemble_advdiff(mesh, fields, opts, findrm, colm, colouring=None):
    ctx = _Ctx(mesh, fields)
    nnz = int(findrm[-1] - 1)
    val = np.zeros(nnz)
    rhs = np.zeros(mesh.n_nodes)
    findrm = np.ascontiguousarray(findrm, dtype=np.int32)
    colm = np.ascontiguousarray(colm, dtype=np.int32)
    if colouring is not None:
        cptr = np.ascontiguousarray(colouring[0], dtype=np.int32)
        cel = np.ascontiguousarray(colouring[1], dtype=np.int32)
        nc = len(cptr) - 1
    else:
        cptr = cel = None
        nc = 0
    af = ctx.adv()
    st = lib().orc_assemble_advdiff(C.byref(ctx.mesh), C.byref(af), C.byref(opts), _ip(findrm),
                                    _ip(colm), C.c_int(nc), _ip(cptr), _ip(cel), _dp(val), _dp(rhs))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return dict(matrix=val, rhs=rhs)


def block_addto(findrm, colm, rows, cols, vals, val):
    """Accumulates vals (nrows, ncols) at (rows, cols) (1-based) into val in place."""
    findrm = np.ascontiguousarray(findrm, dtype=np.int32)
    colm = np.ascontiguousarray(colm, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    v = np.asfortranarray(vals, dtype=np.float64)
    lib().orc_block_addto(_ip(findrm), _ip(colm), C.c_int(len(rows)), _ip(rows), C.c_int(len(cols)),
                          _ip(cols), v.ctypes.data_as(c_dp), _dp(val))


def halo_copy(block_size, field_p, sends_p_to_q, field_q, recvs_q_from_p):
    s = np.ascontiguousarray(sends_p_to_q, dtype=np.int32)
    r = np.ascontiguousarray(recvs_q_from_p, dtype=np.int32)
    assert len(s) == len(r)
    lib().orc_halo_copy(C.c_int(block_size), _dp(field_p), _ip(s), C.c_int(len(s)), _dp(field_q),
                        _ip(r))


# ---- surface loops and Dirichlet conditions -------------------------------------------
class _Surface(C.Structure):
    _fields_ = [("n_faces", C.c_int), ("sloc", C.c_int), ("sngi", C.c_int), ("sndgln", c_ip), ("face_ele", c_ip),
                ("n_f", c_dp), ("dn_f", c_dp), ("weight_f", c_dp)]


def face_tables(dim):
    """n_f, dn_f, weight_f of the face element in the raw column-major layout (restated rule)."""
    sloc = dim
    l = np.zeros(4 * sloc)
    w = np.zeros(4)
    sngi = lib().orc_quadrature_face_degree3(C.c_int(dim), _dp(l), _dp(w))
    n = np.zeros(sloc * sngi)
    dn = np.zeros(sloc * sngi * (dim - 1))
    lib().orc_shape_p1(C.c_int(dim - 1), C.c_int(sngi), _dp(l), _dp(n), _dp(dn))
    return n, dn, w[:sngi].copy()


def transform_facet_to_physical(dim, X_f, X_val):
    """X_f (sloc, dim), X_val (loc, dim) rows = node positions. Returns detwei_f (sngi), normal (sngi, dim)."""
    n, dn, w = face_tables(dim)
    sngi = len(w)
    detwei = np.zeros(sngi)
    normal = np.zeros(dim * sngi)
    lib().orc_transform_facet_to_physical(C.c_int(dim), C.c_int(dim), C.c_int(sngi),
                                          _dp(np.ascontiguousarray(X_f, dtype=np.float64)),
                                          _dp(np.ascontiguousarray(X_val, dtype=np.float64)), _dp(dn), _dp(w),
                                          _dp(detwei), _dp(normal))
    return detwei, normal.reshape(sngi, dim)


class _SurfCtx(_Ctx):
    def __init__(self, mesh, fields, sndgln, face_ele):
        super().__init__(mesh, fields)
        n, dn, w = face_tables(mesh.dim)
        sn = np.ascontiguousarray(sndgln, dtype=np.int32)
        fe = np.ascontiguousarray(face_ele, dtype=np.int32)
        self.keep += [n, dn, w, sn, fe]
        self.n_faces = len(fe)
        self.surface = _Surface(len(fe), mesh.dim, len(w), _ip(sn), _ip(fe), _dp(n), _dp(dn), _dp(w))


def advdiff_face(mesh, fields, opts, sndgln, face_ele, face, bc_type, t_bc=None, t_bc_2=None):
    """matrix_addto (sloc, sloc), rhs_addto (sloc) of one face (1-based)."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    sloc = mesh.dim
    A = np.zeros(sloc * sloc)
    r = np.zeros(sloc)
    b1 = np.ascontiguousarray(t_bc if t_bc is not None else np.zeros(sloc), dtype=np.float64)
    b2 = np.ascontiguousarray(t_bc_2 if t_bc_2 is not None else np.zeros(sloc), dtype=np.float64)
    af = ctx.adv()
    st = lib().orc_advdiff_face(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(af), C.byref(opts), C.c_int(face),
                                C.c_int(bc_type), _dp(b1), _dp(b2), _dp(A), _dp(r))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return A.reshape(sloc, sloc).T.copy(), r


def assemble_advdiff_surface(mesh, fields, opts, findrm, colm, sndgln, face_ele, bc_type, t_bc, t_bc_2, matrix, rhs):
    """ADDS the face loop to matrix (nnz) / rhs (n_nodes) in place. t_bc, t_bc_2: (n_faces, sloc) or None."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    bt = np.ascontiguousarray(bc_type, dtype=np.int32)
    b1 = np.ascontiguousarray(t_bc, dtype=np.float64) if t_bc is not None else None
    b2 = np.ascontiguousarray(t_bc_2, dtype=np.float64) if t_bc_2 is not None else None
    af = ctx.adv()
    st = lib().orc_assemble_advdiff_surface(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(af), C.byref(opts),
                                            _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                            _ip(np.ascontiguousarray(colm, dtype=np.int32)), _ip(bt), _dp(b1), _dp(b2),
                                            _dp(matrix), _dp(rhs))
    if st:
        raise RuntimeError("oracle status %d" % st)


def apply_dirichlet_scalar(nodes, values, field, dt, rhs, inactive=None):
    """In place on rhs (and the int32 `inactive` flags). dt None: the rhs receives the value itself."""
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    lib().orc_apply_dirichlet_scalar(C.c_int(len(nd)), _ip(nd), _dp(np.ascontiguousarray(values, dtype=np.float64)),
                                     _dp(np.ascontiguousarray(field, dtype=np.float64)), C.c_int(0 if dt is None else 1),
                                     C.c_double(dt or 0.0), _dp(rhs), _ip(inactive))


def momentum_face(mesh, fields, opts, sndgln, face_ele, face, velocity_bc_type, velocity_bc=None, want_masslump=False):
    """big_m_addto (dim, sloc, sloc), rhs_addto (dim, sloc) of one face (+ masslump_addto (dim, sloc) if asked: the
    free-surface stabilisation adds to it). velocity_bc (sloc, dim)."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    dim = sloc = mesh.dim
    B = np.zeros(dim * sloc * sloc)
    r = np.zeros(dim * sloc)
    ml = np.zeros(dim * sloc)
    bt = np.ascontiguousarray(velocity_bc_type, dtype=np.int32)
    bv = np.ascontiguousarray(velocity_bc if velocity_bc is not None else np.zeros((sloc, dim)), dtype=np.float64)
    mf = ctx.mom()
    st = lib().orc_momentum_face_ml(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(mf), C.byref(opts), C.c_int(face),
                                    _ip(bt), _dp(bv), _dp(B), _dp(r), _dp(ml))
    if st:
        raise RuntimeError("oracle status %d" % st)
    out = (B.reshape(sloc, sloc, dim).transpose(2, 1, 0).copy(), r.reshape(sloc, dim).T.copy())
    return out + (ml.reshape(sloc, dim).T.copy(),) if want_masslump else out


def assemble_momentum_surface(mesh, fields, opts, findrm, colm, sndgln, face_ele, velocity_bc_type, velocity_bc,
                              big_m, rhs, pressure_bc_type=None, masslump=None):
    """ADDS the surface loop to big_m (dim, nnz) / rhs (n_nodes, dim) / masslump (n_nodes, dim; free-surface
    stabilisation only) in place. velocity_bc_type (n_faces, dim), velocity_bc (n_faces, sloc, dim)."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    bt = np.ascontiguousarray(velocity_bc_type, dtype=np.int32)
    bv = np.ascontiguousarray(velocity_bc, dtype=np.float64)
    pt = np.ascontiguousarray(pressure_bc_type, dtype=np.int32) if pressure_bc_type is not None else None
    mf = ctx.mom()
    st = lib().orc_assemble_momentum_surface_ml(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(mf), C.byref(opts),
                                                _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                                _ip(np.ascontiguousarray(colm, dtype=np.int32)), _ip(bt), _dp(bv), _ip(pt),
                                                _dp(big_m), _dp(rhs), _dp(masslump))
    if st:
        raise RuntimeError("oracle status %d" % st)


# ---- lumped-mass pressure matrix ----------------------------------------------------------
def make_sparsity_mult(n_nodes, findrm, colm):
    """Second-order sparsity (make_sparsity_mult, P1-P1): findrm2, colm2, 1-based."""
    f, c = c_ip(), c_ip()
    lib().orc_make_sparsity_mult.restype = C.c_int
    nnz = lib().orc_make_sparsity_mult(C.c_int(n_nodes), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(colm, dtype=np.int32)), C.byref(f), C.byref(c))
    findrm2 = np.ctypeslib.as_array(f, shape=(n_nodes + 1,)).copy()
    colm2 = np.ctypeslib.as_array(c, shape=(max(nnz, 1),)).copy()[:nnz]
    lib().orc_free(f)
    lib().orc_free(c)
    return findrm2, colm2


def mult_div_vector_div_T(findrm, colm, ct1, ct2, vfield, findrm2, colm2):
    """product = ct1 . diag(vfield) . ct2^T on the second-order sparsity. ct (dim, nnz); vfield (n_nodes, dim)."""
    ct1 = np.ascontiguousarray(ct1, dtype=np.float64)
    ct2 = np.ascontiguousarray(ct2, dtype=np.float64)
    v = np.ascontiguousarray(vfield, dtype=np.float64)
    dim, n_nodes = ct1.shape[0], len(findrm) - 1
    out = np.zeros(len(colm2))
    lib().orc_mult_div_vector_div_T(C.c_int(dim), C.c_int(n_nodes), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                    _ip(np.ascontiguousarray(colm, dtype=np.int32)), _dp(ct1), _dp(ct2), _dp(v),
                                    _ip(np.ascontiguousarray(findrm2, dtype=np.int32)),
                                    _ip(np.ascontiguousarray(colm2, dtype=np.int32)), _dp(out))
    return out


# ---- P1-P1 pressure stabilisation (assemble_kmk_matrix) ---------------------------------------------------------
def simplex_tensor(pos_ele):
    """pos_ele (loc, dim) -> the element's metric tensor (dim, dim)."""
    pos = np.ascontiguousarray(pos_ele, dtype=np.float64)
    dim = pos.shape[1]
    m = np.zeros((dim, dim))
    lib().orc_simplex_tensor.restype = C.c_int
    if lib().orc_simplex_tensor(C.c_int(dim), _dp(pos), _dp(m)):
        raise RuntimeError("degenerate element")
    return m


def edge_length_from_metric(metric):
    m = np.ascontiguousarray(metric, dtype=np.float64)
    out = np.zeros_like(m)
    lib().orc_edge_length_from_metric(C.c_int(m.shape[0]), _dp(m), _dp(out))
    return out


def assemble_kt(mesh, findrm, colm):
    """The pressure diffusion matrix kt (nnz) of assemble_kmk_matrix and the lumped pressure mass (n_nodes)."""
    ctx = _Ctx(mesh, None)
    nnz = int(findrm[-1]) - 1
    kt, ml = np.zeros(nnz), np.zeros(mesh.n_nodes)
    lib().orc_assemble_kt.restype = C.c_int
    st = lib().orc_assemble_kt(C.byref(ctx.mesh), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                               _ip(np.ascontiguousarray(colm, dtype=np.int32)), _dp(kt), _dp(ml))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return kt, ml


def mult_div_invscalar_div_T(findrm, colm, m1, sfield, m2, findrm2, colm2):
    out = np.zeros(len(colm2))
    lib().orc_mult_div_invscalar_div_T(C.c_int(len(findrm) - 1), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(colm, dtype=np.int32)),
                                       _dp(np.ascontiguousarray(m1, dtype=np.float64)),
                                       _dp(np.ascontiguousarray(sfield, dtype=np.float64)),
                                       _dp(np.ascontiguousarray(m2, dtype=np.float64)),
                                       _ip(np.ascontiguousarray(findrm2, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(colm2, dtype=np.int32)), _dp(out))
    return out


def assemble_kmk(mesh, findrm, colm, findrm2, colm2, theta_pg=1.0):
    """kmk = kt diag(1 / (theta_pg p_masslump)) kt^T on the second-order sparsity (Momentum_CG.F90:2755-2763)."""
    kt, ml = assemble_kt(mesh, findrm, colm)
    s = ml if abs(theta_pg - 1.0) < np.finfo(float).eps else ml * theta_pg
    return mult_div_invscalar_div_T(findrm, colm, kt, s, kt, findrm2, colm2), kt, ml


def momentum_face_ct(mesh, fields, opts, sndgln, face_ele, face, velocity_bc_type, velocity_bc=None, pressure_bc_type=0,
                     pressure_bc=None, hb_pressure=None, include_pressure_and_continuity_bcs=False):
    """Continuity half of the momentum surface element: ct_addto (dim, sloc [p], sloc [u]), ct_rhs_addto (sloc),
    rhs_addto (dim, sloc)."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    dim = sloc = mesh.dim
    Cb, cr, r = np.zeros(dim * sloc * sloc), np.zeros(sloc), np.zeros(dim * sloc)
    bt = np.ascontiguousarray(velocity_bc_type, dtype=np.int32)
    bv = np.ascontiguousarray(velocity_bc if velocity_bc is not None else np.zeros((sloc, dim)), dtype=np.float64)
    pb = np.ascontiguousarray(pressure_bc, dtype=np.float64) if pressure_bc is not None else None
    hb = np.ascontiguousarray(hb_pressure, dtype=np.float64) if hb_pressure is not None else None
    st = lib().orc_momentum_face_ct(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(opts), C.c_int(face), _ip(bt), _dp(bv),
                                    C.c_int(pressure_bc_type), _dp(pb), _dp(hb),
                                    C.c_int(1 if include_pressure_and_continuity_bcs else 0), _dp(Cb), _dp(cr), _dp(r))
    if st:
        raise RuntimeError("oracle status %d" % st)
    return Cb.reshape(sloc, sloc, dim).transpose(2, 1, 0).copy(), cr, r.reshape(sloc, dim).T.copy()


def assemble_ct_surface(mesh, fields, opts, findrm, colm, sndgln, face_ele, velocity_bc_type, ct_m, pressure_bc_type=None):
    """ADDS the continuity boundary blocks to ct_m (dim, nnz) in place."""
    ctx = _SurfCtx(mesh, fields, sndgln, face_ele)
    bt = np.ascontiguousarray(velocity_bc_type, dtype=np.int32)
    pt = np.ascontiguousarray(pressure_bc_type, dtype=np.int32) if pressure_bc_type is not None else None
    st = lib().orc_assemble_ct_surface(C.byref(ctx.mesh), C.byref(ctx.surface), C.byref(opts),
                                       _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(colm, dtype=np.int32)), _ip(bt), _ip(pt), _dp(ct_m))
    if st:
        raise RuntimeError("oracle status %d" % st)


def correct_masslumped_velocity(findrm, colm, ct_m, inverse_masslump, delta_p, u):
    """u (n_nodes, dim) corrected IN PLACE: u_d += inverse_masslump_d * (ct_m_d^T delta_p) (Momentum_CG.F90:2544-2575)."""
    n, dim = u.shape
    lib().orc_correct_masslumped_velocity(C.c_int(dim), C.c_int(n), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                          _ip(np.ascontiguousarray(colm, dtype=np.int32)),
                                          _dp(np.ascontiguousarray(ct_m)), _dp(np.ascontiguousarray(inverse_masslump)),
                                          _dp(np.ascontiguousarray(delta_p)), _dp(u))
    return u


def lift_boundary_conditions(findrm, colm, big_m, rhs, nodes, comps):
    """Strong Dirichlet rows of big_m (dim, nnz) / rhs (n_nodes, dim), IN PLACE; rhs already holds the boundary values in
    the listed (node, component) entries (collect_vector_dirichlet_conditions)."""
    n, dim = rhs.shape
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    comps = np.ascontiguousarray(comps, dtype=np.int32)
    lib().orc_lift_boundary_conditions(C.c_int(dim), C.c_int(n), _ip(np.ascontiguousarray(findrm, dtype=np.int32)),
                                       _ip(np.ascontiguousarray(colm, dtype=np.int32)), _dp(big_m), _dp(rhs),
                                       C.c_int(len(nodes)), _ip(nodes), _ip(comps))
