"""Host-side binding of libcgasm.so (the C ABI of include/cgasm.h) for Python callers.

This is the Python twin of the Fortran shim in fluidity_b200/fortran/cgasm_fortran.F90: it
only marshals numpy buffers (already laid out like the reference's Fortran arrays, see
synthetic.py) into the C calls. There is no computation here and no CPU path: if the
library or a B200 is missing every call raises.
"""
import ctypes as C
import os
import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# CGASM_LIB: another build of the same library (A/B timing of two builds on one GPU box)
LIB_PATH = os.environ.get("CGASM_LIB") or os.path.join(_HERE, "libcgasm.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class CgasmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cgasm error %d: %s" % (code, msg))
        self.code = code


def load():
    """Loads libcgasm.so (built in-tree by fluidity_b200/csrc/Makefile). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            "%s not built: run `make -C fluidity_b200/csrc` (or __graft_entry__.build()); "
            "there is no fallback path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.cgasm_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _check(st):
    if st != 0:
        raise CgasmError(st, load().cgasm_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def cmc_sparsity_host(findrm, colm):
    """cgasm_cmc_sparsity_host: second-order pattern (1-based) of a first-order one, on the host."""
    lib = load()
    f = np.ascontiguousarray(findrm, dtype=np.int32)
    c = np.ascontiguousarray(colm, dtype=np.int32)
    n = len(f) - 1
    f2 = np.zeros(n + 1, dtype=np.int32)
    need = C.c_longlong(0)
    _check(lib.cgasm_cmc_sparsity_host(C.c_int(n), _ip(f), _ip(c), _ip(f2), None, C.c_longlong(0), C.byref(need)))
    c2 = np.zeros(need.value, dtype=np.int32)
    _check(lib.cgasm_cmc_sparsity_host(C.c_int(n), _ip(f), _ip(c), _ip(f2), _ip(c2), C.c_longlong(need.value), C.byref(need)))
    return f2, c2


def cmc_expand_plan_host(findrm, colm, findrm2, colm2):
    """cgasm_cmc_expand_plan_host: (tpos, pptr, slots, n2max) or None if the patterns admit no expansion plan."""
    lib = load()
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (findrm, colm, findrm2, colm2)]
    n = len(a[0]) - 1
    tpos = np.zeros(len(a[1]), dtype=np.int32)
    pptr = np.zeros(n + 1, dtype=np.int64)
    need, n2max = C.c_longlong(0), C.c_int(0)
    args = [C.c_int(n)] + [_ip(x) for x in a] + [_ip(tpos), pptr.ctypes.data_as(C.POINTER(C.c_longlong))]
    _check(lib.cgasm_cmc_expand_plan_host(*args, None, C.c_longlong(0), C.byref(need), C.byref(n2max)))
    if need.value < 0:
        return None
    slots = np.zeros(need.value, dtype=np.uint16)
    _check(lib.cgasm_cmc_expand_plan_host(*args, slots.ctypes.data_as(C.POINTER(C.c_ushort)), C.c_longlong(need.value),
                                          C.byref(need), C.byref(n2max)))
    return tpos, pptr, slots, n2max.value


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(load().cgasm_nccl_unique_id(buf))
    return buf.raw


class Assembler:
    """One handle = one mesh resident on one GPU (cgasm_create ... cgasm_destroy)."""

    def __init__(self, mesh, tables, device=-1):
        """mesh: synthetic.Mesh-like (dim, ndglno (E,loc) 1-based int32, X (N,dim));
        tables: (n, dn, weight) raw column-major element_type tables."""
        self.lib = load()
        self.dim, self.loc = mesh.dim, mesh.dim + 1
        self.n_nodes, self.n_elements = mesh.n_nodes, mesh.n_elements
        n, dn, w = tables
        nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
        ident = C.c_int(0)
        _check(self.lib.cgasm_create(C.byref(ident), C.c_int(device), C.c_int(self.dim),
                                     C.c_int(self.loc), C.c_int(len(w)), C.c_int(self.n_nodes),
                                     C.c_int(self.n_elements), _ip(nd),
                                     _dp(np.ascontiguousarray(n)), _dp(np.ascontiguousarray(dn)),
                                     _dp(np.ascontiguousarray(w))))
        self.id = ident.value
        self.nnz = None
        self.set_coordinates(mesh.X)

    def close(self):
        if getattr(self, "id", 0):
            self.lib.cgasm_destroy(C.c_int(self.id))
            self.id = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mesh-level state -------------------------------------------------------------
    def set_coordinates(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        assert X.shape == (self.n_nodes, self.dim)
        _check(self.lib.cgasm_set_coordinates(C.c_int(self.id), _dp(X)))

    def build_sparsity(self):
        nnz = C.c_int(0)
        _check(self.lib.cgasm_build_sparsity(C.c_int(self.id), C.byref(nnz)))
        self.nnz = nnz.value
        return self.nnz

    def get_sparsity(self):
        findrm = np.zeros(self.n_nodes + 1, dtype=np.int32)
        colm = np.zeros(self.nnz, dtype=np.int32)
        centrm = np.zeros(self.n_nodes, dtype=np.int32)
        _check(self.lib.cgasm_get_sparsity(C.c_int(self.id), _ip(findrm), _ip(colm), _ip(centrm)))
        return findrm, colm, centrm

    def set_sparsity(self, findrm, colm):
        findrm = np.ascontiguousarray(findrm, dtype=np.int32)
        colm = np.ascontiguousarray(colm, dtype=np.int32)
        _check(self.lib.cgasm_set_sparsity(C.c_int(self.id), C.c_int(len(findrm) - 1),
                                           C.c_int(len(colm)), _ip(findrm), _ip(colm)))
        self.nnz = len(colm)

    def build_colouring(self):
        nc = C.c_int(0)
        _check(self.lib.cgasm_build_colouring(C.c_int(self.id), C.byref(nc)))
        return nc.value

    def get_colouring(self, ncolours):
        ptr = np.zeros(ncolours + 1, dtype=np.int32)
        els = np.zeros(self.n_elements, dtype=np.int32)
        _check(self.lib.cgasm_get_colouring(C.c_int(self.id), _ip(ptr), _ip(els)))
        return ptr, els

    def set_colouring(self, colour_ptr, colour_elements):
        ptr = np.ascontiguousarray(colour_ptr, dtype=np.int32)
        els = np.ascontiguousarray(colour_elements, dtype=np.int32)
        _check(self.lib.cgasm_set_colouring(C.c_int(self.id), C.c_int(len(ptr) - 1), _ip(ptr), _ip(els)))

    def set_scatter(self, variant):
        _check(self.lib.cgasm_set_scatter(C.c_int(self.id), C.c_int(variant)))

    # -- fields -----------------------------------------------------------------------
    def set_field(self, slot, val, field_type=abi.FIELD_NORMAL):
        val = np.ascontiguousarray(val, dtype=np.float64)
        n_val_nodes = 1 if field_type == abi.FIELD_CONSTANT else self.n_nodes
        _check(self.lib.cgasm_set_field(C.c_int(self.id), C.c_int(slot), C.c_int(abi.FIELD_RANK[slot]),
                                        C.c_int(field_type), _dp(val), C.c_int(n_val_nodes)))

    def set_fields(self, fieldset):
        for slot, (val, ft) in fieldset.items():
            self.set_field(slot, val, ft)

    def get_field(self, slot, shape):
        out = np.zeros(shape, dtype=np.float64)
        _check(self.lib.cgasm_get_field(C.c_int(self.id), C.c_int(slot), _dp(out), C.c_int(shape[0])))
        return out

    # -- the two element loops ------------------------------------------------------------
    def momentum(self, opts, out=None):
        """Host-buffer call (what the Fortran shim makes). Returns dict of big_m (dim,nnz),
        rhs (N,dim), masslump (N,dim)|None, ct_m (dim,nnz)|None."""
        dim, nnz, nn = self.dim, self.nnz, self.n_nodes
        out = out or {}
        big_m = out.get("big_m") if out.get("big_m") is not None else np.empty((dim, nnz))
        rhs = out.get("rhs") if out.get("rhs") is not None else np.empty((nn, dim))
        ml = ct = None
        if opts.assemble_inverse_masslump:
            ml = out.get("masslump") if out.get("masslump") is not None else np.empty((nn, dim))
        if opts.assemble_ct_matrix_here:
            ct = out.get("ct_m") if out.get("ct_m") is not None else np.empty((dim, nnz))
        _check(self.lib.cgasm_momentum(C.c_int(self.id), C.byref(opts), _dp(big_m), _dp(rhs), _dp(ml), _dp(ct)))
        return dict(big_m=big_m, rhs=rhs, masslump=ml, ct_m=ct)

    def advdiff(self, opts, out=None):
        out = out or {}
        val = out.get("matrix") if out.get("matrix") is not None else np.empty(self.nnz)
        rhs = out.get("rhs") if out.get("rhs") is not None else np.empty(self.n_nodes)
        _check(self.lib.cgasm_advdiff(C.c_int(self.id), C.byref(opts), _dp(val), _dp(rhs)))
        return dict(matrix=val, rhs=rhs)

    def momentum_dev(self, opts):
        _check(self.lib.cgasm_momentum_dev(C.c_int(self.id), C.byref(opts)))

    def advdiff_dev(self, opts):
        _check(self.lib.cgasm_advdiff_dev(C.c_int(self.id), C.byref(opts)))

    def momentum_advdiff_dev(self, mopts, aopts):
        """Both element loops in one call (one fused kernel for the common STRIP option sets)."""
        _check(self.lib.cgasm_momentum_advdiff_dev(C.c_int(self.id), C.byref(mopts), C.byref(aopts)))

    def momentum_fetch(self, want_masslump=True, want_ct=False):
        dim, nnz, nn = self.dim, self.nnz, self.n_nodes
        big_m = np.empty((dim, nnz))
        rhs = np.empty((nn, dim))
        ml = np.empty((nn, dim)) if want_masslump else None
        ct = np.empty((dim, nnz)) if want_ct else None
        _check(self.lib.cgasm_momentum_fetch(C.c_int(self.id), _dp(big_m), _dp(rhs), _dp(ml), _dp(ct)))
        return dict(big_m=big_m, rhs=rhs, masslump=ml, ct_m=ct)

    def momentum_mass_fetch(self):
        """The `mass` matrix (opts.assemble_mass_matrix): (dim, nnz) diagonal blocks."""
        mass = np.empty((self.dim, self.nnz))
        _check(self.lib.cgasm_momentum_mass_fetch(C.c_int(self.id), _dp(mass)))
        return mass

    def momentum_identical_blocks(self):
        f = C.c_int(0)
        _check(self.lib.cgasm_momentum_identical_blocks(C.c_int(self.id), C.byref(f)))
        return bool(f.value)

    def momentum_fetch_blocks(self, first, nblocks, out):
        _check(self.lib.cgasm_momentum_fetch_blocks(C.c_int(self.id), C.c_int(first), C.c_int(nblocks), _dp(out)))
        return out

    def momentum_host(self, opts, out):
        """Host-buffer momentum call that moves only what is distinct: when the dim diagonal blocks are
        identical (cgasm_momentum_identical_blocks) one block crosses PCIe and out['big_m'] has shape
        (1, nnz); the caller inserts it dim times."""
        self.momentum_dev(opts)
        nb = 1 if self.momentum_identical_blocks() else self.dim
        self.momentum_fetch_blocks(0, nb, out["big_m"])
        ml = out.get("masslump") if opts.assemble_inverse_masslump else None
        _check(self.lib.cgasm_momentum_fetch(C.c_int(self.id), None, _dp(out["rhs"]), _dp(ml), None))
        return nb

    def advdiff_fetch_into(self, out):
        """cgasm_advdiff_fetch into caller-owned (pinned) buffers out['matrix'], out['rhs']."""
        _check(self.lib.cgasm_advdiff_fetch(C.c_int(self.id), _dp(out["matrix"]), _dp(out["rhs"])))
        return out

    def advdiff_fetch(self):
        val = np.empty(self.nnz)
        rhs = np.empty(self.n_nodes)
        _check(self.lib.cgasm_advdiff_fetch(C.c_int(self.id), _dp(val), _dp(rhs)))
        return dict(matrix=val, rhs=rhs)

    # -- surface loops and strong Dirichlet conditions (add to the device-resident result) ----
    def set_surface(self, sndgln, face_ele, face_tables):
        """sndgln (n_faces, sloc) 1-based, face_ele (n_faces,) 1-based, face_tables = (n_f, dn_f, weight_f)."""
        n, dn, w = face_tables
        sn = np.ascontiguousarray(sndgln, dtype=np.int32)
        fe = np.ascontiguousarray(face_ele, dtype=np.int32)
        self.n_faces = len(fe)
        sloc = sn.shape[1] if sn.ndim == 2 else self.dim
        _check(self.lib.cgasm_set_surface(C.c_int(self.id), C.c_int(len(fe)), C.c_int(sloc), C.c_int(len(w)), _ip(sn), _ip(fe),
                                          _dp(np.ascontiguousarray(n)), _dp(np.ascontiguousarray(dn)),
                                          _dp(np.ascontiguousarray(w))))

    def advdiff_surface_dev(self, opts, bc_type, t_bc=None, t_bc_2=None):
        """bc_type (n_faces,), t_bc / t_bc_2 (n_faces, sloc) or None."""
        bt = np.ascontiguousarray(bc_type, dtype=np.int32)
        b1 = np.ascontiguousarray(t_bc, dtype=np.float64) if t_bc is not None else None
        b2 = np.ascontiguousarray(t_bc_2, dtype=np.float64) if t_bc_2 is not None else None
        _check(self.lib.cgasm_advdiff_surface_dev(C.c_int(self.id), C.byref(opts), _ip(bt), _dp(b1), _dp(b2)))

    def advdiff_dirichlet_dev(self, nodes, values, dt=None):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(self.lib.cgasm_advdiff_dirichlet_dev(C.c_int(self.id), C.c_int(len(nd)), _ip(nd), _dp(v),
                                                    C.c_int(0 if dt is None else 1), C.c_double(dt or 0.0)))

    def momentum_surface_dev(self, opts, velocity_bc_type, velocity_bc=None, pressure_bc_type=None):
        """velocity_bc_type (n_faces, dim), velocity_bc (n_faces, sloc, dim) or None, pressure_bc_type (n_faces,) or None."""
        bt = np.ascontiguousarray(velocity_bc_type, dtype=np.int32)
        bv = np.ascontiguousarray(velocity_bc, dtype=np.float64) if velocity_bc is not None else None
        pt = np.ascontiguousarray(pressure_bc_type, dtype=np.int32) if pressure_bc_type is not None else None
        _check(self.lib.cgasm_momentum_surface_dev(C.c_int(self.id), C.byref(opts), _ip(bt), _dp(bv), _ip(pt)))

    def momentum_dirichlet_dev(self, nodes, comps, values):
        """Strong Dirichlet conditions on the resident big_m / rhs: nodes, comps 1-based, values as
        collect_vector_dirichlet_conditions writes them into rhs."""
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        cp = np.ascontiguousarray(comps, dtype=np.int32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(self.lib.cgasm_momentum_dirichlet_dev(C.c_int(self.id), C.c_int(len(nd)), _ip(nd), _ip(cp), _dp(v)))

    def correct_masslumped_velocity(self, inverse_masslump, delta_p, u, ct_m=None):
        """u (n_nodes, dim) corrected in place (returns it)."""
        iml = np.ascontiguousarray(inverse_masslump, dtype=np.float64)
        dp = np.ascontiguousarray(delta_p, dtype=np.float64)
        assert u.flags.c_contiguous and u.dtype == np.float64
        ct = np.ascontiguousarray(ct_m, dtype=np.float64) if ct_m is not None else None
        _check(self.lib.cgasm_correct_masslumped_velocity(C.c_int(self.id), _dp(ct), _dp(iml), _dp(dp), _dp(u)))
        return u

    # -- lumped-mass pressure matrix C M_L^-1 C^T --------------------------------------------
    def cmc_build_sparsity(self):
        n = C.c_longlong(0)
        _check(self.lib.cgasm_cmc_build_sparsity(C.c_int(self.id), C.byref(n)))
        self.nnz2 = n.value
        return self.nnz2

    def cmc_get_sparsity(self):
        f = np.zeros(self.n_nodes + 1, dtype=np.int32)
        c = np.zeros(self.nnz2, dtype=np.int32)
        _check(self.lib.cgasm_cmc_get_sparsity(C.c_int(self.id), _ip(f), _ip(c)))
        return f, c

    def cmc_set_sparsity(self, findrm2, colm2):
        f = np.ascontiguousarray(findrm2, dtype=np.int32)
        c = np.ascontiguousarray(colm2, dtype=np.int32)
        _check(self.lib.cgasm_cmc_set_sparsity(C.c_int(self.id), C.c_int(len(f) - 1), C.c_int(len(c)), _ip(f), _ip(c)))
        self.nnz2 = len(c)

    def cmc_dev(self, ct_m=None, inverse_masslump=None):
        ct = np.ascontiguousarray(ct_m, dtype=np.float64) if ct_m is not None else None
        iv = np.ascontiguousarray(inverse_masslump, dtype=np.float64) if inverse_masslump is not None else None
        _check(self.lib.cgasm_cmc_dev(C.c_int(self.id), _dp(ct), _dp(iv)))

    def cmc_fetch(self):
        out = np.empty(self.nnz2)
        _check(self.lib.cgasm_cmc_fetch(C.c_int(self.id), _dp(out)))
        return out

    def kmk(self, theta_pg=1.0, want_parts=False):
        """assemble_kmk_matrix on the device: kmk (nnz2), optionally kt (nnz) and the lumped pressure mass (n_nodes)."""
        _check(self.lib.cgasm_kmk_dev(C.c_int(self.id), C.c_double(theta_pg)))
        out = np.empty(self.nnz2)
        kt = np.empty(self.nnz) if want_parts else None
        ml = np.empty(self.n_nodes) if want_parts else None
        _check(self.lib.cgasm_kmk_fetch(C.c_int(self.id), _dp(out), _dp(kt), _dp(ml)))
        return (out, kt, ml) if want_parts else out

    def momentum_result_dev(self):
        p = [C.c_void_p() for _ in range(4)]
        _check(self.lib.cgasm_momentum_result_dev(C.c_int(self.id), *[C.byref(x) for x in p]))
        return [x.value for x in p]

    def advdiff_result_dev(self):
        p = [C.c_void_p() for _ in range(2)]
        _check(self.lib.cgasm_advdiff_result_dev(C.c_int(self.id), *[C.byref(x) for x in p]))
        return [x.value for x in p]

    def momentum_element(self, opts, ele):
        dim, loc = self.dim, self.loc
        T = np.zeros(dim * dim * loc * loc)
        r = np.zeros(dim * loc)
        ml = np.zeros(dim * loc)
        gp = np.zeros(dim * loc * loc)
        _check(self.lib.cgasm_momentum_element(C.c_int(self.id), C.byref(opts), C.c_int(ele), _dp(T),
                                               _dp(r), _dp(ml), _dp(gp)))
        return (T.reshape(loc, loc, dim, dim).transpose(3, 2, 1, 0).copy(), r.reshape(loc, dim).T.copy(),
                ml.reshape(loc, dim).T.copy(), gp.reshape(loc, loc, dim).transpose(2, 1, 0).copy())

    def advdiff_element(self, opts, ele):
        loc = self.loc
        A = np.zeros(loc * loc)
        r = np.zeros(loc)
        _check(self.lib.cgasm_advdiff_element(C.c_int(self.id), C.byref(opts), C.c_int(ele), _dp(A), _dp(r)))
        return A.reshape(loc, loc).T.copy(), r

    # -- plumbing -------------------------------------------------------------------------
    def set_async(self, on=True):
        """cgasm_set_async: uploads and result downloads are queued, cgasm_synchronize waits."""
        _check(self.lib.cgasm_set_async(C.c_int(self.id), C.c_int(1 if on else 0)))

    def synchronize(self):
        _check(self.lib.cgasm_synchronize(C.c_int(self.id)))

    def stream(self):
        s = C.c_void_p()
        _check(self.lib.cgasm_stream(C.c_int(self.id), C.byref(s)))
        return s.value

    def launch_count(self):
        n = C.c_longlong(0)
        _check(self.lib.cgasm_launch_count(C.c_int(self.id), C.byref(n)))
        return n.value

    def last_kernel_ms(self):
        ms = C.c_float(0)
        _check(self.lib.cgasm_last_kernel_ms(C.c_int(self.id), C.byref(ms)))
        return ms.value

    # -- halo ---------------------------------------------------------------------------
    def halo_create(self, nprocs, rank, sends, recvs, unique_id):
        """sends/recvs: list (len nprocs) of 1-based node arrays (empty for rank itself)."""
        nsend = np.array([len(s) for s in sends], dtype=np.int32)
        nrecv = np.array([len(r) for r in recvs], dtype=np.int32)
        s_all = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in sends])
                                     if nsend.sum() else np.zeros(0, dtype=np.int32))
        r_all = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in recvs])
                                     if nrecv.sum() else np.zeros(0, dtype=np.int32))
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self.lib.cgasm_halo_create(C.c_int(self.id), C.c_int(nprocs), C.c_int(rank), _ip(nsend),
                                          _ip(s_all), _ip(nrecv), _ip(r_all), uid))

    PATHS = {0: "none", 1: "element", 2: "tiled", 3: "gather_staged", 4: "gather_rows", 5: "strip", 6: "strip_staged"}

    def last_path(self):
        """(momentum, tracer) kernel family of the last assemblies (cgasm_last_path)."""
        m, a = C.c_int(0), C.c_int(0)
        _check(self.lib.cgasm_last_path(C.c_int(self.id), C.byref(m), C.byref(a)))
        return self.PATHS[m.value], self.PATHS[a.value]

    def plan_stats(self):
        """Shape of the row-block plan (cgasm_plan_stats)."""
        st = (C.c_double * 8)()
        _check(self.lib.cgasm_plan_stats(C.c_int(self.id), st))
        keys = ("row_blocks", "rows_per_block", "longest_row", "strip_entries_per_pair", "staged", "max_nodes_per_block",
                "staged_node_capacity", "momentum_smem_bytes_per_block")
        return dict(zip(keys, [float(v) for v in st]))

    # -- device hand-off to PETSc (COO triplets in universal numbering) --------------------------------
    def coo_pattern(self, which, row_gnn2unn, col_gnn2unn=None, compact=False):
        """which: 0 momentum (dim diagonal blocks), 1 tracer. gnn2unn: (n_nodes, nfields) int array as
        petsc_numbering%gnn2unn (0-based, -1 = masked). Returns ncoo; the device pointers stay with the handle."""
        r = np.asfortranarray(np.asarray(row_gnn2unn, dtype=np.int32))
        c = np.asfortranarray(np.asarray(col_gnn2unn, dtype=np.int32)) if col_gnn2unn is not None else None
        n = C.c_longlong(0)
        pi, pj = C.c_void_p(), C.c_void_p()
        _check(self.lib.cgasm_coo_pattern_dev(C.c_int(self.id), C.c_int(which), r.ctypes.data_as(c_ip),
                                              c.ctypes.data_as(c_ip) if c is not None else None, C.c_int(1 if compact else 0),
                                              C.byref(n), C.byref(pi), C.byref(pj)))
        return n.value

    def coo_fetch(self, which, ncoo):
        i = np.empty(ncoo, dtype=np.int32)
        j = np.empty(ncoo, dtype=np.int32)
        v = np.empty(ncoo)
        _check(self.lib.cgasm_coo_fetch(C.c_int(self.id), C.c_int(which), C.c_longlong(ncoo), _ip(i), _ip(j), _dp(v)))
        return i, j, v

    def halo_set_overlap(self, on):
        _check(self.lib.cgasm_halo_set_overlap(C.c_int(self.id), C.c_int(1 if on else 0)))

    def halo_update(self, slots):
        mask = 0
        for s in slots:
            mask |= 1 << s
        _check(self.lib.cgasm_halo_update(C.c_int(self.id), C.c_uint(mask)))
