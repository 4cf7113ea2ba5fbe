"""Synthetic meshes and fields for parity tests and bench.py.

The reference reads its meshes from gmsh/triangle files (femtools/Read_GMSH.F90); the
example meshes are not checked in and gmsh is absent, so the example configs are
represented by their option sets on these deterministic box meshes.

Layouts are the reference's (femtools/Fields_Data_Types.F90:59-233), expressed in numpy
C order so that the raw buffer is the Fortran column-major array:
  ndglno  (n_elements, loc) int32, 1-based      == mesh%ndglno(loc*n_elements)
  X       (n_nodes, dim) float64                == Coordinate%val(dim, n_nodes)
  vector  (n_nodes, dim)                        == val(dim, n_nodes)
  tensor  (n_nodes, dim, dim), [node, b, a]=T(a,b) == val(dim, dim, n_nodes)
A CONSTANT field has one node (preprocessor/Populate_State.F90:1678-1724).
"""
from dataclasses import dataclass, field as _dc_field
import itertools
import numpy as np

from . import _abi as abi


@dataclass
class Mesh:
    dim: int
    ndglno: np.ndarray  # (n_elements, loc) int32, 1-based
    X: np.ndarray       # (n_nodes, dim)
    shape: tuple = ()   # cubes per axis, for box meshes

    @property
    def loc(self):
        return self.dim + 1

    @property
    def n_nodes(self):
        return self.X.shape[0]

    @property
    def n_elements(self):
        return self.ndglno.shape[0]


def _perm_sign(p):
    s = 1
    p = list(p)
    for i in range(len(p)):
        for j in range(i + 1, len(p)):
            if p[i] > p[j]:
                s = -s
    return s


def kuhn_cell_table(dim):
    """Local corner offsets (bit k = +1 along axis k) of the dim! Kuhn simplices of a unit
    cell; every simplex runs corner 0 -> corner 2^dim-1 along one axis permutation. The first
    two vertices of odd permutations are swapped so that det[x_k - x_loc] > 0 for all."""
    tab = []
    for p in itertools.permutations(range(dim)):
        v = [0]
        for a in p:
            v.append(v[-1] | (1 << a))
        if _perm_sign(p) * (-1 if dim == 3 else 1) < 0:
            v[0], v[1] = v[1], v[0]
        tab.append(v)
    return np.array(tab, dtype=np.int64)


def box_mesh(ncells, jitter=0.1, seed=20240601, lengths=None, origin=None, dtype=np.int32):
    """Kuhn box mesh: prod(ncells) cells x dim! simplices, lexicographic node and cell
    numbering (x fastest). Interior nodes are displaced by U(-jitter*h, jitter*h) per axis
    (0.1h keeps every Kuhn simplex positively oriented)."""
    ncells = tuple(int(c) for c in ncells)
    dim = len(ncells)
    assert dim in (2, 3)
    lengths = tuple(lengths) if lengths is not None else (1.0,) * dim
    origin = tuple(origin) if origin is not None else (0.0,) * dim
    npts = tuple(c + 1 for c in ncells)
    # coordinates, x fastest
    axes = [origin[k] + lengths[k] * np.arange(npts[k], dtype=np.float64) / ncells[k] for k in range(dim)]
    grids = np.meshgrid(*axes[::-1], indexing="ij")  # slowest axis first
    X = np.stack([g.ravel() for g in grids[::-1]], axis=1)  # (n_nodes, dim), col k = axis k
    if jitter:
        rng = np.random.default_rng(seed)
        disp = rng.uniform(-jitter, jitter, size=X.shape)
        idx = [np.arange(npts[k]) for k in range(dim)]
        ig = np.meshgrid(*idx[::-1], indexing="ij")
        interior = np.ones(X.shape[0], dtype=bool)
        for k in range(dim):
            ik = ig[::-1][k].ravel()
            interior &= (ik > 0) & (ik < ncells[k])
            disp[:, k] *= lengths[k] / ncells[k]
        X[interior] += disp[interior]
    # connectivity
    strides = [1]
    for k in range(1, dim):
        strides.append(strides[-1] * npts[k - 1])
    cidx = [np.arange(ncells[k], dtype=np.int64) for k in range(dim)]
    cg = np.meshgrid(*cidx[::-1], indexing="ij")
    base = np.zeros(cg[0].size, dtype=np.int64)
    for k in range(dim):
        base += cg[::-1][k].ravel() * strides[k]
    corner_off = np.zeros(1 << dim, dtype=np.int64)
    for c in range(1 << dim):
        corner_off[c] = sum(strides[k] for k in range(dim) if c & (1 << k))
    tab = kuhn_cell_table(dim)            # (dim!, loc) corner ids
    off = corner_off[tab]                 # (dim!, loc) node offsets
    nd = base[:, None, None] + off[None, :, :] + 1
    ndglno = np.ascontiguousarray(nd.reshape(-1, dim + 1).astype(dtype))
    return Mesh(dim=dim, ndglno=ndglno, X=np.ascontiguousarray(X), shape=ncells)


def shuffled(mesh, seed=7):
    """Same mesh with nodes AND elements randomly renumbered (and local node order rotated):
    an 'unstructured' numbering for tests, since real gmsh meshes have no lexicographic order."""
    rng = np.random.default_rng(seed)
    n = mesh.n_nodes
    perm = rng.permutation(n)            # new id of old node i is perm[i]
    X = np.empty_like(mesh.X)
    X[perm] = mesh.X
    nd = perm[mesh.ndglno.astype(np.int64) - 1] + 1
    eperm = rng.permutation(mesh.n_elements)
    nd = nd[eperm]
    # even permutations of local nodes keep orientation: rotate first three
    rot = rng.integers(0, 3, size=nd.shape[0])
    first3 = nd[:, :3].copy()
    for r in range(3):
        sel = rot == r
        nd[sel, :3] = np.roll(first3[sel], r, axis=1)
    return Mesh(dim=mesh.dim, ndglno=np.ascontiguousarray(nd.astype(np.int32)), X=X, shape=())


def delaunay_points(npoints, dim=3, seed=11, graded=True):
    """Points of an unstructured test mesh in the unit box: uniform background plus (graded) a cloud concentrated
    around the centre, like the refinement around the obstacle of examples/flow_past_sphere_Re100. Numbered the way
    a mesh generator leaves them: sorted by a coarse lattice cell (lexicographic), arbitrary inside a cell."""
    rng = np.random.default_rng(seed)
    n_bg = npoints if not graded else (npoints * 3) // 5
    pts = [rng.random((n_bg, dim))]
    if graded:
        c = 0.5 + 0.12 * rng.standard_normal((npoints - n_bg, dim))
        pts.append(np.clip(c, 0.0, 1.0))
    # the corners keep the hull the unit box
    corners = np.array(list(itertools.product((0.0, 1.0), repeat=dim)))
    P = np.concatenate(pts + [corners])
    P = np.unique(P, axis=0)
    cells = max(2, int(round((P.shape[0] / 64.0) ** (1.0 / dim))))
    key = np.zeros(P.shape[0], dtype=np.int64)
    for k in range(dim - 1, -1, -1):
        key = key * cells + np.minimum((P[:, k] * cells).astype(np.int64), cells - 1)
    return np.ascontiguousarray(P[np.argsort(key, kind="stable")])


def delaunay_mesh(npoints, dim=3, seed=11, graded=True, min_quality=0.02):
    """scipy.spatial.Delaunay of `delaunay_points`: an UNSTRUCTURED simplex mesh (node degrees, strip lengths and row
    lengths all vary; ~6.5 tets per point in 3-D). Slivers (|det J| below min_quality x longest edge^dim; a regular
    tetrahedron has 0.71) are dropped, as a mesh generator would -- the Delaunay triangulation of random points has
    many flat ones on the hull; local node order is whatever qhull returns (both orientations occur)."""
    from scipy.spatial import Delaunay
    P = delaunay_points(npoints, dim, seed, graded)
    simp = Delaunay(P).simplices.astype(np.int64)
    V = P[simp]
    det = np.abs(np.linalg.det(V[:, 1:] - V[:, :1]))
    lmax = np.zeros(simp.shape[0])
    for a in range(dim + 1):
        for b in range(a + 1, dim + 1):
            lmax = np.maximum(lmax, np.linalg.norm(V[:, a] - V[:, b], axis=1))
    keep = det > min_quality * lmax ** dim
    nd = simp[keep]
    # drop nodes no kept element uses (none in practice) by compacting the numbering
    used = np.zeros(P.shape[0], dtype=bool)
    used[nd.ravel()] = True
    if not used.all():
        new_id = np.cumsum(used) - 1
        nd = new_id[nd]
        P = P[used]
    return Mesh(dim=dim, ndglno=np.ascontiguousarray((nd + 1).astype(np.int32)), X=np.ascontiguousarray(P), shape=())


def boundary_faces(mesh):
    """Surface mesh of `mesh` the way femtools add_faces leaves it for a mesh without a surface file:
    the facets that belong to exactly one element (femtools/Fields_Allocates.F90:1166-1296), here in
    ascending (element, local facet) order; local facet k is the one opposite local node k
    (femtools/Element_Numbering.F90 boundary numbering of simplices). Returns
    sndgln (n_faces, sloc) int32 1-based global nodes (ascending local-node order) and face_ele
    (n_faces,) int32 1-based owning element."""
    nd = mesh.ndglno.astype(np.int64)
    loc = mesh.loc
    faces = np.stack([np.delete(nd, k, axis=1) for k in range(loc)], axis=1)  # (ne, loc, sloc)
    key = np.sort(faces, axis=2).reshape(-1, loc - 1)
    n = mesh.n_nodes + 1
    if n ** (loc - 1) < 2 ** 62:  # one integer per facet: much faster than a row-wise unique
        k1 = key[:, 0].copy()
        for c in range(1, loc - 1):
            k1 = k1 * n + key[:, c]
        _, inv, cnt = np.unique(k1, return_inverse=True, return_counts=True)
    else:
        _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    on_boundary = (cnt[inv.ravel()] == 1).reshape(nd.shape[0], loc)
    e, k = np.nonzero(on_boundary)
    return np.ascontiguousarray(faces[e, k], dtype=np.int32), (e + 1).astype(np.int32)


@dataclass
class FieldSet:
    """slot -> (values, field_type); values laid out as the module docstring says."""
    data: dict = _dc_field(default_factory=dict)

    def set(self, slot, val, field_type=abi.FIELD_NORMAL):
        self.data[slot] = (np.ascontiguousarray(val, dtype=np.float64), field_type)
        return self

    def get(self, slot):
        return self.data[slot]

    def items(self):
        return self.data.items()


def iso_tensor(dim, value):
    return (value * np.eye(dim))[None, :, :]


def standard_fields(mesh, nodal_viscosity=False):
    """Field set of config S3 / S2: Taylor-Green-like nu/oldu + noise, nodal
    density and buoyancy, constant gravity direction, constant isotropic viscosity and
    diffusivity, nodal tracer. Extra slots (absorption, source, ...) are filled with smooth
    nodal data so option variants can switch them on."""
    dim, n = mesh.dim, mesh.n_nodes
    X = mesh.X
    two_pi = 2.0 * np.pi

    def velocity(seed):
        u = np.zeros((n, dim))
        u[:, 0] = np.sin(two_pi * X[:, 0]) * np.cos(two_pi * X[:, 1])
        u[:, 1] = -np.cos(two_pi * X[:, 0]) * np.sin(two_pi * X[:, 1])
        if dim == 3:
            u[:, 2] = 0.1 * np.sin(two_pi * X[:, 2])
        u += 0.01 * np.random.default_rng(seed).standard_normal((n, dim))
        return u

    fs = FieldSet()
    fs.set(abi.F_NU, velocity(1))
    fs.set(abi.F_OLDU, velocity(2))
    fs.set(abi.F_DENSITY, 1.0 + 0.1 * np.random.default_rng(3).uniform(size=n))
    fs.set(abi.F_BUOYANCY, np.random.default_rng(4).uniform(size=n))
    fs.set(abi.F_HB_DENSITY, 0.5 + 0.25 * X[:, dim - 1])
    g = np.zeros((1, dim))
    g[0, dim - 1] = -1.0
    fs.set(abi.F_GRAVITY, g, abi.FIELD_CONSTANT)
    if nodal_viscosity:
        rng = np.random.default_rng(6)
        A = rng.uniform(-0.2, 0.2, size=(n, dim, dim))
        visc = 1e-3 * (np.eye(dim)[None] + 0.5 * (A + A.transpose(0, 2, 1)))
        fs.set(abi.F_VISCOSITY, visc)
    else:
        fs.set(abi.F_VISCOSITY, iso_tensor(dim, 1e-3), abi.FIELD_CONSTANT)
    absn = np.zeros((n, dim))
    absn[:, 0] = 1.0 + np.sin(two_pi * X[:, 0]) ** 2
    absn[:, 1:] = 0.25 * np.random.default_rng(8).uniform(size=(n, dim - 1))
    fs.set(abi.F_ABSORPTION, absn)
    fs.set(abi.F_SOURCE, 0.3 * velocity(9))
    fs.set(abi.F_T, np.random.default_rng(5).uniform(size=n))
    fs.set(abi.F_T_DIFFUSIVITY, iso_tensor(dim, 1e-3), abi.FIELD_CONSTANT)
    fs.set(abi.F_T_SOURCE, np.cos(two_pi * X[:, 0]) + 0.1 * np.random.default_rng(10).uniform(size=n))
    fs.set(abi.F_T_ABSORPTION, 0.5 + 0.5 * np.random.default_rng(11).uniform(size=n))
    return fs


def aniso_tensor(dim):
    """Constant anisotropic_symmetric viscosity as in examples/flow_past_sphere_Re100."""
    if dim == 3:
        T = np.array([[1.0e-2, 2.0e-3, 0.0], [2.0e-3, 5.0e-3, 1.0e-3], [0.0, 1.0e-3, 2.0e-2]])
    else:
        T = np.array([[1.0e-2, 2.0e-3], [2.0e-3, 5.0e-3]])
    return T[None, :, :]


def algorithmic_bytes(dim, n_nodes, n_elements, nnz, what="both"):
    """Compulsory HBM traffic per element of the unfused momentum + tracer passes, SURVEY.md
    8(d): every input read once, every output written once (int32 = 4 B, FP64 = 8 B)."""
    loc = dim + 1
    r = n_nodes / n_elements
    z = nnz / n_elements
    mom = 4 * loc + 3 * 8 * dim * r + 2 * 8 * r + 8 * dim * z + 2 * 8 * dim * r
    tra = 4 * loc + 2 * 8 * dim * r + 8 * r + 8 * z + 8 * r
    return {"momentum": mom, "tracer": tra, "both": mom + tra}[what]
