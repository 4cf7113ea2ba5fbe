"""Node-based domain decomposition in the reference's conventions (SURVEY.md sections 5, 8(e)).

Fluidity decomposes by NODES (fldecomp / flredecomp, external METIS / Zoltan -- absent
here): every rank owns a set of nodes and additionally holds
  level-1 halo nodes  = nodes of elements that contain an owned node, and
  level-2 halo nodes  = nodes of elements that contain an owned or level-1 node,
numbered "trailing receives" (femtools/Halo_Data_Types.F90 HALO_ORDER_TRAILING_RECEIVES:
owned nodes first, then the level-1 receives, then the level-2-only receives; see the real
fixture tests/golden/prectangle_halos.json). The local mesh is every element with an owned or
a level-1 node (fldecomp/fldgmsh.cpp:375-455); each rank assembles ALL of them (no ownership test, SURVEY.md section
0 fact 5) and only the rows of owned nodes are used. halo_update uses the largest (level-2)
halo: sends(p) = owned nodes that rank p receives, in the order of p's receives.

Producers:
  partition_by_owner  any mesh + a node->rank map (numpy; test sizes)
  rcb_owner           a node->rank map by recursive coordinate bisection (the geometric stand-in
                      for Zoltan/METIS; any nprocs, balanced to within one node)
  slab_partition      box meshes cut into slabs along the last axis, generated directly per
                      rank without ever building the global mesh (bench sizes). Same node, element and
                      halo SETS as partition_by_owner (= fldecomp's writer) for the slab owner map
                      (tests/test_partition.py), in its own order: cells lexicographic, receive lists layer
                      by layer (nearest layer first) instead of ascending global id
"""
from dataclasses import dataclass
import numpy as np

from .synthetic import Mesh, box_mesh


@dataclass
class LocalPart:
    mesh: Mesh                 # local mesh, local 1-based numbering
    n_owned: int               # n_private_nodes: owned nodes are local ids 1..n_owned
    global_node: np.ndarray    # (n_local_nodes,) 0-based global node id of each local node
    global_element: np.ndarray  # (n_local_elements,) 0-based global element id
    sends: list                # per process: 1-based LOCAL node ids (level-2 halo)
    recvs: list                # per process: 1-based LOCAL node ids (level-2 halo)
    n_l1: int = 0              # number of level-1 receive nodes (they precede the L2-only ones)
    sndgln: np.ndarray = None  # (n_local_faces, sloc) 1-based LOCAL nodes of the boundary faces held locally
    boundary_ids: np.ndarray = None
    global_face: np.ndarray = None  # 0-based global surface-element id of each local face


def _grow(ndglno0, mask):
    """nodes of all elements touching a node in mask"""
    touched = mask[ndglno0].any(axis=1)
    out = np.zeros_like(mask)
    out[ndglno0[touched].ravel()] = True
    return out


def partition_by_owner(mesh, owner, nprocs, sndgln=None, boundary_ids=None):
    """Returns [LocalPart] for every rank from a node -> rank map, in the conventions of the reference's own
    decomposition writer (fldecomp/fldgmsh.cpp write_partitions_gmsh :310-626; tests/test_formats.py compares
    the two bit for bit through oracle/_ref):
      nodes     owned (ascending global id), then the level-1 halo nodes (ascending), then the level-2-only ones;
      elements  those with an owned node -- first the ones whose lowest-ranked node owner is this rank, then the
                others, each ascending (:375-425) -- then the level-2 elements: no owned node but a level-1 node
                (:431-455);
      halos     receives from p = the halo nodes p owns in ascending GLOBAL id; sends to p = p's receives from us in
                the same order (:589-611); the level-2 lists contain the level-1 ones (:458-459);
      faces     (if sndgln is given) the global surface elements all of whose nodes lie in one local element, in
                global order (:518-556).
    owner: (n_nodes,) rank of each global node; sndgln (n_faces, sloc) 1-based, boundary_ids (n_faces,)."""
    nd0 = mesh.ndglno.astype(np.int64) - 1
    owner = np.asarray(owner)
    eown = owner[nd0]                                   # (n_elements, loc)
    emin = eown.min(axis=1)
    parts = []
    local_of = []  # per rank: global -> local (0-based) map, -1 if absent
    l2_sets = []
    for r in range(nprocs):
        own = owner == r
        n_own_in_e = (eown == r).sum(axis=1)
        e_owned = n_own_in_e > 0
        l1 = np.zeros(mesh.n_nodes, dtype=bool)
        l1[nd0[e_owned].ravel()] = True
        l1 &= ~own
        e_halo2 = (~e_owned) & l1[nd0].any(axis=1)
        l2 = np.zeros(mesh.n_nodes, dtype=bool)
        l2[nd0[e_halo2].ravel()] = True
        l2 &= ~l1                                       # (no owned node in these elements)
        g_own, g_l1, g_l2 = np.flatnonzero(own), np.flatnonzero(l1), np.flatnonzero(l2)
        gl = np.concatenate([g_own, g_l1, g_l2])
        g2l = -np.ones(mesh.n_nodes, dtype=np.int64)
        g2l[gl] = np.arange(len(gl))
        eo = np.flatnonzero(e_owned)
        ge = np.concatenate([eo[emin[eo] == r], eo[emin[eo] != r], np.flatnonzero(e_halo2)])
        lnd = (g2l[nd0[ge]] + 1).astype(np.int32)
        lm = Mesh(dim=mesh.dim, ndglno=np.ascontiguousarray(lnd), X=np.ascontiguousarray(mesh.X[gl]))
        halo_g = np.flatnonzero(l1 | l2)                 # ascending global id, level 1 and 2 together
        recvs = [(g2l[halo_g[owner[halo_g] == p]] + 1).astype(np.int32) for p in range(nprocs)]
        lp = LocalPart(mesh=lm, n_owned=len(g_own), global_node=gl, global_element=ge,
                       sends=[None] * nprocs, recvs=recvs, n_l1=len(g_l1))
        if sndgln is not None:
            lp.sndgln, lp.boundary_ids, lp.global_face = _local_faces(lm, g2l, np.asarray(sndgln), boundary_ids)
        parts.append(lp)
        local_of.append(g2l)
    # sends(p) on rank r = what p receives from r, in p's receive order, in r's numbering
    for r in range(nprocs):
        for p in range(nprocs):
            gp = parts[p].global_node[parts[p].recvs[r] - 1] if len(parts[p].recvs[r]) else np.zeros(0, dtype=np.int64)
            parts[r].sends[p] = (local_of[r][gp] + 1).astype(np.int32)
    return parts


def _local_faces(lmesh, g2l, sndgln, boundary_ids):
    """Global surface elements whose nodes all belong to one local volume element (fldgmsh.cpp:518-556)."""
    lf = g2l[sndgln.astype(np.int64) - 1]               # (n_faces, sloc) local 0-based, -1 = absent
    cand = np.flatnonzero((lf >= 0).all(axis=1))
    loc = lmesh.loc
    nd = lmesh.ndglno.astype(np.int64) - 1
    facets = np.stack([np.sort(np.delete(nd, k, axis=1), axis=1) for k in range(loc)], axis=1).reshape(-1, loc - 1)
    n = lmesh.n_nodes + 1

    def key(a):
        k = a[:, 0].copy()
        for c in range(1, a.shape[1]):
            k = k * n + a[:, c]
        return k

    present = np.isin(key(np.sort(lf[cand], axis=1)), key(facets))
    keep = cand[present]
    bid = np.asarray(boundary_ids)[keep].astype(np.int32) if boundary_ids is not None else None
    return (lf[keep] + 1).astype(np.int32), bid, keep


def rcb_owner(X, nprocs):
    """Recursive coordinate bisection: split the node set at the weighted median of its longest
    axis into groups of floor(p/2) and ceil(p/2) ranks, recurse. Deterministic (stable sorts, ties
    broken by node id). Returns (n_nodes,) int64 ranks. Zoltan's RCB makes the same cuts up to its
    tie-breaking; partition QUALITY parity with the reference's graph partitioners is unpinned
    (externals absent), `partition_quality` reports the figures to compare."""
    X = np.asarray(X, dtype=np.float64)
    owner = np.zeros(X.shape[0], dtype=np.int64)
    stack = [(np.arange(X.shape[0], dtype=np.int64), 0, int(nprocs))]
    while stack:
        ids, r0, p = stack.pop()
        if p == 1 or len(ids) == 0:
            owner[ids] = r0
            continue
        ext = X[ids].max(axis=0) - X[ids].min(axis=0)
        axis = int(np.argmax(ext))
        pl = p // 2
        nl = (len(ids) * pl + p // 2) // p  # nodes in proportion to the ranks on each side
        order = ids[np.lexsort((ids, X[ids, axis]))]
        stack.append((np.sort(order[:nl]), r0, pl))
        stack.append((np.sort(order[nl:]), r0 + pl, p - pl))
    return owner


def partition_quality(mesh, parts):
    """What one would compare against a Zoltan/METIS decomposition of the same mesh: owned-node
    balance, redundant (halo) element assembly, halo_update volume."""
    n_owned = np.array([lp.n_owned for lp in parts], dtype=np.float64)
    n_el = np.array([lp.mesh.n_elements for lp in parts], dtype=np.float64)
    sent = np.array([sum(len(s) for s in lp.sends) for lp in parts], dtype=np.float64)
    nbr = np.array([sum(1 for s in lp.sends if len(s)) for lp in parts])
    return dict(nprocs=len(parts), owned_imbalance=float(n_owned.max() / n_owned.mean()),
                element_redundancy=float(n_el.sum() / mesh.n_elements),
                local_elements_max=int(n_el.max()), halo_nodes_sent_max=int(sent.max()),
                halo_nodes_sent_total=int(sent.sum()), neighbours_max=int(nbr.max()))


def _hash_uniform(ids, seed, k):
    """Counter-based U(0,1) from (global id, component k): identical on every rank
    (splitmix64 finaliser)."""
    with np.errstate(over="ignore"):
        x = (ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(seed) * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(k + 1) * np.uint64(0x94D049BB133111EB))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def slab_layers(n_layers_nodes, nprocs):
    """Owned node-layer ranges [L_r, L_{r+1}) along the last axis."""
    base, rem = divmod(n_layers_nodes, nprocs)
    L = [0]
    for r in range(nprocs):
        L.append(L[-1] + base + (1 if r < rem else 0))
    return L


def slab_partition(ncells_global, nprocs, rank, jitter=0.1, seed=20240601):
    """LocalPart of `rank` for a Kuhn box mesh of ncells_global cells cut into nprocs slabs
    of node layers along the last axis. Generated directly (the global mesh is never
    built). Node jitter is a hash of the GLOBAL node id so shared nodes agree across ranks.
    With nprocs == 1 this is the whole box in lexicographic numbering."""
    ncells_global = tuple(int(c) for c in ncells_global)
    dim = len(ncells_global)
    npts = tuple(c + 1 for c in ncells_global)
    L = slab_layers(npts[-1], nprocs)
    lo_own, hi_own = L[rank], L[rank + 1]              # owned node layers [lo, hi)
    assert nprocs == 1 or hi_own - lo_own >= 2, "slabs must be at least two node layers thick"
    lo = max(lo_own - 2, 0)                            # node layers held locally [lo, hi_all)
    hi = min(hi_own + 2, npts[-1])
    ncl = ncells_global[:-1] + (hi - lo - 1,)
    h = [1.0 / ncells_global[k] for k in range(dim)]
    # the unit box is [0,1]^dim for the GLOBAL mesh (weak scaling keeps the cell count per rank)
    origin = [0.0] * (dim - 1) + [lo * h[-1]]
    lengths = [1.0] * (dim - 1) + [(hi - lo - 1) * h[-1]]
    m = box_mesh(ncl, jitter=0.0, lengths=lengths, origin=origin)
    plane = int(np.prod(npts[:-1]))
    n_local = m.n_nodes
    lidx = np.arange(n_local, dtype=np.int64)
    gid = lidx + lo * plane                            # lexicographic: last axis slowest
    # last-axis coordinate from the GLOBAL layer index so every rank computes identical bits
    m.X[:, dim - 1] = (lidx // plane + lo).astype(np.float64) / ncells_global[-1]
    if jitter:
        layer = lidx // plane + lo
        inplane = lidx % plane
        interior = (layer > 0) & (layer < npts[-1] - 1)
        rem = inplane
        for k in range(dim - 1):
            ik = rem % npts[k]
            rem = rem // npts[k]
            interior &= (ik > 0) & (ik < npts[k] - 1)
        for k in range(dim):
            d = (2.0 * _hash_uniform(gid, seed, k) - 1.0) * jitter * h[k]
            m.X[:, k] += np.where(interior, d, 0.0)
    layer = lidx // plane + lo
    own = (layer >= lo_own) & (layer < hi_own)
    below1 = layer == lo_own - 1
    above1 = layer == hi_own
    below2 = layer == lo_own - 2
    above2 = layer == hi_own + 1
    order = np.concatenate([np.flatnonzero(own), np.flatnonzero(below1), np.flatnonzero(above1),
                            np.flatnonzero(below2), np.flatnonzero(above2)])
    assert len(order) == n_local
    old2new = np.empty(n_local, dtype=np.int64)
    old2new[order] = np.arange(n_local)
    nd = (old2new[m.ndglno.astype(np.int64) - 1] + 1).astype(np.int32)
    X = m.X[order]
    lm = Mesh(dim=dim, ndglno=np.ascontiguousarray(nd), X=np.ascontiguousarray(X), shape=ncl)
    n_owned = int(own.sum())
    n_l1 = int(below1.sum() + above1.sum())
    sends = [np.zeros(0, dtype=np.int32) for _ in range(nprocs)]
    recvs = [np.zeros(0, dtype=np.int32) for _ in range(nprocs)]

    def loc_ids(mask):
        return (old2new[np.flatnonzero(mask)] + 1).astype(np.int32)

    if rank > 0:
        # receive from rank-1: its top two owned layers (our below1 then below2); send it our
        # bottom two owned layers (its above1 = our layer lo_own, its above2 = lo_own+1)
        recvs[rank - 1] = np.concatenate([loc_ids(below1), loc_ids(below2)])
        sends[rank - 1] = np.concatenate([loc_ids(layer == lo_own), loc_ids(layer == lo_own + 1)])
    if rank < nprocs - 1:
        recvs[rank + 1] = np.concatenate([loc_ids(above1), loc_ids(above2)])
        sends[rank + 1] = np.concatenate([loc_ids(layer == hi_own - 1), loc_ids(layer == hi_own - 2)])
    cells_plane = int(np.prod(ncells_global[:-1]))
    nsimp = m.n_elements // int(np.prod(ncl))
    ge = np.arange(m.n_elements, dtype=np.int64) + lo * cells_plane * nsimp
    return LocalPart(mesh=lm, n_owned=n_owned, global_node=gid[order], global_element=ge,
                     sends=sends, recvs=recvs, n_l1=n_l1)


def global_nodal_fields(dim, X, gid, seeds=(1, 2, 3, 4, 5)):
    """The S3 field set evaluated from coordinates + GLOBAL node ids (hash noise), so every rank
    computes identical values on shared nodes. Returns dict name -> array."""
    two_pi = 2.0 * np.pi
    n = X.shape[0]

    def noise(seed, k):
        # Box-Muller from two hashed uniforms: N(0,1)
        u1 = np.maximum(_hash_uniform(gid, seed, 2 * k), 1e-300)
        u2 = _hash_uniform(gid, seed, 2 * k + 1)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(two_pi * u2)

    def velocity(seed):
        u = np.zeros((n, dim))
        u[:, 0] = np.sin(two_pi * X[:, 0]) * np.cos(two_pi * X[:, 1])
        u[:, 1] = -np.cos(two_pi * X[:, 0]) * np.sin(two_pi * X[:, 1])
        if dim == 3:
            u[:, 2] = 0.1 * np.sin(two_pi * X[:, 2])
        for k in range(dim):
            u[:, k] += 0.01 * noise(seed, k)
        return u

    return dict(nu=velocity(seeds[0]), oldu=velocity(seeds[1]),
                density=1.0 + 0.1 * _hash_uniform(gid, seeds[2], 0),
                buoyancy=_hash_uniform(gid, seeds[3], 0), t=_hash_uniform(gid, seeds[4], 0))


def block_grid(nprocs):
    """Process grid (px, py, pz) for a box: powers of two are split along the last axes first
    (1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2, 8 -> 2x2x2); other counts are factorised greedily the same way."""
    p = [1, 1, 1]
    n, axis, f = int(nprocs), 2, 2
    while n > 1:
        while n % f:
            f += 1
        p[axis] *= f
        n //= f
        axis = (axis - 1) % 3
    return tuple(p)


def block_partition(ncells_global, pgrid, rank, jitter=0.1, seed=20240601):
    """LocalPart of `rank` for ONE Kuhn box mesh of ncells_global cells decomposed into pgrid = (px, py, pz) blocks
    of nodes (strong scaling: the mesh is fixed, the blocks shrink). Generated from the block's own neighbourhood --
    the sub-box of node layers within two cells of the owned block -- without ever building the global mesh; on that
    sub-box the sets are made exactly as partition_by_owner makes them on the global mesh (= the reference's
    fldecomp writer, fldecomp/fldgmsh.cpp:310-626): owned nodes, level-1 halo (nodes of elements with an owned node),
    level-2 halo, the elements with an owned or level-1 node, receive lists in ascending global id, send lists in the
    receiver's order. tests/test_partition.py compares the two on small boxes, every rank, every list.
    Ranks are numbered x fastest: rank = rx + px * (ry + py * rz)."""
    ncells_global = tuple(int(c) for c in ncells_global)
    dim = len(ncells_global)
    pgrid = tuple(int(p) for p in pgrid)
    assert len(pgrid) == dim
    nprocs = int(np.prod(pgrid))
    npts = tuple(c + 1 for c in ncells_global)
    L = [slab_layers(npts[k], pgrid[k]) for k in range(dim)]
    rc, rem = [], rank
    for k in range(dim):
        rc.append(rem % pgrid[k])
        rem //= pgrid[k]
    lo_own = [L[k][rc[k]] for k in range(dim)]
    hi_own = [L[k][rc[k] + 1] for k in range(dim)]
    lo = [max(lo_own[k] - 2, 0) for k in range(dim)]
    hi = [min(hi_own[k] + 2, npts[k]) for k in range(dim)]
    ncl = tuple(hi[k] - lo[k] - 1 for k in range(dim))
    m = box_mesh(ncl, jitter=0.0)
    lpts = tuple(c + 1 for c in ncl)
    n_sub = m.n_nodes
    lidx = np.arange(n_sub, dtype=np.int64)
    gi, r_ = [], lidx
    for k in range(dim):
        gi.append(r_ % lpts[k] + lo[k])
        r_ = r_ // lpts[k]
    gid = np.zeros(n_sub, dtype=np.int64)
    for k in reversed(range(dim)):
        gid = gid * npts[k] + gi[k]
    X = np.empty((n_sub, dim))
    interior = np.ones(n_sub, dtype=bool)
    for k in range(dim):
        X[:, k] = gi[k].astype(np.float64) / ncells_global[k]   # from the GLOBAL index: identical bits on every rank
        interior &= (gi[k] > 0) & (gi[k] < npts[k] - 1)
    if jitter:
        for k in range(dim):
            d = (2.0 * _hash_uniform(gid, seed, k) - 1.0) * jitter * (1.0 / ncells_global[k])   # as slab_partition: same bits
            X[:, k] += np.where(interior, d, 0.0)
    # owner of every sub-box node
    owner = np.zeros(n_sub, dtype=np.int64)
    for k in reversed(range(dim)):
        bk = np.searchsorted(np.asarray(L[k]), gi[k], side="right") - 1
        owner = owner * pgrid[k] + bk
    nd0 = m.ndglno.astype(np.int64) - 1
    if nprocs == 1:
        lm = Mesh(dim=dim, ndglno=m.ndglno, X=np.ascontiguousarray(X), shape=ncl)
        return LocalPart(mesh=lm, n_owned=n_sub, global_node=gid, global_element=np.arange(m.n_elements, dtype=np.int64),
                         sends=[np.zeros(0, dtype=np.int32)], recvs=[np.zeros(0, dtype=np.int32)], n_l1=0)
    eown = owner[nd0]

    def halo_sets(p):
        """(own, l1, l2, e_owned, e_halo2) of rank p, restricted to this sub-box"""
        own = owner == p
        e_owned = (eown == p).any(axis=1)
        l1 = np.zeros(n_sub, dtype=bool)
        l1[nd0[e_owned].ravel()] = True
        l1 &= ~own
        e_halo2 = (~e_owned) & l1[nd0].any(axis=1)
        l2 = np.zeros(n_sub, dtype=bool)
        l2[nd0[e_halo2].ravel()] = True
        l2 &= ~l1
        l2 &= ~own
        return own, l1, l2, e_owned, e_halo2

    own, l1, l2, e_owned, e_halo2 = halo_sets(rank)
    s_own, s_l1, s_l2 = np.flatnonzero(own), np.flatnonzero(l1), np.flatnonzero(l2)
    sl = np.concatenate([s_own, s_l1, s_l2])                # sub-box ids in local order
    s2l = -np.ones(n_sub, dtype=np.int64)
    s2l[sl] = np.arange(len(sl))
    emin = eown.min(axis=1)
    eo = np.flatnonzero(e_owned)
    se = np.concatenate([eo[emin[eo] == rank], eo[emin[eo] != rank], np.flatnonzero(e_halo2)])
    lnd = (s2l[nd0[se]] + 1).astype(np.int32)
    lm = Mesh(dim=dim, ndglno=np.ascontiguousarray(lnd), X=np.ascontiguousarray(X[sl]), shape=ncl)
    # global element ids: cell (global lexicographic) * dim! + simplex
    nsimp = m.n_elements // int(np.prod(ncl))
    cell = se // nsimp
    gc, r_, stride = np.zeros(len(se), dtype=np.int64), cell, 1
    for k in range(dim):
        gc += (r_ % ncl[k] + lo[k]) * stride
        stride *= ncells_global[k]
        r_ = r_ // ncl[k]
    ge = gc * nsimp + se % nsimp
    halo = l1 | l2
    recvs, sends = [], []
    for p in range(nprocs):
        if p == rank:
            recvs.append(np.zeros(0, dtype=np.int32))
            sends.append(np.zeros(0, dtype=np.int32))
            continue
        recvs.append((s2l[np.flatnonzero(halo & (owner == p))] + 1).astype(np.int32))
        if not (owner == p).any():
            sends.append(np.zeros(0, dtype=np.int32))
            continue
        _, pl1, pl2, _, _ = halo_sets(p)
        sends.append((s2l[np.flatnonzero(own & (pl1 | pl2))] + 1).astype(np.int32))
    return LocalPart(mesh=lm, n_owned=len(s_own), global_node=gid[sl], global_element=ge, sends=sends, recvs=recvs,
                     n_l1=len(s_l1))
