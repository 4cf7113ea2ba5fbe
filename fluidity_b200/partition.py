"""Node-based domain decomposition in the reference's conventions (SURVEY.md sections 5, 8(e)).

Fluidity decomposes by NODES (fldecomp / flredecomp, external METIS / Zoltan -- absent
here): every rank owns a set of nodes and additionally holds
  level-1 halo nodes  = nodes of elements that contain an owned node, and
  level-2 halo nodes  = nodes of elements that contain an owned or level-1 node,
numbered "trailing receives" (femtools/Halo_Data_Types.F90 HALO_ORDER_TRAILING_RECEIVES:
owned nodes first, then the level-1 receives, then the level-2-only receives; see the real
fixture tests/golden/prectangle_halos.json). The local mesh is every element whose nodes
all lie in that set; each rank assembles ALL of them (no ownership test, SURVEY.md section
0 fact 5) and only the rows of owned nodes are used. halo_update uses the largest (level-2)
halo: sends(p) = owned nodes that rank p receives, in the order of p's receives.

Producers:
  partition_by_owner  any mesh + a node->rank map (numpy; test sizes)
  rcb_owner           a node->rank map by recursive coordinate bisection (the geometric stand-in
                      for Zoltan/METIS; any nprocs, balanced to within one node)
  slab_partition      box meshes cut into slabs along the last axis, generated directly per
                      rank without ever building the global mesh (bench sizes)
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


def _grow(ndglno0, mask):
    """nodes of all elements touching a node in mask"""
    touched = mask[ndglno0].any(axis=1)
    out = np.zeros_like(mask)
    out[ndglno0[touched].ravel()] = True
    return out


def partition_by_owner(mesh, owner, nprocs):
    """Returns [LocalPart] for every rank. owner: (n_nodes,) rank of each global node."""
    nd0 = mesh.ndglno.astype(np.int64) - 1
    owner = np.asarray(owner)
    parts = []
    local_of = []  # per rank: global -> local (0-based) map, -1 if absent
    for r in range(nprocs):
        own = owner == r
        l1 = _grow(nd0, own) & ~own
        l2 = _grow(nd0, own | l1) & ~own & ~l1
        g_own = np.flatnonzero(own)
        # receives sorted by sending process, then global id (stable, deterministic)
        g_l1 = np.flatnonzero(l1)
        g_l1 = g_l1[np.lexsort((g_l1, owner[g_l1]))]
        g_l2 = np.flatnonzero(l2)
        g_l2 = g_l2[np.lexsort((g_l2, owner[g_l2]))]
        gl = np.concatenate([g_own, g_l1, g_l2])
        g2l = -np.ones(mesh.n_nodes, dtype=np.int64)
        g2l[gl] = np.arange(len(gl))
        present = g2l >= 0
        keep = present[nd0].all(axis=1)
        ge = np.flatnonzero(keep)
        lnd = (g2l[nd0[ge]] + 1).astype(np.int32)
        lm = Mesh(dim=mesh.dim, ndglno=np.ascontiguousarray(lnd), X=np.ascontiguousarray(mesh.X[gl]))
        recvs = []
        halo_g = np.concatenate([g_l1, g_l2])
        for p in range(nprocs):
            sel = halo_g[owner[halo_g] == p]
            recvs.append((g2l[sel] + 1).astype(np.int32))
        parts.append(LocalPart(mesh=lm, n_owned=len(g_own), global_node=gl, global_element=ge,
                               sends=[None] * nprocs, recvs=recvs, n_l1=len(g_l1)))
        local_of.append(g2l)
    # sends(p) on rank r = what p receives from r, in p's receive order, in r's numbering
    for r in range(nprocs):
        for p in range(nprocs):
            gp = parts[p].global_node[parts[p].recvs[r] - 1] if len(parts[p].recvs[r]) else np.zeros(0, dtype=np.int64)
            parts[r].sends[p] = (local_of[r][gp] + 1).astype(np.int32)
    return parts


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
