"""On-disk formats either side of the assembly path.

Host-side tooling only: the Fortran host keeps its own readers. These let the parity tests and
tools consume, without a Fortran toolchain,
  * gmsh .msh meshes, format 2.x ASCII and binary -- the subset femtools/Read_GMSH.F90 reads for
    simplex meshes (read_header :347-433, read_nodes_coords_v2 :586-641,
    read_faces_and_elements_v2 :1204-1316, process_gmsh_elements :1321-1423) -- and write them
    the way femtools/Write_GMSH.F90:165-430 does (version "2.1", faces first with 2 or 4 tags,
    then the volume elements with 2 tags);
  * .halo XML (femtools/Halos_IO.cpp:34-176 ReadHalos, :178-277 WriteHalos), the `_<rank>` file
    naming of decomposed meshes (cHaloReaderSetInput, Halos_IO.cpp:303-311; parallel_filename,
    Read_GMSH.F90:113-118) and the hand-over to `partition.LocalPart`;
  * PETSc binary viewer files as written by `dump_matrix` (femtools/Petsc_Tools.F90:1487-1504
    DumpMatrixEquation: MatView(A), VecView(b), VecView(x0); PETSc's published binary layout:
    big-endian, MAT_FILE_CLASSID 1211216 {M, N, nz, row lengths, 0-based columns, values},
    VEC_FILE_CLASSID 1211214 {n, values}), and the reference's numberings that map a
    block_csr / petsc_csr matrix to PETSc rows (femtools/Petsc_Tools.F90:141-199).
Format 4.x gmsh files are refused (the reference reads them through its entity maps; none of
the in-tree fixtures on this path uses them).
"""
from dataclasses import dataclass, field as _field
import os
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .synthetic import Mesh

# gmsh element types (femtools/GMSH_Common.F90:44-60): nodes per element
GMSH_LINE, GMSH_TRIANGLE, GMSH_QUAD, GMSH_TET, GMSH_HEX, GMSH_NODE = 1, 2, 3, 4, 5, 15
_NUM_NODES = {GMSH_LINE: 2, GMSH_TRIANGLE: 3, GMSH_QUAD: 4, GMSH_TET: 4, GMSH_HEX: 8, GMSH_NODE: 1}


class FormatError(ValueError):
    """Malformed or unsupported file (the reference FLExit()s with the same messages)."""


@dataclass
class GmshMesh:
    """What read_gmsh_simple hands to the rest of femtools (Read_GMSH.F90:270-335)."""
    mesh: Mesh                      # dim, ndglno (1-based), X (n_nodes, dim)
    sndgln: np.ndarray              # (n_faces, sloc) int32, 1-based
    boundary_ids: np.ndarray = None  # (n_faces,) first tag of every face, or None
    element_owner: np.ndarray = None  # (n_faces,) fourth tag (faces with 4 tags), or None
    region_ids: np.ndarray = None   # (n_elements,) first tag of every element, or None
    version: tuple = (2, 1)
    binary: bool = False


def _read_line(f):
    line = f.readline()
    if not line:
        raise FormatError("unexpected end of gmsh file")
    return line.decode("ascii", "replace").strip()


def _expect(f, tag):
    try:
        line = _read_line(f)
        while line == "":
            line = _read_line(f)
    except FormatError:
        line = None
    if line != tag:
        raise FormatError("Error: cannot find '%s' in GMSH mesh file" % tag)


def read_gmsh(path, coordinate_dim=None):
    """Read_GMSH.F90:72-339 for format 2.x simplex meshes. `coordinate_dim` = the optional `mdim`
    argument (:262-271): how many coordinates to keep; default = the topological dimension."""
    with open(path, "rb") as f:
        _expect(f, "$MeshFormat")
        hdr = _read_line(f).split()
        ver = tuple(int(x) for x in hdr[0].split(".")) + (0,)
        major, minor = ver[0], ver[1]
        if major < 2 or major == 3 or (major == 4 and minor > 1) or major > 4:
            raise FormatError("Error: GMSH mesh version must be 2.x or 4.x")  # Read_GMSH.F90:382-387
        if major == 4:
            raise FormatError("gmsh format 4.x is not read by this tool; convert with gmsh -format msh2")
        binary = int(hdr[1]) == 1
        if int(hdr[2]) != 8:
            raise FormatError("Error: GMSH data size does not equal 8")  # :391-393
        endian = "<"
        if binary:
            one = f.read(4)
            if struct.unpack("<i", one)[0] == 1:
                endian = "<"
            elif struct.unpack(">i", one)[0] == 1:
                endian = ">"
            else:
                raise FormatError("Error: GMSH binary endianness marker is not 1")
            f.readline()
        _expect(f, "$EndMeshFormat")

        # nodes: Read_GMSH.F90:586-641
        _expect(f, "$Nodes")
        n_nodes = int(_read_line(f))
        if n_nodes < 2:
            raise FormatError("Error: GMSH number of nodes field < 2")
        if binary:
            rec = np.dtype([("id", endian + "i4"), ("x", endian + "f8", (3,))])
            raw = np.frombuffer(f.read(rec.itemsize * n_nodes), dtype=rec)
            if raw.shape[0] != n_nodes:
                raise FormatError("truncated $Nodes section")
            ids, xyz = raw["id"].astype(np.int64), raw["x"].astype(np.float64)
            f.readline()
        else:
            tab = np.array([_read_line(f).split() for _ in range(n_nodes)])
            ids, xyz = tab[:, 0].astype(np.int64), tab[:, 1:4].astype(np.float64)
        _expect(f, "$EndNodes")

        # elements: Read_GMSH.F90:1204-1316
        _expect(f, "$Elements")
        n_all = int(_read_line(f))
        if n_all < 1:
            raise FormatError("Error: number of elements in GMSH file < 1")
        types, tags, nodes = [], [], []
        if binary:
            e = 0
            while e < n_all:
                gtype, gcount, gtags = struct.unpack(endian + "3i", f.read(12))
                if e + gcount > n_all:
                    raise FormatError("GMSH element group contains more than the total")
                if gtype not in _NUM_NODES:
                    raise FormatError("Unsupported element type in gmsh .msh file")
                w = 1 + gtags + _NUM_NODES[gtype]
                blk = np.frombuffer(f.read(4 * w * gcount), dtype=endian + "i4").reshape(gcount, w)
                for row in blk:
                    types.append(gtype)
                    tags.append(row[1:1 + gtags].astype(np.int64))
                    nodes.append(row[1 + gtags:].astype(np.int64))
                e += gcount
            f.readline()
        else:
            for _ in range(n_all):
                tok = [int(t) for t in _read_line(f).split()]
                gtype, ntags = tok[1], tok[2]
                if gtype not in _NUM_NODES:
                    raise FormatError("Unsupported element type in gmsh .msh file")
                types.append(gtype)
                tags.append(np.array(tok[3:3 + ntags], dtype=np.int64))
                nodes.append(np.array(tok[3 + ntags:3 + ntags + _NUM_NODES[gtype]], dtype=np.int64))
        _expect(f, "$EndElements")

    # process_gmsh_elements (:1321-1423): which type is the element, which the face
    types = np.array(types)
    count = {t: int((types == t).sum()) for t in _NUM_NODES}
    if count[GMSH_QUAD] or count[GMSH_HEX]:
        raise FormatError("quadrilateral/hexahedral meshes are outside the P1 simplex path")
    if count[GMSH_TET]:
        etype, ftype, dim = GMSH_TET, GMSH_TRIANGLE, 3
    elif count[GMSH_TRIANGLE]:
        etype, ftype, dim = GMSH_TRIANGLE, GMSH_LINE, 2
    elif count[GMSH_LINE]:
        etype, ftype, dim = GMSH_LINE, GMSH_NODE, 1
    else:
        raise FormatError("Unsupported mixture of face/element types")
    el = [i for i in range(len(types)) if types[i] == etype]
    fa = [i for i in range(len(types)) if types[i] == ftype]

    def _uniform_tags(idx, what):
        nt = {len(tags[i]) for i in idx}
        if len(nt) > 1:
            raise FormatError("Inconsistent number of %s tags" % what)  # :208-226
        return nt.pop() if nt else 0

    n_ftags, n_etags = _uniform_tags(fa, "face"), _uniform_tags(el, "element")
    # nodes(n)%nodeID indexes the coordinate field (:286-292): ids must be 1..n_nodes
    if ids.min() != 1 or ids.max() != n_nodes or len(np.unique(ids)) != n_nodes:
        raise FormatError("gmsh node ids are not a permutation of 1..numNodes")
    cdim = coordinate_dim or dim
    X = np.zeros((n_nodes, cdim))
    X[ids - 1] = xyz[:, :cdim]
    ndglno = np.array([nodes[i] for i in el], dtype=np.int32).reshape(len(el), _NUM_NODES[etype])
    sndgln = np.array([nodes[i] for i in fa], dtype=np.int32).reshape(len(fa), _NUM_NODES[ftype])
    out = GmshMesh(mesh=Mesh(dim=dim, ndglno=np.ascontiguousarray(ndglno), X=X), sndgln=sndgln,
                   version=(major, minor), binary=binary)
    if n_ftags > 0:
        out.boundary_ids = np.array([tags[i][0] for i in fa], dtype=np.int32)
    if n_ftags == 4:
        out.element_owner = np.array([tags[i][3] for i in fa], dtype=np.int32)
    if n_etags > 0:
        out.region_ids = np.array([tags[i][0] for i in el], dtype=np.int32)
    return out


def write_gmsh(path, gm, binary=False, style="femtools"):
    """style "femtools": Write_GMSH.F90:165-430: "2.1 <0|1> 8", faces first (2 tags, or 4 with element
    owners), then the volume elements with 2 tags (region id, 0). ASCII coordinates use repr() instead of
    the reference's F0.10 so that a round trip is exact.
    style "fldecomp": the files fldecomp writes for every partition (fldecomp/fldgmsh.cpp
    write_part_main_mesh :99-300): binary, ONE tag per face (the boundary id; none if there are no ids)
    and one per element (the region id), a face group header even when there is no face."""
    if style == "fldecomp":
        return _write_gmsh_fldecomp(path, gm)
    m = gm.mesh
    etype = {1: GMSH_LINE, 2: GMSH_TRIANGLE, 3: GMSH_TET}[m.dim]
    ftype = {1: GMSH_NODE, 2: GMSH_LINE, 3: GMSH_TRIANGLE}[m.dim]
    nf, ne = len(gm.sndgln), m.n_elements
    bid = gm.boundary_ids if gm.boundary_ids is not None else np.zeros(nf, dtype=np.int32)
    rid = gm.region_ids if gm.region_ids is not None else np.zeros(ne, dtype=np.int32)
    xyz = np.zeros((m.n_nodes, 3))
    xyz[:, :m.X.shape[1]] = m.X
    with open(path, "wb") as f:
        f.write(b"$MeshFormat\n")
        f.write(("2.1 %d 8\n" % (1 if binary else 0)).encode())
        if binary:
            f.write(struct.pack("<i", 1) + b"\n")
        f.write(b"$EndMeshFormat\n$Nodes\n")
        f.write(("%d\n" % m.n_nodes).encode())
        if binary:
            rec = np.zeros(m.n_nodes, dtype=[("id", "<i4"), ("x", "<f8", (3,))])
            rec["id"] = np.arange(1, m.n_nodes + 1)
            rec["x"] = xyz
            f.write(rec.tobytes() + b"\n")
        else:
            for i in range(m.n_nodes):
                f.write(("%d %r %r %r\n" % (i + 1, float(xyz[i, 0]), float(xyz[i, 1]), float(xyz[i, 2]))).encode())
        f.write(b"$EndNodes\n$Elements\n")
        f.write(("%d\n" % (nf + ne)).encode())
        if gm.element_owner is not None:
            ftags = np.stack([bid, np.zeros_like(bid), np.zeros_like(bid), gm.element_owner], axis=1)
        else:
            ftags = np.stack([bid, np.zeros_like(bid)], axis=1)
        etags = np.stack([rid, np.zeros_like(rid)], axis=1)
        fid = np.arange(1, nf + 1)[:, None]
        eid = np.arange(nf + 1, nf + ne + 1)[:, None]
        if binary:
            if nf:
                f.write(struct.pack("<3i", ftype, nf, ftags.shape[1]))
                f.write(np.hstack([fid, ftags, gm.sndgln]).astype("<i4").tobytes())
            f.write(struct.pack("<3i", etype, ne, 2))
            f.write(np.hstack([eid, etags, m.ndglno]).astype("<i4").tobytes() + b"\n")
        else:
            for k in range(nf):
                row = [k + 1, ftype, ftags.shape[1]] + list(ftags[k]) + list(gm.sndgln[k])
                f.write((" ".join(str(int(v)) for v in row) + "\n").encode())
            for k in range(ne):
                row = [nf + k + 1, etype, 2] + list(etags[k]) + list(m.ndglno[k])
                f.write((" ".join(str(int(v)) for v in row) + "\n").encode())
        f.write(b"$EndElements\n")


def _write_gmsh_fldecomp(path, gm):
    m = gm.mesh
    etype = {2: GMSH_TRIANGLE, 3: GMSH_TET}[m.dim]
    ftype = {2: GMSH_LINE, 3: GMSH_TRIANGLE}[m.dim]
    nf, ne = len(gm.sndgln), m.n_elements
    xyz = np.zeros((m.n_nodes, 3))
    xyz[:, :m.X.shape[1]] = m.X
    with open(path, "wb") as f:
        f.write(b"$MeshFormat\n2.1 1 8\n" + struct.pack("<i", 1) + b"\n$EndMeshFormat\n$Nodes\n")
        f.write(("%d\n" % m.n_nodes).encode())
        rec = np.zeros(m.n_nodes, dtype=[("id", "<i4"), ("x", "<f8", (3,))])
        rec["id"] = np.arange(1, m.n_nodes + 1)
        rec["x"] = xyz
        f.write(rec.tobytes() + b"\n$EndNodes\n$Elements\n")
        f.write(("%d\n" % (nf + ne)).encode())
        have_bid = gm.boundary_ids is not None and len(gm.boundary_ids) > 0
        f.write(struct.pack("<3i", ftype, nf, 1 if have_bid else 0))
        cols = [np.arange(1, nf + 1)[:, None]] + ([np.asarray(gm.boundary_ids)[:, None]] if have_bid else []) + [gm.sndgln]
        f.write(np.hstack(cols).astype("<i4").tobytes())
        rid = gm.region_ids if gm.region_ids is not None else np.zeros(ne, dtype=np.int32)
        f.write(struct.pack("<3i", etype, ne, 1))
        f.write(np.hstack([np.arange(nf + 1, nf + ne + 1)[:, None], np.asarray(rid)[:, None], m.ndglno]).astype("<i4").tobytes())
        f.write(b"\n$EndElements\n")


# ---- .halo ---------------------------------------------------------------------------------------------
@dataclass
class HaloLevel:
    n_private_nodes: int
    sends: list     # per process: 1-based local node ids
    receives: list  # per process: 1-based local node ids


@dataclass
class Halos:
    process: int
    nprocs: int
    levels: dict = _field(default_factory=dict)  # level -> HaloLevel


def _ints(el):
    txt = el.text if el is not None and el.text else ""
    return np.array([int(t) for t in txt.split()], dtype=np.int32)


def read_halo(path):
    """Halos_IO.cpp:34-176. Same validity rules: process/nprocs attributes, `level` (or the legacy
    `tag`), n_private_nodes >= 0, halo_data process in range and not repeated, empty lists allowed."""
    try:
        with open(path, "rb") as f:
            # TinyXML accepts comments / white space in front of the declaration and ReadHalos skips to
            # it (Halos_IO.cpp:45-49; the serial fixture tests/data/cube-parallel_0.halo starts with a comment)
            root = ET.fromstring(re.sub(rb"<\?xml[^>]*\?>", b"", f.read(), count=1).strip())
    except (ET.ParseError, OSError) as e:
        raise FormatError("Error reading halo file %s: %s" % (path, e))
    if root.tag != "halos" or root.get("process") is None or root.get("nprocs") is None:
        raise FormatError("Invalid .halo file")
    process, nprocs = int(root.get("process")), int(root.get("nprocs"))
    if process < 0 or process >= nprocs:
        raise FormatError("Invalid .halo file")
    out = Halos(process=process, nprocs=nprocs)
    for h in root.findall("halo"):
        lv = h.get("level", h.get("tag"))
        npn = h.get("n_private_nodes")
        if lv is None or npn is None or int(npn) < 0:
            raise FormatError("Invalid .halo file")
        sends = [np.zeros(0, dtype=np.int32) for _ in range(nprocs)]
        recvs = [np.zeros(0, dtype=np.int32) for _ in range(nprocs)]
        seen = set()
        for d in h.findall("halo_data"):
            p = d.get("process")
            if p is None or not 0 <= int(p) < nprocs or int(p) in seen:
                raise FormatError("Invalid .halo file")
            p = int(p)
            seen.add(p)
            sends[p] = _ints(d.find("send"))
            recvs[p] = _ints(d.find("receive"))
        out.levels[int(lv)] = HaloLevel(int(npn), sends, recvs)
    return out


def write_halo(path, halos):
    """Halos_IO.cpp:178-277: one <halo> per level, one <halo_data> per process, ids followed by a
    blank, XML declaration version 1.0 / utf-8."""
    lines = ['<?xml version="1.0" encoding="utf-8" ?>', '<halos process="%d" nprocs="%d">' % (halos.process, halos.nprocs)]
    for lv in sorted(halos.levels):
        h = halos.levels[lv]
        lines.append('    <halo level="%d" n_private_nodes="%d">' % (lv, h.n_private_nodes))
        for p in range(halos.nprocs):
            lines.append('        <halo_data process="%d">' % p)
            lines.append("            <send>%s</send>" % "".join("%d " % v for v in h.sends[p]))
            lines.append("            <receive>%s</receive>" % "".join("%d " % v for v in h.receives[p]))
            lines.append("        </halo_data>")
        lines.append("    </halo>")
    lines.append("</halos>")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def parallel_filename(basename, rank, ext):
    """`<basename>_<rank><ext>` (Halos_IO.cpp:309, femtools/Parallel_Tools.F90 parallel_filename)."""
    return "%s_%d%s" % (basename, rank, ext)


def trailing_receives_consistent(h):
    """femtools/Halos_Debug / Halo_Data_Types.F90 HALO_ORDER_TRAILING_RECEIVES: every receive node
    has an id above n_private_nodes, every send node at or below it, and no receive node appears
    twice."""
    rec = np.concatenate([np.asarray(r) for r in h.receives]) if h.receives else np.zeros(0, dtype=np.int32)
    snd = np.concatenate([np.asarray(s) for s in h.sends]) if h.sends else np.zeros(0, dtype=np.int32)
    return bool((rec > h.n_private_nodes).all() and (snd <= h.n_private_nodes).all() and (snd >= 1).all()
                and len(np.unique(rec)) == len(rec))


def write_decomposition(basename, parts, binary=False, style="femtools", region_ids=None):
    """Writes `<basename>_<rank>.msh` + `.halo` for every LocalPart the way fldecomp/flredecomp
    leave them: local numbering, owned nodes first; level-1 and level-2 halos. The level-1 lists
    are the level-2 lists restricted to the level-1 receive nodes (ids <= n_owned + n_l1) on the
    receiving side and to their images on the sending side. The boundary faces a LocalPart carries
    (partition_by_owner(..., sndgln, boundary_ids)) are written too. style "fldecomp" (+ the GLOBAL
    region_ids, default 0) gives byte for byte the files of fldecomp/fldgmsh.cpp for the same node ->
    partition map (tests/test_formats.py, against the compiled reference in oracle/_ref)."""
    from .partition import LocalPart  # noqa: F401  (documented type of `parts`)
    nprocs = len(parts)
    # position of the level-1 entries inside each rank's level-2 receive list
    l1pos = [[np.flatnonzero(np.asarray(parts[r].recvs[p]) <= parts[r].n_owned + parts[r].n_l1) for p in range(nprocs)]
             for r in range(nprocs)]
    for r, lp in enumerate(parts):
        gm = GmshMesh(mesh=lp.mesh, sndgln=np.zeros((0, lp.mesh.dim), dtype=np.int32))
        if getattr(lp, "sndgln", None) is not None:
            gm.sndgln, gm.boundary_ids = lp.sndgln, lp.boundary_ids
        if region_ids is not None and lp.global_element is not None:
            gm.region_ids = np.asarray(region_ids)[lp.global_element].astype(np.int32)
        write_gmsh(parallel_filename(basename, r, ".msh"), gm, binary=binary, style=style)
        l2 = HaloLevel(lp.n_owned, [np.asarray(s) for s in lp.sends], [np.asarray(v) for v in lp.recvs])
        l1 = HaloLevel(lp.n_owned,
                       [np.asarray(lp.sends[p])[l1pos[p][r]] for p in range(nprocs)],
                       [np.asarray(lp.recvs[p])[l1pos[r][p]] for p in range(nprocs)])
        write_halo(parallel_filename(basename, r, ".halo"), Halos(process=r, nprocs=nprocs, levels={1: l1, 2: l2}))


def read_decomposition(basename, rank, coordinate_dim=None):
    """`<basename>_<rank>.msh` + `.halo` (reference-produced or ours) -> partition.LocalPart. The
    halo used for halo_update is the highest level in the file (level 2 when present,
    Halos_Communications.F90:497-567 is called with the mesh's largest halo); global numbers are
    not stored in these files (the reference derives its universal numbering at run time)."""
    from .partition import LocalPart
    gm = read_gmsh(parallel_filename(basename, rank, ".msh"), coordinate_dim)
    hs = read_halo(parallel_filename(basename, rank, ".halo"))
    if hs.process != rank:
        raise FormatError("Unexpected process number in .halo file")  # Halos_IO.cpp:343-346
    top = hs.levels[max(hs.levels)]
    if not trailing_receives_consistent(top):
        raise FormatError("halo is not in trailing-receives order")
    n_l1 = 0
    if 1 in hs.levels:
        n_l1 = int(sum(len(r) for r in hs.levels[1].receives))
    return LocalPart(mesh=gm.mesh, n_owned=top.n_private_nodes, global_node=None, global_element=None,
                     sends=[np.asarray(s, dtype=np.int32) for s in top.sends],
                     recvs=[np.asarray(v, dtype=np.int32) for v in top.receives], n_l1=n_l1), gm, hs


# ---- PETSc binary viewer files ---------------------------------------------------------------------------
MAT_FILE_CLASSID = 1211216
VEC_FILE_CLASSID = 1211214


@dataclass
class PetscMat:
    rows: int
    cols: int
    findrm: np.ndarray  # (rows + 1,) 0-based row starts
    colm: np.ndarray    # (nnz,) 0-based columns
    val: np.ndarray     # (nnz,)


def read_petsc_binary(path, int64=False):
    """Every object in a PETSc binary viewer file, in file order: PetscMat for matrices, 1-d float64
    arrays for vectors. `int64`: the file was written by a --with-64-bit-indices build."""
    it = ">i8" if int64 else ">i4"
    isz = 8 if int64 else 4
    out = []
    with open(path, "rb") as f:
        data = f.read()
    off = 0

    def take(dtype, n):
        nonlocal off
        nb = np.dtype(dtype).itemsize * n
        if off + nb > len(data):
            raise FormatError("truncated PETSc binary file")
        a = np.frombuffer(data, dtype=dtype, count=n, offset=off)
        off += nb
        return a

    while off < len(data):
        cid = int(take(it, 1)[0])
        if cid == MAT_FILE_CLASSID:
            M, N, nz = (int(v) for v in take(it, 3))
            if nz < 0:
                raise FormatError("dense PETSc matrices are not written by dump_matrix")
            rl = take(it, M).astype(np.int64)
            if rl.sum() != nz:
                raise FormatError("PETSc matrix row lengths do not add up to nz")
            colm = take(it, nz).astype(np.int64)
            val = take(">f8", nz).astype(np.float64)
            findrm = np.concatenate([[0], np.cumsum(rl)])
            out.append(PetscMat(M, N, findrm, colm, val))
        elif cid == VEC_FILE_CLASSID:
            n = int(take(it, 1)[0])
            out.append(take(">f8", n).astype(np.float64))
        else:
            raise FormatError("unknown PETSc class id %d at byte %d" % (cid, off - isz))
    return out


def write_petsc_binary(path, objects, int64=False):
    it = ">i8" if int64 else ">i4"
    with open(path, "wb") as f:
        for o in objects:
            if isinstance(o, PetscMat):
                f.write(np.array([MAT_FILE_CLASSID, o.rows, o.cols, len(o.val)], dtype=it).tobytes())
                f.write(np.diff(o.findrm).astype(it).tobytes())
                f.write(np.asarray(o.colm).astype(it).tobytes())
                f.write(np.asarray(o.val).astype(">f8").tobytes())
            else:
                v = np.asarray(o, dtype=np.float64).ravel()
                f.write(np.array([VEC_FILE_CLASSID, len(v)], dtype=it).tobytes())
                f.write(v.astype(">f8").tobytes())


def petsc_row_numbering(n_nodes, nfields, group_size=1):
    """Serial gnn2unn (femtools/Petsc_Tools.F90:184-199), 0-based: (n_nodes, nfields). group_size 1
    (the default) numbers field by field: row = f*n_nodes + node; group_size = g interleaves the g
    fields of a group at every node: row = start + g*node + f."""
    fpg = group_size
    if nfields % fpg:
        raise ValueError("nfields must be a multiple of group_size")
    out = np.zeros((n_nodes, nfields), dtype=np.int64)
    start = 0
    node = np.arange(n_nodes)
    for g in range(nfields // fpg):
        for f in range(fpg):
            out[:, g * fpg + f] = start + fpg * node + f
        start += n_nodes * fpg
    return out


def blocks_to_petsc(findrm, colm, blocks, n_nodes, group_size=1, diagonal=True, keep_zeros=True):
    """The PETSc AIJ matrix that Sparse_Tools_Petsc.F90:848-879 builds from per-block additions.
    findrm/colm: 1-based femtools sparsity. blocks: diagonal=True -> (dim, nnz) = the (d,d) blocks
    (block_mask diagonal only, Momentum_CG.F90:1293-1300); else (dim, dim, nnz) indexed [bi][bj].
    Rows come out sorted by column, as MatView writes them."""
    blocks = np.asarray(blocks)
    dim = blocks.shape[0]
    num = petsc_row_numbering(n_nodes, dim, group_size)
    findrm0 = np.asarray(findrm, dtype=np.int64) - 1
    col0 = np.asarray(colm, dtype=np.int64) - 1
    row_of = np.repeat(np.arange(n_nodes), np.diff(findrm0))
    R, Cc, V = [], [], []
    for bi in range(dim):
        for bj in range(dim):
            if diagonal and bi != bj:
                continue
            v = blocks[bi] if diagonal else blocks[bi, bj]
            R.append(num[row_of, bi])
            Cc.append(num[col0, bj])
            V.append(np.asarray(v, dtype=np.float64))
    R, Cc, V = np.concatenate(R), np.concatenate(Cc), np.concatenate(V)
    if not keep_zeros:
        k = V != 0.0
        R, Cc, V = R[k], Cc[k], V[k]
    order = np.lexsort((Cc, R))
    R, Cc, V = R[order], Cc[order], V[order]
    nrows = n_nodes * dim
    fr = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(fr, R + 1, 1)
    return PetscMat(nrows, nrows, np.cumsum(fr), Cc, V)


def field_to_petsc(val, group_size=1):
    """A scalar (n_nodes,) or vector (n_nodes, dim) field as the PETSc Vec field2petsc builds with the serial
    numbering above (femtools/Petsc_Tools.F90 field2petsc): entry gnn2unn(node, component)."""
    val = np.asarray(val, dtype=np.float64)
    if val.ndim == 1:
        return val.copy()
    n, dim = val.shape
    num = petsc_row_numbering(n, dim, group_size)
    out = np.zeros(n * dim)
    out[num.ravel()] = val.ravel()
    return out


def csr_to_petsc(findrm, colm, val, n_nodes):
    """A femtools csr_matrix (1-based sparsity) as the PETSc matrix csr2petsc produces for a scalar
    field (femtools/Petsc_Tools.F90:1154-1301): same rows, 0-based sorted columns."""
    return blocks_to_petsc(findrm, colm, np.asarray(val)[None, :], n_nodes)


def compare_petsc_mats(a, ref, rtol=1e-12):
    """SURVEY.md 8(c) parity metric (ii) between two PetscMat with possibly different stored-zero
    patterns: max-norm relative error over the whole matrix and per row. Returns a dict; `ok` is
    the verdict."""
    import scipy.sparse as sp
    A = sp.csr_matrix((a.val, a.colm, a.findrm), shape=(a.rows, a.cols))
    B = sp.csr_matrix((ref.val, ref.colm, ref.findrm), shape=(ref.rows, ref.cols))
    if A.shape != B.shape:
        return {"ok": False, "reason": "shape %s vs %s" % (A.shape, B.shape)}
    D = abs(A - B).tocsr()
    scale = abs(B).max() if B.nnz else 0.0
    block = (D.max() / scale) if scale > 0 else float(D.max() if D.nnz else 0.0)
    rmax_ref = np.asarray(abs(B).max(axis=1).todense()).ravel()
    rmax_d = np.asarray(D.max(axis=1).todense()).ravel()
    nz = rmax_ref > 0
    row = float((rmax_d[nz] / rmax_ref[nz]).max()) if nz.any() else 0.0
    stray = float(rmax_d[~nz].max()) if (~nz).any() else 0.0
    return {"ok": bool(block <= rtol and row <= rtol and stray == 0.0), "block_rel": float(block), "row_rel": row,
            "entries_in_empty_reference_rows": stray}


_DUMP_RE = re.compile(r"^(?P<field>.+)_(?P<index>\d+)$")


def dump_name_parts(path):
    """dump_matrix file names are `matrixdump` or `<name>_<index>` (femtools/Solvers.F90:1496-1518: `<filename>_<dump_matrix_index>`)."""
    base = os.path.basename(path)
    m = _DUMP_RE.match(base)
    return (m.group("field"), int(m.group("index"))) if m else (base, None)


# ---- .stat files (diagnostics/Diagnostic_variables.F90 writes them; python/fluidity_tools.py stat_parser reads) -----
def read_stat(path, subsample=1):
    """A Fluidity `.stat` file: an XML `<header>` naming every column (`<field column= name= statistic=
    [material_phase=] [components=]/>`, `<constant name= type= value=/>`), then one line of reals per dump -- or, when the
    header holds the constant format = "binary", raw reals of `real_size` bytes in `<path>.dat`. Returns the hierarchy
    the reference's stat_parser builds: result[material_phase][field][statistic] (or result[field][statistic] for
    columns without a phase) -> array over the dumps, (components, dumps) for multi-component columns; the constants
    are kept under result["__constants__"]."""
    import xml.dom.minidom
    assert subsample > 0
    with open(path, "rb") as f:
        xml_lines = []
        while True:
            line = f.readline()
            if not line:
                raise FormatError("%s: no </header> in the .stat file" % path)
            xml_lines.append(line.decode())
            if "</header>" in xml_lines[-1]:
                break
        body = f.read().decode()
    dom = xml.dom.minidom.parseString("".join(xml_lines))
    constants = {}
    for el in dom.getElementsByTagName("constant"):
        constants[el.getAttribute("name")] = (el.getAttribute("type"), el.getAttribute("value"))
    fields = dom.getElementsByTagName("field")
    ncol = sum(int(el.getAttribute("components") or 1) for el in fields)
    if constants.get("format", ("", ""))[1] == "binary":
        real_size = int(constants.get("real_size", ("integer", "8"))[1])
        if real_size not in (4, 8):
            raise FormatError("%s: unexpected real size %d" % (path, real_size))
        raw = np.fromfile(path + ".dat", dtype=np.float32 if real_size == 4 else np.float64)
        rows = raw[: (raw.size // ncol) * ncol].reshape(-1, ncol)  # an incomplete last line is ignored
        columns = rows[::subsample].T.astype(np.float64)
    else:
        rows = []
        for n, line in enumerate(l for l in body.splitlines()):
            entries = line.split()
            if len(entries) != ncol:
                raise FormatError("%s: incomplete line %d: expected %d columns, got %d" % (path, n, ncol, len(entries)))
            if n % subsample == 0:
                rows.append([float(e) for e in entries])
        columns = np.array(rows, dtype=np.float64).reshape(len(rows), ncol).T
    out = {"__constants__": constants}
    for el in fields:
        phase, name = el.getAttribute("material_phase"), el.getAttribute("name")
        col, stat, comps = int(el.getAttribute("column")), el.getAttribute("statistic"), el.getAttribute("components")
        d = out.setdefault(phase, {}) if phase else out
        d.setdefault(name, {})[stat] = columns[col - 1: col - 1 + int(comps)] if comps else columns[col - 1]
    return out
