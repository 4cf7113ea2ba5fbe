"""Regenerates tests/golden/*.npz|json from the reference's own in-tree fixtures.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Inputs (reference data files, not source code):
    tests/data/cube.1.msh, tests/data/cube-parallel.msh      ASCII gmsh 2.x, tets
    tests/data/square-cavity-2d.msh                            the test_colouring mesh
    tests/data/2d_square.msh
    tests/meshconv_test/src/prectangle_{0,1}.halo             a real 2-rank L1+L2 halo pair
    tests/meshconv_test/src/prectangle_{0,1}.msh              its two local meshes (binary gmsh)
Only node coordinates and the volume elements (gmsh type 4 = tet, type 2 = triangle on 2-D
meshes) are kept, as float64 / int32 arrays.
"""
import json
import os
import re
import sys
import xml.etree.ElementTree as ET
import numpy as np

REF = os.environ.get("FLUIDITY_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def read_gmsh_ascii(path, dim):
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    i = lines.index("$Nodes")
    nn = int(lines[i + 1])
    nodes = np.array([[float(x) for x in lines[i + 2 + k].split()[1:4]] for k in range(nn)])
    ids = np.array([int(lines[i + 2 + k].split()[0]) for k in range(nn)])
    assert (ids == np.arange(1, nn + 1)).all()
    i = lines.index("$Elements")
    ne = int(lines[i + 1])
    want = 4 if dim == 3 else 2
    nloc = dim + 1
    eles = []
    for k in range(ne):
        t = [int(x) for x in lines[i + 2 + k].split()]
        etype, ntags = t[1], t[2]
        if etype == want:
            eles.append(t[3 + ntags:3 + ntags + nloc])
    return nodes[:, :dim].copy(), np.array(eles, dtype=np.int32)


def read_halo(path):
    root = ET.parse(path).getroot()
    out = {"process": int(root.get("process")), "nprocs": int(root.get("nprocs")), "levels": {}}
    for h in root.findall("halo"):
        lvl = h.get("level") or h.get("tag")
        ent = {"n_private_nodes": int(h.get("n_private_nodes")), "sends": {}, "receives": {}}
        for hd in h.findall("halo_data"):
            p = hd.get("process")
            ent["sends"][p] = [int(x) for x in (hd.find("send").text or "").split()]
            ent["receives"][p] = [int(x) for x in (hd.find("receive").text or "").split()]
        out["levels"][lvl] = ent
    return out


def main():
    data = os.path.join(REF, "tests", "data")
    for name, dim in (("cube.1", 3), ("cube-parallel", 3), ("square-cavity-2d", 2), ("2d_square", 2)):
        X, nd = read_gmsh_ascii(os.path.join(data, name + ".msh"), dim)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, ndglno=nd, dim=dim)
        print(name, X.shape, nd.shape)
    halos = {}
    for r in (0, 1):
        halos[str(r)] = read_halo(os.path.join(REF, "tests", "meshconv_test", "src",
                                               "prectangle_%d.halo" % r))
    with open(os.path.join(OUT, "prectangle_halos.json"), "w") as f:
        json.dump(halos, f, indent=1)
    print("halos", {k: {l: v["n_private_nodes"] for l, v in h["levels"].items()} for k, h in halos.items()})
    # the matching decomposed meshes (binary gmsh 2.1 with 4 face tags) and the serial mesh they
    # were cut from, through the package's own reader (fluidity_b200/formats.py; its ASCII path
    # is cross-checked against read_gmsh_ascii above by tests/test_formats.py)
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from fluidity_b200 import formats
    src = os.path.join(REF, "tests", "meshconv_test", "src")
    for r in (0, 1):
        g = formats.read_gmsh(os.path.join(src, "prectangle_%d.msh" % r))
        np.savez_compressed(os.path.join(OUT, "prectangle_%d.npz" % r), X=g.mesh.X, ndglno=g.mesh.ndglno, dim=g.mesh.dim,
                            sndgln=g.sndgln, boundary_ids=g.boundary_ids, element_owner=g.element_owner,
                            region_ids=g.region_ids)
        print("prectangle_%d" % r, g.mesh.X.shape, g.mesh.ndglno.shape, g.sndgln.shape)


if __name__ == "__main__":
    sys.exit(main())
