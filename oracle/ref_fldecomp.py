"""ctypes access to oracle/_ref/libref_fldecomp.so: the REFERENCE'S OWN decomposition writer
(fldecomp/fldgmsh.cpp write_partitions_gmsh, compiled unmodified by `make -C oracle ref` with the glue
oracle/ref_fldecomp_shim.cpp). Given a node -> partition map it writes <name>_<part>.msh + .halo exactly as
`fldecomp` does after METIS has produced the map. TEST INFRASTRUCTURE (tests/test_formats.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_fldecomp.so")
REF = os.environ.get("FLUIDITY_REFERENCE", "/root/reference")
_LIB = None


def available():
    return os.path.exists(_PATH) or os.path.isdir(os.path.join(REF, "fldecomp"))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-s", "-C", _HERE, "ref", "REF=" + REF], check=True)
        _LIB = C.CDLL(_PATH)
    return _LIB


def write_partitions(basename, mesh, owner, nparts, sndgln=None, boundary_ids=None, region_ids=None):
    """mesh: synthetic.Mesh (1-based ndglno); owner (n_nodes,) 0-based partition of every node."""
    X = np.ascontiguousarray(mesh.X, dtype=np.float64)
    nd = np.ascontiguousarray(mesh.ndglno, dtype=np.int32)
    dec = np.ascontiguousarray(owner, dtype=np.int32)
    rid = np.ascontiguousarray(region_ids if region_ids is not None else np.zeros(mesh.n_elements), dtype=np.int32)
    sn = np.ascontiguousarray(sndgln if sndgln is not None else np.zeros((0, mesh.dim)), dtype=np.int32)
    bid = np.ascontiguousarray(boundary_ids if boundary_ids is not None else np.zeros(len(sn)), dtype=np.int32)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    st = lib().ref_write_partitions_gmsh(basename.encode(), C.c_int(nparts), C.c_int(mesh.n_nodes), C.c_int(mesh.dim),
                                         X.ctypes.data_as(dp), dec.ctypes.data_as(ip), C.c_int(mesh.loc),
                                         C.c_int(mesh.n_elements), nd.ctypes.data_as(ip), rid.ctypes.data_as(ip),
                                         C.c_int(mesh.dim), C.c_int(len(sn)), sn.ctypes.data_as(ip), bid.ctypes.data_as(ip))
    if st:
        raise RuntimeError("reference write_partitions_gmsh failed")
