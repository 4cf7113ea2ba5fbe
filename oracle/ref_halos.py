"""ctypes access to oracle/_ref/libref_halos_io.so: the REFERENCE'S OWN .halo reader and writer
(femtools/Halos_IO.cpp, compiled unmodified from /root/reference by `make -C oracle ref`), through the
extern "C" entry points its Fortran binds (femtools/Halos_IO.F90: chalo_reader_*, chalo_writer_*).
TEST INFRASTRUCTURE: used by tests/test_formats.py to validate fluidity_b200/formats.py; never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_halos_io.so")
REF = os.environ.get("FLUIDITY_REFERENCE", "/root/reference")
_LIB = None


def available():
    """True if the library exists or can be built (the reference tree is present)."""
    return os.path.exists(_PATH) or os.path.isdir(os.path.join(REF, "femtools"))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-s", "-C", _HERE, "ref", "REF=" + REF], check=True)
        _LIB = C.CDLL(_PATH)
    return _LIB


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def read(basename, process, nprocs, levels=(1, 2)):
    """cHaloReaderSetInput + QueryOutput + GetOutput: {level: (n_private_nodes, sends[p], receives[p])}."""
    L = lib()
    name = basename.encode()
    ln, pr, np_ = C.c_int(len(name)), C.c_int(process), C.c_int(nprocs)
    errors = L.chalo_reader_set_input_(C.c_char_p(name), C.byref(ln), C.byref(pr), C.byref(np_))
    if errors:
        L.chalo_reader_reset_()
        raise ValueError("reference halo reader reports %d error(s) for %s_%d.halo" % (errors, basename, process))
    out = {}
    for level in levels:
        lv = C.c_int(level)
        ns, nr = np.zeros(nprocs, dtype=np.int32), np.zeros(nprocs, dtype=np.int32)
        L.chalo_reader_query_output_(C.byref(lv), C.byref(np_), _ip(ns), _ip(nr))
        send, recv = np.zeros(max(int(ns.sum()), 1), dtype=np.int32), np.zeros(max(int(nr.sum()), 1), dtype=np.int32)
        npn = C.c_int(0)
        L.chalo_reader_get_output_(C.byref(lv), C.byref(np_), _ip(ns), _ip(nr), C.byref(npn), _ip(send), _ip(recv))
        so, ro = np.concatenate([[0], np.cumsum(ns)]), np.concatenate([[0], np.cumsum(nr)])
        out[level] = (npn.value, [send[so[p]:so[p + 1]].copy() for p in range(nprocs)],
                      [recv[ro[p]:ro[p + 1]].copy() for p in range(nprocs)])
    L.chalo_reader_reset_()
    return out


def write(basename, process, nprocs, levels):
    """cHaloWriterInitialise + SetInput (per level) + Write. levels: {level: (n_private_nodes, sends[p], receives[p])}."""
    L = lib()
    pr, np_ = C.c_int(process), C.c_int(nprocs)
    L.chalo_writer_initialise_(C.byref(pr), C.byref(np_))
    for level, (npn, sends, recvs) in sorted(levels.items()):
        ns = np.array([len(s) for s in sends], dtype=np.int32)
        nr = np.array([len(r) for r in recvs], dtype=np.int32)
        send = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in sends] + [np.zeros(0, np.int32)]))
        recv = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in recvs] + [np.zeros(0, np.int32)]))
        if len(send) == 0:
            send = np.zeros(1, dtype=np.int32)
        if len(recv) == 0:
            recv = np.zeros(1, dtype=np.int32)
        lv, n = C.c_int(level), C.c_int(npn)
        L.chalo_writer_set_input_(C.byref(lv), C.byref(np_), _ip(ns), _ip(nr), C.byref(n), _ip(send), _ip(recv))
    name = basename.encode()
    ln = C.c_int(len(name))
    L.chalo_writer_write_.restype = C.c_int
    st = L.chalo_writer_write_(C.c_char_p(name), C.byref(ln))
    L.chalo_writer_reset_()
    if st:
        raise OSError("reference halo writer failed for %s_%d.halo" % (basename, process))
