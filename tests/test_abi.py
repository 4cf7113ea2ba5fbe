"""CPU-only checks of the drop-in boundary: libcgasm.so loads, exports every symbol that
include/cgasm.h declares, the ctypes structs match the header, and (without a GPU) the
library refuses to work instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest

from conftest import ROOT
from fluidity_b200 import _abi as abi, cgasm, tables, synthetic as syn

HEADER = os.path.join(ROOT, "include", "cgasm.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cgasm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = cgasm.load()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libcgasm.so lacks " + s


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cgasm.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(cgasm_momentum_opts),'
                   'sizeof(cgasm_advdiff_opts), offsetof(cgasm_momentum_opts, lump_mass),'
                   'offsetof(cgasm_momentum_opts, integrate_continuity_by_parts),'
                   'offsetof(cgasm_advdiff_opts, have_mass),'
                   'offsetof(cgasm_advdiff_opts, equation_type_not_advdiff));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    M, A = abi.MomentumOpts, abi.AdvDiffOpts
    want = [C.sizeof(M), C.sizeof(A), M.lump_mass.offset, M.integrate_continuity_by_parts.offset,
            A.have_mass.offset, A.equation_type_not_advdiff.offset]
    assert got == want


def test_enums_match_header():
    txt = open(HEADER).read()
    for name, val in (("CGASM_F_T_ABSORPTION", abi.F_T_ABSORPTION), ("CGASM_F_NSLOTS", abi.F_NSLOTS),
                      ("CGASM_SCATTER_TILED", abi.SCATTER_TILED), ("CGASM_ENODEVICE", abi.ENODEVICE)):
        m = re.search(name + r"\s*=\s*(\d+)", txt)
        assert m and int(m.group(1)) == val, name


@pytest.mark.parametrize("dim", [2, 3])
def test_product_tables_equal_oracle_tables(orc, dim):
    n, dn, w = tables.p1_tables(dim)
    on, odn, ow = orc.tables(dim)
    assert (n == on).all() and (dn == odn).all() and (w == ow).all()


def test_no_cpu_path_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mesh = syn.box_mesh((2, 2, 2))
    with pytest.raises(cgasm.CgasmError) as ei:
        cgasm.Assembler(mesh, tables.p1_tables(3))
    assert ei.value.code in (abi.ENODEVICE, abi.ECUDA)


def test_bad_arguments_are_status_codes_not_aborts():
    lib = cgasm.load()
    ident = C.c_int(0)
    n, dn, w = tables.p1_tables(3)
    nd = np.ones(4, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    st = lib.cgasm_create(C.byref(ident), -1, 4, 5, 5, 4, 1, nd.ctypes.data_as(C.POINTER(C.c_int)), dp(n), dp(dn), dp(w))
    assert st == abi.EUNSUPPORTED
    st = lib.cgasm_create(C.byref(ident), -1, 3, 4, 11, 4, 1, nd.ctypes.data_as(C.POINTER(C.c_int)), dp(n), dp(dn), dp(w))
    assert st == abi.EUNSUPPORTED  # degree-4 quadrature: caller keeps the Fortran loop
    assert lib.cgasm_destroy(12345) == abi.EHANDLE
    assert lib.cgasm_momentum_dev(777, None) == abi.EHANDLE
    assert b"handle" in lib.cgasm_last_error()


# ---- the uncompiled Fortran half of the boundary must not drift from the header ----------------------
FORTRAN = os.path.join(ROOT, "fluidity_b200", "fortran", "cgasm_fortran.F90")
# diagnostics / device-pointer accessors a Fortran caller has no use for
_NOT_BOUND_IN_FORTRAN = {"cgasm_advdiff_result_dev", "cgasm_momentum_result_dev", "cgasm_last_kernel_ms",
                         "cgasm_launch_count", "cgasm_row_blocks_host", "cgasm_strip_plan_host", "cgasm_stream",
                         "cgasm_plan_host_timing", "cgasm_plan_host_stats", "cgasm_last_path", "cgasm_plan_stats", "cgasm_cmc_result_dev", "cgasm_kmk_result_dev", "cgasm_cmc_sparsity_host", "cgasm_cmc_expand_plan_host"}


def _fortran_type_fields(txt, name):
    body = re.search(r"type,\s*bind\(c\)\s*::\s*%s(.*?)end type" % name, txt, flags=re.S | re.I).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("!")[0]
        m = re.match(r"\s*(real\(c_double\)|integer\(c_int\))\s*::\s*(.*)", line)
        if m:
            fields += [(m.group(1), f.strip()) for f in m.group(2).split(",") if f.strip()]
    return fields


def test_fortran_module_binds_the_header():
    txt = open(FORTRAN).read()
    bound = set(re.findall(r'name\s*=\s*"(cgasm_[a-z0-9_]+)"', txt))
    declared = set(_declared_symbols())
    assert bound <= declared, bound - declared
    assert declared - bound == _NOT_BOUND_IN_FORTRAN
    # derived types: same members, same order, same kinds as the C structs (= the ctypes twins)
    for tname, cls in (("cgasm_momentum_opts", abi.MomentumOpts), ("cgasm_advdiff_opts", abi.AdvDiffOpts)):
        want = [("real(c_double)" if t is C.c_double else "integer(c_int)", k) for k, t in cls._fields_]
        assert _fortran_type_fields(txt, tname) == want, tname
    # enumerators repeated as Fortran parameters
    hdr = open(HEADER).read()
    for name, val in re.findall(r"\b(CGASM_[A-Z0-9_]+)\s*=\s*(\d+)", txt):
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, hdr)
        assert m and m.group(1) == val, name


def test_integration_guide_only_names_declared_symbols():
    """Every cgasm_* identifier the maintainer-facing documents mention exists in the header (or is one of its types)."""
    declared = set(_declared_symbols()) | {"cgasm_momentum_opts", "cgasm_advdiff_opts", "cgasm_interface", "cgasm_fortran",
                                           "cgasm_last_error_string"}  # (a Fortran-side wrapper of cgasm_last_error)
    for doc in ("INTEGRATION.md", "README.md", "DESIGN.md"):
        txt = open(os.path.join(ROOT, doc)).read()
        used = set(re.findall(r"\bcgasm_[a-z0-9_]+\b", txt))
        # prefixes written with a wildcard in prose (`cgasm_cmc_*`, `cgasm_*_surface_dev`) leave a trailing underscore
        used = {u for u in used if not u.endswith("_")}
        unknown = sorted(u for u in used if u not in declared and not any(d.startswith(u) for d in declared))
        assert not unknown, (doc, unknown)


def test_plain_c_client_compiles_links_and_is_refused_without_a_gpu(tmp_path):
    """examples/c_client.c: the ABI from C99 with nothing but the header and the shared library. In this container
    (no GPU) the library must refuse at cgasm_create with CGASM_ENODEVICE -- there is no CPU path to fall into."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the client would run")
    except ImportError:
        pass
    exe = tmp_path / "c_client"
    libdir = os.path.join(ROOT, "fluidity_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_client.c"), "-L", libdir, "-lcgasm", "-Wl,-rpath," + libdir, "-o", str(exe)],
                   check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == abi.ENODEVICE and "no CPU path" in p.stderr and p.stdout == ""


@pytest.mark.gpu
def test_plain_c_client_runs_on_the_gpu_and_prints_the_oracles_matrix(tmp_path, orc):
    """examples/c_client.c built with gcc against the header and the shared library only, RUN on the B200: the matrix and
    right-hand side it prints (two triangles, default tracer options) are the oracle's."""
    exe = tmp_path / "c_client"
    libdir = os.path.join(ROOT, "fluidity_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_client.c"), "-L", libdir, "-lcgasm", "-Wl,-rpath," + libdir, "-o", str(exe)],
                   check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    rows = [ln for ln in p.stdout.splitlines() if ln.startswith("row ")]
    assert len(rows) == 4
    mesh = syn.Mesh(dim=2, ndglno=np.array([[1, 2, 3], [2, 4, 3]], dtype=np.int32),
                    X=np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float64))
    fs = syn.FieldSet()
    fs.set(abi.F_T, np.array([0.0, 1.0, 0.5, 0.25]))
    fs.set(abi.F_NU, np.array([[1.0, 0.0]] * 4))
    fs.set(abi.F_T_DIFFUSIVITY, np.array([[[1e-3, 0.0], [0.0, 1e-3]]]), abi.FIELD_CONSTANT)
    o = abi.common_advdiff_opts()
    findrm, colm, _ = orc.make_sparsity(mesh)
    ref = orc.assemble_advdiff(mesh, fs, o, findrm, colm)
    got_vals, got_rhs, got_cols = [], [], []
    for ln in rows:
        body, rhs = ln.split("| rhs")
        got_rhs.append(float(rhs))
        for col, val in re.findall(r"\((\d+)\)\s+(\S+)", body):
            got_cols.append(int(col))
            got_vals.append(float(val))
    assert got_cols == list(colm)
    # the client prints 7 significant digits
    assert np.abs(np.array(got_vals) - ref["matrix"]).max() <= 1e-6 * np.abs(ref["matrix"]).max()
    assert np.abs(np.array(got_rhs) - ref["rhs"]).max() <= 1e-6 * np.abs(ref["rhs"]).max()
