"""The per-face arithmetic the device kernels run (fluidity_b200/csrc/surface_math.h) compiled for the host
with g++ and compared with the oracle, face by face, for every boundary-condition branch. CPU only: this
checks the product's formulae without a GPU; tests/test_surface_gpu.py checks the kernels around them."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from conftest import ROOT, load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, tables

TOL = 1e-12
c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(c_dp)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("harness") / "libsurface_harness.so"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-Wall", "-Werror",
                    os.path.join(ROOT, "tests", "surface_harness.cpp"), "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    lib.harness_momentum_face.restype = C.c_int
    lib.harness_momentum_face_fs.restype = C.c_int
    lib.harness_csr_pos0.restype = C.c_int
    return lib


def _cases():
    return {"box2": syn.box_mesh((5, 4), seed=3), "box3": syn.box_mesh((3, 4, 2), seed=5), "cube.1": load_golden_mesh("cube.1")}


@pytest.mark.parametrize("name", ["box2", "box3", "cube.1"])
def test_tracer_face_math_equals_the_oracle(orc, harness, name):
    mesh = _cases()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    sn, fe = syn.boundary_faces(mesh)
    n, dn, w = tables.p1_face_tables(dim)
    rng = np.random.default_rng(31)
    T, U = fs.get(abi.F_T)[0], fs.get(abi.F_NU)[0]
    variants = [(abi.common_advdiff_opts(), abi.TBC_NEUMANN), (abi.common_advdiff_opts(), abi.TBC_ROBIN),
                (abi.common_advdiff_opts(), abi.TBC_NONE), (abi.common_advdiff_opts(theta=0.0), abi.TBC_ROBIN),
                (abi.common_advdiff_opts(integrate_advection_by_parts=1), abi.TBC_NONE),
                (abi.common_advdiff_opts(integrate_advection_by_parts=1), abi.TBC_ROBIN),
                (abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0), abi.TBC_WEAKDIRICHLET),
                (abi.common_advdiff_opts(integrate_advection_by_parts=1, have_diffusivity=0, theta=0.0), abi.TBC_WEAKDIRICHLET),
                (abi.common_advdiff_opts(have_advection=0), abi.TBC_NEUMANN)]
    for f in rng.choice(len(fe), size=min(10, len(fe)), replace=False):
        Xf = np.ascontiguousarray(mesh.X[sn[f] - 1])
        Xc = mesh.X[mesh.ndglno[fe[f] - 1] - 1].mean(0)
        Tf, Uf = np.ascontiguousarray(T[sn[f] - 1]), np.ascontiguousarray(U[sn[f] - 1])
        bc, bc2 = rng.uniform(size=dim), rng.uniform(0.5, 2, size=dim)
        for o, bt in variants:
            A, r = np.zeros((dim, dim)), np.zeros(dim)
            harness.harness_advdiff_face(C.c_int(dim), C.c_int(len(w)), _dp(n), _dp(dn), _dp(w), C.byref(o), C.c_int(bt),
                                         _dp(Xf), _dp(Xc), _dp(Tf), _dp(Uf), _dp(bc), _dp(bc2), _dp(A), _dp(r))
            oA, orr = orc.advdiff_face(mesh, fs, o, sn, fe, f + 1, bt, bc, bc2)
            assert rel_err(A, oA) < TOL and rel_err(r, orr) < TOL, (f, bt)


@pytest.mark.parametrize("name", ["box2", "box3"])
def test_momentum_face_math_equals_the_oracle(orc, harness, name):
    mesh = _cases()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    sn, fe = syn.boundary_faces(mesh)
    n, dn, w = tables.p1_face_tables(dim)
    rng = np.random.default_rng(32)
    U, O, R = fs.get(abi.F_NU)[0], fs.get(abi.F_OLDU)[0], fs.get(abi.F_DENSITY)[0]
    opts = [abi.common_momentum_opts(integrate_advection_by_parts=1), abi.common_momentum_opts(),
            abi.common_momentum_opts(integrate_advection_by_parts=1, exclude_advection=1)]
    types = [[0] * dim, [abi.VBC_WEAKDIRICHLET] * dim, [abi.VBC_WEAKDIRICHLET] + [0] * (dim - 1),
             [0] * (dim - 1) + [abi.VBC_FLUX], [abi.VBC_NO_NORMAL_FLOW] + [abi.VBC_FLUX] * (dim - 1),
             [abi.VBC_FREE_SURFACE] * dim]
    for f in rng.choice(len(fe), size=min(8, len(fe)), replace=False):
        Xf = np.ascontiguousarray(mesh.X[sn[f] - 1])
        Xc = mesh.X[mesh.ndglno[fe[f] - 1] - 1].mean(0)
        Uf, Of, rho = (np.ascontiguousarray(a[sn[f] - 1]) for a in (U, O, R))
        bc = rng.uniform(size=(dim, dim))  # [face node, component]
        for o in opts:
            for bt in types:
                B, r = np.zeros((dim, dim, dim)), np.zeros((dim, dim))
                bta = np.array(bt, dtype=np.int32)
                skipped = harness.harness_momentum_face(C.c_int(dim), C.c_int(len(w)), _dp(n), _dp(dn), _dp(w), C.byref(o),
                                                        bta.ctypes.data_as(c_ip), C.c_int(0), _dp(Xf), _dp(Xc), _dp(Uf), _dp(Of),
                                                        _dp(rho), _dp(bc), _dp(B), _dp(r))
                assert skipped == 0
                oB, orr = orc.momentum_face(mesh, fs, o, sn, fe, f + 1, bta, bc)
                scale = max(np.abs(oB).max(), 1e-300)
                assert np.abs(B - oB).max() <= TOL * scale and rel_err(r, orr) < TOL, (f, bt)
    # the skip rule (Momentum_CG.F90:799-803)
    z = np.zeros(dim * dim)
    for bt, ptype, want in (([abi.VBC_NO_NORMAL_FLOW] + [0] * (dim - 1), 0, 1), ([abi.VBC_NO_NORMAL_FLOW] + [0] * (dim - 1), 1, 0),
                            ([0] * (dim - 1) + [abi.VBC_INTERNAL], 0, 1), ([abi.VBC_NO_NORMAL_FLOW] + [abi.VBC_FLUX] * (dim - 1), 0, 0)):
        bta = np.array(bt, dtype=np.int32)
        got = harness.harness_momentum_face(C.c_int(dim), C.c_int(len(w)), _dp(n), _dp(dn), _dp(w), C.byref(opts[0]),
                                            bta.ctypes.data_as(c_ip), C.c_int(ptype), _dp(z), _dp(z), _dp(z), _dp(z), _dp(z), _dp(z),
                                            _dp(np.zeros(dim ** 3)), _dp(np.zeros(dim * dim)))
        assert got == want, (bt, ptype)


def test_csr_position_search(orc, harness):
    mesh = load_golden_mesh("cube-parallel")
    findrm, colm, _ = orc.make_sparsity(mesh)
    f0, c0 = np.ascontiguousarray(findrm - 1, dtype=np.int32), np.ascontiguousarray(colm - 1, dtype=np.int32)
    rng = np.random.default_rng(0)
    for i in rng.integers(0, mesh.n_nodes, size=200):
        row = c0[f0[i]:f0[i + 1]]
        for j in list(row[[0, -1, len(row) // 2]]) + [int(rng.integers(0, mesh.n_nodes))]:
            got = harness.harness_csr_pos0(f0.ctypes.data_as(c_ip), c0.ctypes.data_as(c_ip), C.c_int(int(i)), C.c_int(int(j)))
            hit = np.flatnonzero(row == j)
            assert got == (f0[i] + hit[0] if len(hit) else -1)


@pytest.mark.parametrize("name", ["box2", "box3"])
@pytest.mark.parametrize("case", ["lumped", "lumped_pressure_corrected", "consistent"])
def test_free_surface_stabilisation_face_math_equals_the_oracle(orc, harness, name, case):
    """Momentum_CG.F90:1108-1176 on faces of type FREE_SURFACE: shape_shape_vector with dt g fs_sf (n . k) k, lumped onto
    the diagonal (lump_mass; into masslump too with pressure-corrected absorption) or as a full face matrix."""
    mesh = _cases()[name]
    dim = mesh.dim
    fs = syn.standard_fields(mesh)
    rng = np.random.default_rng(5)
    nodal_g = rng.normal(size=(mesh.n_nodes, dim))
    nodal_g /= np.linalg.norm(nodal_g, axis=1)[:, None]
    sn, fe = syn.boundary_faces(mesh)
    n, dn, w = tables.p1_face_tables(dim)
    U, O, R = fs.get(abi.F_NU)[0], fs.get(abi.F_OLDU)[0], fs.get(abi.F_DENSITY)[0]
    o = abi.common_momentum_opts(have_surface_fs_stabilisation=1, fs_sf=0.7, lump_mass=int(case != "consistent"),
                                 pressure_corrected_absorption=int(case == "lumped_pressure_corrected"),
                                 have_absorption=int(case == "lumped_pressure_corrected"),
                                 lump_absorption=int(case == "lumped_pressure_corrected"))
    bta = np.array([abi.VBC_FREE_SURFACE] * dim, dtype=np.int32)
    nonzero = 0
    for gravity, ftype in ((np.eye(dim)[dim - 1:dim] * -1.0, abi.FIELD_CONSTANT), (nodal_g, abi.FIELD_NORMAL)):
        fs.set(abi.F_GRAVITY, np.ascontiguousarray(gravity), ftype)
        for f in rng.choice(len(fe), size=min(10, len(fe)), replace=False):
            Xf = np.ascontiguousarray(mesh.X[sn[f] - 1])
            Xc = mesh.X[mesh.ndglno[fe[f] - 1] - 1].mean(0)
            Uf, Of, rho = (np.ascontiguousarray(a[sn[f] - 1]) for a in (U, O, R))
            Gf = np.ascontiguousarray(gravity[sn[f] - 1] if ftype == abi.FIELD_NORMAL else np.repeat(gravity, dim, axis=0))
            B, r, ml = np.zeros((dim, dim, dim)), np.zeros((dim, dim)), np.zeros((dim, dim))
            bc = np.zeros((dim, dim))
            assert harness.harness_momentum_face_fs(C.c_int(dim), C.c_int(len(w)), _dp(n), _dp(dn), _dp(w), C.byref(o),
                                                    bta.ctypes.data_as(c_ip), C.c_int(0), _dp(Xf), _dp(Xc), _dp(Uf), _dp(Of),
                                                    _dp(rho), _dp(bc), _dp(Gf), _dp(B), _dp(r), _dp(ml)) == 0
            oB, orr, oml = orc.momentum_face(mesh, fs, o, sn, fe, f + 1, bta, want_masslump=True)
            scale = max(np.abs(oB).max(), 1e-300)
            assert np.abs(B - oB).max() <= TOL * scale and np.abs(r - orr).max() <= TOL * max(np.abs(orr).max(), 1e-300)
            assert np.abs(ml - oml).max() <= TOL * scale
            nonzero += int(np.abs(oB).max() > 0)
            if case == "lumped_pressure_corrected":
                assert np.abs(oml).max() > 0 or np.abs(oB).max() == 0
            else:
                assert np.abs(oml).max() == 0
            if case != "consistent":   # lumped: diagonal only
                for d in range(dim):
                    assert np.abs(oB[d] - np.diag(np.diag(oB[d]))).max() == 0
    assert nonzero > 0  # (faces whose normal is orthogonal to gravity contribute nothing)
