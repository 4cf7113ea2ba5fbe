"""Lumped-mass pressure matrix C M_L^-1 C^T next to the path: second-order sparsity
(make_sparsity_mult) bit-exact, values against the oracle and against an independent scipy triple product.
CPU part: oracle, the library's host pattern builder, and the device entry function compiled for the host."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, load_golden_mesh, rel_err
from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables


def cases():
    return {"box3": syn.box_mesh((4, 3, 5), seed=2), "box2": syn.box_mesh((7, 6), seed=4),
            "cube-parallel": load_golden_mesh("cube-parallel"), "cavity": load_golden_mesh("square-cavity-2d"),
            "cube.1": load_golden_mesh("cube.1"), "shuffled": syn.shuffled(syn.box_mesh((5, 5, 5)), seed=3)}


def scipy_reference(mesh, findrm, colm, ct, w, findrm2, colm2):
    n = mesh.n_nodes
    Cd = [sp.csr_matrix((ct[d], colm - 1, findrm - 1), shape=(n, n)) for d in range(mesh.dim)]
    M = sum(Cd[d] @ sp.diags(w[:, d]) @ Cd[d].T for d in range(mesh.dim)).tocsr()
    return np.asarray(M[np.repeat(np.arange(n), np.diff(findrm2)), colm2 - 1]).ravel()


def inputs(orc, mesh):
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    res = orc.assemble_momentum(mesh, fs, abi.common_momentum_opts(assemble_ct_matrix_here=1), findrm, colm, want_ct=True)
    return fs, findrm, colm, res["ct_m"], 1.0 / res["masslump"]


@pytest.mark.parametrize("name", list(cases()))
def test_second_order_sparsity_bit_exact(orc, name):
    mesh = cases()[name]
    findrm, colm, _ = orc.make_sparsity(mesh)
    f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    n = mesh.n_nodes
    S = sp.csr_matrix((np.ones(len(colm)), colm - 1, findrm - 1), shape=(n, n))
    P = (S @ S).tocsr()
    P.sort_indices()
    assert (P.indptr + 1 == f2).all() and (P.indices + 1 == c2).all()      # the restated list algorithm
    g2, d2 = cgasm.cmc_sparsity_host(findrm, colm)                          # the library's builder
    assert g2.dtype == np.int32 and (g2 == f2).all() and (d2 == c2).all()


@pytest.mark.parametrize("name", list(cases()))
def test_cmc_values_oracle_vs_scipy(orc, name):
    mesh = cases()[name]
    fs, findrm, colm, ct, w = inputs(orc, mesh)
    f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    got = orc.mult_div_vector_div_T(findrm, colm, ct, ct, w, f2, c2)
    ref = scipy_reference(mesh, findrm, colm, ct, w, f2, c2)
    assert rel_err(got, ref) < 1e-12
    # symmetric positive semi-definite with the constants in its kernel (C^T 1 = 0 away from the boundary is
    # not exact with boundaries, so only symmetry and the sign of the diagonal are size-independent here)
    n = mesh.n_nodes
    M = sp.csr_matrix((got, c2 - 1, f2 - 1), shape=(n, n))
    assert abs(M - M.T).max() <= 1e-12 * np.abs(got).max() and (M.diagonal() > 0).all()


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("cmc") / "libcmc_harness.so"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wall", "-Werror",
                    os.path.join(ROOT, "tests", "cmc_harness.cpp"), "-o", str(out)], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("name", ["box3", "cavity", "shuffled"])
def test_device_entry_function_on_the_host_is_bitwise_the_oracle(orc, harness, name):
    mesh = cases()[name]
    fs, findrm, colm, ct, w = inputs(orc, mesh)
    f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    ref = orc.mult_div_vector_div_T(findrm, colm, ct, ct, w, f2, c2)
    i32 = lambda a: np.ascontiguousarray(a - 1, dtype=np.int32)
    f0, c0, g0, d0 = i32(findrm), i32(colm), i32(f2), i32(c2)
    out = np.zeros(len(c2))
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    ctc, wc = np.ascontiguousarray(ct), np.ascontiguousarray(w)
    harness.harness_cmc(C.c_int(mesh.dim), C.c_int(mesh.n_nodes), f0.ctypes.data_as(ip), c0.ctypes.data_as(ip),
                        ctc.ctypes.data_as(dp), C.c_longlong(len(colm)), wc.ctypes.data_as(dp), g0.ctypes.data_as(ip),
                        d0.ctypes.data_as(ip), out.ctypes.data_as(dp))
    assert (out == ref).all()  # same operations in the same order, no contraction


@pytest.mark.parametrize("name", list(cases()))
def test_expansion_plan_and_kernel_emulation_are_bitwise_the_oracle(orc, harness, name):
    """The default device kernel expands (k in row i) x (j in row k) with a plan built by the library on the host:
    plan invariants, and a lane-level emulation of the kernel with that plan against the oracle, bit for bit."""
    mesh = cases()[name]
    fs, findrm, colm, ct, w = inputs(orc, mesh)
    f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    plan = cgasm.cmc_expand_plan_host(findrm, colm, f2, c2)
    assert plan is not None
    tpos, pptr, slots, n2max = plan
    f0, c0 = findrm - 1, colm - 1
    rows = np.repeat(np.arange(mesh.n_nodes), np.diff(f0))
    assert (c0[tpos] == rows).all() and (rows[tpos] == c0).all()          # transposed positions
    assert n2max == np.diff(f2).max() and pptr[-1] == len(slots)
    per_row = np.array([np.diff(f0)[c0[f0[i]:f0[i + 1]]].sum() for i in range(mesh.n_nodes)])
    assert (np.diff(pptr) == per_row).all()
    i = mesh.n_nodes // 2                                                   # slots of one row, spelled out
    want = [np.searchsorted(c2[f2[i] - 1:f2[i + 1] - 1], colm[p]) for a in range(f0[i], f0[i + 1])
            for p in range(f0[c0[a]], f0[c0[a] + 1])]
    assert slots[pptr[i]:pptr[i + 1]].tolist() == want
    ref = orc.mult_div_vector_div_T(findrm, colm, ct, ct, w, f2, c2)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    out = np.zeros(len(c2))
    ctc, wc, f0c, c0c, g0 = np.ascontiguousarray(ct), np.ascontiguousarray(w), i32(f0), i32(c0), i32(f2 - 1)
    harness.harness_cmc_expand.restype = C.c_longlong
    bad = harness.harness_cmc_expand(C.c_int(mesh.dim), C.c_int(mesh.n_nodes), f0c.ctypes.data_as(ip), c0c.ctypes.data_as(ip),
                                     ctc.ctypes.data_as(dp), C.c_longlong(len(colm)), wc.ctypes.data_as(dp),
                                     tpos.ctypes.data_as(ip), pptr.ctypes.data_as(C.POINTER(C.c_longlong)),
                                     slots.ctypes.data_as(C.POINTER(C.c_ushort)), g0.ctypes.data_as(ip), C.c_int(n2max),
                                     out.ctypes.data_as(dp))
    assert bad == 0          # no two lanes of a step share a slot; the plan covers each row exactly
    assert (out == ref).all()


def test_expansion_plan_is_refused_for_patterns_it_cannot_serve(orc):
    mesh = cases()["box2"]
    findrm, colm, _ = orc.make_sparsity(mesh)
    f2, c2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    # a second-order pattern that lacks an entry of S.S
    keep = np.ones(len(c2), dtype=bool)
    keep[f2[5] - 1 + 1] = False
    g2 = f2.copy()
    g2[6:] -= 1
    assert cgasm.cmc_expand_plan_host(findrm, colm, g2, c2[keep]) is None
    # a structurally asymmetric first-order pattern (entry (1, j) dropped, (j, 1) kept)
    drop = np.ones(len(colm), dtype=bool)
    drop[findrm[0] - 1 + 1] = False
    gf = findrm.copy()
    gf[1:] -= 1
    assert cgasm.cmc_expand_plan_host(gf, colm[drop], f2, c2) is None


# ---- CUDA path ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases()))
def test_cmc_on_the_device(orc, name):
    mesh = cases()[name]
    fs = syn.standard_fields(mesh)
    asm = cgasm.Assembler(mesh, tables.p1_tables(mesh.dim))
    asm.build_sparsity()
    asm.set_fields(fs)
    findrm, colm, _ = asm.get_sparsity()
    nnz2 = asm.cmc_build_sparsity()
    f2, c2 = asm.cmc_get_sparsity()
    of2, oc2 = orc.make_sparsity_mult(mesh.n_nodes, findrm, colm)
    assert nnz2 == len(oc2) and (f2 == of2).all() and (c2 == oc2).all()
    o = abi.common_momentum_opts(assemble_ct_matrix_here=1)
    # resident inputs: ct_m and the lumped mass of the momentum loop that just ran
    asm.momentum_dev(o)
    asm.cmc_dev()
    got = asm.cmc_fetch()
    ref_m = orc.assemble_momentum(mesh, fs, o, findrm, colm, want_ct=True)
    ref = orc.mult_div_vector_div_T(findrm, colm, ref_m["ct_m"], ref_m["ct_m"], 1.0 / ref_m["masslump"], of2, oc2)
    assert rel_err(got, ref) < 1e-12
    # caller-provided inputs (inverse mass with strong Dirichlet rows zeroed, Momentum_CG.F90:873-876),
    # on a pattern adopted from the reference: same operations in the same order as the oracle
    w = 1.0 / ref_m["masslump"]
    w[:: 7] = 0.0
    asm.cmc_set_sparsity(of2, oc2)
    asm.cmc_dev(ref_m["ct_m"], w)
    got = asm.cmc_fetch()
    ref = orc.mult_div_vector_div_T(findrm, colm, ref_m["ct_m"], ref_m["ct_m"], w, of2, oc2)
    assert rel_err(got, ref) < 1e-12
    assert (got == ref).all()
    # the merge kernel (fallback for patterns without an expansion plan) gives the same bits
    os.environ["CGASM_CMC_MERGE"] = "1"
    try:
        asm.cmc_dev(ref_m["ct_m"], w)
        assert (asm.cmc_fetch() == ref).all()
    finally:
        del os.environ["CGASM_CMC_MERGE"]


@pytest.mark.gpu
def test_cmc_call_order_and_arguments():
    mesh = syn.box_mesh((3, 3), seed=1)
    asm = cgasm.Assembler(mesh, tables.p1_tables(2))
    asm.build_sparsity()
    asm.set_fields(syn.standard_fields(mesh))

    def code(fn, *a):
        with pytest.raises(cgasm.CgasmError) as ei:
            fn(*a)
        return ei.value.code

    assert code(asm.cmc_dev) == abi.ESTATE                         # no second-order sparsity yet
    asm.cmc_build_sparsity()
    assert code(asm.cmc_dev) == abi.ESTATE                         # nothing resident to read
    asm.momentum_dev(abi.common_momentum_opts())
    assert code(asm.cmc_dev) == abi.ESTATE                         # ct_m was not assembled
    f2, c2 = asm.cmc_get_sparsity()
    bad = c2.copy()
    bad[0], bad[1] = bad[1], bad[0]
    assert code(asm.cmc_set_sparsity, f2, bad) == abi.EARG         # unsorted row
    assert code(asm.cmc_set_sparsity, f2[:-1], c2) == abi.EARG     # wrong number of rows
