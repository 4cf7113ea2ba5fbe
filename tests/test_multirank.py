"""N>1 path. CPU: two real processes over gloo exchange the slab-partition halos with the same
send/receive lists the NCCL path uses (host logic of SURVEY.md 8(e)). GPU (needs >= 2 devices):
two ranks run cgasm_halo_update over NCCL and assemble; owned rows must equal the oracle's
global assembly."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

CELLS = (4, 3, 12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fields_for(lp):
    from fluidity_b200 import partition as part
    return part.global_nodal_fields(3, lp.mesh.X, lp.global_node)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from fluidity_b200 import partition as part
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp = part.slab_partition(CELLS, world, rank)
    F = _fields_for(lp)
    ok = True
    for name in ("nu", "t"):
        want = F[name].copy()
        have = want.copy()
        have[lp.n_owned:] = -777.0  # only owned values are known before the halo update
        flat = have.reshape(have.shape[0], -1)
        reqs, bufs = [], []
        for p in range(world):
            if len(lp.sends[p]):
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(flat[lp.sends[p] - 1])), dst=p))
            if len(lp.recvs[p]):
                b = torch.empty((len(lp.recvs[p]), flat.shape[1]), dtype=torch.float64)
                bufs.append((p, b))
                reqs.append(dist.irecv(b, src=p))
        for r in reqs:
            r.wait()
        for p, b in bufs:
            flat[lp.recvs[p] - 1] = b.numpy()
        ok = ok and bool((have == want).all())
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == world
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_lists_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res)


def _decomposed_worker(rank, world, port, q, base):
    """Every rank reads ITS files of a decomposition in the reference's formats (<base>_<rank>.msh + .halo,
    fluidity_b200/formats.py) and runs the halo update of a coordinate-derived field with those lists."""
    sys.path.insert(0, ROOT)
    from fluidity_b200 import formats as fmt
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp, gm, hs = fmt.read_decomposition(base, rank)
    want = np.stack([np.sin(3 * lp.mesh.X[:, 0]) + lp.mesh.X[:, 1] ** 2, np.cos(2 * lp.mesh.X[:, -1])], axis=1)
    have = want.copy()
    have[lp.n_owned:] = -777.0
    reqs, bufs = [], []
    for p in range(world):
        if len(lp.sends[p]):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(have[lp.sends[p] - 1])), dst=p))
        if len(lp.recvs[p]):
            b = torch.empty((len(lp.recvs[p]), 2), dtype=torch.float64)
            bufs.append((p, b))
            reqs.append(dist.irecv(b, src=p))
    for r in reqs:
        r.wait()
    for p, b in bufs:
        have[lp.recvs[p] - 1] = b.numpy()
    ok = bool((have == want).all()) and hs.nprocs == world and sorted(hs.levels) == [1, 2]
    # owned-node counts add up to the global mesh
    t = torch.tensor([float(lp.n_owned)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    q.put((rank, ok, int(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rcb_decomposition_through_the_reference_formats_over_gloo(world, tmp_path):
    sys.path.insert(0, ROOT)
    from conftest import load_golden_mesh
    from fluidity_b200 import formats as fmt, partition as part
    mesh = load_golden_mesh("cube-parallel")  # unstructured gmsh mesh from the reference's tests/data
    parts = part.partition_by_owner(mesh, part.rcb_owner(mesh.X, world), world)
    base = str(tmp_path / "cube")
    fmt.write_decomposition(base, parts, binary=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_decomposed_worker, args=(r, world, port, q, base)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _, _ in res) == list(range(world))
    assert all(ok for _, ok, _ in res) and all(n == mesh.n_nodes for _, _, n in res)


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from fluidity_b200 import partition as part, cgasm, tables, _abi as abi, synthetic as syn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp = part.slab_partition(CELLS, world, rank)
    F = _fields_for(lp)
    asm = cgasm.Assembler(lp.mesh, tables.p1_tables(3), device=rank)
    asm.build_sparsity()
    g = np.zeros((1, 3)); g[0, 2] = -1.0
    asm.set_field(abi.F_GRAVITY, g, abi.FIELD_CONSTANT)
    asm.set_field(abi.F_VISCOSITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_T_DIFFUSIVITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    slots = [(abi.F_NU, "nu"), (abi.F_OLDU, "oldu"), (abi.F_DENSITY, "density"), (abi.F_BUOYANCY, "buoyancy"), (abi.F_T, "t")]
    for s, name in slots:
        a = F[name].copy()
        a[lp.n_owned:] = 0.0
        asm.set_field(s, a)
    uid = [cgasm.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.halo_create(world, rank, lp.sends, lp.recvs, uid[0])
    asm.halo_update([s for s, _ in slots])
    ok = True
    for s, name in slots:
        got = asm.get_field(s, F[name].shape)
        ok = ok and bool((got == F[name]).all())
    asm.set_scatter(abi.SCATTER_STRIP)  # after the halo update: the received nodes' records were repacked
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    out_m, out_a = asm.momentum(om), asm.advdiff(oa)
    # the same step with the exchange overlapped (cgasm_halo_set_overlap): halo values zeroed again, the update runs
    # on its own stream beside the row blocks that read no received node -- results must be bitwise the same
    asm.halo_set_overlap(True)
    for s, name in slots:
        a = F[name].copy()
        a[lp.n_owned:] = 0.0
        asm.set_field(s, a)
    asm.halo_update([s for s, _ in slots])
    o2m, o2a = asm.momentum(om), asm.advdiff(oa)
    for k in ("big_m", "rhs", "masslump"):
        ok = ok and bool((o2m[k] == out_m[k]).all())
    for k in ("matrix", "rhs"):
        ok = ok and bool((o2a[k] == out_a[k]).all())
    findrm, colm, _ = asm.get_sparsity()
    q.put((rank, ok, lp.n_owned, lp.global_node, findrm, colm, out_m["big_m"], out_m["rhs"], out_m["masslump"],
           out_a["matrix"], out_a["rhs"]))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_halo_update_and_owned_rows(orc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from fluidity_b200 import partition as part, _abi as abi, synthetic as syn
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    # global reference: the nprocs=1 slab is the whole mesh
    g = part.slab_partition(CELLS, 1, 0)
    F = part.global_nodal_fields(3, g.mesh.X, g.global_node)
    fs = syn.standard_fields(g.mesh)
    for s, name in ((abi.F_NU, "nu"), (abi.F_OLDU, "oldu"), (abi.F_DENSITY, "density"), (abi.F_BUOYANCY, "buoyancy"), (abi.F_T, "t")):
        fs.set(s, F[name])
    gf, gc, _ = orc.make_sparsity(g.mesh)
    gm = orc.assemble_momentum(g.mesh, fs, abi.common_momentum_opts(), gf, gc)
    ga = orc.assemble_advdiff(g.mesh, fs, abi.common_advdiff_opts(), gf, gc)
    for rank, ok, n_owned, gnode, findrm, colm, big_m, rhs, ml, mat, arhs in res:
        assert ok, "halo_update did not reproduce the owners' values, or the overlapped step differs from the plain one"
        own = gnode[:n_owned]
        assert np.abs(rhs[:n_owned] - gm["rhs"][own]).max() <= 1e-12 * np.abs(gm["rhs"]).max()
        assert np.abs(ml[:n_owned] - gm["masslump"][own]).max() <= 1e-12 * np.abs(gm["masslump"]).max()
        assert np.abs(arhs[:n_owned] - ga["rhs"][own]).max() <= 1e-12 * np.abs(ga["rhs"]).max()
        for i in range(0, n_owned, 7):
            gi = own[i]
            lrow = slice(findrm[i] - 1, findrm[i + 1] - 1)
            grow = slice(gf[gi] - 1, gf[gi + 1] - 1)
            perm = np.argsort(gnode[colm[lrow] - 1])
            assert (gnode[colm[lrow] - 1][perm] == gc[grow] - 1).all()
            for d in range(3):
                assert np.abs(big_m[d][lrow][perm] - gm["big_m"][d][grow]).max() <= 1e-12 * np.abs(gm["big_m"][d]).max()
            assert np.abs(mat[lrow][perm] - ga["matrix"][grow]).max() <= 1e-12 * np.abs(ga["matrix"]).max()
