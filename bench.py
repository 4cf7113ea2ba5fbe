#!/usr/bin/env python
"""bench.py -- CG momentum + tracer element assembly throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]

A step = one pass of the hot path over the mesh: the momentum element loop
(assemble/Momentum_CG.F90:726-752) followed by the tracer element loop
(assemble/Advection_Diffusion_CG.F90:574-598) with the common option set of the four example
configs. Workload at N=1: S3 = 256^3 x 6 = 100 663 296 Kuhn tets; at N>1
every rank holds its own 256^3-cell slab partition (weak scaling), exchanges the halo of
nu / oldu / T / density / buoyancy with NCCL p2p (cgasm_halo_update) and assembles all local
elements, exactly like the reference's MPI ranks.

value  = elements assembled by all ranks / device time, inputs resident in HBM.
e2e    = same through the host-buffer C-ABI calls cgasm_set_field / cgasm_momentum /
         cgasm_advdiff: per step the changing fields go host->device from pinned memory
         and every assembled array comes back device->host.
--impl reference times the CPU restatement of the reference loops (oracle/, OpenMP over the
reference's own colouring, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import resource
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CG momentum+tracer assembly Melements/s (whole job)"
UNIT = "Melements/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--cells", type=int, default=256, help="cells per axis per GPU (256 = S3)")
    ap.add_argument("--scatter", default="best", choices=["best", "atomic", "coloured", "warpagg", "tiled", "gather", "strip"])
    ap.add_argument("--cpu-cells", type=int, default=64, help="cells per axis of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---- clocks sampling (B200_PROFILING.md) ------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region. The S3 step is ~8 ms, so the timed region is
    shorter than one `nvidia-smi -lms` period: NVML is polled from a thread every few ms instead
    (nvidia_ml_py), with the nvidia-smi loop of the profiling recipe as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.smax = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.003)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml:
            self._stop.set()
            self.thread.join(timeout=1)
            nv = self.nvml
            masks = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [a for a, _ in self.samples]
            reasons = sorted(nm for nm, m in masks.items() if any(rs & m for _, rs in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 3 ms poll during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ---- reference arm / cpu baseline ------------------------------------------------------------------
def cpu_assembly_rate(cells, steps, warmup):
    """Times the oracle port (OpenMP, reference colouring) on a cells^3 x 6 sample.
    Returns (Melements/s, threads, ms_per_step, n_elements)."""
    from fluidity_b200 import synthetic as syn, _abi as abi
    from oracle import oracle as orc
    mesh = syn.box_mesh((cells,) * 3)
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    col, nc = orc.colour_elements(mesh)
    sets = orc.colour_sets(col, nc)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    threads = len(os.sched_getaffinity(0))
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.assemble_momentum(mesh, fs, om, findrm, colm, colouring=sets)
        orc.assemble_advdiff(mesh, fs, oa, findrm, colm, colouring=sets)
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    ms = 1e3 * float(np.mean(times))
    return mesh.n_elements / (ms * 1e-3) / 1e6, threads, ms, mesh.n_elements


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone with all the
    # host threads it can use (set before the OpenMP runtime of the oracle library starts)
    if "WORLD_SIZE" in os.environ and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    cells = args.cpu_cells
    rate, threads, ms, ne = cpu_assembly_rate(cells, max(1, args.steps), max(0, min(args.warmup, 1)))
    sample = "%d^3 x 6 = %d Kuhn tets per step (bounded sample of S3), OpenMP over the reference greedy colouring" % (cells, ne)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Fortran+PETSc and cannot be built in this image; this is the C restatement (oracle/) of its element loops",
    }
    emit(line)
    return 0


def workload_config(args, n):
    c = args.cells
    return {"workload": "S3: %d^3 x 6 Kuhn tets per GPU (%d elements/GPU), P1, degree-3 quadrature, "
                        "momentum (lumped mass, advection, isotropic viscosity, buoyancy, inverse lumped mass) "
                        "+ tracer (consistent mass, advection, isotropic diffusivity)" % (c, 6 * c ** 3),
            "cells_per_axis_per_gpu": c, "partition": "slab along z, one partition per GPU, L1+L2 halos" if n > 1 else "single partition",
            "l2_policy": "inputs+outputs per step (>8 GB at S3) exceed the 126 MB L2; no explicit flush",
            "parallelism": "dp%d" % n}


# ---- graft arm --------------------------------------------------------------------------------------
def run_graft(args):
    import torch
    import torch.distributed as dist
    from fluidity_b200 import _abi as abi, cgasm, tables, partition as part, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcgasm has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # torchrun pins OMP_NUM_THREADS=1; the once-per-mesh host preprocessing (sparsity, plans) is
        # OpenMP code, so give every rank its share of the host cores
        if os.environ.get("OMP_NUM_THREADS", "1") == "1":
            share = max(1, len(os.sched_getaffinity(0)) // world)
            os.environ["OMP_NUM_THREADS"] = str(share)
            # libcgasm.so resolves libgomp.so.1 to the copy torch bundles, which read OMP_NUM_THREADS=1 when torch
            # was imported: the environment alone no longer reaches it (measured: 390 s of single-threaded plan
            # building per rank at N=2 against 45 s at N=1), so set the thread count through the runtime as well
            torch.set_num_threads(share)

    c = args.cells
    t_setup = time.perf_counter()
    lp = part.slab_partition((c, c, c * world), world, rank)
    mesh = lp.mesh
    F = part.global_nodal_fields(3, mesh.X, lp.global_node)
    asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=local_rank)
    nnz = asm.build_sparsity()
    g = np.zeros((1, 3)); g[0, 2] = -1.0
    asm.set_field(abi.F_GRAVITY, g, abi.FIELD_CONSTANT)
    asm.set_field(abi.F_VISCOSITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_T_DIFFUSIVITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    dyn = [(abi.F_NU, F["nu"]), (abi.F_OLDU, F["oldu"]), (abi.F_DENSITY, F["density"]),
           (abi.F_BUOYANCY, F["buoyancy"]), (abi.F_T, F["t"])]
    if world > 1:
        # ranks only know their owned values; the halo arrives through cgasm_halo_update
        for _, a in dyn:
            a[lp.n_owned:] = 0.0
    for slot, a in dyn:
        asm.set_field(slot, a)
    halo_slots = [s for s, _ in dyn]
    if world > 1:
        uid = [cgasm.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        asm.halo_create(world, rank, lp.sends, lp.recvs, uid[0])
    scatter = {"atomic": abi.SCATTER_ATOMIC, "coloured": abi.SCATTER_COLOURED, "warpagg": abi.SCATTER_WARPAGG,
               "tiled": abi.SCATTER_TILED, "gather": abi.SCATTER_GATHER, "strip": abi.SCATTER_STRIP}
    chosen = args.scatter
    if chosen == "best":
        try:
            asm.set_scatter(abi.SCATTER_STRIP)
            chosen = "strip"
        except cgasm.CgasmError:
            asm.set_scatter(abi.SCATTER_ATOMIC)
            chosen = "atomic"
    else:
        asm.set_scatter(scatter[chosen])
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    setup_s = time.perf_counter() - t_setup
    # the library holds its own copies; drop the generator's arrays before the pinned e2e buffers
    n_el_local, n_nodes_local, n_owned = mesh.n_elements, mesh.n_nodes, lp.n_owned
    n_sent = int(sum(len(s) for s in lp.sends))
    mesh.ndglno = None
    mesh.X = None
    lp.global_node = lp.global_element = None
    import gc
    gc.collect()

    stream = torch.cuda.ExternalStream(asm.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world > 1:
            asm.halo_update(halo_slots)
        asm.momentum_dev(om)
        asm.advdiff_dev(oa)

    # ---- value: inputs resident in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = asm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mom_ms, adv_ms = [], []
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = asm.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device times (separate pass so the event queries do not perturb the timed region)
    halo_ms = None
    if world > 1:
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0.record(stream)
        for _ in range(5):
            asm.halo_update(halo_slots)
        h1.record(stream)
        barrier()
        halo_ms = h0.elapsed_time(h1) / 5
    for _ in range(max(3, min(args.steps, 5))):
        if world > 1:
            asm.halo_update(halo_slots)
        asm.momentum_dev(om)
        mom_ms.append(asm.last_kernel_ms())
        asm.advdiff_dev(oa)
        adv_ms.append(asm.last_kernel_ms())
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    tot_el = torch.tensor([float(n_el_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_el, op=dist.ReduceOp.SUM)
    total_elements = float(tot_el.item())
    value = total_elements / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers through the C-ABI calls --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(shape):
            return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        host_in = []
        for slot, a in dyn:
            b = pinned(a.shape)
            b[...] = a
            host_in.append((slot, b))
        # the common option set has no absorption: the 3 diagonal blocks are identical and the library
        # says so (cgasm_momentum_identical_blocks), so ONE block crosses PCIe and the shim inserts it 3x
        out_m = dict(big_m=pinned((1, nnz)), rhs=pinned((n_nodes_local, 3)), masslump=pinned((n_nodes_local, 3)))
        out_a = dict(matrix=pinned((nnz,)), rhs=pinned((n_nodes_local,)))
        h2d = sum(b.nbytes for _, b in host_in)
        d2h = sum(v.nbytes for v in out_m.values()) + sum(v.nbytes for v in out_a.values())

        # Asynchronous host flavour (cgasm_set_async), in the order of Fluidity's time step (scalar fields
        # before momentum): the tracer's inputs go up, its loop runs and its matrix starts down the copy
        # stream while the momentum inputs come up the other direction of the link and the momentum loop
        # runs; one synchronize at the end of the step, as the solves need both systems on the host.
        by_slot = dict(host_in)
        tracer_in = [abi.F_NU, abi.F_T]
        momentum_in = [sl for sl, _ in host_in if sl not in tracer_in]

        def step_e2e():
            for sl in tracer_in:
                asm.set_field(sl, by_slot[sl])
            if world > 1:
                asm.halo_update(tracer_in)
            asm.advdiff_dev(oa)
            asm.advdiff_fetch_into(out_a)
            for sl in momentum_in:
                asm.set_field(sl, by_slot[sl])
            if world > 1:
                asm.halo_update(momentum_in)
            nb = asm.momentum_host(om, out_m)
            assert nb == 1
            asm.synchronize()

        asm.set_async(True)
        n_e2e = max(2, min(args.steps, 3))
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_e2e()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        asm.set_async(False)
        e2e = {"value": total_elements / float(dt.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(dt.item()), "steps": n_e2e,
               "checksum": float(out_a["rhs"][:n_owned].sum())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (momentum) -------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bytes_mom = syn.algorithmic_bytes(3, n_nodes_local, n_el_local, nnz, "momentum")
    bytes_tra = syn.algorithmic_bytes(3, n_nodes_local, n_el_local, nnz, "tracer")
    m_ms, a_ms = float(np.mean(mom_ms)), float(np.mean(adv_ms))
    ach = bytes_mom * n_el_local / (m_ms * 1e-3) / 1e9
    # measured DRAM bytes per launch of the same kernel (ncu capture of this command, profiles/)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f)["kernels"]["momentum"]
        if chosen == "strip" and world == 1 and tr["elements"] == n_el_local:
            traffic = tr["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "momentum assembly (%s scatter)" % chosen, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_mom * n_el_local,
                "algorithmic_bytes_per_element": bytes_mom, "kernel_ms": m_ms,
                "frac_of_8TBs": ach / 8000.0,
                "tracer": {"achieved": bytes_tra * n_el_local / (a_ms * 1e-3) / 1e9,
                           "algorithmic_bytes_per_element": bytes_tra, "kernel_ms": a_ms},
                "combined_frac": (bytes_mom + bytes_tra) * n_el_local / ((m_ms + a_ms) * 1e-3) / 1e9 / peak}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, threads, ms, ne = cpu_assembly_rate(args.cpu_cells, 3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d^3 x 6 = %d Kuhn tets, 3 timed passes of momentum+tracer, OpenMP over the reference colouring (%.0f ms/pass)" % (args.cpu_cells, ne, ms)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "per_gpu_melements_s": value / world, "scatter": chosen, "setup_s": setup_s,
        "elements_total": total_elements, "nnz_rank0": nnz, "n_nodes_rank0": n_nodes_local,
        "host_peak_rss_gb_rank0": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "halo_update_ms_rank0": halo_ms, "halo_nodes_sent_rank0": n_sent,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # stdout carries exactly one JSON line. Native libraries write there too (NCCL prints its version banner
    # when the first communicator is made), so fd 1 points at stderr for the duration of the run and the
    # JSON line goes to a duplicate of the original stdout.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_graft(args)


if __name__ == "__main__":
    sys.exit(main())
