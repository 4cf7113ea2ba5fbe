#!/usr/bin/env python
"""bench.py -- CG momentum + tracer element assembly throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]

A step = one pass of the hot path over the mesh: the momentum element loop
(assemble/Momentum_CG.F90:726-752) followed by the tracer element loop
(assemble/Advection_Diffusion_CG.F90:574-598) with the common option set of the four example
configs. Workload: S3 = 256^3 x 6 = 100 663 296 Kuhn tets. At N>1 the default is STRONG
scaling, BASELINE.json's configs[4]: the ONE 100 M-tet mesh is decomposed into N node blocks
(1x1x2, 1x2x2, 2x2x2) with the reference's fldecomp conventions (partition.block_partition: owned nodes +
level-1/2 halos, trailing receives), one partition per GPU; every rank exchanges the halo of
nu / oldu / T / density / buoyancy with NCCL p2p (cgasm_halo_update, overlapped with the row blocks that read
no received node) and assembles ALL its local elements, like the reference's MPI ranks. `--scaling weak`
keeps round 1's layout (every rank its own 256^3-cell slab).

value  = OWNED elements (= the mesh's elements) / device time (max over ranks), inputs resident in HBM.
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
    ap.add_argument("--cpu-cells", type=int, default=128, help="cells per axis of the CPU-baseline sample")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1: strong = one cells^3 mesh split into N blocks (default); weak = cells^3 per GPU")
    ap.add_argument("--no-overlap", action="store_true", help="halo exchange in front of the kernels instead of beside them")
    ap.add_argument("--step", default="separate", choices=["separate", "fused"],
                    help="timed step = cgasm_momentum_dev + cgasm_advdiff_dev (the reference's two loops, default) or the "
                         "one-call cgasm_momentum_advdiff_dev (one fused kernel); the other one is reported beside it")
    ap.add_argument("--no-configs", action="store_true", help="skip the example-config option sets (N=1 only)")
    ap.add_argument("--delaunay-points", type=int, default=1500000,
                    help="points of the unstructured (scipy Delaunay) mesh of the configs block; 0 = skip it")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity check against the oracle (N>1)")
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

    def mark(self):
        """Forget what was sampled so far (warm-up): only the timed region is reported."""
        self.samples = []
        self.lines = []

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
def cpu_assembly_rate(cells, steps, warmup, threads=None):
    """Times the oracle port (OpenMP, reference colouring) on a cells^3 x 6 sample with the TIMING build of the oracle
    (oracle/liborc_fast.so: -O3 -march=native, FMA contraction allowed, compiled on this host; parity tests keep the
    strict -ffp-contract=off build). Returns (Melements/s, threads, ms_per_step, n_elements)."""
    from fluidity_b200 import synthetic as syn, _abi as abi
    from oracle import oracle as orc
    orc.select("fast")
    mesh = syn.box_mesh((cells,) * 3)
    fs = syn.standard_fields(mesh)
    findrm, colm, _ = orc.make_sparsity(mesh)
    col, nc = orc.colour_elements(mesh)
    sets = orc.colour_sets(col, nc)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    avail = len(os.sched_getaffinity(0))
    threads = avail if threads is None else threads
    orc.set_threads(threads)
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


def cpu_baseline_block(cells):
    """All-core and single-thread rates of the CPU restatement on bounded samples of S3."""
    rate, threads, ms, ne = cpu_assembly_rate(cells, 3, 1)
    c1 = max(16, cells // 2)
    r1, _, ms1, ne1 = cpu_assembly_rate(c1, 1, 0, threads=1)
    # the oracle and libcgasm share one OpenMP runtime: give the threads back, or every host set-up that follows
    # (the configs block) runs single-threaded -- it did until GPU call 48 of round 2 found it (library_setup_s 4-17 s
    # where the same calls take 0.5-2 s)
    from oracle import oracle as orc
    orc.set_threads(len(os.sched_getaffinity(0)))
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "build": "oracle/liborc_fast.so: gcc -O3 -march=native -fopenmp (FMA contraction on), built on this host",
            "sample": "%d^3 x 6 = %d Kuhn tets, 3 timed passes of momentum+tracer, OpenMP over the reference colouring "
                      "(%.0f ms/pass)" % (cells, ne, ms),
            "single_thread": {"value": r1, "unit": UNIT, "cores": 1,
                              "sample": "%d^3 x 6 = %d Kuhn tets, 1 pass (%.0f ms)" % (c1, ne1, ms1)}}


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
        "scaling": scaling_of(args, args.gpus), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "build": "oracle/liborc_fast.so: gcc -O3 -march=native -fopenmp, built on this host"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Fortran+PETSc and cannot be built in this image; this is the C restatement (oracle/) of its element loops",
    }
    emit(line)
    return 0


def scaling_of(args, n):
    """"strong": the total work (one cells^3 mesh) is fixed as N grows; "weak": cells^3 per GPU."""
    return args.scaling


def workload_config(args, n):
    c = args.cells
    strong = args.scaling == "strong" or n == 1
    from fluidity_b200 import partition as part
    what = ("S3: ONE %d^3 x 6 Kuhn tet mesh (%d elements), " % (c, 6 * c ** 3)) if strong else \
           ("S3 per GPU: %d^3 x 6 Kuhn tets per GPU (%d elements/GPU), " % (c, 6 * c ** 3))
    if n == 1:
        partition = "single partition"
    elif strong:
        partition = ("node blocks %s (x, y, z), one partition per GPU, fldecomp conventions: owned + level-1/2 halo nodes, "
                     "every element with an owned or level-1 node" % "x".join(str(p) for p in part.block_grid(n)))
    else:
        partition = "slab along z, one partition per GPU, L1+L2 halos"
    return {"workload": what + "P1, degree-3 quadrature, momentum (lumped mass, advection, isotropic viscosity, buoyancy, "
                        "inverse lumped mass) + tracer (consistent mass, advection, isotropic diffusivity)",
            "cells_per_axis": c, "partition": partition,
            "l2_policy": "inputs+outputs per step (>1 GB per GPU) exceed the 126 MB L2; no explicit flush",
            "parallelism": "dp%d" % n}


# ---- graft arm --------------------------------------------------------------------------------------
def bind_to_gpu_numa(index):
    """Pins this rank's host threads to the CPUs of the NUMA node its GPU hangs off (sysfs numa_node of the GPU's PCI
    function), BEFORE the pinned staging buffers of the e2e leg are allocated and first touched. Round 1's e2e
    collapsed with N (918 -> 218 Melements/s per GPU at N = 8): every rank's pinned memory sat on NUMA node 0 and all
    host<->device traffic crossed one socket's memory controllers and the inter-socket link. Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = bus.lower()
        if len(dev.split(":")[0]) == 8:        # NVML prints an 8-digit domain, sysfs a 4-digit one
            dev = dev[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % dev) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False, "why": "the platform reports no NUMA affinity for the GPU"}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {"numa_node": node, "bound": False, "why": "none of the node's CPUs is in this process's affinity mask"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": True, "cpus": len(allowed)}
    except Exception as exc:  # no NVML / sysfs: run unbound
        return {"bound": False, "why": "%s: %s" % (type(exc).__name__, exc)}


def setup_fields(asm, abi, syn, F, n_owned, world):
    g = np.zeros((1, 3)); g[0, 2] = -1.0
    asm.set_field(abi.F_GRAVITY, g, abi.FIELD_CONSTANT)
    asm.set_field(abi.F_VISCOSITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    asm.set_field(abi.F_T_DIFFUSIVITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    dyn = [(abi.F_NU, F["nu"]), (abi.F_OLDU, F["oldu"]), (abi.F_DENSITY, F["density"]),
           (abi.F_BUOYANCY, F["buoyancy"]), (abi.F_T, F["t"])]
    if world > 1:
        # ranks only know their owned values; the halo arrives through cgasm_halo_update
        for _, a in dyn:
            a[n_owned:] = 0.0
    for slot, a in dyn:
        asm.set_field(slot, a)
    return dyn


def multi_gpu_parity(args, world, rank, local_rank, dist, overlap):
    """N>1 correctness inside the bench (the GPU test lease has one GPU): a small box, 24 cells per block and axis,
    decomposed exactly like the timed mesh; every rank assembles its partition (halo values arrive through
    cgasm_halo_update only), rank 0 assembles the GLOBAL mesh with the CPU oracle and compares every owned row.
    Returns the largest relative error over big_m, rhs, masslump, tracer matrix and rhs (rank 0), else None."""
    from fluidity_b200 import _abi as abi, cgasm, tables, partition as part, synthetic as syn
    pg = part.block_grid(world)
    shape = tuple(24 * p for p in pg)
    lp = part.block_partition(shape, pg, rank)
    mesh = lp.mesh
    F = part.global_nodal_fields(3, mesh.X, lp.global_node)
    asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=local_rank)
    asm.build_sparsity()
    dyn = setup_fields(asm, abi, syn, F, lp.n_owned, world)
    uid = [cgasm.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.halo_create(world, rank, lp.sends, lp.recvs, uid[0])
    asm.halo_set_overlap(overlap)
    asm.set_scatter(abi.SCATTER_STRIP)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    asm.halo_update([s for s, _ in dyn])
    got_m = asm.momentum(om)
    got_a = asm.advdiff(oa)
    findrm, colm, _ = asm.get_sparsity()
    no = lp.n_owned
    g = lp.global_node
    rows = np.repeat(np.arange(mesh.n_nodes), np.diff(findrm))
    keep = rows < no
    nn_glob = int(np.prod([c + 1 for c in shape]))
    key = g[rows[keep]] * nn_glob + g[colm[keep] - 1]
    mine = dict(key=key, big_m=got_m["big_m"][0][keep], matrix=got_a["matrix"][keep], gnode=g[:no],
                rhs=got_m["rhs"][:no], masslump=got_m["masslump"][:no], arhs=got_a["rhs"][:no])
    asm.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    from oracle import oracle as orc
    whole = part.block_partition(shape, (1, 1, 1), 0)
    gm = whole.mesh
    GF = part.global_nodal_fields(3, gm.X, whole.global_node)
    fs = syn.standard_fields(gm)
    for slot, name in ((abi.F_NU, "nu"), (abi.F_OLDU, "oldu"), (abi.F_DENSITY, "density"), (abi.F_BUOYANCY, "buoyancy"),
                       (abi.F_T, "t")):
        fs.set(slot, GF[name])
    gvec = np.zeros((1, 3)); gvec[0, 2] = -1.0
    fs.set(abi.F_GRAVITY, gvec, abi.FIELD_CONSTANT)
    fs.set(abi.F_VISCOSITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    fs.set(abi.F_T_DIFFUSIVITY, syn.iso_tensor(3, 1e-3), abi.FIELD_CONSTANT)
    gf, gc, _ = orc.make_sparsity(gm)
    ref_m = orc.assemble_momentum(gm, fs, om, gf, gc)
    ref_a = orc.assemble_advdiff(gm, fs, oa, gf, gc)
    grows = np.repeat(np.arange(gm.n_nodes), np.diff(gf))
    gkey = grows.astype(np.int64) * nn_glob + (gc.astype(np.int64) - 1)   # ascending: rows ascending, columns sorted
    worst, seen = 0.0, 0
    for pr in parts:
        pos = np.searchsorted(gkey, pr["key"])
        assert (gkey[pos] == pr["key"]).all(), "a partition holds an entry the global sparsity lacks"
        seen += len(pr["gnode"])
        for got, ref in ((pr["big_m"], ref_m["big_m"][0][pos]), (pr["matrix"], ref_a["matrix"][pos]),
                         (pr["rhs"], ref_m["rhs"][pr["gnode"]]), (pr["masslump"], ref_m["masslump"][pr["gnode"]]),
                         (pr["arhs"], ref_a["rhs"][pr["gnode"]])):
            worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    assert seen == gm.n_nodes, "the partitions' owned nodes do not cover the mesh"
    return worst


DELAUNAY_CODE = """
import sys, numpy as np
sys.path.insert(0, %r)
from fluidity_b200 import synthetic as syn
m = syn.delaunay_mesh(int(sys.argv[1]))
np.save(sys.argv[2] + '.X.npy', m.X); np.save(sys.argv[2] + '.nd.npy', m.ndglno)
"""


def start_delaunay(points):
    """The unstructured leg of the configs block needs a scipy (qhull) Delaunay triangulation of `points` graded random
    points (~6.5 tets per point: 1.5 M points = 10 M tets, ~90 s on one core): built by a child process beside the
    S3 set-up and timing, which do not need that core."""
    import tempfile
    base = os.path.join(tempfile.mkdtemp(prefix="cgasm_delaunay_"), "mesh")
    proc = subprocess.Popen([sys.executable, "-c", DELAUNAY_CODE % ROOT, str(points), base],
                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return proc, base, time.perf_counter()


def example_configs(device, cells3, cells2, reps=5, delaunay=None, peak_gbs=None):
    """Kernel time of the two element loops for the option sets of BASELINE.json configs[0..3] (their meshes are not
    in the reference tree: option sets on synthetic meshes of the same dimension), plus the S3 option set on a randomly
    renumbered mesh. STRIP variant, CUDA events on the handle's stream."""
    import statistics
    from fluidity_b200 import synthetic as syn, _abi as abi, cgasm, tables
    cm, ca = abi.common_momentum_opts, abi.common_advdiff_opts
    cases = [
        ("driven_cavity: 2-D, nodal density, constant isotropic viscosity, no gravity", 2, cm(have_gravity=0), ca(), None, False),
        ("lock_exchange: 2-D, Boussinesq, nodal buoyancy, gravity; tracer", 2, cm(), ca(), "const_density", False),
        ("backward_facing_step_3d: 3-D, constant density, nodal vector absorption", 3,
         cm(have_absorption=1, have_gravity=0), ca(), "const_density", False),
        ("flow_past_sphere_Re100: 3-D, constant anisotropic viscosity", 3,
         cm(viscosity_shape=abi.TENSOR_FULL, have_gravity=0), ca(), "aniso", False),
        ("S3 option set, tracer with nodal absorption and source", 3, cm(), ca(have_absorption=1, have_source=1), None, False),
        ("S3 option set at this size", 3, cm(), ca(), None, False),
        ("S3 option set, nodes and elements randomly renumbered", 3, cm(), ca(), None, True),
        ("S3 option set with streamline-upwind stabilisation (nu_bar optimal) in both loops: per-element nu_bar at every "
         "quadrature point, so the element-owner two-pass path by design", 3,
         cm(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND), ca(stabilisation_scheme=abi.STAB_STREAMLINE_UPWIND), None, False),
    ]
    if delaunay is not None:
        cases.append(("flow_past_sphere_Re100 option set on an UNSTRUCTURED mesh: Delaunay triangulation of graded random "
                      "points (scipy/qhull), generator-like numbering", 3,
                      cm(viscosity_shape=abi.TENSOR_FULL, have_gravity=0), ca(), "aniso", "delaunay"))
    meshes, out = {}, []
    for name, dim, om, oa, tweak, shuffle in cases:
        keym = (dim, shuffle)
        if keym not in meshes:
            if shuffle == "delaunay":
                proc, base, t_start = delaunay
                try:
                    proc.wait(timeout=300)   # it has been running beside everything above; a slow host must not stall the line
                except subprocess.TimeoutExpired:
                    proc.kill()
                    out.append({"config": name, "error": "Delaunay triangulation not ready after %.0f s: skipped" % (time.perf_counter() - t_start)})
                    continue
                if proc.returncode != 0:
                    out.append({"config": name, "error": "Delaunay child process failed (%d)" % proc.returncode})
                    continue
                mesh = syn.Mesh(dim=3, ndglno=np.load(base + ".nd.npy"), X=np.load(base + ".X.npy"))
                for suffix in (".nd.npy", ".X.npy"):
                    os.remove(base + suffix)
            else:
                c = cells3 if dim == 3 else cells2
                mesh = syn.box_mesh((c,) * dim)
                if shuffle:
                    mesh = syn.shuffled(mesh)
            meshes[keym] = (mesh, syn.standard_fields(mesh))
        mesh, fs0 = meshes[keym]
        t0 = time.perf_counter()
        asm = cgasm.Assembler(mesh, tables.p1_tables(dim), device=device)
        asm.build_sparsity()
        asm.set_fields(fs0)
        if tweak == "const_density":
            asm.set_field(abi.F_DENSITY, np.ones(1), abi.FIELD_CONSTANT)
        if tweak == "aniso":
            asm.set_field(abi.F_VISCOSITY, syn.aniso_tensor(dim), abi.FIELD_CONSTANT)
        asm.set_scatter(abi.SCATTER_STRIP)
        setup = time.perf_counter() - t0
        mom, adv = [], []
        l0 = asm.launch_count()
        for i in range(reps + 2):
            asm.momentum_dev(om)
            m = asm.last_kernel_ms()
            asm.advdiff_dev(oa)
            a = asm.last_kernel_ms()
            if i >= 2:
                mom.append(m)
                adv.append(a)
        launches = (asm.launch_count() - l0) / (reps + 2)
        paths = asm.last_path()
        mm, aa = statistics.median(mom), statistics.median(adv)
        rec = {"config": name, "dim": dim, "elements": mesh.n_elements, "momentum_ms": mm, "tracer_ms": aa,
               "gel_s": mesh.n_elements / ((mm + aa) * 1e-3) / 1e9, "kernel_launches_per_step": launches,
               "momentum_path": paths[0], "tracer_path": paths[1],
               "library_setup_s": setup}
        if peak_gbs:
            # the same roofline as the headline's: algorithmic bytes of this mesh (synthetic.algorithmic_bytes) / kernel time / peak
            bm = syn.algorithmic_bytes(dim, mesh.n_nodes, mesh.n_elements, asm.nnz, "momentum")
            bt = syn.algorithmic_bytes(dim, mesh.n_nodes, mesh.n_elements, asm.nnz, "tracer")
            rec["roofline_frac"] = {"momentum": bm * mesh.n_elements / (mm * 1e-3) / 1e9 / peak_gbs,
                                    "tracer": bt * mesh.n_elements / (aa * 1e-3) / 1e9 / peak_gbs,
                                    "both": (bm + bt) * mesh.n_elements / ((mm + aa) * 1e-3) / 1e9 / peak_gbs,
                                    "algorithmic_bytes_per_element": [bm, bt]}
        if shuffle:
            nnz = asm.nnz
            rec.update({"nodes": mesh.n_nodes, "nnz": int(nnz), "mean_row_length": nnz / mesh.n_nodes,
                        "elements_per_node": mesh.n_elements * (dim + 1) / mesh.n_nodes,
                        "plan": asm.plan_stats()})
        out.append(rec)
        asm.close()
    return out


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
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"bound": False, "why": "single rank"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # torchrun pins OMP_NUM_THREADS=1; the once-per-mesh host preprocessing (sparsity, plans) is
        # OpenMP code, so give every rank its share of the host cores
        if os.environ.get("OMP_NUM_THREADS", "1") == "1":
            share = max(1, len(os.sched_getaffinity(0)) // (1 if numa.get("bound") else world))
            if numa.get("bound"):
                # the node's CPUs are shared by the ranks bound to it
                share = max(1, share // max(1, world // 2))
            os.environ["OMP_NUM_THREADS"] = str(share)
            # libcgasm.so resolves libgomp.so.1 to the copy torch bundles, which read OMP_NUM_THREADS=1 when torch
            # was imported: the environment alone no longer reaches it, so set the thread count through the runtime too
            torch.set_num_threads(share)
    strong = args.scaling == "strong" or world == 1
    overlap = world > 1 and not args.no_overlap
    delaunay = None
    if world == 1 and not args.no_configs and args.delaunay_points > 0:
        delaunay = start_delaunay(args.delaunay_points)

    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(args, world, rank, local_rank, dist, overlap)

    c = args.cells
    t_gen = time.perf_counter()
    if strong:
        lp = part.block_partition((c, c, c), part.block_grid(world), rank)
    else:
        lp = part.slab_partition((c, c, c * world), world, rank)
    mesh = lp.mesh
    F = part.global_nodal_fields(3, mesh.X, lp.global_node)
    mesh_gen_s = time.perf_counter() - t_gen
    t_setup = time.perf_counter()
    asm = cgasm.Assembler(mesh, tables.p1_tables(3), device=local_rank)
    nnz = asm.build_sparsity()
    dyn = setup_fields(asm, abi, syn, F, lp.n_owned, world)
    halo_slots = [s for s, _ in dyn]
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
    setup_s = time.perf_counter() - t_setup
    if world > 1:
        uid = [cgasm.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        asm.halo_create(world, rank, lp.sends, lp.recvs, uid[0])
        asm.halo_set_overlap(overlap)
    om, oa = abi.common_momentum_opts(), abi.common_advdiff_opts()
    # the library holds its own copies; drop the generator's arrays before the pinned e2e buffers
    n_el_local, n_nodes_local, n_owned = mesh.n_elements, mesh.n_nodes, lp.n_owned
    n_sent = int(sum(len(s) for s in lp.sends))
    n_neighbours = int(sum(1 for s in lp.sends if len(s)))
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

    marks = None   # inside the timed region: CUDA events on the handle's stream around each of the two calls of a step

    def step_resident():
        if world > 1:
            asm.halo_update(halo_slots)
        if args.step == "fused":
            asm.momentum_advdiff_dev(om, oa)
        elif marks is None:
            asm.momentum_dev(om)
            asm.advdiff_dev(oa)
        else:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            asm.momentum_dev(om)
            e[1].record(stream)
            asm.advdiff_dev(oa)
            e[2].record(stream)
            marks.append(e)

    # ---- value: inputs resident in HBM ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # before the barrier: NVML start-up on rank 0 must not skew the ranks' entry into the timed region
    for _ in range(args.warmup):
        step_resident()
    barrier()
    # one more untimed step AFTER the barrier + synchronize: the ranks leave the host barrier microseconds to
    # milliseconds apart, and the first halo exchange makes the early ones wait for the late ones -- that wait belongs
    # to the barrier, not to the K timed steps. Its exchange lines the ranks up on the device; ev0 follows on-stream.
    if world > 1:
        step_resident()
    l0 = asm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mom_ms, adv_ms = [], []
    if rank == 0:
        sampler.mark()
    marks = [] if args.step == "separate" else None
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    my_total_ms = ev0.elapsed_time(ev1)
    in_region = None
    if marks:
        # the roofline's kernel times: average launch duration INSIDE the K timed steps (events only, no host wait)
        in_region = (float(np.mean([e[0].elapsed_time(e[1]) for e in marks])), float(np.mean([e[1].elapsed_time(e[2]) for e in marks])))
    marks = None
    launches = asm.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device times (separate pass so the event queries do not perturb the timed region)
    halo_ms = None
    if world > 1:
        # the exchange alone (no kernels beside it): pack + NCCL group + unpack, joined back onto the compute stream
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0.record(stream)
        for _ in range(5):
            asm.halo_update(halo_slots)
            asm.synchronize()
        h1.record(stream)
        barrier()
        halo_ms = h0.elapsed_time(h1) / 5
    for _ in range(max(3, min(args.steps, 5))):
        if world > 1:
            asm.halo_update(halo_slots)
            asm.synchronize()
        asm.momentum_dev(om)
        mom_ms.append(asm.last_kernel_ms())
        asm.advdiff_dev(oa)
        adv_ms.append(asm.last_kernel_ms())
    # the one-call flavour of the step (one fused kernel for this option set), timed alone
    fused_ms = []
    for _ in range(max(3, min(args.steps, 5))):
        if world > 1:
            asm.halo_update(halo_slots)
            asm.synchronize()
        asm.momentum_advdiff_dev(om, oa)
        fused_ms.append(asm.last_kernel_ms())
    fused_ms = float(np.mean(fused_ms))
    t = torch.tensor([my_total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # units of work: the mesh's elements (strong: every element is owned by exactly one block's rows; the halo
    # elements a rank assembles on top are redundant work, not throughput) / weak: every rank's own mesh
    if strong:
        total_elements = float(6 * c ** 3)
    else:
        tot_el = torch.tensor([float(n_el_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot_el, op=dist.ReduceOp.SUM)
        total_elements = float(tot_el.item())
    value = total_elements / (ms_per_step * 1e-3) / 1e6
    # per-kernel times: from inside the timed region where the step is the two calls; else the separate pass (median:
    # the first launch after an idle gap is an outlier that a mean of five does not absorb)
    alone_ms = (float(np.median(mom_ms)), float(np.median(adv_ms)))
    if rank == 0:
        print("bench: kernels timed alone (ms): momentum %s  tracer %s" % (["%.3f" % x for x in mom_ms], ["%.3f" % x for x in adv_ms]), file=sys.stderr)
    m_ms, a_ms = in_region if in_region else alone_ms
    per_rank = None
    if world > 1:
        mine = dict(rank=rank, step_ms=my_total_ms / args.steps, momentum_ms=m_ms, tracer_ms=a_ms, fused_ms=fused_ms, halo_ms=halo_ms,
                    local_elements=n_el_local, local_nodes=n_nodes_local, owned_nodes=n_owned, halo_nodes_sent=n_sent,
                    neighbours=n_neighbours, library_setup_s=setup_s, mesh_gen_s=mesh_gen_s, numa=numa)
        allr = [None] * world if rank == 0 else None
        dist.gather_object(mine, allr, dst=0)
        per_rank = allr

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
        tb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e2e = {"value": total_elements / float(dt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()),
               "per_gpu_melements_s": total_elements / float(dt.item()) / 1e6 / world,
               "ms_per_step": 1e3 * float(dt.item()), "steps": n_e2e,
               "checksum": float(out_a["rhs"][:n_owned].sum())}
    asm.close()

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
    ach = bytes_mom * n_el_local / (m_ms * 1e-3) / 1e9
    # measured DRAM bytes per launch of the same kernel (ncu capture of this command, profiles/)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tr = json.load(f)["kernels"]["momentum"]
        if chosen == "strip" and world == 1 and tr["elements"] == n_el_local:
            traffic = tr["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "momentum assembly (%s scatter), rank 0's partition" % chosen, "achieved": ach,
                "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_mom * n_el_local,
                "algorithmic_bytes_per_element": bytes_mom, "kernel_ms": m_ms,
                "frac_of_8TBs": ach / 8000.0,
                "tracer": {"achieved": bytes_tra * n_el_local / (a_ms * 1e-3) / 1e9,
                           "frac": bytes_tra * n_el_local / (a_ms * 1e-3) / 1e9 / peak,
                           "algorithmic_bytes_per_element": bytes_tra, "kernel_ms": a_ms},
                "combined_frac": (bytes_mom + bytes_tra) * n_el_local / ((m_ms + a_ms) * 1e-3) / 1e9 / peak}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_block(args.cpu_cells)
    configs = None
    if world == 1 and not args.no_configs:
        configs = example_configs(local_rank, 128, 2048, delaunay=delaunay, peak_gbs=peak)
    elif delaunay is not None:
        delaunay[0].kill()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling_of(args, world), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "per_gpu_melements_s": value / world, "scatter": chosen, "setup_s": setup_s, "mesh_gen_s": mesh_gen_s,
        "elements_total": total_elements, "local_elements_rank0": n_el_local, "nnz_rank0": nnz, "n_nodes_rank0": n_nodes_local,
        "host_peak_rss_gb_rank0": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "step_flavour": args.step, "separate_kernels_ms_rank0": m_ms + a_ms, "fused_kernel_ms_rank0": fused_ms,
        "kernel_ms_source": ("CUDA events on the handle's stream around each call inside the K timed steps (mean)" if in_region
                             else "cgasm_last_kernel_ms of calls timed alone after the timed region (median)"),
        "kernels_timed_alone_ms_rank0": list(alone_ms),
        "halo_update_ms_rank0": halo_ms, "halo_nodes_sent_rank0": n_sent, "halo_overlap": overlap,
        "multi_gpu_parity_max_rel_err": parity, "per_rank": per_rank, "configs": configs,
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
