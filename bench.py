#!/usr/bin/env python
"""bench.py -- elements/sec of one `makeϕrKt(elems, u)` evaluation (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                 our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W         the reference ALGORITHM on the host cores (see below)
  python bench.py --config C1|C2|C3|C4|C5                       the other configurations of BASELINE.json (default C2)
  python bench.py --gpus N --scaling strong                     the SAME 1M-element mesh split over N GPUs

Default workload: config C2 of BASELINE.json (the one `metric` is quoted on) -- 3-D Hex8 NeoHooke(1000, 5000) structured
block, 2x2x2 Gauss, u = 0.02·h·U(-1,1), 100x100x100 (1M elements).  N > 1: z-slab partition, interface rows exchanged
over NVLink (dist.py); `--scaling weak` (default) gives every rank a 100x100x100 slab of a 100x100x(100·N) block,
`--scaling strong` splits the one 100x100x100 block into N slabs (12/13 layers each at N = 8).
A "step" is one full evaluation: u on device -> ϕ, r, CSC(=CSR) values of Kt on device (C4: one staggered pair, the
u-phase then the d-phase assembly with its history update).

The reference is pure Julia and Julia is not installed in this image, so `--impl reference` times the oracle's C++
restatement of the reference algorithm (threaded per-element D2 duals + serial COO->CSC assembly, oracle/cxx) on a
bounded sample of the same workload, with all host threads; it is labelled "port", never "Julia".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "elements/s"

# algorithmic work per element.  flop: canonical generic-tangent count.  Bytes: node ids (int32) + ∇N +
# wgt (+ N values) + u gather (each node once) | Kt values written once + r (+ history read/write).  Staging, scatter
# maps and column indices are implementation overhead and are NOT counted.
CONFIGS = {
    "C1": dict(name="2-D Quad4 Hooke2D plane strain 200x200", shape="quad4", n=(200, 200), flop=1608.0,
               b_elem=16 + 256 + 32 + 16.2, b_asm=289.0 + 16.2, sample=(200, 200),
               k1="k1_u_kernel<2,4,4,HOOKE>", asm="k3_kernel + k2_kernel<2>"),
    "C2": dict(name="3-D Hex8 NeoHooke", shape="hex8", n=(100, 100, 100), flop=31424.0,
               b_elem=32 + 1536 + 64 + 24.7, b_asm=1963.5 + 24.7, sample=(56, 56, 56),
               k1="k1_u_kernel<3,8,8,NEO>", asm="k3_kernel + k2s_tma_kernel3"),
    "C3": dict(name="3-D Tet4 MooneyRivlin Kuhn lattice 70x70x68", shape="tet4", n=(70, 70, 68), flop=2032.0,
               b_elem=16 + 96 + 8 + 4.2, b_asm=183.7 + 4.2, sample=(33, 33, 32),
               k1="k1_u_kernel_w1<3,4,MOONEY> (thread per element, warp-cooperative sorted stores)", asm="k3_kernel + k2s_tma_kernel3"),
    "C4": dict(name="2-D QuadP PhaseField(NeoHooke) 500x500, staggered u/d assembly with history", shape="quadp",
               n=(500, 500), flop=4 * (32 + 32 + 288 + 400 + 60) + 800.0,
               b_elem=2 * (16 + 256 + 32 + 128 + 16.2 + 8) + 256, b_asm=(288.4 + 16.1) + (72.1 + 8.0), sample=(200, 200),
               k1="k1_u_kernel<2,4,4,PFNEO,PF> + k1_d_kernel<2,4,4>", asm="k3_kernel + k2_kernel<2> / <1> (u- and d-phase)"),
    "C5": dict(name="3-D Hex8 Ogden (200x200x25 slab per GPU; 8 GPUs = the 200^3 mesh)", shape="hex8", n=(200, 200, 25),
               flop=40224.0, b_elem=32 + 1536 + 64 + 24.7, b_asm=1963.5 + 24.7, sample=(200, 200, 1),
               k1="k1_u_kernel<3,8,8,OGDEN>", asm="k3_kernel + k2s_tma_kernel3"),
}


def metric_name(cfg):
    return "makeϕrKt elements/sec (Hex8 NeoHooke)" if cfg == "C2" else f"makeϕrKt elements/sec ({CONFIGS[cfg]['name']})"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def _ncu_traffic(cfg):
    """DRAM bytes per launch of the dominant kernels, read from the committed ncu --set full captures (profiles/
    ncu_traffic.json names the capture each number comes from).  Not a live measurement: ncu cannot run inside a bench."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(cfg)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one streaming nvidia-smi process, 50 ms period)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self._p = None
        self._t = None

    def _read(self):
        for line in self._p.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def __enter__(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu),
                                        "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
            time.sleep(0.15)  # first sample before the load starts is discarded below
            self._skip = len(self.rows)
        except Exception:
            self._p = None
            self._skip = 0
        return self

    def __exit__(self, *a):
        if self._p is not None:
            time.sleep(0.06)
            self._p.terminate()
            try:
                self._p.wait(timeout=3)
            except Exception:
                self._p.kill()
            self._t.join(timeout=3)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows[self._skip:] or self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- workloads ----------------------------------------------------------------------------------------------------
def material(cfg, override=None):
    from ad4sm_b200 import Materials as Ma
    if override == "neohooke" or (override is None and cfg == "C2"):
        return Ma.NeoHooke(1000.0, 5000.0)                              # README.md:150-151 parameters
    if override == "ogden" or (override is None and cfg == "C5"):
        return Ma.Ogden(3.0, 1000.0, 5000.0)                           # SURVEY §8d suggestion
    if cfg == "C1":
        return Ma.Hooke2D(210000.0, 0.3, plane_stress=False)            # README.md:124-125
    if cfg == "C3":
        return Ma.MooneyRivlin(800.0, 200.0, 5000.0)
    if cfg == "C4":
        return Ma.PhaseField(1e-2, 100.0, Ma.NeoHooke(1000.0, 5000.0), 2)  # README.md:148-153
    raise ValueError(cfg)


class Workload:
    """One configuration at one size: the SoA batch, u (D x nNodes, column-major), and for C4 d and the history."""

    def __init__(self, cfg, dims=None, jitter=0.0, seed=20261021, mat=None, device_ctor=False):
        """device_ctor: the product builds its plan from connectivity + coordinates (batched constructors on the GPU,
        `*_batch(..., device=True)`); the host constructors' arrays are then only made when the oracle asks for them."""
        from ad4sm_b200 import Elements as El, mesh
        c = CONFIGS[cfg]
        dims = tuple(dims or c["n"])
        self.cfg, self.dims, self.shape = cfg, dims, c["shape"]
        m = material(cfg, mat)
        self.plan_opts = {}
        self.d = self.hist = None
        if self.shape == "hex8":
            coords, conn = mesh.hex_block(*dims, jitter=jitter, seed=seed)
            ctor, self.D = El.Hex08_batch, 3
        elif self.shape == "tet4":
            coords, conn = mesh.tet_kuhn(*dims, jitter=jitter, seed=seed)
            ctor, self.D = El.Tet04_batch, 3
        elif self.shape == "quad4":
            coords, conn = mesh.quad_grid(*dims, jitter=jitter, seed=seed)
            ctor, self.D = El.Quad_batch, 2
        else:
            coords, conn = mesh.quad_grid(*dims, jitter=jitter, seed=seed)
            ctor, self.D = El.QuadP_batch, 2
            self.plan_opts = {"nh2d_extension": True}   # C4 as named needs the plane-strain embedding
            self.d = np.random.default_rng(seed + 3).uniform(0.0, 0.5, len(coords))
        self._host_ctor = lambda: ctor(conn, coords, m)
        self._host_batch = self._oracle_batch = None
        self.batch = ctor(conn, coords, m, device=True) if device_ctor else self.host_batch
        if self.shape == "quadp":
            self.hist = np.zeros((self.batch.ne, self.batch.P, 2))
        # amplitude relative to the element size of the FULL configuration's generator (h = 1 / max(dims))
        self.u = mesh.random_u(len(coords), self.D, 1.0 / max(dims), a=0.02, seed=seed + 2)
        self.n_nodes = len(coords)
        self.ne = self.batch.ne

    @property
    def host_batch(self):
        """the element arrays as the host (numpy) constructors build them -- what the oracle evaluates"""
        if self._host_batch is None:
            self._host_batch = self._host_ctor()
        return self._host_batch

    @property
    def oracle_batch(self):
        """The element arrays the oracle evaluates: EXACTLY the ∇N / wgt / N the product's plan holds.  With device
        constructors that is the device-built arrays read back (`materialize`), not the host constructors' -- the two differ
        by the conditioning of J (coordinates O(1), differences O(h): 1e-16 / h relative, 2e-14 at h = 1/200), which is a
        property of the constructors (tests/test_gpu_construct.py), not of makeϕrKt."""
        from ad4sm_b200 import Elements as El
        if isinstance(self.batch, El.MeshBatch):
            if self._oracle_batch is None:
                self._oracle_batch = self.batch.materialize()
            return self._oracle_batch
        return self.batch

    def oracle(self, nthreads=0):
        """the reference-algorithm restatement on this workload: list of (ϕ, r, (colptr, rowval, nzval)) and seconds"""
        from oracle import cxx_binding as X
        hb = self.oracle_batch
        t0 = time.perf_counter()
        if self.shape == "quadp":
            kw = dict(nh2d_extension=True)
            res = [X.make_phi_r_Kt([hb], self.u, mode=X.M_PF_U, d=self.d, nthreads=nthreads, **kw),
                   X.make_phi_r_Kt([hb], self.u, mode=X.M_PF_D, d=self.d, hist=[self.hist.copy()], nthreads=nthreads, **kw)]
        else:
            res = [X.make_phi_r_Kt([hb], self.u, nthreads=nthreads)]
        return res, time.perf_counter() - t0

    def product(self):
        """the CUDA path through the reference-facing host API on this workload (same list layout as oracle())"""
        from ad4sm_b200 import Elements as El
        plan = El.Plan(self.batch, self.n_nodes, self.D, **self.plan_opts)
        if self.shape == "quadp":
            a = plan.makeϕrKt(self.u, self.d)
            b = plan.makeϕrKt_d(self.u, self.d, self.hist.copy())
            out = [a, b]
        else:
            out = [plan.makeϕrKt(self.u)]
        plan.close()
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def sample_desc(cfg, dims, ne, dt=None):
    t = "" if dt is None else f", {dt:.1f} s"
    return f"{'x'.join(map(str, dims))} block of the {cfg} generator ({ne} elements{t}), same material and u amplitude"


def run_reference(args):
    """The reference arm: the reference ALGORITHM (C++ restatement) with every host thread, OpenMP thread count set
    explicitly (torchrun exports OMP_NUM_THREADS=1, which would silently make this single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    dims = tuple(args.cpu_sample) if args.cpu_sample else CONFIGS[cfg]["sample"]
    w = Workload(cfg, dims, mat=args.material)
    nthreads = host_threads()
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = w.oracle(nthreads=nthreads)
        if i >= args.warmup:
            times.append(dt)
    dt = float(np.mean(times))
    v = w.ne / dt
    sample = sample_desc(cfg, dims, w.ne) + " per step"
    full = CONFIGS[cfg]["n"]
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg} {CONFIGS[cfg]['name']} {'x'.join(map(str, full))}, one makeϕrKt per step", "sample": sample,
                   "note": "reference ALGORITHM (threaded element duals + serial COO->CSC) restated in C++; the Julia reference cannot run here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference algorithm (oracle/cxx); Julia is not installed"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line, ensure_ascii=False))


# ---- untimed parity checks ---------------------------------------------------------------------------------------
def parity_check_single(cfg, dims, mat, oracle_res=None, workload=None):
    """The CUDA path against the oracle on a block of the benchmarked configuration: pattern bit-exact, every stored
    value to relative 1e-12 (oracle/check.py).  Reuses the oracle result of the cpu_baseline leg when there is one."""
    from oracle.check import assert_result_parity
    w = workload or Workload(cfg, dims, mat=mat, jitter=0.1, device_ctor=True)
    ref = oracle_res if oracle_res is not None else w.oracle(nthreads=host_threads())[0]
    got = w.product()
    worst = 0.0
    # Ogden has no reference derivative: product (spectral tangent) and oracle (duals through a matrix
    # square root) are different formulations, held to 1e-11; everything else to the north-star 1e-12
    tol = 1e-11 if type(w.batch.mat).__name__ == "Ogden" else 1e-12
    for (phi, r, Kt), rf in zip(got, ref):
        _, mr = assert_result_parity(phi, r, Kt.colptr, Kt.rowval, Kt.nzval, rf, tol)
        worst = max(worst, mr)
    return {"status": "ok", "against": "oracle (reference-algorithm restatement)", "sample": sample_desc(cfg, w.dims, w.ne),
            "tolerance": tol, "max_rel_diff_per_entry": worst}


def parity_check_partitioned(world, rank, local, interface, mat):
    """Every rank: a small jittered Hex8 mesh partitioned over the N ranks with the SAME interface transport as the timed
    run must reproduce, bit for bit on the rows this rank owns, a single-GPU evaluation of the whole mesh; rank 0 also
    checks that single-GPU result against the oracle."""
    import torch
    import torch.distributed as dist
    from ad4sm_b200 import dist as AD, mesh, Elements as El
    coords, conn = mesh.hex_block(10, 9, 6 * world, jitter=0.1, seed=77)
    b = El.Hex08_batch(conn, coords, mat)
    u = mesh.random_u(len(coords), 3, 1.0 / (6 * world), a=0.02, seed=78)
    plan1 = El.Plan(b, len(coords), 3, device=local)
    phi0, r0, Kt0 = plan1.makeϕrKt(u)
    plan1.close()
    p = AD.build_partition(conn, AD.contiguous_elem_ranks(len(conn), world), rank, world)
    own = El.ElemBatch("CE", 3, p.local_conn(conn[p.own_elems]), b.gradN[p.own_elems], b.wgt[p.own_elems], b.mat)
    ul = torch.from_numpy(np.ascontiguousarray(u[:, p.local_nodes - 1].T)).cuda()
    if interface == "lib":
        lp = AD.LibDistPlan(own, p.local_conn(conn[p.halo_elems]), p, 3, AD.torch_share_id, device=local)
        plan = lp.plan
        plan.set_stream(torch.cuda.current_stream().cuda_stream or 1)

        def run(x):
            lp.step_device(x.data_ptr())
            v = plan.device_view()
            return AD._wrap(torch, v.phi_dev, (1,))[0], AD._wrap(torch, v.r_dev, (v.n_dof,))
    else:
        be = AD.GpuBackend(own, p.local_conn(conn[p.halo_elems]), p, 3, device=local)
        if interface != "nccl":
            be.enable_peer_staging(p, push=(interface == "push"))
        dp = AD.DistPlan(be, p)
        plan = be.plan

        def run(x):
            return dp.step(x), be.r_tensor()
    ok = 1
    try:
        for f in (0.5, 1.0):   # two evaluations: the second runs on the other receive buffer of the push protocol
            phi, r_t = run(ul * f)
        torch.cuda.synchronize()
        Kt = plan.fetch_csc()
        r = r_t.cpu().numpy()
        assert abs(float(phi.item()) - phi0) <= 1e-13 * abs(phi0)
        loc = np.where(p.owned)[0]
        g = p.local_nodes[loc] - 1
        for c in range(3):
            lc, gc = loc * 3 + c, g * 3 + c
            assert np.array_equal(r[lc], r0[gc])
            assert np.array_equal(Kt.colptr[lc + 1] - Kt.colptr[lc], Kt0.colptr[gc + 1] - Kt0.colptr[gc])
            for a, bb in zip(lc, gc):
                assert np.array_equal(Kt.nzval[Kt.colptr[a] - 1:Kt.colptr[a + 1] - 1], Kt0.nzval[Kt0.colptr[bb] - 1:Kt0.colptr[bb + 1] - 1])
    except AssertionError:
        ok = 0
    flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"status": "ok" if int(flag.item()) else "FAILED",
           "partitioned_vs_single_gpu": f"bitwise on owned rows, {world} ranks, 10x9x{6 * world} jittered Hex8, interface={interface}"}
    if rank == 0:
        from oracle import cxx_binding as X
        from oracle.check import assert_result_parity
        _, mr = assert_result_parity(phi0, r0, Kt0.colptr, Kt0.rowval, Kt0.nzval, X.make_phi_r_Kt([b], u, nthreads=host_threads()), 1e-12)
        out["against"] = "oracle (reference-algorithm restatement), single-GPU result of the same mesh"
        out["max_rel_diff_per_entry"] = mr
    return out


# ---- our arm -------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ad4sm_b200 import Elements as El, _capi as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    cfg = args.config
    conf = CONFIGS[cfg]
    if world > 1 and conf["shape"] != "hex8":
        raise SystemExit(f"{cfg} is a single-GPU configuration (BASELINE.json partitions only the Hex8 meshes); run it with --gpus 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dims = list(conf["n"])
    if args.n:
        dims[0] = dims[1] = args.n
        if len(dims) == 3:
            dims[2] = args.nz or args.n
    elif args.nz and len(dims) == 3:
        dims[2] = args.nz
    dims = tuple(dims)
    strong = args.scaling == "strong" and world > 1
    t_build0 = time.perf_counter()
    w = None
    if world == 1:
        w = Workload(cfg, dims, mat=args.material, device_ctor=not args.host_constructors)
        t_mesh = time.perf_counter() - t_build0
        t_plan0 = time.perf_counter()
        plan = El.Plan(w.batch, w.n_nodes, w.D, device=local, **w.plan_opts)
        t_plan = time.perf_counter() - t_plan0
        batch, u, runner = w.batch, w.u, None
    else:
        from ad4sm_b200 import dist as D
        t_mesh = 0.0
        t_plan0 = time.perf_counter()
        runner = D.SlabRunner(dims[0], dims[2], rank, world, device=local,
                              peer_staging=(False if args.interface == "nccl" else args.interface),
                              mat=material(cfg, args.material), nz_total=(dims[2] if strong else None))
        t_plan = time.perf_counter() - t_plan0
        plan, batch, u = runner.plan, runner.batch, runner.u_local
    ne_local = batch.ne
    ne_t = torch.tensor([ne_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ne_t)
    ne_total = int(ne_t.item())
    dpn = u.shape[0]
    pf = conf["shape"] == "quadp"

    # a non-default torch stream: the plan launches on it, so torch.cuda.Event timing sees the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    plan.set_stream(stream.cuda_stream)   # library kernels, torch ops and NCCL all order on this stream
    u_pin = torch.from_numpy(np.ascontiguousarray(u.T)).pin_memory()      # [node][comp] == column-major dpn x nNodes
    u_dev = u_pin.cuda(non_blocking=True)
    d_pin = d_dev = h_dev = h_pin = None
    if pf:
        d_pin = torch.from_numpy(w.d.copy()).pin_memory()
        d_dev = d_pin.cuda(non_blocking=True)
        h_pin = torch.zeros(w.hist.shape, dtype=torch.float64).pin_memory()
        h_dev = torch.zeros(w.hist.shape, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        if runner is not None:
            runner.step_device(u_dev)
        elif pf:
            plan.eval_device(capi.MODE_PF_U, u_dev.data_ptr(), d_dev.data_ptr())
            plan.eval_device(capi.MODE_PF_D, u_dev.data_ptr(), d_dev.data_ptr(), h_dev.data_ptr())
        else:
            plan.eval_device(capi.MODE_ELASTIC, u_dev.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- first call (plan build is paid once per `elems`; a one-shot makeϕrKt(elems, u) is plan-bound) --------------
    t0 = time.perf_counter()
    step_device()
    barrier()
    t_first = time.perf_counter() - t0

    # ---- device-resident timing (value) ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    plan.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        if runner is not None and getattr(runner, "dplan", None) is not None and os.environ.get("AD4SM_DIST_PROF"):
            print(f"DISTPROF rank {rank}: {runner.dplan.prof_report(skip=max(args.warmup, 3) + 1)}", file=sys.stderr)   # developer tool
        prof_ms, prof_n = plan.profile_read()
        plan.profile_enable(False)
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / args.steps
        # the timed region above lasts tens of ms: keep the same load running (untimed) for ~0.4 s so that the 50 ms
        # clock sampler records the clocks / throttle reasons this workload runs at.  The step count comes from the
        # all-reduced time, so every rank runs the same number of (collective) steps.
        for _ in range(max(1, min(2000, int(400.0 / max(ms_step, 1e-3) / 5)))):
            for _ in range(5):
                step_device()
            torch.cuda.synchronize()
    value = ne_total / (ms_step * 1e-3)

    # ---- end-to-end through the reference-facing call with host buffers (e2e) -------------------
    n_r = u.size
    r_host = torch.empty(n_r, dtype=torch.float64).pin_memory().numpy()  # pinned, like u (bench contract)
    rd_host = torch.empty(u.shape[1], dtype=torch.float64).pin_memory().numpy() if pf else None
    u_np = u_pin.numpy()
    phi, nnz = C.c_double(), C.c_int64()
    L = capi.lib()
    hp = (C.c_void_p * 1)(h_pin.numpy().ctypes.data) if pf else None

    def step_e2e():
        if runner is not None:
            runner.step_host(u_np, r_host)
        elif pf:
            capi.check(L.ad4sm_eval_pf_u(plan._h, u_np.ctypes.data, d_pin.numpy().ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))
            capi.check(L.ad4sm_eval_pf_d(plan._h, u_np.ctypes.data, d_pin.numpy().ctypes.data, hp, C.byref(phi), rd_host.ctypes.data, C.byref(nnz)))
        else:
            capi.check(L.ad4sm_eval(plan._h, u_np.ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))

    def timed(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps

    t_e2e = timed(step_e2e, args.steps)
    e2e_value = ne_total / t_e2e
    h2d = int(u.nbytes) + (int(w.d.nbytes) * 2 + int(w.hist.nbytes) if pf else 0)
    d2h = int(r_host.nbytes + 8 + 8) + (int(rd_host.nbytes + 16 + w.hist.nbytes) if pf else 0)

    # ---- what the Julia drop-in pays when it also wants Kt on the host: + ad4sm_fetch_csc (values only; the pattern is
    # cached on the host after the first call), pinned destination ---------------------------------------------------
    e2e_full = None
    if runner is None and not args.no_e2e_full:
        nd, nz_struct = plan.pattern_size(capi.FIELD_U)
        nz_host = torch.empty(nz_struct, dtype=torch.float64).pin_memory().numpy()
        nzd_host = torch.empty(plan.pattern_size(capi.FIELD_D)[1], dtype=torch.float64).pin_memory().numpy() if pf else None

        def step_full():
            if pf:
                capi.check(L.ad4sm_eval_pf_u(plan._h, u_np.ctypes.data, d_pin.numpy().ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))
                capi.check(L.ad4sm_fetch_csc(plan._h, None, None, nz_host.ctypes.data))
                capi.check(L.ad4sm_eval_pf_d(plan._h, u_np.ctypes.data, d_pin.numpy().ctypes.data, hp, C.byref(phi), rd_host.ctypes.data, C.byref(nnz)))
                capi.check(L.ad4sm_fetch_csc(plan._h, None, None, nzd_host.ctypes.data))
            else:
                capi.check(L.ad4sm_eval(plan._h, u_np.ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))
                capi.check(L.ad4sm_fetch_csc(plan._h, None, None, nz_host.ctypes.data))

        t_full = timed(step_full, max(2, min(args.steps, 5)))
        e2e_full = {"value": ne_total / t_full, "unit": UNIT, "ms_per_step": t_full * 1e3,
                    "d2h_bytes_per_step": d2h + int(nz_host.nbytes) + (int(nzd_host.nbytes) if pf else 0),
                    "call": "ad4sm_eval + ad4sm_fetch_csc(nzval only, pinned; colptr/rowval cached by the caller after the first "
                            "call): the cost of handing the WHOLE Kt to a host solver every evaluation.  The device-side Newton "
                            "glue (ad4sm_newton_*, DESIGN.md) exists so that a Newton loop does not have to pay this."}
        del nz_host

    # ---- device-side Newton glue: res / Kt[ifree,ifree] values / Kt[ifree,icnst]*Δu on the device, only those to the host
    newton = None
    if runner is None and not pf and hasattr(plan, "newton_setup") and not args.no_newton:
        newton = bench_newton(plan, w, u_np, timed, ne_total, torch)

    # ---- untimed parity checks (rank 0 prints the verdict) -----------------------------------------------------------
    cpu = None
    oracle_res = sample_w = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sdims = tuple(args.cpu_sample) if args.cpu_sample else conf["sample"]
        sample_w = Workload(cfg, sdims, mat=args.material, jitter=0.1, device_ctor=not args.host_constructors)
        nthreads = host_threads()
        oracle_res, dt = sample_w.oracle(nthreads=nthreads)
        cpu = {"value": sample_w.ne / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": sample_desc(cfg, sdims, sample_w.ne, dt),
               "note": "C++ restatement of the reference algorithm (oracle/cxx, OpenMP element loop + serial COO->CSC); Julia not installed"}
    parity = None
    if not args.no_parity_check:
        try:
            if world == 1:
                pdims = sample_w.dims if sample_w is not None else tuple(max(2, x // 5) for x in conf["sample"])
                parity = parity_check_single(cfg, pdims, args.material, oracle_res, sample_w)
            else:
                parity = parity_check_partitioned(world, rank, local, args.interface, material(cfg, args.material))
        except AssertionError as e:
            parity = {"status": "FAILED", "error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    peaks, peak_kind = _peaks()
    hbm = peaks["hbm_gbs"]
    k1_ms = prof_ms[capi.PROF_ELEMENT] / max(1, args.steps)
    k2_ms = prof_ms[capi.PROF_ASSEMBLE] / max(1, args.steps)
    red_ms = prof_ms[capi.PROF_REDUCE] / max(1, args.steps)
    tf, mhz = C.c_double(), C.c_double()
    capi.check(L.ad4sm_measure_fp64_peak(local, C.byref(tf), C.byref(mhz)))
    fp64_peak = tf.value
    traffic = _ncu_traffic(cfg) if tuple(dims) == tuple(conf["n"]) and args.material is None else None
    # K1: FP64-bound when W/fp64 > B/hbm, else HBM-bound (Quad4 / Tet4 / QuadP)
    t_fp, t_bw = conf["flop"] / (fp64_peak * 1e12), conf["b_elem"] / (hbm * 1e9)
    k1_fp64 = t_fp >= t_bw
    if k1_fp64:
        ach = conf["flop"] * ne_local / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0
        roof_k1 = {"kernel": conf["k1"] + " (element energy/gradient/Hessian)", "bound": "fp64", "achieved": ach, "peak": fp64_peak,
                   "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                   "peak_source": "DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                   "algorithmic_flop_per_element": conf["flop"]}
    else:
        ach = conf["b_elem"] * ne_local / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
        roof_k1 = {"kernel": conf["k1"] + " (element energy/gradient/Hessian)", "bound": "hbm", "achieved": ach, "peak": hbm,
                   "unit": "GB/s", "frac": ach / hbm, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"}
    roof_k1.update({"algorithmic_bytes_per_launch": conf["b_elem"] * ne_local, "algorithmic_bytes_per_element": conf["b_elem"],
                    "ms_per_launch": k1_ms, "traffic": traffic["k1"] if traffic else None})
    k2_gbs = conf["b_asm"] * ne_local / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    roof_k2 = {"kernel": conf["asm"] + " (deterministic assembly)", "bound": "hbm", "achieved": k2_gbs, "peak": hbm, "unit": "GB/s",
               "frac": k2_gbs / hbm, "traffic": traffic["asm"] if traffic else None,
               "algorithmic_bytes_per_launch": conf["b_asm"] * ne_local, "algorithmic_bytes_per_element": conf["b_asm"],
               "ms_per_launch": k2_ms, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"}
    for rf in (roof_k1, roof_k2):
        if rf["traffic"]:
            rf["traffic_unit"] = "bytes/launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture " \
                                 f"({traffic['source']}), not a live read"
            rf["wasted_traffic_ratio"] = rf["traffic"] / rf["algorithmic_bytes_per_launch"]
    dominant = roof_k1 if k1_ms >= k2_ms else roof_k2
    # whole-path roofline: binding time per element = max(W_alg / FP64 peak, B_alg / HBM peak)
    b_alg = conf["b_elem"] + conf["b_asm"]
    t_bind = max(t_fp, b_alg / (hbm * 1e9))
    path = {"roofline_elements_per_s_per_gpu": 1.0 / t_bind, "frac": (value / world) * t_bind,
            "binding": "fp64" if t_fp > b_alg / (hbm * 1e9) else "hbm",
            "algorithmic_bytes_per_element": b_alg, "algorithmic_flop_per_element": conf["flop"]}
    if traffic:
        path["dram_traffic_per_step"] = traffic["k1"] + traffic["asm"]
        path["wasted_traffic_ratio"] = (traffic["k1"] + traffic["asm"]) / (b_alg * ne_local)

    if world == 1:
        part = "single GPU"
    elif strong:
        part = f"strong scaling: the {'x'.join(map(str, dims))} mesh split into {world} z-slabs ({ne_local} elements on rank 0)"
    else:
        part = f"weak scaling: z-slabs over {world} ranks, {'x'.join(map(str, dims))} per rank"
    zt = dims[-1] * (world if (world > 1 and not strong) else 1)
    gdims = tuple(dims[:-1]) + (zt,)
    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{cfg} {conf['name']} {'x'.join(map(str, gdims))} ({ne_total} elements), one makeϕrKt per step",
                   "elements_per_gpu": ne_local, "partition": part,
                   "interface": None if world == 1 else (
                       "ad4sm_dist_eval: grouped ncclSend/ncclRecv of the staged interface rows + ncclAllGather(ϕ), issued inside the "
                       "library on the plan's stream" if runner.lib is not None else
                       "copy-engine push of the interface rows into the owner's double-buffered HALO rows over NVLink, overlapped "
                       "with K1, + barrier" if runner.backend.push else
                       "K1 bulk stores into the owner's staging over NVLink peer memory + barrier" if runner.peer else
                       "NCCL send/recv of staged element blocks"),
                   "l2": ("working set (∇N, staging, Kt values: GBs per GPU) >> 126 MB L2; no flush needed" if ne_local * b_alg > 5e8 else
                          f"working set {ne_local * b_alg / 1e6:.0f} MB: partly L2-resident between steps (not flushed; launch-latency-bound config)"),
                   "first_call": {"mesh_and_element_constructors_s": t_mesh, "plan_build_s": t_plan, "first_evaluation_s": t_first,
                                  "constructors": ("host (numpy) + upload" if args.host_constructors or world > 1 else
                                                   "batched on the device inside plan_build (ad4sm_batch.mesh)"),
                                  "note": "paid once per `elems`; every later makeϕrKt on the same elems costs ms_per_step"}},
        "clocks": dict(clk.summary(), note="sampled every 50 ms over the timed region and ~0.4 s more of the same load"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": ("ad4sm_eval(plan, u_host, &phi, r_host, &nnz)" if world == 1 and not pf else
                         "ad4sm_eval_pf_u + ad4sm_eval_pf_d (u, d, history in; ϕ, r, history out)" if pf else
                         "per rank: ad4sm_eval_elements_host(u_host) -> interface exchange -> ad4sm_eval_assemble_device -> "
                         "ad4sm_fetch_phi_r_async(r_host) -> ϕ all-gather -> sync")
                        + ": u H2D (node chunks, overlapped with K1), ϕ and r D2H (overlapped with the Kt assembly) inside the timed "
                          "region, pinned host buffers; Kt stays on the device (see e2e_full / newton for the host-side consumers)"},
        "e2e_full": e2e_full, "newton": newton,
        "gpu_launches": int(sum(prof_n)),
        "roofline": dominant, "roofline_kernels": [roof_k1, roof_k2], "roofline_path": path,
        "kernel_ms": {"k1_element": k1_ms, "k2k3_assembly": k2_ms, "phi_reduce": red_ms,
                      "interface_exchange": prof_ms[capi.PROF_GATHER] / max(1, args.steps)},
        "cpu_baseline": cpu, "parity_check": parity,
    }
    print(json.dumps(line, ensure_ascii=False))
    if world > 1:
        dist.destroy_process_group()


def bench_newton(plan, w, u_np, timed, ne_total, torch):
    """Newton-step glue on the device (solvers.jl:211-224, 257-291): bottom face clamped, top face
    displacement-driven.  One corrector iteration = ad4sm_eval (u up; ϕ, r down) + ad4sm_newton_reduce (res, the VALUES of
    Kt[ifree,ifree] and the two norms down, pinned): everything a host `lu(Kt[ifree,ifree]) \\ res` needs, with no host-side
    sparse slicing.  Also the device time of the reduction alone."""
    import ctypes as C
    from ad4sm_b200 import _capi as capi
    L = capi.lib()
    nn = w.n_nodes
    npl = int(np.prod([k + 1 for k in w.dims[:w.D - 1]]))   # nodes of one plane (3-D) / one row (2-D)
    bfree = np.ones((w.D, nn), dtype=np.uint8, order="F")
    bfree[:, :npl] = 0               # bottom plane / row (nodes are numbered x-fastest, last direction slowest): clamped
    bfree[w.D - 1, nn - npl:] = 0    # top plane / row: driven in the last direction
    t0 = time.perf_counter()
    nfree, ncnst, nnz_ff = plan.newton_setup(bfree)
    plan.newton_pattern()
    t_setup = time.perf_counter() - t0
    res = torch.empty(nfree, dtype=torch.float64).pin_memory().numpy()
    nzff = torch.empty(nnz_ff, dtype=torch.float64).pin_memory().numpy()
    norms = np.empty(2)
    r_host = torch.empty(nn * w.D, dtype=torch.float64).pin_memory().numpy()
    phi, nnz = C.c_double(), C.c_int64()

    def step():
        capi.check(L.ad4sm_eval(plan._h, u_np.ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))
        capi.check(L.ad4sm_newton_reduce(plan._h, capi.FIELD_U, 1, None, None, res.ctypes.data, nzff.ctypes.data, norms.ctypes.data))

    t_it = timed(step, 4)
    # device time of the reduction alone (gather of the values + res + norms), CUDA events on the plan's stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.newton_reduce_device(1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        plan.newton_reduce_device(1)
    e1.record()
    torch.cuda.synchronize()
    return {"value": ne_total / t_it, "unit": UNIT, "ms_per_iteration": t_it * 1e3, "reduce_device_ms": e0.elapsed_time(e1) / 5,
            "nfree": nfree, "ncnst": ncnst, "nnz_ff": nnz_ff, "setup_s": t_setup,
            "d2h_bytes_per_iteration": int(res.nbytes + nzff.nbytes + 16 + r_host.nbytes + 16),
            "call": "ad4sm_eval + ad4sm_newton_reduce(corrector): res, nzval of Kt[ifree,ifree], norm(res), norm(fi[icnst]) to pinned "
                    "host memory; the sparse solve itself stays on the host (out of scope)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C2 = the headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one full-size slab per rank; strong = the single-GPU mesh split over the ranks")
    ap.add_argument("--mesh-n", "--n", dest="n", type=int, default=0, help="override: elements per x/y direction")
    ap.add_argument("--mesh-nz", "--nz", dest="nz", type=int, default=0, help="override: element layers (per GPU for weak scaling)")
    ap.add_argument("--material", default=None, choices=["neohooke", "ogden"], help="override the Hex8 configurations' material")
    ap.add_argument("--cpu-sample", type=int, nargs="+", default=None, help="dims of the CPU sample block (default per config)")
    ap.add_argument("--host-constructors", action="store_true",
                    help="build ∇N / wgt with the host (numpy) constructors and upload them, instead of the batched device constructors")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-e2e-full", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--interface", default="push", choices=["nccl", "push", "k1", "lib"],
                    help="multi-GPU interface transport: nccl = send/recv of the staged rows after K1; push = the copy engine "
                         "pushes them into the owner's HALO rows over NVLink (CUDA IPC) while K1 still runs; k1 = K1's own "
                         "bulk stores into the owner's memory (element-major path); lib = the whole exchange inside the library "
                         "(ad4sm_dist_eval: ncclSend/Recv + ϕ all-gather issued by the C ABI, what a Julia host calls)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
