#!/usr/bin/env python
"""bench.py -- elements/sec of one `makeϕrKt(elems, u)` evaluation (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the reference ALGORITHM on the host cores (see below)

Workload at every N: config C2 of BASELINE.json -- 3-D Hex8 NeoHooke(1000, 5000) structured block, 2x2x2 Gauss,
u = 0.02·h·U(-1,1).  N = 1: 100x100x100 (1M elements).  N > 1: weak scaling, each rank owns a 100x100x100 slab of a
100x100x(100·N) block (z-slab partition), interface rows exchanged over NCCL (dist.py).
A "step" is one full evaluation: u on device -> ϕ, r, CSR/CSC values of Kt on device.

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

METRIC = "makeϕrKt elements/sec (Hex8 NeoHooke)"
UNIT = "elements/s"
# algorithmic work per Hex8 NeoHooke element
W_ALG_FLOP = 31424.0   # canonical generic-tangent count (F, Bᵀ·P, two-stage symmetric Bᵀ·A·B, invariant-level material)
B_ALG_BYTES = 3645.0   # nodes(int32) + ∇N + wgt + u gather + Kt values + r
# K2 (assembly) algorithmic bytes per element: Kt values written once + r (the staging it reads is overhead)
B_ALG_K2 = 1963.5 + 24.7


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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


def material(name):
    from ad4sm_b200 import Materials as Ma
    return {"neohooke": lambda: Ma.NeoHooke(1000.0, 5000.0),          # config C2 (README.md:150-151 parameters)
            "ogden": lambda: Ma.Ogden(3.0, 1000.0, 5000.0)}[name]()    # config C5 (SURVEY §8d suggestion)


def build_workload(n, nz, jitter=0.0, seed=20261021, mat="neohooke"):
    from ad4sm_b200 import Elements as El, mesh
    coords, conn = mesh.hex_block(n, n, nz, jitter=jitter, seed=seed)
    batch = El.Hex08_batch(conn, coords, material(mat))
    u = mesh.random_u(len(coords), 3, 1.0 / n, a=0.02, seed=seed + 2)
    return batch, u


def cpu_reference_rate(sample_n, nthreads=0, repeats=1):
    """elements/s of the reference-algorithm restatement (oracle C++) on a sample_n^3 block of the same workload."""
    from oracle import cxx_binding as X
    batch, u = build_workload(sample_n, sample_n)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        X.make_phi_r_Kt([batch], u, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cores = X.max_threads() if nthreads <= 0 else nthreads
    return batch.ne / best, cores, batch.ne, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_sample
    rates = []
    cores = ne = 0
    for i in range(args.warmup + args.steps):
        rate, cores, ne, dt = cpu_reference_rate(n)
        if i >= args.warmup:
            rates.append((rate, dt))
    v = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([d for _, d in rates]) * 1e3)
    sample = f"{n}x{n}x{n} Hex8 NeoHooke block ({ne} elements) per step, same generator/material/u as the GPU arm"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2 Hex8 NeoHooke {args.n}x{args.n}x{args.n * args.gpus} ({args.n ** 3 * args.gpus} elements), 2x2x2 Gauss, "
                               "one makeϕrKt per step", "sample": sample,
                   "note": "reference ALGORITHM (threaded element duals + serial COO->CSC) restated in C++; the Julia reference cannot run here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference algorithm (oracle/cxx); Julia is not installed"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line, ensure_ascii=False))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ad4sm_b200 import Elements as El, _capi as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    t_build0 = time.perf_counter()
    if world == 1:
        batch, u = build_workload(n, args.nz or n, mat=args.material)
        plan = El.Plan(batch, u.shape[1], 3, device=local)
        runner = None
    else:
        from ad4sm_b200 import dist as D
        runner = D.SlabRunner(n, args.nz or n, rank, world, device=local, peer_staging=("k1" if args.peer_staging else (False if args.interface == "nccl" else args.interface)),
                              mat=material(args.material))
        plan, batch, u = runner.plan, runner.batch, runner.u_local
    t_build = time.perf_counter() - t_build0
    ne_local = batch.ne
    ne_total = ne_local * world

    # a non-default torch stream: the plan launches on it, so torch.cuda.Event timing sees the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    plan.set_stream(stream.cuda_stream)   # library kernels, torch ops and NCCL all order on this stream
    u_pin = torch.from_numpy(np.ascontiguousarray(u.T)).pin_memory()      # [node][comp] == column-major 3 x nNodes
    u_dev = u_pin.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def step_device():
        if runner is None:
            plan.eval_device(capi.MODE_ELASTIC, u_dev.data_ptr())
        else:
            runner.step_device(u_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        prof_ms, prof_n = plan.profile_read()
        plan.profile_enable(False)
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / args.steps
        # the timed region above lasts tens of ms: keep the same load running (untimed) for ~0.4 s so that the 50 ms
        # clock sampler records the clocks / throttle reasons this workload runs at.  The step count comes from the
        # all-reduced time, so every rank runs the same number of (collective) steps.
        for _ in range(max(1, int(400.0 / max(ms_step, 1e-3) / 5))):
            for _ in range(5):
                step_device()
            torch.cuda.synchronize()
    value = ne_total / (ms_step * 1e-3)

    # ---- end-to-end through the reference-facing call with host buffers (e2e) -------------------
    r_host = torch.empty(u.size, dtype=torch.float64).pin_memory().numpy()  # pinned, like u (bench contract)
    u_np = u_pin.numpy()
    import ctypes as C
    phi, nnz = C.c_double(), C.c_int64()

    def step_e2e():
        if runner is None:
            capi.check(capi.lib().ad4sm_eval(plan._h, u_np.ctypes.data, C.byref(phi), r_host.ctypes.data, C.byref(nnz)))
        else:
            runner.step_host(u_np, r_host)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = ne_total / (float(t_e2e.item()) / args.steps)
    h2d = int(u.nbytes)
    d2h = int(r_host.nbytes + 8 + 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    peaks, peak_kind = _peaks()
    k1_ms = prof_ms[capi.PROF_ELEMENT] / max(1, args.steps)
    k2_ms = prof_ms[capi.PROF_ASSEMBLE] / max(1, args.steps)
    red_ms = prof_ms[capi.PROF_REDUCE] / max(1, args.steps)
    tf, mhz = C.c_double(), C.c_double()
    capi.check(capi.lib().ad4sm_measure_fp64_peak(local, C.byref(tf), C.byref(mhz)))
    fp64_peak = tf.value
    k1_tflops = W_ALG_FLOP * ne_local / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0
    k2_gbs = B_ALG_K2 * ne_local / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    # DRAM traffic per launch from the committed ncu --set full capture of this same command (profiles/
    # r01_ncu_k1_k3_k2s_sorted_path.txt): dram__bytes_read.sum + dram__bytes_write.sum, 1M-element workload only
    at_profiled_size = (n == 100 and (args.nz or n) == 100 and args.material == "neohooke")
    roof_k1 = {"kernel": "k1_u_kernel<3,8,8,NEO> (element energy/gradient/Hessian)", "bound": "fp64", "achieved": k1_tflops,
               "peak": fp64_peak, "unit": "TFLOP/s", "frac": k1_tflops / fp64_peak if fp64_peak else None,
               "traffic": (1.905480e9 + 3.114725e9) if at_profiled_size else None, "traffic_unit": "bytes/launch (ncu)",
               "algorithmic_bytes_per_launch": (32 + 1536 + 64 + 24.7 + 144 + 36 * 80 + 192 + 8) * ne_local,
               "ms_per_launch": k1_ms, "peak_source": "DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
               "algorithmic_flop_per_element": W_ALG_FLOP}
    roof_k2 = {"kernel": "k3_kernel + k2s_tma_kernel3 (deterministic assembly from the destination-sorted staging)", "bound": "hbm",
               "achieved": k2_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": k2_gbs / peaks["hbm_gbs"],
               "traffic": (3.202257e9 + 1.993955e9) if at_profiled_size else None, "traffic_unit": "bytes/launch (ncu, k2s_tma_kernel3)",
               "ms_per_launch": k2_ms, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "algorithmic_bytes_per_element": B_ALG_K2}
    dominant = roof_k1 if k1_ms >= k2_ms else roof_k2
    # whole-path roofline: binding time per element = max(W_alg / FP64 peak, B_alg / HBM peak)
    t_bind = max(W_ALG_FLOP / (fp64_peak * 1e12), B_ALG_BYTES / (peaks["hbm_gbs"] * 1e9)) if fp64_peak else None
    path = {"roofline_elements_per_s_per_gpu": 1.0 / t_bind if t_bind else None,
            "frac": (value / world) * t_bind if t_bind else None,
            "binding": "fp64" if fp64_peak and W_ALG_FLOP / (fp64_peak * 1e12) > B_ALG_BYTES / (peaks["hbm_gbs"] * 1e9) else "hbm"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, ne_s, dt = cpu_reference_rate(args.cpu_sample)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_sample}^3 Hex8 NeoHooke block ({ne_s} elements, {dt:.1f} s) of the same workload",
               "note": "C++ restatement of the reference algorithm (oracle/cxx, OpenMP element loop + serial COO->CSC); Julia not installed"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{'C2' if args.material == 'neohooke' else 'C5'} Hex8 {'NeoHooke' if args.material == 'neohooke' else 'Ogden'} "
                               f"{n}x{n}x{(args.nz or n) * world} ({ne_total} elements), 2x2x2 Gauss, one makeϕrKt per step",
                   "elements_per_gpu": ne_local, "partition": "single GPU" if world == 1 else f"z-slabs over {world} ranks",
                   "interface": None if world == 1 else (
                       "copy-engine push of the interface rows into the owner's double-buffered HALO rows over NVLink, overlapped "
                       "with K1, + barrier" if runner.backend.push else
                       "K1 bulk stores into the owner's staging over NVLink peer memory + barrier" if runner.peer else
                       "NCCL send/recv of staged element blocks"),
                   "l2": "working set (∇N 1.5 GB + staging 2.9 GB + Kt 2.0 GB per GPU) >> 126 MB L2; no flush needed",
                   "plan_build_s": t_build},
        "clocks": dict(clk.summary(), note="sampled every 50 ms over the timed region and ~0.4 s more of the same load"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": ("ad4sm_eval(plan, u_host, &phi, r_host, &nnz)" if world == 1 else
                         "per rank: ad4sm_eval_elements_host(u_host) -> interface exchange -> ad4sm_eval_assemble_device -> "
                         "ad4sm_fetch_phi_r_async(r_host) -> ϕ all-gather -> sync")
                        + ": u H2D (node chunks, overlapped with K1), ϕ and r D2H (overlapped with the Kt assembly) inside the timed "
                          "region, pinned host buffers; Kt stays on the device (the reference's consumer, lu(Kt[ifree,ifree]), is "
                          "out of scope)"},
        "gpu_launches": int(sum(prof_n)),
        "roofline": dominant, "roofline_kernels": [roof_k1, roof_k2], "roofline_path": path,
        "kernel_ms": {"k1_element": k1_ms, "k2k3_assembly": k2_ms, "phi_reduce": red_ms,
                      "interface_exchange": prof_ms[capi.PROF_GATHER] / max(1, args.steps)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line, ensure_ascii=False))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh-n", "--n", dest="n", type=int, default=100, help="elements per x/y direction (100 -> 1M Hex8 per GPU with --nz 100)")
    ap.add_argument("--mesh-nz", "--nz", dest="nz", type=int, default=0, help="element layers per GPU (default: --n); config C5 = --mesh-n 200 --mesh-nz 25 --material ogden on 8 GPUs (torchrun swallows a bare --n)")
    ap.add_argument("--material", default="neohooke", choices=["neohooke", "ogden"])
    ap.add_argument("--cpu-sample", type=int, default=56, help="edge of the CPU sample block (56 -> 175616 elements, a few seconds on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--interface", default="push", choices=["nccl", "push", "k1"],
                    help="multi-GPU interface transport: nccl = send/recv of the staged rows after K1; push = the copy engine "
                         "pushes them into the owner's HALO rows over NVLink (CUDA IPC) while K1 still runs; k1 = K1's own "
                         "bulk stores into the owner's memory (element-major path)")
    ap.add_argument("--peer-staging", action="store_true", help="alias of --interface k1 (round-1 name)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
