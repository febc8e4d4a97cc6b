#!/usr/bin/env python3
"""Benchmark of the D3Q19 BGK collide-and-stream path on B200 (and of the reference on the host CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one lattice iteration (one launch of the fused collide-and-stream kernel over the whole
lattice; 3 launches per rank when the cube is split into z-slabs).  Metric = MLUPS as the reference
defines it (lbmcl.hpp:604-620): wet cells (DIM-2)^3 x iterations / time.

Workloads (BASELINE.json configs): N = 1 -> LDC 256^3 fp32 (config 3, the one the metric is quoted
on); N = 2, 4, 8 -> LDC 1024^3 fp32 split into z-slabs (config 5; it does not fit one GPU), the five
crossing populations per face exchanged with NCCL send/recv over NVLink (issued by the library on a
high-priority stream), overlapped with the interior update.  The lattices (2.5 GB at 256^3) are far larger than the 126 MB L2, so every step streams from
HBM; no explicit flush is needed.

The JSON line also carries
  roofline     achieved algorithmic GB/s (152 B/cell fp32, 304 B/cell fp64) of the step kernel,
               from CUDA events on the launching stream, against MEASURED_PEAKS.json's HBM copy rate;
  cpu_baseline the reference's own kernels.cl compiled as host C++ (oracle/_ref) timed on this box's
               host cores on a bounded sample of the same workload (rank 0, N = 1 only);
  e2e          the same metric for the whole job through the C ABI with host buffers:
               lbm_init + K iterations + blocking read-back of rho/u into pinned host memory;
  clocks       SM clock / throttle reasons sampled through NVML during the timed region.
`--impl reference` times the reference arm alone (oracle/_ref on all host cores).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS, fused D3Q19 BGK collide-and-stream (wet cells x iterations / time, lbmcl.hpp:604-620)"
BYTES_PER_CELL = {"f32": 152, "f64": 304}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dim", type=int, default=0, help="cube edge (default 256 at N=1, 1024 at N>1)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--stride", type=int, default=32)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--block", default="")
    ap.add_argument("--fast-math", type=int, default=0, help="1 = the reference's -o switch")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dense-halos", action="store_true",
                    help="N > 1: exchange packed halos with ncclSend/ncclRecv instead of fused peer stores")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(key):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own kernel source on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(dim, precision, stride, steps, warmup, seconds):
    """Time `steps` iterations (bounded by `seconds`) of the reference kernel on the host CPU.
    Uses oracle/_ref (reference kernels.cl compiled as host C++) when that configuration was built,
    else the oracle restatement ("port").  OpenMP over z, all host cores."""
    from oracle import Oracle, RefKernel, ref_available

    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    use_dim = dim
    # the reference's 32-bit index arithmetic stops at 256^3
    while use_dim > 256:
        use_dim //= 2
    nu, u_lid = 0.0089, 0.05
    if ref_available(precision, use_dim, stride):
        kind = "reference"
        ref = RefKernel(precision, use_dim, stride)
        st = ref.alloc()
        ref.init(st)

        def step(it):
            ref.step(st, it, 0)
    else:
        kind = "port"
        orc = Oracle(precision)
        st = orc.alloc(use_dim)
        orc.init(st, use_dim, stride, nu, u_lid)

        def step(it):
            orc.step(st, use_dim, stride, nu, u_lid, it, 0)

    it = 1
    t0 = time.perf_counter()
    step(it)
    it += 1
    t_one = time.perf_counter() - t0
    w = max(0, min(warmup, int(0.2 * seconds / max(t_one, 1e-9))))
    for _ in range(w):
        step(it)
        it += 1
    n = max(1, min(steps, int(seconds / max(t_one, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(n):
        step(it)
        it += 1
    dt = time.perf_counter() - t0
    wet = (use_dim - 2) ** 3
    mlups = wet * n / dt / 1e6
    return {
        "value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
        "sample": f"{n} iterations of LDC {use_dim}^3 {precision} stride {stride} after {w + 1} warm-up, "
                  f"OpenMP over z, {dt:.2f} s" + ("" if use_dim == dim else f" (workload is {dim}^3; the reference's "
                                                  "int indexing stops at 256^3)"),
        "ms_per_step": dt / n * 1e3, "steps": n,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dim = a.dim or (256 if a.gpus == 1 else 1024)
    r = cpu_reference_run(dim, a.precision, a.stride, a.steps, a.warmup, max(a.cpu_seconds, 60.0))
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MLUPS", "n_gpus": a.gpus,
        "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": {"workload": f"LDC {dim}^3 {a.precision}, nu 0.0089, U 0.05, stride {a.stride}, -e 0"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample_once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:  # noqa: BLE001
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.sample_once()
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lbmcl_b200.capi import Simulation
    from lbmcl_b200.slabs import connect_slabs, slab_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    dim = a.dim or (256 if world == 1 else 1024)
    z0, z1 = slab_range(dim, world, rank)
    nz = z1 - z0
    block = tuple(int(v) for v in a.block.split(",")) if a.block else (256, 1, 1)
    esize = 4 if a.precision == "f32" else 8
    npdtype = np.float32 if a.precision == "f32" else np.float64
    tdtype = torch.float32 if a.precision == "f32" else torch.float64

    sim = Simulation(dim=dim, precision=a.precision, stride=a.stride, block=block, variant=a.variant,
                     fast_math=bool(a.fast_math), device=local_rank, z_range=(z0, z1))
    eff_block, vec = sim.block_shape
    # an explicit stream: torch's default stream has handle 0, which lbm_set_stream reads as "own stream"
    main = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main)
    sim.set_stream(main.cuda_stream)
    has_lo, has_hi = rank > 0, rank < world - 1

    if world > 1:
        # NCCL communicator owned by the library; torch.distributed only bootstraps it (128-byte id,
        # CUDA IPC handles of the neighbours' lattices)
        transport = connect_slabs(sim, rank, world, dev, fused=not a.dense_halos)

    def run_steps(n, every=0):
        # N = 1: n launches of the step kernel.  N > 1: per iteration the boundary planes run on a
        # high-priority stream and hand the 5 crossing populations per face to the neighbours (peer
        # stores over NVLink + an NCCL token, or packed ncclSend/ncclRecv), interior planes concurrently
        # on `main`; lbm_run drives it (include/lbm_b200.h, transports 2b / 2c)
        sim.run(n, every)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-timed throughput: inputs resident in HBM ----
    sim.init()
    run_steps(a.warmup)
    sync_all()
    l0 = sim.launch_count
    sampler = ClockSampler(local_rank)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    start.record(main)
    run_steps(a.steps)
    stop.record(main)
    sync_all()
    sampler.stop_flag.set()
    sampler.join()
    ms = start.elapsed_time(stop)
    launches = sim.launch_count - l0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    wet = (dim - 2) ** 3
    mlups = wet * a.steps / (ms * 1e3)
    bpc = BYTES_PER_CELL[a.precision]
    achieved = mlups * 1e6 * bpc / 1e9
    peak, peak_src = measured_peak()

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not a.no_e2e:
        n_own = nz * dim * dim
        rho_h = torch.empty(n_own, dtype=tdtype, pin_memory=True).numpy()        # this rank's planes only
        u_h = torch.empty(3 * n_own, dtype=tdtype, pin_memory=True).numpy().reshape(3, n_own)
        sync_all()
        t0 = time.perf_counter()
        sim.init()
        run_steps(a.steps, a.steps)
        sim.read_macros_slab(rho_h, u_h)  # blocking D2H of this rank's planes into pinned memory
        sync_all()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert np.isfinite(rho_h[(nz // 2) * dim * dim + (dim // 2) * dim + dim // 2])
        e2e = {"value": wet * a.steps / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 4 * n_own * esize * world / a.steps,
               "note": "whole job: lbm_init + K iterations + blocking rho/u read-back into pinned host memory; "
                       "the lid-driven cavity has no host-side input (the reference initialises on the device)"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(dim, a.precision, a.stride, 10 ** 9, 1, a.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        key = f"{a.precision}_{dim}_vec{vec}_fast{a.fast_math}"
        out = {
            "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {
                "workload": f"LDC {dim}^3 {a.precision}, nu 0.0089, U 0.05, stride {a.stride}, -e 0"
                            + (f", {world} z-slabs of {nz} planes, halo transport {transport}" if world > 1 else ""),
                "kernel": {8: "in-place AA pattern (one lattice)", 16: "two-lattice pull, TMA-fed"}.get(
                    a.variant, "two-lattice pull") + f", {vec} cell(s)/thread, block {list(eff_block)}, "
                          + ("fast math (-o)" if a.fast_math else "strict IEEE operation order"),
                "l2": "lattices (2 x %.2f GB per GPU) exceed the 126 MB L2; no flush needed"
                      % (19 * (nz + 2) * dim * dim * esize / 1e9),
            },
            "roofline": {"bound": "hbm", "achieved": achieved / world, "peak": peak, "unit": "GB/s",
                         "frac": achieved / world / peak, "traffic": recorded_traffic(key),
                         "peak_source": peak_src, "bytes_per_cell": bpc,
                         "note": "per-GPU algorithmic bytes of the step kernel / event time"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.result(),
            "device": sim.device_name,
        }
        print(json.dumps(out), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def run_b200_single_process(a):
    """`python bench.py --gpus N` WITHOUT torchrun: one process drives N devices through the same-process
    z-slab group (lbm_group_*: peer stores between the slabs, event ordering, no NCCL).  The driver's
    contract launches N > 1 under torchrun (run_b200); this path is what the CLI's -G N uses."""
    import numpy as np
    import torch
    from lbmcl_b200.capi import Group

    if torch.cuda.device_count() < a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but {torch.cuda.device_count()} device(s) visible")
    dim = a.dim or 1024
    esize = 4 if a.precision == "f32" else 8
    block = tuple(int(v) for v in a.block.split(",")) if a.block else (256, 1, 1)
    g = Group(list(range(a.gpus)), dim=dim, precision=a.precision, stride=a.stride, block=block, variant=a.variant,
              fast_math=bool(a.fast_math))
    g.init()
    g.run(a.warmup, 0)
    g.sync()
    sampler = ClockSampler(0)
    sampler.start()
    _, k0 = g.time_ms()
    g.run(a.steps, 0)
    _, k1 = g.time_ms()          # slowest slab's device time of the batch (CUDA events on its stream)
    sampler.stop_flag.set()
    sampler.join()
    ms = k1 - k0
    wet = (dim - 2) ** 3
    mlups = wet * a.steps / (ms * 1e3)
    bpc = BYTES_PER_CELL[a.precision]
    achieved = mlups * 1e6 * bpc / 1e9
    peak, peak_src = measured_peak()
    e2e = None
    if not a.no_e2e:
        npd = np.float32 if a.precision == "f32" else np.float64
        rho_h = np.empty(dim ** 3, dtype=npd)
        t0 = time.perf_counter()
        g.init()
        g.run(a.steps, a.steps)
        g.read_macros(rho_h, None)
        dt = time.perf_counter() - t0
        e2e = {"value": wet * a.steps / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": dim ** 3 * esize / a.steps,
               "note": "whole job: init + K iterations + blocking rho read-back (pageable host memory)"}
    out = {
        "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic",
        "config": {"workload": f"LDC {dim}^3 {a.precision}, nu 0.0089, U 0.05, stride {a.stride}, -e 0, {a.gpus} z-slabs "
                               "in one process, halo transport peer-stores+events",
                   "kernel": "two-lattice pull, strict IEEE operation order" if not a.fast_math else "two-lattice pull, -o"},
        "roofline": {"bound": "hbm", "achieved": achieved / a.gpus, "peak": peak, "unit": "GB/s",
                     "frac": achieved / a.gpus / peak, "traffic": None, "peak_source": peak_src, "bytes_per_cell": bpc},
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": 3 * a.steps * a.gpus, "clocks": sampler.result(),
    }
    print(json.dumps(out), flush=True)
    g.close()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_b200_single_process(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
