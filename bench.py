#!/usr/bin/env python3
"""Benchmark of the D3Q19 BGK collide-and-stream path on B200 (and of the reference on the host CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one lattice iteration = ONE launch of the fused collide-and-stream kernel per GPU (also with
z-slabs: the launch's grid starts with the slab's boundary planes, which store the 5 crossing populations
per face straight into the neighbours' halo planes over NVLink and order themselves with in-kernel epoch
flags; include/lbm_b200.h transport 2c).  Metric = MLUPS as the reference defines it (lbmcl.hpp:604-620):
wet cells (DIM-2)^3 x iterations / time.

Workloads (BASELINE.json configs): N = 1 -> LDC 256^3 fp32 (config 3, the one the metric is quoted on);
N = 2, 4, 8 -> LDC 1024^3 fp32 split into z-slabs (config 5; two lattices of it do not fit one GPU).  The
lattices (2.5 GB at 256^3) are far larger than the 126 MB L2, so every step streams from HBM; no flush.

The JSON line also carries
  roofline     achieved algorithmic GB/s (152 B/cell fp32, 304 B/cell fp64) of the step kernel, from CUDA
               events on the launching stream, against MEASURED_PEAKS.json's HBM copy rate;
  cpu_baseline the reference's own kernels.cl compiled as host C++ (oracle/_ref) timed on this box's host
               cores on a bounded sample of the same workload (rank 0, N = 1 only);
  e2e          the same metric for the whole job through the C ABI with host buffers: lbm_init + K
               iterations + read-back of rho/u into pinned host memory; `with_vti` = the reference-facing
               CLI (`lbmcl -i K -e K`) including its VTI files, i.e. the reference's "Total MLUPS";
  extra        (N = 1) config 4 (512^3 fp64) and config 5 on ONE GPU with the in-place kernels (1024^3 fp32);
  parity_gate  (N > 1) a small cavity run over the same ranks and transport, compared bit for bit on rank 0
               with a single-context run and with the CPU oracle BEFORE anything is timed;
  efficiency_same_workload (N > 1) value / (N x the one-GPU in-place 1024^3 rate measured on rank 0);
  clocks       SM clock / throttle reasons sampled through NVML during the timed region.
`--impl reference` times the reference arm alone (oracle/_ref on all host cores).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS, fused D3Q19 BGK collide-and-stream (wet cells x iterations / time, lbmcl.hpp:604-620)"
BYTES_PER_CELL = {"f32": 152, "f64": 304}
LOCKSTEP = 3   # last warm-up iterations, enqueued right before the timed ones (no host sync in between)


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
    ap.add_argument("--no-extra", action="store_true", help="skip the config 4 / config 5 side measurements")
    ap.add_argument("--no-parity-gate", action="store_true")
    ap.add_argument("--transport", default="flags", choices=["flags", "token", "dense"],
                    help="N > 1: in-kernel epoch flags (default), NCCL token, or dense NCCL halos")
    ap.add_argument("--dense-halos", action="store_true", help="alias of --transport dense")
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


def workload_text(dim, precision, stride):
    return f"LDC {dim}^3 {precision}, nu 0.0089, U 0.05, stride {stride}, -e 0"


def l2_text(dim, planes, esize):
    return ("lattices (2 x %.2f GB per GPU) exceed the 126 MB L2; no flush needed"
            % (19 * planes * dim * dim * esize / 1e9))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own kernel source on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(dim, precision, stride, steps, warmup, seconds):
    """Time `steps` iterations (bounded by `seconds`) of the reference kernel on the host CPU.
    Uses oracle/_ref (reference kernels.cl compiled as host C++) when that configuration was built,
    else the oracle restatement ("port").  OpenMP over z on all host cores: the thread count is set
    through the library (launchers such as torchrun export OMP_NUM_THREADS=1) and the number really in
    effect is what gets reported."""
    from oracle import Oracle, RefKernel, host_cores, ref_available

    want = host_cores()
    use_dim = dim
    # the reference's 32-bit index arithmetic stops at 256^3 (kernels.cl:64-67)
    while use_dim > 256:
        use_dim //= 2
    nu, u_lid = 0.0089, 0.05
    if ref_available(precision, use_dim, stride):
        kind = "reference"
        ref = RefKernel(precision, use_dim, stride)
        cores = ref.set_threads(want)
        st = ref.alloc()
        ref.init(st)

        def step(it):
            ref.step(st, it, 0)
    else:
        kind = "port"
        orc = Oracle(precision)
        cores = orc.set_threads(want)
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
                  f"OpenMP over z with {cores} threads ({want} cores available), {dt:.2f} s"
                  + ("" if use_dim == dim else f" (the B200 arm runs {dim}^3; the reference's int indexing stops at 256^3)"),
        "ms_per_step": dt / n * 1e3, "steps": n, "dim": use_dim,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dim = a.dim or (256 if a.gpus == 1 else 1024)
    esize = 4 if a.precision == "f32" else 8
    r = cpu_reference_run(dim, a.precision, a.stride, a.steps, a.warmup, max(a.cpu_seconds, 60.0))
    used = r["dim"]
    workload = workload_text(used, a.precision, a.stride)
    if used != dim:
        workload += (f" (the largest cube the reference can index, kernels.cl:64-67; the B200 arm at {a.gpus} GPUs runs "
                     f"{dim}^3 in z-slabs)")
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MLUPS", "n_gpus": a.gpus,
        "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": {"workload": workload, "l2": l2_text(used, used, esize)},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs.  Constructed (NVML
    initialisation, handle lookup) well BEFORE the timed region; start() only starts the polling thread."""

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
            self.sample_once()          # first NVML query paid here, not in the timed region
            self.samples.clear()
            self.reasons.clear()
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
# pieces of the B200 arm
# ------------------------------------------------------------------------------------------------
def timed_run(sim, stream, steps, warmup, sync_all, sampler=None):
    """W warm-up iterations, then exactly `steps` iterations between two CUDA events on the launching
    stream.  The last LOCKSTEP warm-up iterations are enqueued right before the timed ones, after the
    barrier: with z-slabs they pull the ranks into lock step on the device (a rank cannot run ahead of a
    neighbour by more than one iteration), so host-side skew between the ranks is absorbed before the start
    event instead of inside the timed region."""
    import torch
    lock = min(LOCKSTEP, warmup)
    sim.run(warmup - lock)
    sync_all()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.start()
    sim.run(lock)
    l0 = sim.launch_count
    start.record(stream)
    sim.run(steps)
    stop.record(stream)
    launches = sim.launch_count - l0
    sync_all()
    if sampler is not None:
        sampler.stop_flag.set()
        sampler.join()
    sim.sync()          # reports a transport time-out instead of a silently wrong number
    return start.elapsed_time(stop), launches


def single_gpu_side_run(dim, precision, stride, variant, steps, warmup, device_index):
    """A short device-timed run of another configuration on one GPU (the `extra` entries)."""
    import torch
    from lbmcl_b200.capi import Simulation

    dev = torch.device("cuda", device_index)
    stream = torch.cuda.Stream(device=dev)
    with Simulation(dim=dim, precision=precision, stride=stride, variant=variant, device=device_index) as sim:
        sim.set_stream(stream.cuda_stream)
        sim.init()
        ms, _ = timed_run(sim, stream, steps, warmup, lambda: torch.cuda.synchronize(dev))
        gb = sim.device_bytes / 1e9
    wet = (dim - 2) ** 3
    mlups = wet * steps / (ms * 1e3)
    achieved = mlups * 1e6 * BYTES_PER_CELL[precision] / 1e9
    peak, _ = measured_peak()
    return {"workload": workload_text(dim, precision, stride), "kernel": "in-place AA pattern (one lattice)" if variant == 8
            else "two-lattice pull", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "value": mlups,
            "unit": "MLUPS", "achieved_gbs": achieved, "frac": achieved / peak, "device_gb": gb}


def cli_with_vti(dim, precision, stride, steps, device_index):
    """The reference-facing end-to-end job: the C++ host program `lbmcl -i K -e K` (same CLI, same VTI files
    as the reference), Total MLUPS as the reference reports it (lbmcl.hpp:604-607: init + iterations +
    read-backs + ASCII VTI writing).  Files go to a scratch directory that is removed afterwards."""
    exe = os.path.join(ROOT, "lbmcl_b200", "host", "lbmcl")
    if not os.path.exists(exe):
        return {"unavailable": "lbmcl_b200/host/lbmcl not built"}
    out = tempfile.mkdtemp(prefix="lbmcl_bench_")
    try:
        cmd = [exe, "-D", str(device_index), "-d", str(dim), "-i", str(steps), "-e", str(steps), "-s", str(stride),
               "-v", out, "-p", out] + (["-F"] if precision == "f64" else [])
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": f"lbmcl exited {r.returncode}: {r.stderr[-200:]}"}
        m = re.search(r"Total MLUPS:\s*([0-9.eE+-]+)", r.stdout)
        t = re.search(r"Total time:\s*([0-9.eE+-]+)", r.stdout)
        size = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
        return {"value": float(m.group(1)) if m else None, "unit": "MLUPS", "total_ms": float(t.group(1)) if t else None,
                "process_wall_s": wall, "vti_bytes_written": size, "files": len(os.listdir(out)),
                "command": f"lbmcl -d {dim} -i {steps} -e {steps} -s {stride}" + (" -F" if precision == "f64" else "") + " -v <tmp>",
                "note": "Total MLUPS of the C++ host program (lbmcl.hpp:604-607): init + iterations + asynchronous pinned "
                        "read-back + ASCII VTI files at iteration 0 and K"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    finally:
        shutil.rmtree(out, ignore_errors=True)


def parity_gate(rank, world, local_rank, dev, transport):
    """Before anything is timed: small cavities over the SAME ranks and the SAME transport, gathered on
    rank 0 and compared bit for bit with a single-context run on rank 0's GPU and with the CPU oracle
    (the checker, used outside every timed region)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lbmcl_b200.capi import Simulation
    from lbmcl_b200.slabs import connect_slabs, slab_range

    cases = []
    ok_all = True
    used = None
    for precision, dim, stride, its in (("f32", 128, 32, 21), ("f64", 64, 32, 7)):
        if dim % world != 0:
            continue
        npd = np.float32 if precision == "f32" else np.float64
        z0, z1 = slab_range(dim, world, rank)
        sim = Simulation(dim=dim, precision=precision, stride=stride, device=local_rank, z_range=(z0, z1))
        used = connect_slabs(sim, rank, world, dev, transport=transport)
        sim.init()
        sim.run(its - 3, its)
        sim.run(3, its)
        rho, u = sim.read_macros_slab()
        sim.sync()
        bits = torch.from_numpy(np.concatenate([rho, u.reshape(-1)]).view(np.uint8).copy()).to(dev)
        out = [torch.empty_like(bits) for _ in range(world)] if rank == 0 else None
        dist.gather(bits, out, dst=0)
        dist.barrier()          # nobody frees a lattice a neighbour may still be storing into
        sim.close()
        if rank == 0:
            from oracle import Oracle, host_cores
            Oracle(precision).set_threads(host_cores())   # torchrun exports OMP_NUM_THREADS=1
            n_slab = (dim // world) * dim * dim
            parts = [o.cpu().numpy().view(npd) for o in out]
            g_rho = np.concatenate([p[:n_slab] for p in parts])
            g_u = np.concatenate([p[n_slab:].reshape(3, n_slab) for p in parts], axis=1)
            with Simulation(dim=dim, precision=precision, stride=stride, device=local_rank) as one:
                one.init()
                one.run(its, its)
                s_rho, s_u = one.read_macros()
            exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, its)
            same_single = g_rho.tobytes() == s_rho.tobytes() and g_u.tobytes() == s_u.tobytes()
            same_oracle = g_rho.tobytes() == exp["rho"][1].tobytes() and g_u.tobytes() == exp["u"][1].tobytes()
            cases.append({"case": f"{precision} {dim}^3 x{its}", "vs_single_gpu": same_single, "vs_oracle": same_oracle})
            ok_all = ok_all and same_single and same_oracle
    flag = torch.tensor([1 if ok_all else 0], device=dev, dtype=torch.int32)
    dist.broadcast(flag, 0)
    return {"ok": bool(flag.item()), "transport": used, "ranks": world, "cases": cases,
            "compared": "rho and u of every cell, bit for bit"}


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lbmcl_b200.capi import Simulation, pinned_array
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
    transport_req = "dense" if a.dense_halos else a.transport

    dim = a.dim or (256 if world == 1 else 1024)
    z0, z1 = slab_range(dim, world, rank)
    nz = z1 - z0
    block = tuple(int(v) for v in a.block.split(",")) if a.block else (256, 1, 1)
    esize = 4 if a.precision == "f32" else 8

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # NVML is initialised here, long before the timed region (8 ranks serialise on the driver lock)
    sampler = ClockSampler(local_rank)

    gate = None
    if world > 1 and not a.no_parity_gate:
        gate = parity_gate(rank, world, local_rank, dev, transport_req)
        if not gate["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_gate": gate,
                                  "error": "z-slab run differs from the single-GPU run / the oracle; nothing was timed"}),
                      flush=True)
            dist.destroy_process_group()
            raise SystemExit(3)

    sim = Simulation(dim=dim, precision=a.precision, stride=a.stride, block=block, variant=a.variant,
                     fast_math=bool(a.fast_math), device=local_rank, z_range=(z0, z1))
    eff_block, vec = sim.block_shape
    # an explicit stream: torch's default stream has handle 0, which lbm_set_stream reads as "own stream"
    main = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main)
    sim.set_stream(main.cuda_stream)
    transport = connect_slabs(sim, rank, world, dev, transport=transport_req) if world > 1 else None

    # ---- device-timed throughput: inputs resident in HBM ----
    sim.init()
    ms, launches = timed_run(sim, main, a.steps, a.warmup, sync_all, sampler)
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
        npd = np.float32 if a.precision == "f32" else np.float64
        rho_h = pinned_array((n_own,), npd)          # this rank's planes only, page-locked (lbm_host_alloc)
        u_h = pinned_array((3, n_own), npd)
        sync_all()
        t0 = time.perf_counter()
        sim.init()
        sim.run(a.steps, a.steps)
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
               "inputs": "none: the lid-driven cavity is initialised on the device (kernels.cl:277-318), so no "
                         "host->device copy exists on this path",
               "note": "whole job through the C ABI: lbm_init + K iterations + blocking rho/u read-back into pinned host memory"}
    sim.sync()
    device_name = sim.device_name
    if world > 1:
        dist.barrier()
    sim.close()

    cpu = None
    extra = None
    same = None
    if rank == 0 and world == 1:
        if e2e is not None:
            e2e["with_vti"] = cli_with_vti(dim, a.precision, a.stride, a.steps, local_rank)
        if not a.no_cpu_baseline:
            r = cpu_reference_run(dim, a.precision, a.stride, 10 ** 9, 1, a.cpu_seconds)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not a.no_extra:
            extra = {}
            for key, (d_, p_, v_) in {"config4_512_f64": (512, "f64", 0), "config5_1024_f32_one_gpu": (1024, "f32", 8)}.items():
                try:
                    extra[key] = single_gpu_side_run(d_, p_, a.stride, v_, 20, 5, local_rank)
                except Exception as e:  # noqa: BLE001
                    extra[key] = {"unavailable": str(e)[:200]}
    if world > 1 and not a.no_extra and dim == 1024 and a.precision == "f32":
        # the same-workload denominator of the scaling curve: 1024^3 on ONE GPU needs the in-place kernels
        if rank == 0:
            try:
                one = single_gpu_side_run(1024, "f32", a.stride, 8, 20, 5, local_rank)
                same = {"one_gpu": one, "value": mlups / (world * one["value"]),
                        "note": "value / (N x the in-place 1024^3 rate of one GPU, measured on rank 0 in this run)"}
            except Exception as e:  # noqa: BLE001
                same = {"unavailable": str(e)[:200]}
        dist.barrier()

    if rank == 0:
        key = f"{a.precision}_{dim}_vec{vec}_fast{a.fast_math}"
        out = {
            "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {
                "workload": workload_text(dim, a.precision, a.stride)
                            + (f", {world} z-slabs of {nz} planes, halo transport {transport}" if world > 1 else ""),
                "l2": l2_text(dim, nz + (2 if world > 1 else 0), esize),
            },
            "kernel": {8: "in-place AA pattern (one lattice)", 16: "two-lattice pull, TMA-fed",
                       32: "two-lattice pull, NVRTC-specialised"}.get(a.variant, "two-lattice pull")
                      + f", {vec} cell(s)/thread, block {list(eff_block)}, "
                      + ("fast math (-o)" if a.fast_math else "strict IEEE operation order"),
            "roofline": {"bound": "hbm", "achieved": achieved / world, "peak": peak, "unit": "GB/s",
                         "frac": achieved / world / peak, "traffic": recorded_traffic(key),
                         "peak_source": peak_src, "bytes_per_cell": bpc,
                         "note": "per-GPU algorithmic bytes of the step kernel / event time; traffic = ncu "
                                 "dram__bytes of one launch (profiles/traffic.json, single-GPU captures)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.result(),
            "device": device_name,
        }
        if extra is not None:
            out["extra"] = extra
        if gate is not None:
            out["parity_gate"] = gate
        if same is not None:
            out["efficiency_same_workload"] = same
        print(json.dumps(out), flush=True)
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
    sampler = ClockSampler(0)
    g = Group(list(range(a.gpus)), dim=dim, precision=a.precision, stride=a.stride, block=block, variant=a.variant,
              fast_math=bool(a.fast_math))
    g.init()
    g.run(a.warmup, 0)
    g.sync()
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
        "config": {"workload": workload_text(dim, a.precision, a.stride) + f", {a.gpus} z-slabs in one process, halo transport "
                               "peer-stores+events",
                   "l2": l2_text(dim, dim // a.gpus + 2, esize)},
        "kernel": "two-lattice pull, strict IEEE operation order" if not a.fast_math else "two-lattice pull, -o",
        "roofline": {"bound": "hbm", "achieved": achieved / a.gpus, "peak": peak, "unit": "GB/s",
                     "frac": achieved / a.gpus / peak, "traffic": None, "peak_source": peak_src, "bytes_per_cell": bpc},
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": 2 * a.steps * a.gpus, "clocks": sampler.result(),
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
