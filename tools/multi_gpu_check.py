#!/usr/bin/env python3
"""Run under torchrun: z-slab runs over WORLD_SIZE ranks (one process each; the same code path as bench.py)
compared bit for bit with the CPU oracle on rank 0, for every transport of include/lbm_b200.h:

    host     split-phase ABI, dense halos exchanged by the host through torch.distributed   (2a)
    dense    library-owned NCCL communicator, dense halos                                    (2b)
    flags    CUDA IPC peer stores + in-kernel epoch flags, one launch per iteration          (2c)
    token    CUDA IPC peer stores + NCCL token                                               (2d)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [--backend nccl|gloo] [--same-device] [--only flags]

`--backend gloo --same-device` puts every rank on GPU 0 and bootstraps over gloo: only the NCCL-free
transport ("flags") can run then -- the ranks' kernels are time-sliced on the one GPU and meet through the
flag words; used by tests/test_gpu_multiproc.py on a one-GPU box.  Exit status 1 on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lbmcl_b200.capi import Simulation  # noqa: E402
from lbmcl_b200.slabs import connect_slabs, exchange_halos, slab_range  # noqa: E402


class DevBuf:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


CASES = (("host", "f32", 32, 32, 10), ("host", "f64", 64, 32, 6),
         ("dense", "f32", 32, 32, 10), ("dense", "f64", 64, 32, 7), ("dense", "f32", 128, 32, 20),
         ("token", "f32", 32, 32, 10), ("token", "f64", 64, 32, 7), ("token", "f32", 128, 32, 21),
         ("flags", "f32", 32, 32, 10), ("flags", "f64", 64, 32, 7), ("flags", "f32", 128, 32, 21),
         ("flags", "f32", 64, 4096, 9),      # DIM < stride < cells: the row-base addressing on a slab
         ("flags", "f64", 64, 1, 6))         # AoS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if a.same_device:
        lr = 0
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
        where = dev
    else:
        dist.init_process_group("gloo")
        where = torch.device("cpu")
    ok = True
    for transport, precision, dim, stride, its in CASES:
        if a.only and transport not in a.only.split(","):
            continue
        if a.backend == "gloo" and transport != "flags":
            continue
        if dim % world != 0 or dim // world < 1:
            continue
        z0, z1 = slab_range(dim, world, rank)
        sim = Simulation(dim=dim, precision=precision, stride=stride, device=lr, z_range=(z0, z1))
        main_s = torch.cuda.Stream(device=dev)
        sim.set_stream(main_s.cuda_stream)
        has_lo, has_hi = rank > 0, rank < world - 1
        if transport == "host":
            ts = "<f4" if precision == "f32" else "<f8"
            n_h = sim.halo_elems
            send = [torch.as_tensor(DevBuf(sim.halo_send_ptr(f), n_h, ts), device=dev) if ok_ else None
                    for f, ok_ in ((0, has_lo), (1, has_hi))]
            recv = [torch.as_tensor(DevBuf(sim.halo_recv_ptr(f), n_h, ts), device=dev) if ok_ else None
                    for f, ok_ in ((0, has_lo), (1, has_hi))]
            sim.init()
            # split-phase ABI, exchange posted by the host through torch.distributed
            with torch.cuda.stream(main_s):
                for it in range(1, its + 1):
                    sim.step_planes(z0, z1, it == its)
                    sim.advance()
                    sim.halo_pack()
                    for r in exchange_halos(send, recv, world, rank):
                        r.wait()
                    sim.halo_unpack()
        else:
            used = connect_slabs(sim, rank, world, dev, transport=transport)
            if rank == 0:
                print(f"  transport requested {transport}: in use {used}", flush=True)
            sim.init()
            sim.run(its - 3, its)
            sim.step(False)        # the per-launch entry point goes through the same schedule
            sim.run(2, its)
        rho, u = sim.read_macros_slab()
        sim.sync()
        # gather the slabs on rank 0 (NaN-safe: ship raw bits)
        bits = torch.from_numpy(np.concatenate([rho, u.reshape(-1)]).view(np.uint8).copy()).to(where)
        out = [torch.empty_like(bits) for _ in range(world)] if rank == 0 else None
        dist.gather(bits, out, dst=0)
        dist.barrier()             # nobody frees a lattice a neighbour may still be storing into
        sim.close()
        if rank == 0:
            from oracle import Oracle
            npd = np.float32 if precision == "f32" else np.float64
            exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, its)
            n_slab = (dim // world) * dim * dim
            parts = [o.cpu().numpy().view(npd) for o in out]
            full_rho = np.concatenate([p[:n_slab] for p in parts])
            full_u = np.concatenate([p[n_slab:].reshape(3, n_slab) for p in parts], axis=1)
            same = full_rho.tobytes() == exp["rho"][1].tobytes() and full_u.tobytes() == exp["u"][1].tobytes()
            print(f"multi_gpu_check [{transport}] {precision} {dim}^3 stride {stride} x{its} on {world} ranks"
                  f"{' (one GPU, time-sliced)' if a.same_device else ''}: {'bit-identical' if same else 'MISMATCH'}",
                  flush=True)
            ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
