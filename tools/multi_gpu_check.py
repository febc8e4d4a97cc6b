#!/usr/bin/env python3
"""Run under torchrun: z-slab run over WORLD_SIZE GPUs (one process each, NCCL halo exchange, the
same code path as bench.py) compared bit for bit with the oracle on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py"""
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


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for transport, precision, dim, stride, its in (("host", "f32", 32, 32, 10), ("host", "f64", 64, 32, 6),
                                                   ("nccl-in-library", "f32", 32, 32, 10),
                                                   ("nccl-in-library", "f64", 64, 32, 7),
                                                   ("nccl-in-library", "f32", 128, 32, 20),
                                                   ("fused", "f32", 32, 32, 10), ("fused", "f64", 64, 32, 7),
                                                   ("fused", "f32", 128, 32, 21)):
        z0, z1 = slab_range(dim, world, rank)
        sim = Simulation(dim=dim, precision=precision, stride=stride, device=lr, z_range=(z0, z1))
        main_s = torch.cuda.Stream(device=dev)
        sim.set_stream(main_s.cuda_stream)
        has_lo, has_hi = rank > 0, rank < world - 1
        ts = "<f4" if precision == "f32" else "<f8"
        n_h = sim.halo_elems
        send = [torch.as_tensor(DevBuf(sim.halo_send_ptr(f), n_h, ts), device=dev) if ok_ else None
                for f, ok_ in ((0, has_lo), (1, has_hi))]
        recv = [torch.as_tensor(DevBuf(sim.halo_recv_ptr(f), n_h, ts), device=dev) if ok_ else None
                for f, ok_ in ((0, has_lo), (1, has_hi))]
        sim.init()
        if transport == "host":
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
            # library-driven: NCCL communicator inside the context, overlapped schedule in lbm_run
            used = connect_slabs(sim, rank, world, dev, fused=(transport == "fused"))
            if rank == 0:
                print(f"  transport requested {transport}: in use {used}")
            sim.run(its - 3, its)
            sim.run(3, its)
        n = dim ** 3
        npd = np.float32 if precision == "f32" else np.float64
        rho = np.full(n, np.nan, dtype=npd)
        u = np.full((3, n), np.nan, dtype=npd)
        sim.read_macros(rho, u)
        sim.close()
        td = torch.float32 if precision == "f32" else torch.float64
        # gather the slabs on rank 0 (NaN-safe: ship raw bits)
        bits = torch.from_numpy(np.concatenate([rho, u.reshape(-1)]).view(np.uint8)).to(dev)
        out = [torch.empty_like(bits) for _ in range(world)] if rank == 0 else None
        dist.gather(bits, out, dst=0)
        if rank == 0:
            from oracle import Oracle
            exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, its)
            full_rho = np.full(n, np.nan, dtype=npd)
            full_u = np.full((3, n), np.nan, dtype=npd)
            for r in range(world):
                a = out[r].cpu().numpy().view(npd)
                a0, a1 = slab_range(dim, world, r)
                sl = slice(a0 * dim * dim, a1 * dim * dim)
                full_rho[sl] = a[:n][sl]
                full_u[:, sl] = a[n:].reshape(3, n)[:, sl]
            same = full_rho.tobytes() == exp["rho"][1].tobytes() and full_u.tobytes() == exp["u"][1].tobytes()
            print(f"multi_gpu_check [{transport}] {precision} {dim}^3 x{its} on {world} ranks: {'bit-identical' if same else 'MISMATCH'}")
            ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
