"""Host-side logic of the N > 1 path on CPU: slab partition and the halo exchange plan, run with
world_size 2 and 3 over gloo (one process per rank, as under torchrun)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lbmcl_b200 import slabs


def test_partition():
    assert slabs.slab_range(1024, 8, 0) == (0, 128) and slabs.slab_range(1024, 8, 7) == (896, 1024)
    assert slabs.slab_range(256, 1, 0) == (0, 256)
    covered = [slabs.slab_range(512, 4, r) for r in range(4)]
    assert [c[0] for c in covered[1:]] == [c[1] for c in covered[:-1]]
    with pytest.raises(ValueError):
        slabs.slab_range(256, 3, 0)
    assert slabs.neighbours(4, 0) == (None, 1) and slabs.neighbours(4, 3) == (2, None) and slabs.neighbours(1, 0) == (None, None)
    assert set(slabs.UP_Q) | set(slabs.DOWN_Q) == {5, 6, 11, 12, 13, 14, 15, 16, 17, 18} and len(slabs.UP_Q) == 5
    # 20 MiB per face at 1024^2 fp32, 10 MiB at 512^2 fp64
    assert slabs.halo_elems(1024) * 4 == 20 * 2 ** 20 and slabs.halo_elems(512) * 8 == 10 * 2 ** 20
    assert slabs.halo_bytes_per_step(1024, 8, 0, 4) == 20 * 2 ** 20 and slabs.halo_bytes_per_step(1024, 8, 3, 4) == 40 * 2 ** 20


def _worker(rank, world, port, dim, steps):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = slabs.halo_elems(dim)
        lo, hi = slabs.neighbours(world, rank)
        for step in range(steps):
            # payload identifies (sender, face, step, element)
            def payload(r, face):
                return torch.arange(n, dtype=torch.float64) + 1e6 * r + 1e5 * face + 1e3 * step
            send = [payload(rank, 0) if lo is not None else None, payload(rank, 1) if hi is not None else None]
            recv = [torch.full((n,), -1.0, dtype=torch.float64) if lo is not None else None,
                    torch.full((n,), -1.0, dtype=torch.float64) if hi is not None else None]
            for r in slabs.exchange_halos(send, recv, world, rank):
                r.wait()
            if lo is not None:   # my low face receives what the rank below sent from its HIGH face
                assert torch.equal(recv[0], payload(lo, 1)), (rank, step)
            if hi is not None:
                assert torch.equal(recv[1], payload(hi, 0)), (rank, step)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_plan_over_gloo(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, 16, 3), nprocs=world, join=True)
