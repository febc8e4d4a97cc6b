"""z-slab decomposition: host-side plan for the one-process-per-GPU mode (new functionality; the
reference is single-device).

The cube is cut into `world` slabs of DIM/world planes.  Per iteration each rank sends, per interior
face, the 5 populations that cross it (packed densely by lbm_halo_pack: [5][DIM][DIM]) and receives
the neighbour's; nothing else is communicated and there is no collective on the data path.
Transport is torch.distributed point-to-point (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

# populations with e_z = +1 / -1 (reference kernels.cl:151-205)
UP_Q = (6, 15, 16, 17, 18)
DOWN_Q = (5, 11, 12, 13, 14)
FACE_LOW, FACE_HIGH = 0, 1


def slab_range(dim: int, world: int, rank: int):
    """Owned global planes [z0, z1) of `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    if dim % world != 0:
        raise ValueError(f"dim {dim} is not divisible into {world} z-slabs")
    nz = dim // world
    return rank * nz, (rank + 1) * nz


def neighbours(world: int, rank: int):
    """(rank below or None, rank above or None)"""
    return (rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None)


def halo_elems(dim: int) -> int:
    return 5 * dim * dim


def halo_bytes_per_step(dim: int, world: int, rank: int, itemsize: int) -> int:
    lo, hi = neighbours(world, rank)
    return sum(1 for n in (lo, hi) if n is not None) * halo_elems(dim) * itemsize


def exchange_halos(send, recv, world: int, rank: int, group=None):
    """Post this rank's sends/receives as one batch (ncclGroupStart/End under NCCL) and return the
    request handles.  `send` / `recv` are [low, high] lists of 1-D tensors (None on a cube face).

    My HIGH face talks to the LOW face of rank+1 and vice versa; posting everything in one batch makes
    the order of the operations irrelevant, so no rank can deadlock on its neighbour."""
    import torch.distributed as dist

    lo, hi = neighbours(world, rank)
    ops = []
    if hi is not None:
        ops.append(dist.P2POp(dist.isend, send[FACE_HIGH], hi, group))
        ops.append(dist.P2POp(dist.irecv, recv[FACE_HIGH], hi, group))
    if lo is not None:
        ops.append(dist.P2POp(dist.isend, send[FACE_LOW], lo, group))
        ops.append(dist.P2POp(dist.irecv, recv[FACE_LOW], lo, group))
    return dist.batch_isend_irecv(ops) if ops else []


TRANSPORTS = ("flags", "token", "dense")


def connect_slabs(sim, rank: int, world: int, device, transport: str = "flags") -> str:
    """Connect the slab context `sim` to its neighbours in the other ranks.  torch.distributed (whatever
    backend the process group has: NCCL on the GPUs, gloo works too) is only the bootstrap channel.

      "flags"  CUDA IPC peer stores + in-kernel epoch flags: one launch per iteration, no NCCL at all
               (include/lbm_b200.h transport 2c);
      "token"  CUDA IPC peer stores + a library-owned NCCL communicator carrying one word per face (2d);
      "dense"  packed halos over the library-owned NCCL communicator (2b).

    If a rank cannot attach a neighbour (no peer access), every rank detaches again and the dense transport is
    used.  Returns the transport in use: "peer-stores+flags", "peer-stores+token" or "nccl-dense"."""
    import torch
    import torch.distributed as dist
    from .capi import FUSED_FLAGS, FUSED_TOKEN, LbmError, Simulation

    if transport not in TRANSPORTS:
        raise ValueError(f"transport {transport!r}: expected one of {TRANSPORTS}")
    where = device if dist.get_backend() == "nccl" else torch.device("cpu")

    def comm_init():
        uid = torch.zeros(128, dtype=torch.uint8, device=where)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Simulation.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        sim.comm_init(uid.cpu().numpy().tobytes(), rank, world)

    if world < 2:
        return "single"
    if transport == "dense":
        comm_init()
        return "nccl-dense"
    if transport == "token":
        comm_init()
    ok = 1
    try:
        blob = torch.frombuffer(bytearray(sim.ipc_export()), dtype=torch.uint8).to(where)
    except LbmError:
        blob = torch.zeros(Simulation.IPC_BYTES, dtype=torch.uint8, device=where)
        ok = 0
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    lo, hi = neighbours(world, rank)
    if ok:
        try:
            if lo is not None:
                sim.ipc_attach(FACE_LOW, blobs[lo].cpu().numpy().tobytes())
            if hi is not None:
                sim.ipc_attach(FACE_HIGH, blobs[hi].cpu().numpy().tobytes())
        except LbmError:
            ok = 0
    flag = torch.tensor([ok], device=where, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        sim.comm_fused(FUSED_FLAGS if transport == "flags" else FUSED_TOKEN)
        return "peer-stores+flags" if transport == "flags" else "peer-stores+token"
    # some rank could not attach: nobody keeps a mapping (a half-attached slab must not store into its
    # neighbour while the dense halos travel as well), everybody falls back to the dense transport
    sim.ipc_detach()
    if transport == "flags":
        comm_init()
    return "nccl-dense"
