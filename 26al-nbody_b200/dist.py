"""Multi-GPU plumbing: one process per GPU, `torch.distributed` only for the bootstrap.

The data path is NCCL called from inside libal26b200.so (all-gather of the predicted j-set and an
8-byte min-reduce of the next block time, both captured in the block-step CUDA graph); this module
just gets rank 0's ncclUniqueId to every rank and tells the context its rank.  Stands in for
`number_of_workers=8` MPI ranks of the reference's worker (al26_nbody.py:57,1711-1720).
"""
import os

from . import _lib


def slice_of(n, rank, world):
    """Contiguous index range [i0, i1) owned by `rank` -- must match slice_of() in csrc/api.cu."""
    return n * rank // world, n * (rank + 1) // world


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_unique_id(make_uid, rank, device="cpu", group=None):
    """rank 0 calls make_uid() -> 128 bytes; every rank returns those bytes.  Works on any
    torch.distributed backend (gloo with CPU tensors, nccl with CUDA tensors)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = make_uid()
        if len(uid) != 128:
            raise ValueError("ncclUniqueId must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(buf, 0, group=group)
    return buf.cpu().numpy().tobytes()


def allgather_bytes(blob, world, device="cpu", group=None):
    """All-gather one fixed-size byte string per rank (the CUDA IPC handles of the staging slabs)."""
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return [p.cpu().numpy().tobytes() for p in parts]


def init_context(ctx, rank, world, device="cuda", group=None, mode="p2p", split_min=0):
    """Join `ctx` (an al26 Context on this rank's GPU) to the job: NCCL communicator (energies, and the
    all-gather path when mode == "nccl") plus, in the default peer-memory mode, the hook that swaps the
    staging slabs' IPC handles after every commit."""
    if world == 1:
        return ctx
    uid = broadcast_unique_id(_lib.dist_unique_id, rank, device=device, group=group)
    ctx.dist_init(rank, world, uid, mode=mode, split_min=split_min,
                  exchange=lambda blob: allgather_bytes(blob, world, device=device, group=group))
    return ctx
