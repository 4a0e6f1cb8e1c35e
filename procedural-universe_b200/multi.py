"""Host-side plumbing for one-process-per-GPU runs (torch.distributed).  No compute here.

The data path has exactly one exchange per step -- the position all-gather that the library runs
itself over NCCL (csrc/nccl_dl.cpp).  torch.distributed only carries what a launcher must carry:
the 128-byte NCCL unique id from rank 0 to everyone, barriers, and the max-over-ranks of timings.
The same helpers run on the gloo backend, which is how they are tested without GPUs.
"""
import ctypes as C
import os

from . import binding


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n, rank, world):
    """[first, first+count) owned by `rank` -- nb_shard_range in the C ABI."""
    first, count = C.c_size_t(), C.c_size_t()
    L = binding.load()
    L.nb_shard_range.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    rc = L.nb_shard_range(n, rank, world, C.byref(first), C.byref(count))
    if rc != 0:
        raise binding.NBodyError(L.nb_last_error().decode())
    return first.value, count.value


def broadcast_bytes(payload, src=0):
    """Every rank gets rank `src`'s bytes object (used for the NCCL unique id)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def max_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_rows(local_rows, n_total, rank, world):
    """In-place all-gather layout of the library's exchange, for host arrays: every rank contributes
    the rows it owns and receives everybody else's, concatenated in rank order."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_rows
    pieces = [None] * world
    dist.all_gather_object(pieces, np.ascontiguousarray(local_rows))
    # not np.concatenate: it repacks a padded structured dtype (104-byte Particle -> 100 bytes)
    out = np.empty(n_total, dtype=local_rows.dtype)
    at = 0
    for piece in pieces:
        out[at:at + len(piece)] = piece
        at += len(piece)
    assert at == n_total
    return out


def connect(sim, rank):
    """Creates the library-owned NCCL communicator on every rank."""
    uid = binding.Sim.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, src=0)
    sim.comm_init(uid)


def connect_p2p(sim, rank, world):
    """Fused kick-drift + exchange over NVLink peer memory: all-gather the CUDA IPC handles in rank
    order (torch.distributed carries 384 opaque bytes per rank), then attach."""
    import torch.distributed as dist
    mine = sim.p2p_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    sim.p2p_attach(handles)
    dist.barrier()
