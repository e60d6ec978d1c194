"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Walkers shard trivially: rank r owns global walkers [r*M, (r+1)*M) and the Philox subsequence is the
GLOBAL walker id, so results do not depend on the number of ranks.  The only
exchange on the path is shared-hill metadynamics: every deposition stride each rank writes its M new
hill records into its slot of the next table segment and the segment is completed by ONE in-place
allgather (M x hill_stride reals per rank).
"""
import os

import numpy as np


def shard_config(cfg: dict, n_walkers_per_rank: int = None) -> dict:
    """Fill device / walker_offset / n_walkers_global from the torchrun environment."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    out = dict(cfg)
    M = int(n_walkers_per_rank or out.get("n_walkers", 1))
    out.update(n_walkers=M, walker_offset=rank * M, n_walkers_global=world * M, device=local_rank)
    return out


class _DevSegment:
    """__cuda_array_interface__ view of raw device memory (lets torch wrap the engine's hill table)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2}


def exchange_segment(seg, bytes_per_rank, rank_offset_bytes, group=None):
    """In-place allgather of one hill segment.  ``seg`` is a uint8 tensor of ``world * bytes_per_rank``
    bytes (device memory under NCCL, host memory under gloo in the CPU tests); this rank's records
    already sit at ``rank_offset_bytes``.  Afterwards every rank holds every rank's records, in global
    walker order -- the layout the step kernel reads."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert seg.numel() == world * bytes_per_rank
    assert rank_offset_bytes == dist.get_rank(group) * bytes_per_rank
    mine = seg[rank_offset_bytes:rank_offset_bytes + bytes_per_rank].clone()
    dist.all_gather_into_tensor(seg, mine, group=group)
    return seg


def attach_hill_exchange(engine, group=None):
    """Complete each new hill segment with torch.distributed.all_gather_into_tensor (NCCL)."""
    import torch
    import torch.distributed as dist

    real = torch.float32 if engine.cfg.precision == 0 else torch.float64

    def exchange(ptr, bytes_per_rank, rank_offset_bytes):
        world = dist.get_world_size(group)
        if rank_offset_bytes < 0:  # mtd_grid: sum the grid increments of all ranks in place
            buf = torch.as_tensor(_DevSegment(ptr, bytes_per_rank), device="cuda").view(real)
            dist.all_reduce(buf, group=group)
        else:
            seg = torch.as_tensor(_DevSegment(ptr, bytes_per_rank * world), device="cuda")
            exchange_segment(seg, bytes_per_rank, rank_offset_bytes, group)
        torch.cuda.current_stream().synchronize()
        return 0

    engine.set_exchange(exchange)
    return engine


def attach_native_hill_exchange(engine, group=None):
    """Native NCCL exchange (``nds_set_comm``): rank 0 creates an ncclUniqueId, torch.distributed only carries
    its 128 bytes to the other ranks; from then on the engine runs its allgather / all-reduce itself, on its
    own stream, with no host synchronisation and no Python in the deposition path."""
    import torch.distributed as dist
    from .engine import nccl_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.set_comm(unique_id=box[0], rank=rank, nranks=world)
    return engine


def attach_peer_hill_exchange(engine, group=None, device_barrier=None):
    """NVLink peer-memory exchange: every rank maps every other rank's hill table + exchange window (CUDA IPC)
    and the deposition kernel stores new records (or, with mtd_grid, its grid increment) straight into all of
    them.  The barrier between the ranks runs on the device (one warp signalling / polling flags in peer memory)
    when every rank has a GPU of its own; ranks sharing a GPU fall back to a 1-element all-reduce on the host."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, engine.hills_ipc_handle(), group=group)
    engine.open_peer_tables(handles, rank)
    token = torch.zeros(1, device="cuda")
    if device_barrier is None:
        ids = [None] * world
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        dist.all_gather_object(ids, (os.uname().nodename, str(getattr(props, "uuid", torch.cuda.current_device()))), group=group)
        device_barrier = len(set(ids)) == world and os.environ.get("NDS_HOST_BARRIER") != "1"
    if device_barrier:
        dist.barrier(group=group)  # every rank has mapped every window before the first flag is written
        engine.set_peer_barrier(True)
        return engine

    def barrier(ptr, bytes_per_rank, rank_offset_bytes):
        assert bytes_per_rank == 0
        dist.all_reduce(token, group=group)
        torch.cuda.current_stream().synchronize()
        return 0

    engine.set_exchange(barrier)
    return engine


def gather_commit_counts(basin: np.ndarray, n_basins: int, group=None):
    """Sum of per-basin commit counts over ranks (the only reduction C3 needs)."""
    import torch
    import torch.distributed as dist
    counts = np.array([(basin == b).sum() for b in range(-1, n_basins)], dtype=np.int64)
    if dist.is_available() and dist.is_initialized():
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(counts).to(dev)
        dist.all_reduce(t, group=group)
        counts = t.cpu().numpy()
    return counts
