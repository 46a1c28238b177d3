"""Multi-GPU host plumbing: column shards of the core alignment + the one
collective the path needs (sum of per-pair partial core counts, made by the library over NCCL).

One process per GPU (torchrun); every rank holds all individuals for its
[site_begin, site_end) slice and a replica of the (small) accessory matrix.
Parent draws, accessory flips and HGT come from Philox keyed by
(seed, generation, individual, column block), so every rank computes the same
parents and accessory state without communicating; the generation step has no
collective. The distance pass all-reduces the per-pair partial core counts.
"""
from __future__ import annotations

import numpy as np

SITE_ALIGN = 8192          # PANSIM_SITE_ALIGN in include/pansim_b200.h


def column_shards(core_size: int, world: int) -> list[tuple[int, int]]:
    """Split [0, core_size) into `world` contiguous slices whose starts are multiples of
    SITE_ALIGN, as even as the alignment allows (trailing ranks may be empty)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    regions = (core_size + SITE_ALIGN - 1) // SITE_ALIGN
    base, extra = divmod(regions, world)
    out, r0 = [], 0
    for rank in range(world):
        nr = base + (1 if rank < extra else 0)
        b = min(core_size, r0 * SITE_ALIGN)
        e = min(core_size, (r0 + nr) * SITE_ALIGN)
        out.append((b, e))
        r0 += nr
    return out


def allreduce_pair_counts(partial_core_diff, group=None):
    """Sum the per-pair partial core counts over the column shards (in place for tensors).
    Accepts a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array (gloo)."""
    import torch
    import torch.distributed as dist
    if isinstance(partial_core_diff, np.ndarray):
        t = torch.from_numpy(partial_core_diff.astype(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy().astype(np.uint32)
    dist.all_reduce(partial_core_diff, op=dist.ReduceOp.SUM, group=group)
    return partial_core_diff


def broadcast_unique_id(rank: int, group=None) -> bytes:
    """Host plumbing for pansim_comm_init_rank: rank 0 creates the NCCL unique id (128 bytes), the
    process group carries it to the other ranks (any backend; gloo on CPU works)."""
    import torch
    import torch.distributed as dist
    from .population import Pansim
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(Pansim.comm_unique_id()), dtype=torch.uint8).clone()
    dev = None
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class ShardedPansim:
    """This rank's shard of a column-sharded run (wraps `Pansim`). The one collective of the path --
    the sum of the per-pair partial core counts -- runs inside the library (NCCL communicator created
    with pansim_comm_init_rank); torch.distributed only carries the 128-byte unique id."""

    def __init__(self, params, rank: int, world: int, device: int = 0, group=None):
        from .population import Pansim
        self.rank, self.world = rank, world
        self.shards = column_shards(params.core_size, world)
        b, e = self.shards[rank]
        if world == 1:
            b, e = 0, 0
        self.sim = Pansim.from_params(params, device=device, site_begin=b, site_end=e)
        if world > 1:
            self.sim.comm_init_rank(world, rank, broadcast_unique_id(rank, group))

    def __getattr__(self, name):
        return getattr(self.sim, name)

    def pairwise_distances(self, range1, range2):
        cd, it, un = self.sim.pair_counts(range1, range2)        # whole-alignment counts on every rank
        return self.sim.distances_from_counts(cd, it, un)
