"""Sharded runs over the GPUs of one box: one process per GPU (torchrun), `torch.distributed` for the plumbing.

The engine keeps the run state replicated on every rank, deals the chains of a generation k % world and
exchanges the new live points / covariance statistics through CUDA-IPC-mapped exchange blocks over NVLink inside
its persistent kernel (include/polychord_b200.h, pc_mgpu_*).  This module only moves the 64-byte IPC handles
between the ranks and sums the per-rank evaluation counts; it replaces the reference's MPI administrator/worker
scheme (src/polychord/mpi_utils.F90, nested_sampling.F90:262-303, 420-500).
"""
from . import _capi


def chains_of_rank(K, rank, world):
    """Chains of a generation that rank runs: k % world == rank (the device code uses the same rule)."""
    return list(range(rank, K, world))


def attach(settings, group=None, capi=_capi):
    """Create this rank's exchange block, all-gather the IPC handles, map the peers.  Collective."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world < 2:
        raise ValueError("a sharded run needs at least two ranks")
    handle = capi.mgpu_create(settings, world)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    if any(len(h) != 64 for h in handles):
        raise RuntimeError("bad IPC handle received")
    capi.mgpu_attach(rank, world, handles)
    dist.barrier(group)
    return rank, world


def detach(group=None, capi=_capi):
    import torch.distributed as dist
    dist.barrier(group)   # nobody unmaps while a peer's kernel may still be storing
    capi.mgpu_destroy()


def run_sharded(settings, group=None, capi=_capi, **run_kwargs):
    """One nested-sampling run sharded over the ranks of `group`.  Every rank must pass identical settings.
    Returns (info, dumps, nlike_total): info.nlike counts this rank's evaluations only."""
    import torch
    import torch.distributed as dist
    attach(settings, group, capi)
    try:
        info, dumps = capi.run(settings, **run_kwargs)
        t = torch.tensor([int(info.nlike)], dtype=torch.int64,
                         device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    finally:
        detach(group, capi)
    return info, dumps, int(t.item())
