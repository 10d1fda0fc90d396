"""
Index sharding of sweeps over the ranks of a ``torch.distributed`` job.

The reference treats every grid point / start time as an independent loop iteration
(reference qnmfits/qnmfits.py:1271-1281, :1391-1410), so the flat fit index is split
into one contiguous slab per rank with no data-path exchange; the only collective is
the all-gather of the per-fit mismatches (NCCL over NVLink on GPUs; gloo in the CPU
tests of this logic).  A fit's arithmetic never depends on which rank or CTA runs
it, so the gathered result is bit-identical to the single-GPU one.
"""
import os


def world():
    """(rank, world_size) of the active process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized() and \
            os.environ.get("QNMFITS_B200_NO_SHARD", "0") != "1":
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank, world_size):
    """Contiguous slab [lo, hi) of rank and the common padded slab length."""
    per = -(-int(n_items) // int(world_size)) if n_items > 0 else 0
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi, per


def all_gather_slabs(slab, n_items):
    """Concatenate equally sized (padded) 1-D slabs from all ranks; trim to n_items."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    full = torch.empty(ws * slab.numel(), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(full, slab.contiguous())
    return full[:n_items]
