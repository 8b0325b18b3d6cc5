"""Multi-GPU plumbing (one process per GPU, torch.distributed): rays are independent units, so rendering shards
contiguous row blocks with no collective and training all-reduces one flat 2.45 MB gradient buffer."""
import torch
import torch.distributed as dist


def shard_rows(height: int, rank: int, world: int):
    """Contiguous row block [row0, row0+nrows) of an H-row panorama for `rank` (sizes differ by at most one)."""
    base, rem = divmod(height, world)
    nrows = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, nrows


def all_reduce_flat_(flat_grad: torch.Tensor) -> float:
    """Sum the flat gradient buffer over ranks in place (NCCL over NVLink on GPUs, gloo in the CPU tests) and return
    the 1/world factor the Adam kernel applies (DDP semantics: mean of the per-rank mean losses, train.py:92)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0
