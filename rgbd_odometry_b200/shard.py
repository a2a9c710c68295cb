"""Frame-pair partitioning across ranks and the pose gather: each rank owns a contiguous block of pair
indices, no image data crosses GPUs, and the only collective is one all-gather of the 12-double pose records."""
import numpy as np


def partition(total_pairs, world_size, rank):
    """Contiguous block [start, start+count) of pair indices owned by `rank` (blocks differ by at most one pair)."""
    base, rem = divmod(total_pairs, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def gather_poses(local_poses, total_pairs, group=None):
    """All-gather per-pair pose records (count x 12 float64) into the (total_pairs x 12) array, in pair order.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors); ragged blocks are padded to the largest block."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = local_poses if isinstance(local_poses, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_poses, np.float64))
    counts = [partition(total_pairs, world, r)[1] for r in range(world)]
    mx = max(counts)
    pad = torch.zeros((mx, 12), dtype=torch.float64, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty((world * mx, 12), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * mx: r * mx + counts[r]] for r in range(world)]
    return torch.cat(parts, 0)
