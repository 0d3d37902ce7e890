"""Frame sharding over ranks and the one exchange at the end of the path.

The hot path has no collective inside it: frames are independent (the reference runs them in a
process pool, see/surface_completion/sc_multiproc.py:81-85).  Each rank takes a contiguous frame
range; when the consumer wants every rank's results, ``all_gather_v`` collects the ragged per-rank
tensors (completed clouds, voxel tensors) with one count all-gather + one padded all-gather over
NCCL — the role pcdet.utils.commu_utils.all_gather (detector3d/pcdet/utils/commu_utils.py:50-111)
plays for pickled results in the reference.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous range [lo, hi) of items owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_v(t, group=None):
    """Concatenate ``t`` (n_r, ...) from every rank along dim 0 -> (sum n_r, ...), plus the counts list."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t, [t.shape[0]]
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[: t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, pad, group=group) if t.is_cuda else dist.all_gather(list(out.chunk(world)), pad, group=group)
    parts = [out[r * mx: r * mx + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0), counts


def rebase_batch_index(coords, frame_offset):
    """voxel_coords (M,4) [b,z,y,x] with a rank-local batch index -> global frame index."""
    out = coords.clone()
    out[:, 0] += frame_offset
    return out
