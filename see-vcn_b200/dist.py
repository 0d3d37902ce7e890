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


class PaddedGather:
    """Result of ``all_gather_padded``: ``out`` (world, capacity, ...) and ``counts`` (world,) int64, both on the
    tensor's device; valid after ``wait()`` (which only orders the current stream after the collective)."""

    def __init__(self, out, counts, works, keep):
        self.out, self.counts, self._works, self._keep = out, counts, works, keep

    def wait(self):
        for w in self._works:
            if w is not None:
                w.wait()
        self._works, self._keep = [], None
        return self

    def parts(self):
        """Host-side view: list of per-rank tensors (reads the counts -> synchronises)."""
        self.wait()
        return [self.out[r, : int(c)] for r, c in enumerate(self.counts.tolist())]


def all_gather_padded(t, capacity, group=None, async_op=False):
    """The streaming form of ``all_gather_v``: no host synchronisation.  Every rank contributes ``t`` (n_r <= capacity
    rows) inside a fixed-capacity slot, so the collective sizes are static and the per-rank counts travel as a device
    tensor next to the payload.  With ``async_op`` the collectives run on the process group's stream and the caller's
    stream is only ordered behind them in ``wait()`` — the gather of batch i overlaps the kernels of batch i+1."""
    n = t.shape[0]
    assert n <= capacity, f"{n} rows > capacity {capacity}"
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    cnt = torch.full((1,), n, device=t.device, dtype=torch.int64)
    if world == 1:
        out = torch.empty((1, capacity) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        out[0, :n] = t
        return PaddedGather(out, cnt, [], None)
    pad = torch.empty((capacity,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[:n] = t
    out = torch.empty((world, capacity) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    counts = torch.empty((world,), device=t.device, dtype=torch.int64)
    if t.is_cuda:
        w1 = dist.all_gather_into_tensor(counts, cnt, group=group, async_op=async_op)
        w2 = dist.all_gather_into_tensor(out.view(world * capacity, *t.shape[1:]), pad, group=group, async_op=async_op)
    else:   # gloo (CPU tests)
        w1 = dist.all_gather(list(counts.split(1)), cnt, group=group, async_op=async_op)
        w2 = dist.all_gather(list(out.unbind(0)), pad, group=group, async_op=async_op)
    return PaddedGather(out, counts, [w1, w2] if async_op else [], (pad, cnt))


def rebase_batch_index(coords, frame_offset):
    """voxel_coords (M,4) [b,z,y,x] with a rank-local batch index -> global frame index."""
    out = coords.clone()
    out[:, 0] += frame_offset
    return out
