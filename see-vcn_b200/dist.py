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


class FrameGather:
    """The one exchange at the end of the path, per pipeline batch: every rank's completed clouds and voxel tensors on
    every rank ("collect for the detector"; reference analogue: commu_utils.all_gather, pickle + pad-to-max).

    A rank packs its batch into one contiguous record

        [header: n_obj, M, frame_offset, C_vox | clustered (n_obj, S, 3) f32 | voxel_coords (M, 4) i32
         | voxel_features (M, 3) f32 | voxel_num_points (M,) i32]

    and delivers it into slot ``rank`` of every peer's receive buffer.

    backend "peer": the receive buffers live in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM
      allocations mapped into every rank of the node); a record is delivered with ONE device-to-device copy per peer
      over NVLink, issued on a side stream — the copy engines move it, no SM is taken from the persistent tcgen05
      kernels of the next batch — followed by a signal-pad barrier.  Only the bytes a batch really has travel.
    backend "nccl": all_gather_into_tensor of fixed-capacity slots (the fallback when symmetric memory is unavailable;
      also what the CPU / gloo tests run).
    Two generations of receive buffers alternate: the records of push k stay valid from ``wait()`` until the rank's
    push k + 2 is issued.  Batch indices in voxel_coords are rank-local; ``frame_offset`` travels in the header and
    ``parts()`` rebases them."""

    HEADER = 16   # int32 words

    def __init__(self, dev, world, rank, max_obj, rows_per_obj, frames, max_rows, hard=False, hard_pts=5, hard_max_vox=0,
                 backend="auto", group=None):
        self.dev, self.world, self.rank, self.group = dev, world, rank, group
        self.max_obj, self.S = max_obj, rows_per_obj
        self.vox_cap = frames * hard_max_vox if hard else None        # dynamic voxels: sized at the first push
        self.max_rows = max_rows
        self.backend = backend
        self.kind = None
        self._slot_words = None
        self._gen = 0
        self._events = [None, None]
        self._works = []
        self._side = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self.pushes = 0
        self.bytes_sent = 0
        self._hdr_ring = torch.zeros((8, self.HEADER), dtype=torch.int32)      # pinned on CUDA: the header's H2D source
        if dev.type == "cuda":
            self._hdr_ring = self._hdr_ring.pin_memory()

    # -- layout -----------------------------------------------------------------------------------------------------
    def _setup(self, m_now):
        if self.vox_cap is None:
            # every rank must agree on the capacity: 1.5 x the largest M of the first batch, rounded up to 64k rows
            t = torch.tensor([m_now], device=self.dev, dtype=torch.int64)
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.vox_cap = min(self.max_rows, ((int(t.item()) * 3 // 2 + 65535) // 65536) * 65536)
        self._slot_words = self.HEADER + self.max_obj * self.S * 3 + self.vox_cap * 8
        words = self.world * self._slot_words
        self._send = torch.empty((self._slot_words,), dtype=torch.float32, device=self.dev)
        self._hdl = None
        if self.backend in ("auto", "peer") and self.dev.type == "cuda" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm
                g = self.group if self.group is not None else dist.group.WORLD
                self._recv_flat = symm.empty((2 * words,), dtype=torch.float32, device=self.dev)
                self._hdl = symm.rendezvous(self._recv_flat, group=g)
                self._peer = [[self._hdl.get_buffer(r, (2, self.world, self._slot_words), torch.float32, 0) for r in range(self.world)]]
                import ctypes
                # this rank's slot in every peer's receive buffer, per generation, as raw pointers for seevcn_gather_broadcast
                self._peer_dst = [(ctypes.c_void_p * self.world)(*[self._peer[0][r][gen, self.rank].data_ptr() for r in range(self.world)])
                                  for gen in range(2)]
                self.kind = "peer"
            except Exception as e:   # noqa: BLE001
                if self.backend == "peer":
                    raise
                self._hdl = None
                self._why = repr(e)[:200]
        if self._hdl is None:
            self._recv_flat = torch.empty((2 * words,), dtype=torch.float32, device=self.dev)
            self.kind = "nccl" if self.dev.type == "cuda" else "gloo"
        self._recv = self._recv_flat.view(2, self.world, self._slot_words)

    def describe(self):
        return {"backend": self.kind, "slot_mb": None if self._slot_words is None else round(self._slot_words * 4 / 1e6, 2),
                "mb_sent_per_push_per_peer": round(self.bytes_sent / max(self.pushes, 1) / 1e6, 2)}

    # -- producer side ----------------------------------------------------------------------------------------------
    def push(self, out, frame_offset=0, sync=True):
        """Queue the collection of one finalized pipeline batch (dict of CompletionPipeline.run / run_stream).
        backend "peer": with ``sync`` (default) a signal-pad barrier follows the copies on the side stream, so the ranks
        stay within one push of each other — that is flow control as much as an arrival guarantee: without it
        (``sync=False``: arrival is only established by the barrier in ``wait()``) ranks drift apart and an 8-GPU node
        measured 1.62 M objects/s against 2.14 M with the barrier, the copies of the ranks that run ahead piling up
        on the ones that lag."""
        clustered = out.get("clustered", out.get("surface"))
        coords, feats, nums = out["voxel_coords"], out["voxel_features"], out["voxel_num_points"]
        n_obj, m, c = int(clustered.shape[0]), int(coords.shape[0]), int(feats.shape[1])
        if self._slot_words is None:
            self._setup(m)
        if n_obj > self.max_obj or m > self.vox_cap or c != 3:
            raise RuntimeError(f"FrameGather: batch ({n_obj} objects, {m} voxels) exceeds the slot ({self.max_obj}, {self.vox_cap})")
        g = self._gen
        self._gen ^= 1
        used = self.HEADER + n_obj * self.S * 3 + m * 8
        done = out.get("_done")
        side = self._side
        if side is not None:
            if done is not None:
                side.wait_event(done)
            else:
                side.wait_stream(torch.cuda.current_stream(self.dev))
        if self.dev.type == "cuda" and self.kind == "peer" and not sync:
            # the common case, kept short (this sits on the host's critical path once per batch): raw stream handle, two C calls
            from . import _abi
            L, st = _abi.lib(), _abi.c_void_p(side.cuda_stream)
            clustered, coords, feats, nums = clustered.contiguous(), coords.contiguous(), feats.contiguous(), nums.contiguous()
            _abi.check(L.seevcn_gather_pack(self.HEADER, n_obj, m, int(frame_offset), c, n_obj * self.S * 3, _abi.ptr(clustered),
                                            _abi.ptr(coords), _abi.ptr(feats), _abi.ptr(nums), _abi.ptr(self._send), st))
            _abi.check(L.seevcn_gather_broadcast(_abi.ptr(self._send), used * 4, self._peer_dst[g], self.world, st))
            self._unsynced = True
            self._keep = (getattr(self, "_keep", None) or [])[-3:] + [(out, None, clustered, coords, feats, nums)]
            self.pushes += 1
            self.bytes_sent += used * 4
            return g
        ctx = torch.cuda.stream(side) if side is not None else _nullctx()
        with ctx:
            send = self._send
            hdr = None
            if self.dev.type == "cuda":
                # one pack kernel + (peer) one device-to-device copy per peer, issued from C: ~15 python-level copies per
                # push cost the host 0.12 ms per batch, which is what an 8-rank node is short of
                from . import _abi
                clustered, coords, feats, nums = clustered.contiguous(), coords.contiguous(), feats.contiguous(), nums.contiguous()
                _abi.check(_abi.lib().seevcn_gather_pack(self.HEADER, n_obj, m, int(frame_offset), c, n_obj * self.S * 3,
                                                         _abi.ptr(clustered), _abi.ptr(coords), _abi.ptr(feats), _abi.ptr(nums),
                                                         _abi.ptr(send), _abi.stream()))
                if self.kind == "peer":
                    _abi.check(_abi.lib().seevcn_gather_broadcast(_abi.ptr(send), used * 4, self._peer_dst[g], self.world,
                                                                  _abi.stream()))
                    if sync:
                        self._hdl.barrier(channel=g)             # every rank's copies of this generation have landed
                    else:
                        self._unsynced = True
            else:
                hdr = self._hdr_ring[self.pushes % 8]
                hdr[0], hdr[1], hdr[2], hdr[3] = n_obj, m, frame_offset, c
                send[: self.HEADER].view(torch.int32).copy_(hdr, non_blocking=True)
                o = self.HEADER
                send[o: o + n_obj * self.S * 3].copy_(clustered.reshape(-1), non_blocking=True); o += n_obj * self.S * 3
                send[o: o + m * 4].view(torch.int32).copy_(coords.reshape(-1), non_blocking=True); o += m * 4
                send[o: o + m * 3].copy_(feats.reshape(-1), non_blocking=True); o += m * 3
                send[o: o + m].view(torch.int32).copy_(nums.reshape(-1), non_blocking=True)
            if self.kind != "peer":
                dst = self._recv[g]
                if self.dev.type == "cuda":
                    w = dist.all_gather_into_tensor(dst.view(-1), send, group=self.group, async_op=True)
                else:
                    w = dist.all_gather(list(dst.unbind(0)), send, group=self.group, async_op=True)
                self._works.append(w)
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(side)
                self._events[g] = ev
        # sources stay referenced for a few pushes: the pack runs on the side stream, behind the allocator's back
        self._keep = (getattr(self, "_keep", None) or [])[-3:] + [(out, hdr, clustered, coords, feats, nums)]
        self.pushes += 1
        self.bytes_sent += used * 4 if self.kind == "peer" else self._slot_words * 4
        return g

    def wait(self):
        """Order the current stream after every queued collection."""
        for w in self._works:
            w.wait()
        self._works = []
        if self.kind == "peer" and getattr(self, "_unsynced", False):
            with torch.cuda.stream(self._side):
                self._hdl.barrier(channel=0)                     # behind every copy this rank queued; returns when all ranks' have landed
                ev = torch.cuda.Event()
                ev.record(self._side)
            self._events[0] = ev
            self._unsynced = False
        if self._side is not None:
            for ev in self._events:
                if ev is not None:
                    torch.cuda.current_stream(self.dev).wait_event(ev)
        return self

    # -- consumer side ----------------------------------------------------------------------------------------------
    def parts(self, gen):
        """Host-side view of generation ``gen`` (synchronises): list over ranks of dicts with clustered, voxel_coords
        (batch index rebased to the global frame index), voxel_features, voxel_num_points."""
        self.wait()
        if self._side is not None:
            torch.cuda.current_stream(self.dev).synchronize()
        res = []
        for r in range(self.world):
            slot = self._recv[gen, r]
            n_obj, m, off, c = slot[:4].view(torch.int32).tolist()
            o = self.HEADER
            clustered = slot[o: o + n_obj * self.S * 3].view(n_obj, self.S, 3); o += n_obj * self.S * 3
            coords = rebase_batch_index(slot[o: o + m * 4].view(torch.int32).view(m, 4), off); o += m * 4
            feats = slot[o: o + m * 3].view(m, 3); o += m * 3
            nums = slot[o: o + m].view(torch.int32)
            res.append({"clustered": clustered, "voxel_coords": coords, "voxel_features": feats, "voxel_num_points": nums})
        return res


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
