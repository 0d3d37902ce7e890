"""Frame-level driver of the hot path: crop -> resample -> VCN -> kNN surface -> voxelize.

Mirrors what the reference spreads over ``SEE_VCN.isolate_gt_pts`` / ``complete_gt_pts``
(see/surface_completion/SEE_VCN.py:61-115), ``VCN.inference`` (see/surface_completion/models/VCN.py:43-104)
and the detector's voxelization front end (detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:37-76),
but keeps every tensor on the device between stages: the reference goes through open3d crops on
the host, a D2H copy + cKDTree per object and a .pcd file.
"""
import numpy as np
import torch

from .pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils as roi
from .pcdet.models.backbones_3d.vfe.dynamic_mean_vfe import dynamic_voxelize
from .see.surface_completion.models.vcn.models.build import MODELS
from .see.surface_completion.models.vcn.utils.sampling import get_partial_mesh_batch, get_largest_cluster_batch

WAYMO_VOXEL_CFG = ([-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], [0.1, 0.1, 0.15], [1504, 1504, 40])   # sc_waymo_dataset.yaml:4,39-45


def resample_choice(count, n_points, rng):
    """Indices into the tiled cloud, as ResamplePoints draws them (data_transforms.py:254-262)."""
    reps = int(np.ceil(n_points / count))
    return rng.permutation(reps * count)[:n_points].astype(np.int32)


class CompletionPipeline:
    def __init__(self, model_name="VCN_VC", state_dict=None, device=None, sel_k=10, min_lidar_pts=30, resample_num=1024,
                 voxel_cfg=WAYMO_VOXEL_CFG, precision="bf16", host_rng=False, cluster_eps=None):
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model = MODELS.build({"NAME": model_name}, precision=precision)
        self.state_dict = state_dict
        if state_dict is not None:
            self.model.load_state_dict({k.replace("module.", ""): v for k, v in state_dict.items()})
        self.model.to(self.device).eval()
        self.model_name = model_name
        self.sel_k = sel_k                    # SURFACE_COMPLETION.VCN.SEL_K_NEAREST
        self.min_lidar_pts = min_lidar_pts    # SURFACE_COMPLETION.MIN_LIDAR_PTS
        self.resample_num = resample_num
        self.voxel_cfg = voxel_cfg
        self.host_rng = host_rng              # True: numpy permutation per object on the host (reference's draw)
        self.cluster_eps = cluster_eps        # SURFACE_COMPLETION.VCN.CLUSTER_EPS; None skips the largest-cluster filter

    @torch.no_grad()
    def complete(self, points, boxes, seed=0):
        """points (F,P,3), boxes (F,T,7) CUDA -> dict with the per-object clouds.  One host sync (box counts)."""
        F, P, _ = points.shape
        T = boxes.shape[1]
        idx, counts, offsets, lists = roi.crop_points_in_boxes(points, boxes)
        cnt = counts.cpu().numpy()                                   # F*T ints: the only D2H before the results
        keep = np.argwhere(cnt >= self.min_lidar_pts)                # SEE_VCN.py:71
        out = {"box_idxs_of_pts": idx, "box_counts": counts, "obj_frame": keep[:, 0].astype(np.int32),
               "obj_box": keep[:, 1].astype(np.int32)}
        if len(keep) == 0:
            empty = torch.empty((0, self.resample_num, 3), device=points.device)
            out.update(input=empty, coarse=empty, surface=empty)
            return out
        O = len(keep)
        if self.host_rng:
            # the reference's own draw (numpy permutation on the host), for bit-for-bit comparisons
            rng = np.random.default_rng(seed)
            choice = np.stack([resample_choice(int(cnt[f, k]), self.resample_num, rng) for f, k in keep])
            h = torch.from_numpy(np.concatenate([keep.astype(np.int32).T.reshape(-1), choice.reshape(-1)])).pin_memory()
            d = h.to(points.device, non_blocking=True)
            obj_frame, obj_box, d_choice = d[:O], d[O:2 * O], d[2 * O:].view(O, self.resample_num)
            inp = roi.resample_gather(points, counts, offsets, lists, obj_frame.contiguous(), obj_box.contiguous(),
                                      d_choice.contiguous())
        else:
            d = torch.from_numpy(np.ascontiguousarray(keep.astype(np.int32).T)).pin_memory().to(points.device, non_blocking=True)
            obj_frame, obj_box = d[0], d[1]
            inp = roi.resample_gather_rng(points, counts, offsets, lists, obj_frame, obj_box, self.resample_num, seed)
        in_dict = {"input": inp}
        if self.model_name == "VCN_CN":
            in_dict["gt_boxes"] = boxes[obj_frame.long(), obj_box.long()].contiguous()
        coarse = self.model(in_dict)["coarse"]
        surface, sel_count = get_partial_mesh_batch(inp, coarse, k=self.sel_k, surface_pts=self.resample_num,
                                                    return_count=True)
        out.update(input=inp, coarse=coarse, surface=surface, surface_count=sel_count, obj_frame_dev=obj_frame)
        if self.cluster_eps is not None:   # models/VCN.py:95-98
            out["clustered"] = get_largest_cluster_batch(surface, eps=self.cluster_eps, min_points=2,
                                                         total_pts=coarse.shape[1], period=sel_count)
        return out

    @torch.no_grad()
    def run(self, points, boxes, seed=0):
        """complete() + dynamic voxelization of [frame points ++ completed surfaces] -> detector input."""
        out = self.complete(points, boxes, seed)
        F, P, _ = points.shape
        fid = torch.arange(F, device=points.device, dtype=torch.float32).view(F, 1, 1).expand(F, P, 1)
        rows = [torch.cat((fid, points), dim=2).view(F * P, 4)]
        completed = out.get("clustered", out["surface"])
        if completed.shape[0] > 0:
            O, S, _ = completed.shape
            ofid = out["obj_frame_dev"].to(torch.float32).view(O, 1, 1).expand(O, S, 1)
            rows.append(torch.cat((ofid, completed), dim=2).view(O * S, 4))
        vox_pts = torch.cat(rows, dim=0).contiguous()
        coords, feats, nums = dynamic_voxelize(vox_pts, *self.voxel_cfg, sort=True, batch_size=F)
        out.update(voxel_points=vox_pts, voxel_coords=coords, voxel_features=feats, voxel_num_points=nums)
        return out


class HostStream:
    """The public host-buffer entry of the path: pinned HOST frames in, pinned HOST results out.

    The reference's driver (``sc_multiproc.py:60-94``) reads a frame from disk, completes its objects and
    writes a .pcd; here a batch of frames arrives in pinned memory and the completed clouds + voxel tensors
    leave in pinned memory.  Three CUDA streams: the H2D copy of batch i+1 and the D2H copy of batch i-1
    overlap the kernels of batch i (the copy engines are otherwise idle; every batch still pays its own
    copies inside the caller's timed region).

        hs = HostStream(pipe, frames, pts_per_frame, boxes_per_frame)
        for res in hs.run(batches):      # batches: iterable of (points_pinned (F,P,3), boxes_pinned (F,T,7))
            res["clustered"], res["voxel_coords"], ...   # pinned host views, valid until `depth` batches later
    """
    KEYS = ("clustered", "voxel_coords", "voxel_features", "voxel_num_points")

    def __init__(self, pipe, frames, pts_per_frame, boxes_per_frame, depth=2):
        self.pipe, self.depth = pipe, depth
        dev = pipe.device
        self.dev = dev
        F, P, T, S = frames, pts_per_frame, boxes_per_frame, pipe.resample_num
        max_obj, max_rows = F * T, F * P + F * T * S
        self.d_pts = [torch.empty((F, P, 3), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.d_boxes = [torch.empty((F, T, 7), dtype=torch.float32, device=dev) for _ in range(depth)]
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()   # noqa: E731
        self.h_out = [{"clustered": pin((max_obj, S, 3), torch.float32), "voxel_coords": pin((max_rows, 4), torch.int32),
                       "voxel_features": pin((max_rows, 3), torch.float32), "voxel_num_points": pin((max_rows,), torch.int32)}
                      for _ in range(depth)]
        self.s_h2d, self.s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # H2D of the slot landed
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]     # kernels finished reading the slot's inputs
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # D2H of the slot's results landed
        self.h2d_bytes = self.d2h_bytes = 0

    def _upload(self, slot, pts_pin, boxes_pin):
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(self.ev_free[slot])
            self.d_pts[slot].copy_(pts_pin, non_blocking=True)
            self.d_boxes[slot].copy_(boxes_pin, non_blocking=True)
            self.ev_in[slot].record(self.s_h2d)
        self.h2d_bytes += pts_pin.numel() * 4 + boxes_pin.numel() * 4

    def _download(self, slot, out):
        views = {}
        compute = torch.cuda.current_stream(self.dev)
        done = torch.cuda.Event(); done.record(compute)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(done)
            for k in self.KEYS:
                src = out.get(k)
                if src is None:
                    src = out["surface"] if k == "clustered" else None
                n = src.shape[0]
                dst = self.h_out[slot][k][:n]
                dst.copy_(src, non_blocking=True)
                views[k] = dst
                self.d2h_bytes += src.numel() * src.element_size()
            self.ev_out[slot].record(self.s_d2h)
        views["_keepalive"] = out      # device tensors stay referenced until the copy has landed
        return views

    def run(self, batches, seed=0):
        compute = torch.cuda.current_stream(self.dev)
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        for e in self.ev_free:
            e.record(compute)
        self._upload(0, *nxt)
        i, pending = 0, None
        while nxt is not None:
            slot = i % self.depth
            nxt = next(it, None)
            if nxt is not None:
                self._upload((i + 1) % self.depth, *nxt)         # overlaps the kernels below
            compute.wait_event(self.ev_in[slot])
            out = self.pipe.run(self.d_pts[slot], self.d_boxes[slot], seed=seed)
            self.ev_free[slot].record(compute)
            res = self._download(slot, out)                       # overlaps the next batch's kernels
            if pending is not None:
                self.ev_out[pending[0]].synchronize()
                pending[1].pop("_keepalive", None)
                yield pending[1]
            pending = (slot, res)
            i += 1
        self.ev_out[pending[0]].synchronize()
        pending[1].pop("_keepalive", None)
        yield pending[1]
