"""DynamicMeanVFE on the B200 path: sort-free bucketed scatter-mean (rows in torch.unique order, deterministic means).

ref: detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:14-76
"""
import torch

from ..... import _abi
from .vfe_template import VFETemplate, require_keys


def dynamic_voxelize(points, point_cloud_range, voxel_size, grid_size, max_voxels=None, sort=True, workspace=None,
                     batch_size=0):
    """points (N, 1+C) float32 CUDA rows [batch_idx, x, y, z, ...] ->
    voxel_coords (M, 4) int32 [b, z, y, x], voxel_features (M, C) float32, voxel_counts (M,) int32.
    Rows always come out ordered by the reference's merge key (torch.unique order); ``sort`` is ignored.
    ``batch_size`` = batch_dict['batch_size']; 0 derives it from the batch column (one extra host sync)."""
    points = points.contiguous()
    _abi.require_cuda(points)
    assert points.dtype == torch.float32 and points.dim() == 2
    N, C1 = points.shape
    C = C1 - 1
    dev = points.device
    if max_voxels is None:
        max_voxels = max(N, 1)
    L = _abi.lib()
    if int(batch_size) <= 0:
        batch_size = int(points[:, 0].max().item()) + 1 if N > 0 else 1
    ws_bytes = L.seevcn_dynamic_voxelize_workspace_bytes(N, C, int(batch_size), _abi.iarray(grid_size))
    if ws_bytes == 0:
        raise RuntimeError("dynamic_voxelize: batch_size x grid too large")
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    coords = torch.empty((max_voxels, 4), dtype=torch.int32, device=dev)
    feats = torch.empty((max_voxels, C), dtype=torch.float32, device=dev)
    counts = torch.empty((max_voxels,), dtype=torch.int32, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_dynamic_voxelize(N, C, _abi.ptr(points), _abi.farray(point_cloud_range),
                                             _abi.farray(voxel_size), _abi.iarray(grid_size), max_voxels,
                                             1 if sort else 0, int(batch_size), _abi.ptr(coords), _abi.ptr(feats), _abi.ptr(counts),
                                             _abi.ptr(num), _abi.ptr(workspace), workspace.numel(), _abi.stream()))
    m = min(int(num.item()), max_voxels)   # the one host sync: M is data dependent, as in the reference
    return coords[:m], feats[:m], counts[:m]


def dynamic_voxelize_frames(frame_pts, obj_pts, obj_frame, point_cloud_range, voxel_size, grid_size, sort=True,
                            frame_keep=None, obj_count=None, max_voxels=None):
    """The frame pipeline's form: frame_pts (F,P,3) rows carry batch index = frame, obj_pts (O,S,3) rows carry
    obj_frame[o]; same result as ``dynamic_voxelize`` on the concatenated [batch_idx,x,y,z] matrix, which is never
    built.  After the splice step (SEE_VCN.py:247-265): ``frame_keep`` (F,P) uint8 drops the replaced frame points and
    ``obj_count`` (O,) int32 limits every object to its distinct rows.  No host sync: returns FULL-capacity tensors
    plus the device scalar M — (coords (N,4), feats (N,3), counts (N,), num_voxels (1,) int32 CUDA); rows >= M are
    undefined.  ``max_voxels``: row capacity of the outputs (default: one row per point, which can never overflow); rows
    beyond a smaller capacity are dropped and ``num_voxels`` still reports the true M."""
    frame_pts = frame_pts.contiguous()
    _abi.require_cuda(frame_pts)
    F, P, _ = frame_pts.shape
    dev = frame_pts.device
    if obj_pts is not None and obj_pts.shape[0] > 0:
        obj_pts = obj_pts.contiguous(); obj_frame = obj_frame.contiguous()
        _abi.require_cuda(obj_pts, obj_frame)
        assert obj_frame.dtype == torch.int32 and obj_frame.shape[0] == obj_pts.shape[0]
        O, S, _ = obj_pts.shape
        if obj_count is not None:
            obj_count = obj_count.contiguous()
            _abi.require_cuda(obj_count)
            assert obj_count.dtype == torch.int32 and obj_count.shape[0] == O
    else:
        obj_pts = obj_frame = obj_count = None
        O = S = 0
    if frame_keep is not None:
        frame_keep = frame_keep.contiguous()
        _abi.require_cuda(frame_keep)
        assert frame_keep.dtype == torch.uint8 and frame_keep.numel() == F * P
    N = F * P + O * S
    L = _abi.lib()
    cap = max(N, 1) if max_voxels is None else max(min(int(max_voxels), N), 1)
    ws = _abi.workspace(dev, L.seevcn_dynamic_voxelize_workspace_bytes(N, 3, max(F, 1), _abi.iarray(grid_size)), "dynvox")
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    feats = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    counts = torch.empty((cap,), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_dynamic_voxelize_spliced(F, P, _abi.ptr(frame_pts), _abi.ptr(frame_keep), O, S, _abi.ptr(obj_pts),
                                                     _abi.ptr(obj_frame), _abi.ptr(obj_count),
                                                     _abi.farray(point_cloud_range), _abi.farray(voxel_size),
                                                     _abi.iarray(grid_size), cap, 1 if sort else 0, _abi.ptr(coords),
                                                     _abi.ptr(feats), _abi.ptr(counts), _abi.ptr(num), _abi.ptr(ws),
                                                     ws.numel(), _abi.stream()))
    return coords, feats, counts, num


class DynamicMeanVFE(VFETemplate):
    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.num_point_features = num_point_features
        self.grid_size = [int(g) for g in grid_size]
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.sort = kwargs.get('sort', True)

    def get_output_feature_dim(self):
        return self.num_point_features

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        """batch_dict['points'] (sum N, 1+C) [batch_idx, x, y, z, ...] ->
        ['voxel_features'] (M, C) mean of all in-voxel points, ['voxel_coords'] (M, 4) int32 [b,z,y,x],
        plus ['voxel_num_points'] (M,) int32 (the reference computes and drops unq_cnt, :63)."""
        require_keys(batch_dict, 'points')
        coords, feats, counts = dynamic_voxelize(batch_dict['points'], self.point_cloud_range, self.voxel_size,
                                                 self.grid_size, sort=self.sort, batch_size=int(batch_dict.get('batch_size', 0)))
        batch_dict['voxel_features'] = feats.contiguous()
        batch_dict['voxel_coords'] = coords.contiguous()
        batch_dict['voxel_num_points'] = counts.contiguous()
        return batch_dict
