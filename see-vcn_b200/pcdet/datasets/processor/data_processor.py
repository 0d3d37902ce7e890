"""VoxelGeneratorWrapper on the B200 path (hard voxelization, spconv semantics).

ref: detector3d/pcdet/datasets/processor/data_processor.py:15-60 (wrapper), :115-143 (caller).
PARITY UNPINNED: the arithmetic lives in third-party spconv (docker/Dockerfile:58, unpinned);
semantics restated from spconv v1's points_to_voxel loop: voxels in first-seen order, the first
``max_num_points_per_voxel`` points of each voxel in point order, at most ``max_num_voxels`` voxels
(points that would open a voxel past the cap are skipped).
"""
import numpy as np
import torch

from .... import _abi


def mask_points_by_range(points, limit_range):
    """x, y inclusive range mask.  ref: detector3d/pcdet/utils/common_utils.py:60-63"""
    return (points[:, 0] >= limit_range[0]) & (points[:, 0] <= limit_range[3]) \
        & (points[:, 1] >= limit_range[1]) & (points[:, 1] <= limit_range[4])


class VoxelGeneratorWrapper():
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels):
        self.vsize_xyz = [float(v) for v in vsize_xyz]
        self.coors_range_xyz = [float(v) for v in coors_range_xyz]
        self.num_point_features = int(num_point_features)
        self.max_num_points_per_voxel = int(max_num_points_per_voxel)
        self.max_num_voxels = int(max_num_voxels)
        grid = (np.array(self.coors_range_xyz[3:6]) - np.array(self.coors_range_xyz[0:3])) / np.array(self.vsize_xyz)
        self.grid_size = np.round(grid).astype(np.int64).tolist()   # data_processor.py:117-118
        self._ws = None

    def generate_device(self, points):
        """points (N, C) float32 CUDA -> voxels (M, T, C), coordinates (M, 3) int32 zyx, num_points (M,) int32 (CUDA)."""
        points = points.contiguous()
        _abi.require_cuda(points)
        N, C = points.shape
        dev = points.device
        L = _abi.lib()
        T, MV = self.max_num_points_per_voxel, self.max_num_voxels
        ws_bytes = L.seevcn_hard_voxelize_workspace_bytes(N, T, MV)
        if self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        voxels = torch.empty((MV, T, C), dtype=torch.float32, device=dev)
        coords = torch.empty((MV, 3), dtype=torch.int32, device=dev)
        num_pts = torch.empty((MV,), dtype=torch.int32, device=dev)
        num = torch.zeros((1,), dtype=torch.int32, device=dev)
        with _abi.device_guard(dev):
            _abi.check(L.seevcn_hard_voxelize(N, C, _abi.ptr(points), _abi.farray(self.coors_range_xyz),
                                              _abi.farray(self.vsize_xyz), _abi.iarray(self.grid_size), T, MV,
                                              _abi.ptr(voxels), _abi.ptr(coords), _abi.ptr(num_pts), _abi.ptr(num),
                                              _abi.ptr(self._ws), self._ws.numel(), _abi.stream()))
        m = min(int(num.item()), MV)
        return voxels[:m], coords[:m], num_pts[:m]

    def generate_frames_device(self, points, counts=None):
        """The batched form of the frame pipeline: points (F, stride, C) float32 CUDA, counts (F,) int32 CUDA or None
        (rows >= counts[f] of frame f are padding) -> voxels (F, MV, T, C), coordinates (F, MV, 4) int32 [frame, z, y, x]
        (what collate_batch builds, dataset.py:193-198), num_points (F, MV) int32, num_voxels (F,) int32 CUDA.  Slots
        >= num_voxels[f] are padding (num_points 0).  No host synchronisation."""
        points = points.contiguous()
        _abi.require_cuda(points, counts)
        F, S, C = points.shape
        dev = points.device
        L = _abi.lib()
        T, MV = self.max_num_points_per_voxel, self.max_num_voxels
        ws = _abi.workspace(dev, L.seevcn_hard_voxelize_frames_workspace_bytes(F, S, T, MV), "hardvox")
        voxels = torch.empty((F, MV, T, C), dtype=torch.float32, device=dev)
        coords = torch.empty((F, MV, 4), dtype=torch.int32, device=dev)
        num_pts = torch.empty((F, MV), dtype=torch.int32, device=dev)
        num = torch.empty((F,), dtype=torch.int32, device=dev)
        with _abi.device_guard(dev):
            _abi.check(L.seevcn_hard_voxelize_frames(F, S, C, _abi.ptr(points), _abi.ptr(counts), _abi.farray(self.coors_range_xyz),
                                                     _abi.farray(self.vsize_xyz), _abi.iarray(self.grid_size), T, MV,
                                                     _abi.ptr(voxels), _abi.ptr(coords), _abi.ptr(num_pts), _abi.ptr(num),
                                                     _abi.ptr(ws), ws.numel(), _abi.stream()))
        return voxels, coords, num_pts, num

    def generate(self, points):
        """numpy (N, C) in -> numpy (voxels, coordinates, num_points) out, like the reference wrapper."""
        if isinstance(points, np.ndarray):
            dev = torch.device("cuda", torch.cuda.current_device())
            d = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).pin_memory().to(dev, non_blocking=True)
            v, c, n = self.generate_device(d)
            return v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy()
        return self.generate_device(points)
