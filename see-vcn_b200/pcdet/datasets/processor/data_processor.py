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


def mask_points_by_range_device(points, limit_range):
    """points (N, C) float32 CUDA -> the rows inside the x/y range, in order (one host sync for the count)."""
    points = points.contiguous()
    _abi.require_cuda(points)
    N, C = points.shape
    dev = points.device
    L = _abi.lib()
    out = torch.empty_like(points)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = _abi.workspace(dev, L.seevcn_mask_points_by_range_workspace_bytes(N), "rangemask")
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_mask_points_by_range(N, C, _abi.ptr(points), _abi.farray(limit_range), _abi.ptr(out), _abi.ptr(cnt),
                                                 _abi.ptr(ws), ws.numel(), _abi.stream()))
    return out[: int(cnt.item())]


def shuffle_points_device(points, seed):
    points = points.contiguous()
    _abi.require_cuda(points)
    out = torch.empty_like(points)
    with _abi.device_guard(points.device):
        _abi.check(_abi.lib().seevcn_shuffle_points(points.shape[0], points.shape[1], int(seed) & 0xffffffff, _abi.ptr(points),
                                                    _abi.ptr(out), _abi.stream()))
    return out


class DataProcessor(object):
    """The point-level part of pcdet's DataProcessor on the device.
    ref: detector3d/pcdet/datasets/processor/data_processor.py:63-143.  ``processor_configs``: list of dicts / objects with
    NAME in {mask_points_and_boxes_outside_range, shuffle_points, transform_points_to_voxels} and the reference's keys
    (SHUFFLE_ENABLED, VOXEL_SIZE, MAX_POINTS_PER_VOXEL, MAX_NUMBER_OF_VOXELS).  ``data_dict['points']`` is a CUDA tensor
    (N, C); the ground-truth box filter of the training mode (REMOVE_OUTSIDE_BOXES) is a training-loop concern and is
    left to the caller."""

    def __init__(self, processor_configs, point_cloud_range, training, num_point_features, seed=0):
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.training = training
        self.num_point_features = num_point_features
        self.mode = 'train' if training else 'test'
        self.grid_size = self.voxel_size = None
        self.voxel_generator = None
        self.seed = seed
        self._calls = 0
        self.data_processor_queue = []
        for cfg in processor_configs:
            name = cfg['NAME'] if isinstance(cfg, dict) else cfg.NAME
            self.data_processor_queue.append((getattr(self, name), cfg))
            if name == 'transform_points_to_voxels':
                vs = self._get(cfg, 'VOXEL_SIZE')
                grid = (np.array(self.point_cloud_range[3:6]) - np.array(self.point_cloud_range[0:3])) / np.array(vs)
                self.grid_size = np.round(grid).astype(np.int64)
                self.voxel_size = vs

    @staticmethod
    def _get(cfg, key, default=None):
        return cfg.get(key, default) if isinstance(cfg, dict) else getattr(cfg, key, default)

    def mask_points_and_boxes_outside_range(self, data_dict, config):
        if data_dict.get('points', None) is not None:
            data_dict['points'] = mask_points_by_range_device(data_dict['points'], self.point_cloud_range)
        return data_dict

    def shuffle_points(self, data_dict, config):
        if self._get(config, 'SHUFFLE_ENABLED')[self.mode]:
            data_dict['points'] = shuffle_points_device(data_dict['points'], self.seed + self._calls)
        return data_dict

    def transform_points_to_voxels(self, data_dict, config):
        if self.voxel_generator is None:
            self.voxel_generator = VoxelGeneratorWrapper(
                vsize_xyz=self._get(config, 'VOXEL_SIZE'), coors_range_xyz=self.point_cloud_range,
                num_point_features=self.num_point_features,
                max_num_points_per_voxel=self._get(config, 'MAX_POINTS_PER_VOXEL'),
                max_num_voxels=self._get(config, 'MAX_NUMBER_OF_VOXELS')[self.mode])
        voxels, coordinates, num_points = self.voxel_generator.generate_device(data_dict['points'])
        if not data_dict.get('use_lead_xyz', True):
            voxels = voxels[..., 3:]
        data_dict['voxels'] = voxels
        data_dict['voxel_coords'] = coordinates
        data_dict['voxel_num_points'] = num_points
        return data_dict

    def forward(self, data_dict):
        for fn, cfg in self.data_processor_queue:
            data_dict = fn(data_dict, cfg)
        self._calls += 1
        return data_dict
