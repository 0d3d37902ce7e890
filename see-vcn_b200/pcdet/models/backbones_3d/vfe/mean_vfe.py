"""MeanVFE on the B200 path.  ref: detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py:6-31"""
import torch

from ..... import _abi
from .vfe_template import VFETemplate, require_keys


class MeanVFE(VFETemplate):
    def __init__(self, model_cfg, num_point_features, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        """batch_dict['voxels'] (M, T, C), ['voxel_num_points'] (M,) -> ['voxel_features'] (M, C)
        = voxels.sum(1) / clamp_min(num_points, 1)"""
        require_keys(batch_dict, 'voxels', 'voxel_num_points')
        voxels = batch_dict['voxels'].contiguous()
        num = batch_dict['voxel_num_points'].contiguous()
        if num.dtype not in (torch.float32, torch.int32):      # int32 counts go to the kernel as they are
            num = num.to(torch.float32)
        _abi.require_cuda(voxels, num)
        assert voxels.dtype == torch.float32
        M, T, C = voxels.shape
        out = torch.empty((M, C), dtype=torch.float32, device=voxels.device)
        fn = _abi.lib().seevcn_mean_vfe if num.dtype == torch.float32 else _abi.lib().seevcn_mean_vfe_int
        with _abi.device_guard(voxels.device):
            _abi.check(fn(M, T, C, _abi.ptr(voxels), _abi.ptr(num), _abi.ptr(out), _abi.stream()))
        batch_dict['voxel_features'] = out
        return batch_dict
