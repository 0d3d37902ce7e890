"""VCN inference wrapper on the B200 path.

ref: see/surface_completion/models/VCN.py:14-104.  Same constructor knobs (cfg.MODEL, NORM_WITH_GT,
SEL_K_NEAREST, CLUSTER_EPS, BATCH_SIZE_LIMIT, CKPT_PATH) and the same ``inference`` contract: host numpy
clouds in, dict of numpy arrays out.  Differences, all on the inside:
  * every object of the call goes through ONE forward (the reference pads to a multiple of
    BATCH_SIZE_LIMIT with all-zero clouds and loops over chunks, :55-83; results are identical because
    objects are independent — ``batch_size_limit`` is accepted and ignored);
  * kNN surface selection and the largest-cluster filter run on the device (the reference copies each
    object to the host for cKDTree / open3d, :94-98);
  * one H2D copy of the resampled clouds, one D2H copy per output.
"""
from pathlib import Path

import numpy as np
import torch

from .vcn.models.build import MODELS
from .vcn.utils.sampling import get_partial_mesh_batch, get_largest_cluster_batch


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class ResamplePoints(object):
    """Drop or duplicate points so that a cloud has exactly n points.
    ref: see/surface_completion/models/vcn/datasets/data_transforms.py:247-262.  ``rng`` (optional,
    numpy Generator) makes the draw reproducible; the reference uses the global np.random state."""

    def __init__(self, parameters, rng=None):
        self.n_points = parameters['n_points']
        self.rng = rng

    def __call__(self, pts):
        reps = int(np.ceil(self.n_points / len(pts)))
        tiled = np.tile(pts, (reps, 1))
        perm = self.rng.permutation(tiled.shape[0]) if self.rng is not None else np.random.permutation(tiled.shape[0])
        return tiled[perm[:self.n_points]]


class VCN:
    def __init__(self, cfg, gpu_id=0, state_dict=None, precision="bf16"):
        self.cfg = cfg
        self.device = torch.device(f'cuda:{gpu_id}')
        torch.cuda.set_device(gpu_id)
        self.precision = precision
        self.model_init(state_dict)

    def model_init(self, state_dict=None):
        self.norm_with_gt = _get(self.cfg, 'NORM_WITH_GT')
        self.surface_sel_k = _get(self.cfg, 'SEL_K_NEAREST')
        self.cluster_eps = _get(self.cfg, 'CLUSTER_EPS')
        self.batch_size_limit = _get(self.cfg, 'BATCH_SIZE_LIMIT', None)
        self.model = MODELS.build({'NAME': _get(self.cfg, 'MODEL')}, precision=self.precision)
        if state_dict is None:
            ckpt = _get(self.cfg, 'CKPT_PATH')
            assert ckpt is not None and Path(ckpt).exists(), f"No ckpt found at {ckpt}"
            state_dict = torch.load(ckpt, map_location='cpu')['base_model']
        self.model.load_state_dict({k.replace("module.", ""): v for k, v in state_dict.items()})
        self.model.to(self.device).eval()

    @torch.no_grad()
    def inference(self, pts, gtboxes=None, batch_size_limit=None, resample_num=1024, k=30, eps=0.4, rng=None):
        """
        pts: np.array (N, 3) or list(np.array) of shape (N,3)
        gtboxes: list(np.array) each of shape (7)
        returns dict of numpy arrays:
            input (B, resample_num, 3), coarse (B, 1024, 3) whole completed surface,
            surface (B, resample_num, 3) k nearest completed points of every input point,
            clustered (B, 1024, 3) largest cluster of the surface points
        """
        resample = ResamplePoints({'n_points': resample_num}, rng)
        clouds = pts if isinstance(pts, list) else [pts]
        resampled = np.stack([resample(pc) for pc in clouds]).astype(np.float32)
        in_pc = torch.from_numpy(resampled).pin_memory().to(self.device, non_blocking=True)
        in_dict = {'input': in_pc}
        if self.norm_with_gt:
            gt = np.vstack(gtboxes)[:, :7].astype(np.float32)
            in_dict['gt_boxes'] = torch.from_numpy(gt).pin_memory().to(self.device, non_blocking=True)
        output = self.model(in_dict)['coarse']
        pred_surface = get_partial_mesh_batch(in_pc, output, k=k, surface_pts=resample_num)
        pred_cluster = get_largest_cluster_batch(pred_surface, eps=eps, min_points=2, total_pts=output.shape[1])
        return {'input': resampled, 'surface': pred_surface.cpu().numpy(), 'clustered': pred_cluster.cpu().numpy(),
                'coarse': output.cpu().numpy()}
