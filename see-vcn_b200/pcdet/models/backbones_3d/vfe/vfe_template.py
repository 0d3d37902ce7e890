"""Common base of the voxel feature encoders on the B200 path.

Interface contract taken from detector3d/pcdet/models/backbones_3d/vfe/vfe_template.py:4-22 and from how
Detector3DTemplate.build_vfe drives it (detectors/detector3d_template.py:52-66): constructed with ``model_cfg`` plus
keyword geometry, asked for ``get_output_feature_dim()``, called with the batch dict.
"""
from torch import nn


def require_keys(batch_dict, *keys):
    """The encoders read fixed keys of the detector's batch dict; name the missing one instead of a bare KeyError."""
    absent = [k for k in keys if k not in batch_dict]
    if absent:
        raise KeyError(f"batch_dict lacks {absent}; present: {sorted(batch_dict)}")


class VFETemplate(nn.Module):
    """Subclasses provide ``get_output_feature_dim`` (width of ``voxel_features``) and ``forward(batch_dict)``."""

    def __init__(self, model_cfg, **geometry):
        nn.Module.__init__(self)
        self.model_cfg = model_cfg

    def get_output_feature_dim(self):
        raise NotImplementedError(f"{type(self).__name__} does not report its output feature width")

    def forward(self, batch_dict=None, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} does not implement forward(batch_dict)")
