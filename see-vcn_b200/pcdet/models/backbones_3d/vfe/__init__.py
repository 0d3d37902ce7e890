"""ref: detector3d/pcdet/models/backbones_3d/vfe/__init__.py:8-15 (registry subset on the SEE-VCN path)"""
from .mean_vfe import MeanVFE
from .dynamic_mean_vfe import DynamicMeanVFE
from .vfe_template import VFETemplate

__all__ = {
    'VFETemplate': VFETemplate,
    'MeanVFE': MeanVFE,
    'DynMeanVFE': DynamicMeanVFE,
}
