"""MODELS.build({'NAME': ...}) — ref: see/surface_completion/models/vcn/models/build.py:4-15"""
from .VCN_VC import VCN_VC
from .VCN_CN import VCN_CN

_REGISTRY = {"VCN_VC": VCN_VC, "VCN_CN": VCN_CN}


class _Models:
    def build(self, cfg, **kwargs):
        name = cfg["NAME"] if isinstance(cfg, dict) else cfg.NAME
        if name not in _REGISTRY:
            raise KeyError(f"{name} is not a registered VCN model ({sorted(_REGISTRY)})")
        return _REGISTRY[name](cfg, **kwargs)


MODELS = _Models()


def build_model_from_cfg(cfg, **kwargs):
    return MODELS.build(cfg, **kwargs)
