"""Host-side logic that needs no GPU: state-dict compatibility, BN folding, op argument checks."""
import numpy as np
import pytest
import torch

import oracle
from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS
from seevcn_b200.see.surface_completion.models.vcn.models._base import fold_conv_bn


@pytest.mark.parametrize("name", ["VCN_VC", "VCN_CN"])
def test_state_dict_keys_match_reference(name):
    model = MODELS.build({"NAME": name})
    sd = oracle.make_state_dict(name, seed=3)     # reference key names / shapes (checked by make_golden.py)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    ckpt = {"module." + k: v for k, v in sd.items()}   # see/surface_completion/models/VCN.py:36 strips 'module.'
    model.load_state_dict({k.replace("module.", ""): v for k, v in ckpt.items()})


def test_bn_folding_equals_conv_then_bn():
    torch.manual_seed(0)
    conv = torch.nn.Conv1d(7, 5, 1)
    bn = torch.nn.BatchNorm1d(5)
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
    bn.eval()
    x = torch.randn(3, 7, 11)
    w, b = fold_conv_bn(conv, bn)
    ref = bn(conv(x))
    got = torch.einsum("oc,bcn->bon", w, x) + b[None, :, None]
    np.testing.assert_allclose(got.detach().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)


def test_ops_refuse_cpu_tensors():
    from seevcn_b200.pcdet.ops.pointnet2.pointnet2_batch import pointnet2_utils
    with pytest.raises(RuntimeError, match="CUDA"):
        pointnet2_utils.furthest_point_sample(torch.zeros(1, 8, 3), 4)


def test_training_mode_is_rejected():
    model = MODELS.build({"NAME": "VCN_CN"})
    model.train()
    with pytest.raises(RuntimeError):
        model._pack()
