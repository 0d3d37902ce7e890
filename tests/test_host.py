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


def test_pcd_wire_format_roundtrip_and_reference_header(tmp_path):
    """Binary PCD as open3d writes it for SEE_VCN.save_pcd (SEE_VCN.py:267-280): our header for N points equals the
    header of the reference's demo frame byte for byte (tests/golden/pcd_header.txt, N = 26715), the payload is
    N x 3 float32, and the reader returns what was written; extra fields and the ascii flavour are read too."""
    import os
    import numpy as np
    from seevcn_b200.see.surface_completion.pcd_io import pcd_header, read_pcd, write_pcd
    here = os.path.dirname(os.path.abspath(__file__))
    ref_header = open(os.path.join(here, "golden", "pcd_header.txt"), "rb").read()
    assert pcd_header(26715).encode("ascii") == ref_header
    first = np.load(os.path.join(here, "golden", "pcd_demo_first64.npy"))
    path = str(tmp_path / "frame.pcd")
    write_pcd(path, np.concatenate([first, np.ones((64, 1), np.float32)], axis=1))   # a 4th column is dropped, as in save_pcd
    raw = open(path, "rb").read()
    assert raw.startswith(pcd_header(64).encode("ascii")) and len(raw) == len(pcd_header(64)) + 64 * 12
    np.testing.assert_array_equal(read_pcd(path), first)
    # x y z intensity, binary: the reader picks the requested columns
    hdr = ("VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 64\nHEIGHT 1\n"
           "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 64\nDATA binary\n")
    four = np.concatenate([first, np.arange(64, dtype=np.float32)[:, None]], axis=1)
    open(path, "wb").write(hdr.encode() + four.astype("<f4").tobytes())
    np.testing.assert_array_equal(read_pcd(path), first)
    np.testing.assert_array_equal(read_pcd(path, ("intensity",))[:, 0], np.arange(64, dtype=np.float32))
    open(path, "w").write(hdr.replace("binary", "ascii") + "\n".join(" ".join(repr(float(v)) for v in row) for row in four) + "\n")
    np.testing.assert_array_equal(read_pcd(path), first)
