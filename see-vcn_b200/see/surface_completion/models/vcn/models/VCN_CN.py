"""VCN_CN — completion network canonicalised with the ground-truth box.

ref: see/surface_completion/models/vcn/models/VCN_CN.py:111-157
"""
import torch

from ._base import VCNBase, Encoder, conv_stack, fold_conv_bn


class VCN_CN(VCNBase):
    viewer_centred = False

    def __init__(self, config=None, precision="bf16"):
        super().__init__(config, precision)
        self.sel_k = 30
        nc = self.number_coarse
        self.encoder = Encoder([3, 128, 256, 512, 512, nc])
        self.shape_fc = conv_stack([("lin", 1024, 1024), ("relu",), ("lin", 1024, 1024), ("relu",),
                                    ("lin", 1024, 3 * nc)])

    def _folded(self):
        e = self.encoder
        return {
            "enc1_0": fold_conv_bn(e.mlp_conv1[0], e.mlp_conv1[1]), "enc1_3": fold_conv_bn(e.mlp_conv1[3]),
            "enc2_0": fold_conv_bn(e.mlp_conv2[0], e.mlp_conv2[1]), "enc2_3": fold_conv_bn(e.mlp_conv2[3]),
            "fc0": fold_conv_bn(self.shape_fc[0]), "fc2": fold_conv_bn(self.shape_fc[2]),
            "fc4": fold_conv_bn(self.shape_fc[4]),
        }

    @torch.no_grad()
    def forward(self, in_dict):
        """in_dict['input'] (B, N, 3), in_dict['gt_boxes'] (B, 7) -> {'coarse' (B,1024,3)}
        ref: VCN_CN.py:142-157"""
        coarse, _, _ = self._run(in_dict["input"], in_dict["gt_boxes"])
        return {"coarse": coarse}
