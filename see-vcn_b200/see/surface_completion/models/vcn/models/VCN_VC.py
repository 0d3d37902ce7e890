"""VCN_VC — viewer-centred completion network, inference forward on the B200 path.

ref: see/surface_completion/models/vcn/models/VCN_VC.py:110-213.  Module names/numbering match
the reference so ``load_state_dict(state_dict['base_model'])`` works unchanged
(see/surface_completion/models/VCN.py:34-39).  ``final_conv`` is kept as a loadable but unused
container: the reference constructs it (:133-141) and never calls it in forward.
"""
import torch

from ._base import VCNBase, Encoder, conv_stack, fold_conv_bn


class VCN_VC(VCNBase):
    viewer_centred = True

    def __init__(self, config=None, precision="bf16"):
        super().__init__(config, precision)
        self.sel_k = 30
        nc = self.number_coarse
        self.pose_encoder = conv_stack([("conv", 3, 64), ("leaky",), ("conv", 64, 128), ("leaky",),
                                        ("conv", 128, 1024), ("pool",)])
        self.pose_fc = conv_stack([("lin", 1024, 512), ("leaky",), ("lin", 512, 9)])
        self.encoder = Encoder([3, 128, 256, 512, 512, nc])
        self.shape_fc = conv_stack([("lin", 1024, 1024), ("relu",), ("lin", 1024, 1024), ("relu",),
                                    ("lin", 1024, 3 * nc)])
        self.final_conv = conv_stack([("conv", 1024 + 3 + 2, 512), ("bn", 512), ("relu",), ("conv", 512, 512),
                                      ("bn", 512), ("relu",), ("conv", 512, 3)])

    def _folded(self):
        e = self.encoder
        return {
            "pose_enc0": fold_conv_bn(self.pose_encoder[0]), "pose_enc2": fold_conv_bn(self.pose_encoder[2]),
            "pose_enc4": fold_conv_bn(self.pose_encoder[4]),
            "pose_fc0": fold_conv_bn(self.pose_fc[0]), "pose_fc2": fold_conv_bn(self.pose_fc[2]),
            "enc1_0": fold_conv_bn(e.mlp_conv1[0], e.mlp_conv1[1]), "enc1_3": fold_conv_bn(e.mlp_conv1[3]),
            "enc2_0": fold_conv_bn(e.mlp_conv2[0], e.mlp_conv2[1]), "enc2_3": fold_conv_bn(e.mlp_conv2[3]),
            "fc0": fold_conv_bn(self.shape_fc[0]), "fc2": fold_conv_bn(self.shape_fc[2]),
            "fc4": fold_conv_bn(self.shape_fc[4]),
        }

    @torch.no_grad()
    def forward(self, in_dict):
        """in_dict['input'] (B, N, 3) -> {'coarse' (B,1024,3), 'reg_rot' (B,3,3), 'reg_centre' (B,3)}
        ref: VCN_VC.py:178-213"""
        coarse, reg_rot, reg_centre = self._run(in_dict["input"], None)
        return {"coarse": coarse, "reg_rot": reg_rot, "reg_centre": reg_centre}
