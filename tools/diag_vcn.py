import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from seevcn_b200 import synth, _abi
from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS
cuda = torch.device("cuda", 0)
def rel_chamfer(a, b):
    cd = oracle.chamfer_l2(a, b)
    scale = ((b - b.mean(axis=1, keepdims=True)) ** 2).sum(-1).mean(-1)
    return cd / scale
L = _abi.lib()
for seed in (3, 5, 7):
    for name, nobj, n in (("VCN_VC", 5, 1024), ("VCN_VC", 3, 1000), ("VCN_CN", 3, 1024)):
        part, _, boxes = synth.make_object_clouds(91, nobj, n, 0)
        sd = oracle.make_state_dict(name, seed=seed)
        gt = boxes[:, :7].astype(np.float32) if name == "VCN_CN" else None
        want = oracle.vcn_forward_ref(sd, part, gt, name)["coarse"].numpy()
        res = {}
        for prec, fused in (("bf16", 1), ("bf16", 0), ("fp32", 0)):
            model = MODELS.build({"NAME": name}, precision=prec); model.load_state_dict(sd); model.to(cuda).eval()
            d = {"input": torch.from_numpy(part).to(cuda)}
            if gt is not None: d["gt_boxes"] = torch.from_numpy(gt).to(cuda)
            model.precision = prec if (fused or prec == "fp32") else "bf16_layerwise"
            res[(prec, fused)] = model(d)["coarse"].cpu().numpy()
        print(seed, name, nobj, n, "fused", rel_chamfer(res[("bf16", 1)], want).round(6), "layer", rel_chamfer(res[("bf16", 0)], want).round(6),
              "fp32", rel_chamfer(res[("fp32", 0)], want).max(), "maxabs f-l", np.abs(res[("bf16", 1)] - res[("bf16", 0)]).max())
