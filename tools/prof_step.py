#!/usr/bin/env python
"""Profiling driver: a few steps of the hot path (same workload as bench.py, no e2e / CPU legs) so
`ncu -k regex:... -s N -c M python tools/prof_step.py` stays short.  Not a benchmark: numbers printed
under a profiler are never reported."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    import torch
    import bench
    from seevcn_b200.pipeline import CompletionPipeline
    from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    ref_model = MODELS.build({"NAME": "VCN_VC"})
    for m in ref_model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    pipe = CompletionPipeline("VCN_VC", ref_model.state_dict(), dev, sel_k=bench.SEL_K, precision=args.precision,
                              cluster_eps=bench.CLUSTER_EPS, splice_thresh=bench.SPLICE_THRESH)
    pts, boxes = bench.make_inputs(args.frames, 1000)
    pts_d, boxes_d = torch.from_numpy(pts).to(dev), torch.from_numpy(boxes).to(dev)
    for _ in range(args.steps):
        out = pipe.run(pts_d, boxes_d, seed=0)
    torch.cuda.synchronize()
    print("objects", out["input"].shape[0], "voxels", out["voxel_coords"].shape[0])


if __name__ == "__main__":
    main()
