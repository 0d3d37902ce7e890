#!/usr/bin/env python
"""How long the HOST needs to issue one batch (python wrappers + ctypes + launches), per stage of the pipeline.
If this approaches the GPU time of a batch the stream runs dry: measure before optimising kernels further."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cProfile
    import pstats
    import torch
    import bench
    from seevcn_b200.pipeline import CompletionPipeline
    import oracle
    dev = torch.device("cuda", 0)
    pipe = CompletionPipeline("VCN_VC", oracle.make_state_dict("VCN_VC", 0), dev, sel_k=bench.SEL_K, cluster_eps=bench.CLUSTER_EPS,
                              splice_thresh=bench.SPLICE_THRESH)
    pts, boxes = bench.make_inputs(8, 1000)
    pts_d, boxes_d = torch.from_numpy(pts).to(dev), torch.from_numpy(boxes).to(dev)
    for _ in range(3):
        pipe.run(pts_d, boxes_d)
    torch.cuda.synchronize()
    n = 20
    t_crop = t_b = t_fin = 0.0
    hs = [pipe.crop_async(pts_d, boxes_d) for _ in range(n)]
    torch.cuda.synchronize()           # counts are on the host: complete_from never blocks below
    outs = []
    for h in hs:
        torch.cuda.synchronize()       # empty launch queue: what is timed is the issue cost, never a full queue
        t0 = time.perf_counter()
        outs.append(pipe.run_from(h, 0, defer=True))
        t_b += time.perf_counter() - t0
    torch.cuda.synchronize()
    for o in outs:
        t0 = time.perf_counter(); pipe.finalize(o); t_fin += time.perf_counter() - t0
    for _ in range(n):
        t0 = time.perf_counter(); pipe.crop_async(pts_d, boxes_d); t_crop += time.perf_counter() - t0
    torch.cuda.synchronize()
    print("host issue cost per batch: crop %.3f ms, stage B %.3f ms, finalize %.3f ms" % (1e3 * t_crop / n, 1e3 * t_b / n, 1e3 * t_fin / n))
    pr = cProfile.Profile()
    hs = [pipe.crop_async(pts_d, boxes_d) for _ in range(n)]
    torch.cuda.synchronize()
    outs = []
    for h in hs:
        torch.cuda.synchronize()
        pr.enable()
        outs.append(pipe.run_from(h, 0, defer=True))
        pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
