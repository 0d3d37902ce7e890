#!/usr/bin/env python
"""cProfile of the e2e (HostStream) loop: time inside event synchronize = the host waiting for the GPU (good);
everything else = host work per step.  If the waits are ~0 the e2e leg is host-bound."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import oracle
    from seevcn_b200.pipeline import CompletionPipeline, HostStream
    dev = torch.device("cuda", 0)
    pipe = CompletionPipeline("VCN_VC", oracle.make_state_dict("VCN_VC", 0), dev, sel_k=bench.SEL_K, cluster_eps=bench.CLUSTER_EPS,
                              splice_thresh=bench.SPLICE_THRESH, streams=int(os.environ.get("STREAMS", "4")))
    pts, boxes = bench.make_inputs(8, 1000)
    pp, bp = torch.from_numpy(pts).pin_memory(), torch.from_numpy(boxes).pin_memory()
    hs = HostStream(pipe, 8, pts.shape[1], boxes.shape[1])
    for _ in hs.run((pp, bp) for _ in range(5)):
        pass
    torch.cuda.synchronize()
    n = 40
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in hs.run((pp, bp) for _ in range(n)):
        pass
    pr.disable()
    torch.cuda.synchronize()
    print("wall per step %.3f ms" % (1e3 * (time.perf_counter() - t0) / n))
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(30)
    st.print_callers("torch.empty")
    print("allocator:", {k: v for k, v in torch.cuda.memory_stats(dev).items() if k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "allocation.all.allocated", "segment.all.allocated")})
    # per-stage GPU time inside the e2e loop (library event scopes) next to the resident loop: which stage slows down?
    from seevcn_b200 import _abi
    pts_d, boxes_d = pp.to(dev), bp.to(dev)
    for mode in ("e2e", "resident"):
        _abi.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "e2e":
            for _ in hs.run((pp, bp) for _ in range(n)):
                pass
        else:
            for _ in pipe.run_stream((pts_d, boxes_d) for _ in range(n)):
                pass
        e1.record(); e1.synchronize()
        _abi.prof_enable(False)
        rep = _abi.prof_report()
        print(mode, "step %.3f ms:" % (e0.elapsed_time(e1) / n), "  ".join("%s %.3f" % (k, v[1] / n) for k, v in rep.items()
                                                                         if k in ("crop", "vcn_forward", "knn_surface_select", "largest_cluster", "splice", "dynamic_voxelize")))


if __name__ == "__main__":
    main()
