#!/usr/bin/env python
"""Box health check for the pipeline's host round trips: tiny D2H latency (copy engine vs copy kernel), launch latency,
and the step time of the resident / e2e loops — some boxes of the pool show multi-millisecond stalls in one of them."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import oracle
    from seevcn_b200 import _abi
    from seevcn_b200.pipeline import CompletionPipeline, HostStream
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    print(torch.cuda.get_device_name(0), "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    try:
        print(open("/proc/loadavg").read().strip())
    except OSError:
        pass
    d = torch.zeros(400, dtype=torch.int32, device=dev)
    h = torch.empty(400, dtype=torch.int32).pin_memory()
    L = _abi.lib()
    big_d = torch.empty(18 << 20, dtype=torch.uint8, device=dev); big_h = torch.empty(18 << 20, dtype=torch.uint8).pin_memory()
    side, bulk = torch.cuda.Stream(), torch.cuda.Stream()

    def lat(fn, n=200, with_bulk=False):
        ts = []
        for _ in range(n):
            if with_bulk:
                with torch.cuda.stream(bulk):
                    big_h.copy_(big_d, non_blocking=True)
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
            if with_bulk:
                bulk.synchronize()
        ts.sort()
        return "median %.1f us  p90 %.1f us  max %.1f us" % (1e6 * ts[n // 2], 1e6 * ts[int(n * 0.9)], 1e6 * ts[-1])

    def dma():
        ev = torch.cuda.Event()
        with torch.cuda.stream(side):
            h.copy_(d, non_blocking=True); ev.record(side)
        ev.synchronize()

    def smk():
        _abi.check(L.seevcn_copy_to_pinned(_abi.ptr(d), _abi.c_void_p(h.data_ptr()), 1600, _abi.stream()))
        ev = torch.cuda.Event(); ev.record(); ev.synchronize()

    def launch():
        d.add_(1)
    for _ in range(20):
        dma(); smk(); launch()
    torch.cuda.synchronize()
    print("tiny D2H via copy engine :", lat(dma), "| with bulk D2H in flight:", lat(dma, 60, True))
    print("tiny D2H via copy kernel :", lat(smk), "| with bulk D2H in flight:", lat(smk, 60, True))
    t0 = time.perf_counter()
    for _ in range(2000):
        launch()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("launch issue cost: %.2f us per kernel" % (1e6 * (t1 - t0) / 2000))

    pipe = CompletionPipeline("VCN_VC", oracle.make_state_dict("VCN_VC", 0), dev, sel_k=bench.SEL_K, cluster_eps=bench.CLUSTER_EPS,
                              splice_thresh=bench.SPLICE_THRESH)
    pts, boxes = bench.make_inputs(8, 1000)
    pp, bp = torch.from_numpy(pts).pin_memory(), torch.from_numpy(boxes).pin_memory()
    pts_d, boxes_d = pp.to(dev), bp.to(dev)
    hs = HostStream(pipe, 8, pts.shape[1], boxes.shape[1])
    for rep in range(2):
        for mode in ("resident", "e2e"):
            n = 30
            src = ((pts_d, boxes_d) for _ in range(n + 3)) if mode == "resident" else ((pp, bp) for _ in range(n + 3))
            it = pipe.run_stream(src) if mode == "resident" else hs.run(src)
            for _ in range(3):
                next(it)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gaps = []
            tl = t0
            for _ in it:
                t = time.perf_counter(); gaps.append(t - tl); tl = t
            torch.cuda.synchronize()
            gaps.sort()
            print("%s: %.3f ms/step (per-yield gap median %.3f, max %.3f ms)" % (mode, 1e3 * (time.perf_counter() - t0) / n,
                                                                                 1e3 * gaps[len(gaps) // 2], 1e3 * gaps[-1]))


if __name__ == "__main__":
    main()
