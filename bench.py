#!/usr/bin/env python
"""bench.py — SEE-VCN object-completion + voxelization hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]

A step = one pass of the hot path (crop -> resample -> VCN forward -> kNN surface select ->
largest-cluster filter -> splice (raw points replaced by the completed ones) -> dynamic voxelization) over a batch of F synthetic Waymo-like frames (BASELINE.json configs[1]:
64 beams x 2812 azimuth steps = 180k pts, 50 car boxes, 1024 pts/object) per GPU.  Frames shard
across ranks with no collective on the data path; one all-gather-v of the completed clouds per
step (static capacity, asynchronous, no host sync) stands for "collect for the detector" when N > 1 (weak scaling: F frames per GPU).

Prints ONE JSON line (rank 0).  `value` = completed objects/s with inputs resident in HBM;
`e2e` = same through the public API from pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` = the dominant kernel (vcn_chain_kernel<2>, the enc2 tcgen05 chain) timed by the library's
CUDA-event scopes inside the timed region, against the measured bf16 peak; `stages` = every launch
group the same way; `cpu_baseline` = the oracle port of the same path timed on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "completed objects/sec"
SEL_K = 20               # SURFACE_COMPLETION.VCN.SEL_K_NEAREST (see/surface_completion/cfgs/WAY-GT_VCN-VC.yaml:13)
CLUSTER_EPS = 0.3        # SURFACE_COMPLETION.VCN.CLUSTER_EPS   (WAY-GT_VCN-VC.yaml:14)
SPLICE_THRESH = 0.1      # replace_with_completed_pts(point_dist_thresh=0.1)          (see/surface_completion/SEE_VCN.py:247)
RESAMPLE = 1024
FLOP_PER_OBJ = 2.0 * (959040 * 1024 + 5771776)   # VCN_VC, N = 1024 -> 1.976 GFLOP
FLOP_ENC2_REF = 2.0 * (512 * 512 + 512 * 1024) * 1024    # enc2 as the reference graph states it
FLOP_ENC2_EXEC = 2.0 * (256 * 512 + 512 * 1024) * 1024   # enc2 as executed: the global half folded into a per-object bias


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Sampled through NVML
    in-process (a polling nvidia-smi subprocess takes ~1 s per call on these hosts and stalls launches);
    falls back to nvidia-smi when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag = index, period, False
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:   # noqa: BLE001
            self.nv, self.source, self.period = None, "nvidia-smi", 0.5

    def sample(self):
        if self.nv is not None:
            nv = self.nv
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(get(self.h))
            self.reasons |= {n for n, b in self.BITS if mask & b}
            return
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) >= 6:
            self.sm.append(float(r[0])); self.mx.append(float(r[1]))
            self.reasons |= {n for (n, _), v in zip(self.BITS, (r[2], r[3], r[4], r[5])) if v.lower().startswith("active")}

    def run(self):
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:   # noqa: BLE001
                pass
            time.sleep(self.period)

    def summary(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def make_inputs(frames, seed0):
    from seevcn_b200 import synth
    return synth.make_stream(frames, first_seed=seed0)


# ---------------------------------------------------------------------------- CPU path --
def cpu_path_once(pts, boxes, sd, threads):
    """The oracle port of the whole step on host cores -> (#objects, seconds, #voxelised points)."""
    import torch
    import oracle
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    idx = oracle.points_in_boxes_gpu(pts, boxes)
    rng = np.random.default_rng(0)
    clouds, fid = [], []
    for f in range(pts.shape[0]):
        cnt = np.bincount(idx[f][idx[f] >= 0], minlength=boxes.shape[1])
        for k in np.nonzero(cnt >= 30)[0]:
            clouds.append(oracle.resample_points(pts[f][idx[f] == k], RESAMPLE, rng)[0]); fid.append(f)
    n_obj = len(clouds)
    frame_sc = [None] * pts.shape[0]
    if n_obj:
        inp = np.stack(clouds).astype(np.float32)
        with torch.no_grad():
            coarse = oracle.vcn_forward_ref(sd, inp, None, "VCN_VC")["coarse"].numpy()
        surf, _ = oracle.get_partial_mesh_batch(inp, coarse, k=SEL_K)
        surf, cnt = oracle.get_largest_cluster_batch(surf, eps=CLUSTER_EPS, min_points=2, total_pts=RESAMPLE)
        fid = np.asarray(fid)
        for f in range(pts.shape[0]):   # SEE_VCN.py:244: all_instances = np.unique(np.vstack(clustered), axis=0)
            sel = np.nonzero(fid == f)[0]
            if len(sel):
                frame_sc[f] = oracle.all_instances(surf[sel], cnt[sel])
    rows = []
    for f in range(pts.shape[0]):       # SEE_VCN.py:247-265 splice, then the [batch_idx, x, y, z] rows of dataset.py:187-192
        merged, _ = oracle.replace_with_completed_pts(pts[f], frame_sc[f] if frame_sc[f] is not None and len(frame_sc[f]) else None,
                                                      SPLICE_THRESH)
        rows.append(np.concatenate([np.full((len(merged), 1), f, np.float32), merged], axis=1))
    vox = np.concatenate(rows).astype(np.float32)
    from seevcn_b200.pipeline import WAYMO_VOXEL_CFG
    oracle.dynamic_voxelize(vox, *WAYMO_VOXEL_CFG)
    return n_obj, time.perf_counter() - t0, len(vox)


def cpu_baseline(sample_frames, sd, threads, budget_s=12.0):
    """Bounded sample: passes of the whole path over `sample_frames` synthetic C2 frames until ~budget_s of CPU work."""
    pts, boxes = make_inputs(sample_frames, 5000)
    cpu_path_once(pts[:1], boxes[:1], sd, threads)   # warm-up (thread pools, page-in)
    n_obj = n_vox = passes = 0
    sec = 0.0
    while sec < budget_s:
        n, s, v = cpu_path_once(pts, boxes, sd, threads)
        n_obj += n; sec += s; n_vox += v; passes += 1
    return {"value": n_obj / sec, "unit": "objects/s", "cores": threads, "kind": "port",
            "sample": f"{passes} passes x {sample_frames} frames x 180k pts, {n_obj} objects in {sec:.1f} s "
                      "(oracle/: C + torch fp32, OpenMP/intra-op threads)",
            "voxelized_mpts_per_s": n_vox / sec / 1e6}


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (oracle port; the reference's python
    cannot travel to the GPU box) on all host threads; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = os.cpu_count() or 1
    sd = oracle.make_state_dict("VCN_VC", seed=0)
    sample_frames = 1
    pts, boxes = make_inputs(sample_frames, 1000)
    for _ in range(max(args.warmup, 1)):
        cpu_path_once(pts, boxes, sd, threads)
    tot_obj, tot_sec, tot_vox = 0, 0.0, 0
    for _ in range(args.steps):
        n, s, v = cpu_path_once(pts, boxes, sd, threads)
        tot_obj += n; tot_sec += s; tot_vox += v
    val = tot_obj / tot_sec
    sample = f"{sample_frames} frame x 180k pts per step ({tot_obj // max(args.steps, 1)} objects), oracle port, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "objects/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: synthetic Waymo-like 64-beam frame (180k pts, 50 car boxes, 1024 pts/object)",
                   "frames_per_step": sample_frames, "sel_k": SEL_K, "cluster_eps": CLUSTER_EPS, "splice_thresh": SPLICE_THRESH},
        "voxelized_mpts_per_sec": tot_vox / tot_sec / 1e6,
        "cpu_baseline": {"value": val, "unit": "objects/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------- GPU path --
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seevcn_b200 import _abi
    from seevcn_b200.pipeline import CompletionPipeline
    from seevcn_b200 import dist as sdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weights: seeded random init of the VCN_VC architecture (no checkpoints offline).  Generated without
    # the oracle package: same generator stream as oracle.make_state_dict, restated in the product pipeline.
    from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS
    torch.manual_seed(0)
    ref_model = MODELS.build({"NAME": "VCN_VC"})
    for m in ref_model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    sd = ref_model.state_dict()
    pipe = CompletionPipeline("VCN_VC", sd, dev, sel_k=SEL_K, precision=args.precision, cluster_eps=CLUSTER_EPS,
                              splice_thresh=SPLICE_THRESH)

    F = args.frames
    pts_h, boxes_h = make_inputs(F, 1000 + rank * F)         # rank r owns frames [r*F, (r+1)*F)
    pts_pin = torch.from_numpy(pts_h).pin_memory()
    boxes_pin = torch.from_numpy(boxes_h).pin_memory()
    pts_d, boxes_d = pts_pin.to(dev), boxes_pin.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from seevcn_b200.pipeline import HostStream
    hs = HostStream(pipe, F, pts_h.shape[1], boxes_h.shape[1])

    def resident_batches(n):
        for _ in range(n):
            flush.fill_(1)                                   # L2 flush before every batch (on the compute stream, timed)
            yield pts_d, boxes_d

    def e2e_batches(n):
        for _ in range(n):
            flush.fill_(1)
            yield pts_pin, boxes_pin

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(ms, count):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        c = torch.tensor([float(count)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)         # max over ranks
            dist.all_reduce(c)
        return t.item(), c.item()

    def timed_resident(steps, warmup):
        """K batches streamed through pipe.run_stream with the inputs resident in HBM; the timed region holds the
        K L2 flushes too.  The library's event scopes (seevcn_prof_*) time each launch group on the launching stream."""
        for out in pipe.run_stream(resident_batches(warmup)):
            pass
        sync_all()
        l0 = _abi.lib().seevcn_launch_count()
        _abi.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_obj = 0
        last = None
        e0.record()
        gathered = None
        for out in pipe.run_stream(resident_batches(steps)):
            if world > 1:   # "collect for the detector": static-capacity all-gather, no host sync, overlaps the next batch
                if gathered is not None:
                    gathered.wait()
                gathered = sdist.all_gather_padded(out.get("clustered", out["surface"]), F * boxes_h.shape[1], async_op=True)
            n_obj += out["input"].shape[0]
            last = out
        if gathered is not None:
            gathered.wait()
        e1.record()
        e1.synchronize()
        # rows actually voxelized (spliced-out points, cyclic repeats and out-of-range points are not): the voxel counts
        # sum to it; every step runs the same frames
        n_pts = int(last["voxel_num_points"].sum().item()) * steps
        last["num_voxel_points"] = n_pts // max(steps, 1)
        _abi.prof_enable(False)
        prof = _abi.prof_report()
        launches = _abi.lib().seevcn_launch_count() - l0
        sync_all()
        ms, objs = reduce_max_sum(e0.elapsed_time(e1), n_obj)
        _, pts = reduce_max_sum(0.0, n_pts)
        return ms, objs, pts, last, launches, prof

    def timed_e2e(steps, warmup):
        """Public host-buffer API: pinned host frames in, pinned host results out, every batch's H2D and D2H
        inside the timed region (copies of neighbouring batches overlap the kernels, see HostStream)."""
        for _ in hs.run(e2e_batches(warmup)):
            pass
        sync_all()
        hs.h2d_bytes = hs.d2h_bytes = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_obj = 0
        e0.record()
        for res in hs.run(e2e_batches(steps)):
            n_obj += res["clustered"].shape[0]
        e1.record()                                          # after the last D2H has landed on the host
        e1.synchronize()
        sync_all()
        ms, objs = reduce_max_sum(e0.elapsed_time(e1), n_obj)
        return ms, objs, hs.h2d_bytes // steps, hs.d2h_bytes // steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, n_obj_res, n_vox_pts, out, launches, prof = timed_resident(args.steps, args.warmup)
    n_voxels = out["voxel_coords"].shape[0]
    obj_rank0 = out["input"].shape[0]
    ms_e2e, n_obj_e2e, h2d, d2h = timed_e2e(args.steps, max(args.warmup, 3))
    clocks = sampler.summary() if rank == 0 else None

    if rank == 0:
        pk = peaks()
        steps = args.steps
        value = n_obj_res / (ms_res / 1e3)
        e2e = n_obj_e2e / (ms_e2e / 1e3)
        step_ms = ms_res / steps
        # per launch group, from the library's CUDA events inside the timed region (rank 0)
        P_pts = pts_h.shape[1]
        alg = {   # algorithmic work per STEP on this rank (SURVEY.md §8d per-unit figures x units), and the bound
            "points_in_boxes_kernel": ("hbm", 16.0 * F * P_pts + 28.0 * F * boxes_h.shape[1]),
            "vcn_chain_pose": ("tensor", 2.0 * (64 * 128 + 128 * 1024) * RESAMPLE * obj_rank0),
            "vcn_chain_enc1": ("tensor", 2.0 * (128 * 256) * RESAMPLE * obj_rank0),
            "vcn_chain_enc2": ("tensor", FLOP_ENC2_EXEC * obj_rank0),
            "vcn_forward": ("tensor", FLOP_PER_OBJ * obj_rank0),
            "dynamic_voxelize": ("hbm", 16.0 * out["num_voxel_points"] + 32.0 * n_voxels),
            "knn_surface_select": ("alu", None), "knn_prepare_kernel": ("alu", None), "knn_scan_kernel": ("alu", None), "knn_emit_kernel": ("hbm", None),
            "largest_cluster": ("alu", None), "crop": ("hbm", None),
            "splice": ("hbm", 13.0 * F * P_pts + 12.0 * RESAMPLE * obj_rank0),
        }
        stages = []
        for name, (cnt, tot_ms) in prof.items():
            bound, work = alg.get(name, (None, None))
            row = {"group": name, "launches_per_step": cnt / steps, "ms_per_step": tot_ms / steps, "share_of_step": tot_ms / ms_res,
                   "bound": bound}
            if work is not None and tot_ms > 0:
                rate = work * steps / (tot_ms / 1e3)
                if bound == "hbm":
                    row.update(achieved=rate / 1e9, unit="GB/s", frac=rate / 1e9 / pk["hbm_gbs"])
                else:
                    row.update(achieved=rate / 1e12, unit="TFLOP/s", frac=rate / 1e12 / pk["bf16_tflops"])
            stages.append(row)
        cnt2, ms2 = prof.get("vcn_chain_enc2", (0, 0.0))
        achieved = FLOP_ENC2_EXEC * obj_rank0 * steps / (ms2 / 1e3) / 1e12 if ms2 > 0 else None
        peak = pk["bf16_tflops"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("vcn_chain_kernel<2>", {}).get("dram_bytes_per_launch")
        import oracle   # cpu_baseline leg only (rank 0, N = 1)
        cpu = cpu_baseline(2, oracle.make_state_dict("VCN_VC", 0), os.cpu_count() or 1) if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": "objects/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "C2: synthetic Waymo-like 64-beam frame (180k pts, 50 car boxes, 1024 pts/object), random-init VCN_VC",
                       "frames_per_step_per_gpu": F, "objects_per_step": int(round(n_obj_res / steps)), "sel_k": SEL_K,
                       "cluster_eps": CLUSTER_EPS, "splice_thresh": SPLICE_THRESH, "l2": "flushed (256 MB write) before every step, inside the timed region",
                       "parallelism": f"frame-sharded x{world}"},
            "voxelized_mpts_per_sec": n_vox_pts / (ms_res / 1e3) / 1e6, "voxels_per_step_rank0": int(n_voxels),
            "e2e": {"value": e2e, "unit": "objects/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "vcn_chain_kernel<2> (enc2: mlp_conv2.0 local half + mlp_conv2.3 + max-pool, tcgen05)",
                         "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if achieved else None, "traffic": traffic,
                         "peak_source": pk["source"] + " burst bf16 (cuBLAS); sustained is %.0f" % pk["bf16_tflops_sustained"],
                         "launches": cnt2, "us_per_launch": 1e3 * ms2 / cnt2 if cnt2 else None,
                         "flop_per_object": FLOP_ENC2_EXEC,
                         "note": "flops as executed (W_local x + per-object bias: 655,360 MAC/pt); the reference graph's "
                                 "cat([global, local]) -> conv form is 786,432 MAC/pt (SURVEY.md §8a6)",
                         "achieved_reference_graph": achieved * FLOP_ENC2_REF / FLOP_ENC2_EXEC if achieved else None},
            "stages": stages,
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)    # ~0.3 s per timed leg: box-level jitter (PCIe, host) averages out
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=8, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SEEVCN_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
