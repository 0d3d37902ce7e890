#!/usr/bin/env python
"""Multi-GPU check of dist.FrameGather (run under torchrun, N >= 2): every rank completes its own frames, pushes three
batches, and compares what landed in EVERY rank's slot with what that rank produced (checksums exchanged with
all_gather_object).  usage: python -m torch.distributed.run --nproc-per-node 2 tools/gather_check.py [peer|nccl]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import bench
    from seevcn_b200.pipeline import CompletionPipeline
    from seevcn_b200 import dist as sdist
    backend = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    F = 2
    pipe = CompletionPipeline("VCN_VC", bench.seeded_state_dict(), dev, sel_k=bench.SEL_K, cluster_eps=bench.CLUSTER_EPS,
                              splice_thresh=bench.SPLICE_THRESH, streams=2)
    pts, boxes = bench.make_inputs(3 * F, 5000 + 100 * rank)
    P, T = pts.shape[1], boxes.shape[1]
    g = sdist.FrameGather(dev, world, rank, max_obj=F * T, rows_per_obj=bench.RESAMPLE, frames=F,
                          max_rows=F * P + F * T * bench.RESAMPLE, backend=backend)
    ok = True
    batches = [(torch.from_numpy(pts[i * F:(i + 1) * F]).to(dev), torch.from_numpy(boxes[i * F:(i + 1) * F]).to(dev)) for i in range(3)]
    for k, out in enumerate(pipe.run_stream(batches)):
        gen = g.push(out, frame_offset=rank * 100 + k * F)
        mine = {"n_obj": int(out["clustered"].shape[0]), "m": int(out["voxel_coords"].shape[0]),
                "clu": float(out["clustered"].double().sum()), "feat": float(out["voxel_features"].double().sum()),
                "nums": int(out["voxel_num_points"].long().sum()),
                "coords": int((out["voxel_coords"].long() * torch.tensor([1000003, 1009, 101, 7], device=dev)).sum()),
                "off": rank * 100 + k * F}
        every = [None] * world
        dist.all_gather_object(every, mine)
        parts = g.parts(gen)
        for r, (want, got) in enumerate(zip(every, parts)):
            c = got["voxel_coords"].long().clone()
            c[:, 0] -= want["off"]                      # parts() rebased the batch index to the global frame index
            have = {"n_obj": int(got["clustered"].shape[0]), "m": int(got["voxel_coords"].shape[0]),
                    "clu": float(got["clustered"].double().sum()), "feat": float(got["voxel_features"].double().sum()),
                    "nums": int(got["voxel_num_points"].long().sum()),
                    "coords": int((c * torch.tensor([1000003, 1009, 101, 7], device=dev)).sum())}
            for key, v in have.items():
                if v != want[key]:
                    ok = False
                    print(f"rank {rank} push {k} slot {r}: {key} {v} != {want[key]}")
        if r == world - 1 and rank == 0:
            print(f"push {k}: gen {gen}, objects {[e['n_obj'] for e in every]}, voxels {[e['m'] for e in every]}")
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("gather_check", g.describe(), "OK" if int(flag.item()) == 0 else "MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
