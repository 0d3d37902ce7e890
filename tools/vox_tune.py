#!/usr/bin/env python
"""Times the dynamic voxelizer's four kernels on the bench frames (library CUDA-event scopes) for every bucket width.
usage (GPU box): python tools/vox_tune.py [--frames 8]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=8); ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import numpy as np, torch
    import bench
    from seevcn_b200 import _abi
    from seevcn_b200.pcdet.models.backbones_3d.vfe.dynamic_mean_vfe import dynamic_voxelize_frames
    from seevcn_b200.pipeline import WAYMO_VOXEL_CFG
    dev = torch.device("cuda", 0)
    pts, boxes = bench.make_inputs(args.frames, 1000)
    d = torch.from_numpy(pts).to(dev)
    rng = np.random.default_rng(0)
    # 300 "collapsed" completed clouds (what random-init weights give) + 300 car-sized ones
    O = 300
    cen = pts.reshape(-1, 3)[rng.integers(0, pts.shape[0] * pts.shape[1], 2 * O)]
    small = cen[:O, None, :] + rng.normal(0, 0.05, (O, 1024, 3))
    big = cen[O:, None, :] + rng.uniform(-1, 1, (O, 1024, 3)) * [2.1, 1.0, 0.8]
    for name, obj in (("no objects", None), ("collapsed objects", small), ("car-sized objects", big)):
        od = of = None
        if obj is not None:
            od = torch.from_numpy(obj.astype(np.float32)).to(dev)
            of = torch.from_numpy(np.sort(rng.integers(0, args.frames, O)).astype(np.int32)).to(dev)
        for logw in (10,):
            for _ in range(3):
                c, f, n, m = dynamic_voxelize_frames(d, od, of, *WAYMO_VOXEL_CFG)
            torch.cuda.synchronize()
            _abi.prof_enable(True)
            for _ in range(args.reps):
                c, f, n, m = dynamic_voxelize_frames(d, od, of, *WAYMO_VOXEL_CFG)
            torch.cuda.synchronize()
            _abi.prof_enable(False)
            prof = _abi.prof_report()
            print(f"{name:20s} logw {logw:2d} voxels {int(m.item()):7d}  " +
                  "  ".join(f"{k.replace('dynvox_', '').replace('_kernel', '')} {1e3 * v[1] / v[0]:7.1f}us" for k, v in prof.items()))


if __name__ == "__main__":
    main()
