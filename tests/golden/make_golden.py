"""Generate the committed golden vectors by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container (needs /root/reference; it is not present on the GPU box):
    python tests/golden/make_golden.py

What runs, unmodified, from /root/reference:
  * see/surface_completion/models/vcn/models/VCN_VC.py / VCN_CN.py   (classes imported; absent
    third-party modules stubbed in sys.modules; Tensor.cuda patched to a no-op so
    normalize_vector's hard .cuda() (VCN_VC.py:15) runs on CPU)
  * see/surface_completion/models/vcn/utils/sampling.py  partial_with_KDTree (scipy cKDTree)
  * detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py  MeanVFE
  * detector3d/pcdet/ops/roiaware_pool3d/src/*.cpp,*.cu   points_in_boxes_cpu, built as a torch
    extension into oracle/_ref/ (skipped with a note if the build fails)
Weights come from oracle.make_state_dict(seed) — a CPU torch.Generator stream, so the GPU box
regenerates the same tensors without the 30 MB of parameters being committed.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def stub_modules():
    for name in ["open3d", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "pointnet2_ops",
                 "pointnet2_ops.pointnet2_utils", "chamfer", "easydict", "transforms3d", "transforms3d.euler", "cv2",
                 "tensorboardX", "shapely", "shapely.geometry", "timm", "timm.scheduler"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["easydict"].EasyDict = dict
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["pointnet2_ops"].pointnet2_utils = sys.modules["pointnet2_ops.pointnet2_utils"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["timm.scheduler"].CosineLRScheduler = object


def read_pcd_xyz(path):
    """binary PCD v0.7, FIELDS x y z float32 (demo/demo_data/pcd)"""
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii", "ignore").strip()
            if line.startswith("POINTS"):
                n = int(line.split()[1])
            if line.startswith("DATA"):
                assert "binary" in line
                break
        return np.frombuffer(f.read(n * 12), dtype=np.float32).reshape(n, 3).copy()


def synth_boxes(points, rng, n_boxes):
    """car-sized boxes centred on frame points so they are non-empty"""
    r = np.linalg.norm(points[:, :2], axis=1)
    cand = points[(r > 5) & (r < 50)]
    boxes = []
    while len(boxes) < n_boxes:
        c = cand[rng.integers(len(cand))]
        size = np.array([4.2, 2.0, 1.6]) * (1 + 0.1 * rng.standard_normal(3))
        box = np.concatenate([c + [0, 0, 0.3], size, [rng.uniform(-np.pi, np.pi)]])
        if all(np.linalg.norm(box[:2] - b[:2]) > 6.0 for b in boxes):
            boxes.append(box)
    return np.asarray(boxes, dtype=np.float32)


def main():
    stub_modules()
    torch.Tensor.cuda = lambda self, *a, **k: self   # VCN_VC.py:15 hard-codes .cuda()
    sys.path.insert(0, os.path.join(REF, "see", "surface_completion"))
    from models.vcn.models.VCN_VC import VCN_VC as RefVC   # noqa: E402
    from models.vcn.models.VCN_CN import VCN_CN as RefCN   # noqa: E402
    from models.vcn.utils.sampling import partial_with_KDTree   # noqa: E402

    rng = np.random.default_rng(1234)
    gold = {}

    # ---- VCN_VC / VCN_CN forward on car-like partial clouds ------------------------
    B, N = 3, 1024
    gt = np.stack([np.array([12.0, 3.0, -0.8, 4.3, 1.9, 1.6, 0.4]), np.array([-20.0, -8.0, -1.0, 4.6, 2.1, 1.7, -2.1]),
                   np.array([35.0, 14.0, -0.6, 3.9, 1.8, 1.5, 2.9])]).astype(np.float32)
    pts = []
    for b in range(B):   # two visible faces of the box, in the sensor frame
        l, w, h = gt[b, 3:6]
        u = rng.uniform(-0.5, 0.5, (N, 3)) * [l, w, h]
        face = rng.integers(0, 2, N)
        u[face == 0, 0] = -l / 2
        u[face == 1, 1] = -w / 2
        c, s = np.cos(gt[b, 6]), np.sin(gt[b, 6])
        R = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
        pts.append(u @ R + gt[b, :3])
    pts = np.asarray(pts, dtype=np.float32)
    gold["vcn_input"], gold["vcn_gt_boxes"] = pts, gt
    for name, cls in (("VCN_VC", RefVC), ("VCN_CN", RefCN)):
        sd = oracle.make_state_dict(name, seed=0)
        model = cls({})
        model.load_state_dict(sd)
        model.eval()
        with torch.no_grad():
            ret = model({"input": torch.from_numpy(pts), "gt_boxes": torch.from_numpy(gt)})
        for k, v in ret.items():
            gold[f"{name}.{k}"] = v.numpy().astype(np.float32)
        mine = oracle.vcn_forward_ref(sd, pts, gt, name)
        for k, v in ret.items():
            err = (mine[k] - v).abs().max().item()
            print(f"{name}.{k}: oracle restatement vs reference class max|diff| = {err:.3e}")
            assert err < 1e-4, err

    # ---- kNN surface selection: reference partial_with_KDTree -----------------------
    coarse = gold["VCN_VC.coarse"]
    for k in (10, 30):
        surf = np.stack([partial_with_KDTree(torch.from_numpy(pts[b]), torch.from_numpy(coarse[b]), k=k) for b in range(B)])
        gold[f"surface_k{k}"] = surf.astype(np.float32)

    # ---- MeanVFE (reference module, loaded by path to avoid importing spconv) ---------
    vfe_dir = os.path.join(REF, "detector3d", "pcdet", "models", "backbones_3d", "vfe")
    pkg = types.ModuleType("refvfe"); pkg.__path__ = [vfe_dir]; sys.modules["refvfe"] = pkg
    for mod in ("vfe_template", "mean_vfe"):
        spec = importlib.util.spec_from_file_location(f"refvfe.{mod}", os.path.join(vfe_dir, f"{mod}.py"))
        m = importlib.util.module_from_spec(spec); sys.modules[f"refvfe.{mod}"] = m; spec.loader.exec_module(m)
    MeanVFE = sys.modules["refvfe.mean_vfe"].MeanVFE
    M, T, C = 257, 5, 3
    num = rng.integers(0, T + 1, M).astype(np.float32)
    vox = rng.standard_normal((M, T, C)).astype(np.float32) * 20
    vox *= (np.arange(T)[None, :, None] < num[:, None, None])
    out = MeanVFE(model_cfg={}, num_point_features=C)({"voxels": torch.from_numpy(vox), "voxel_num_points": torch.from_numpy(num)})
    gold["meanvfe_voxels"], gold["meanvfe_num"], gold["meanvfe_out"] = vox, num, out["voxel_features"].numpy()

    # ---- points_in_boxes_cpu: the reference C++ built as-is -------------------------
    frame = read_pcd_xyz(os.path.join(REF, "demo", "demo_data", "pcd", "000001.pcd"))
    sub = frame[rng.permutation(len(frame))[:8192]]
    boxes = synth_boxes(frame, rng, 10)
    gold["pib_points"], gold["pib_boxes"] = sub, boxes
    try:
        from torch.utils.cpp_extension import load
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        src = os.path.join(REF, "detector3d", "pcdet", "ops", "roiaware_pool3d", "src")
        bdir = os.path.join(ROOT, "oracle", "_ref", "roiaware_pool3d_cuda")
        os.makedirs(bdir, exist_ok=True)
        ext = load(name="roiaware_pool3d_cuda", sources=[os.path.join(src, "roiaware_pool3d.cpp"),
                                                         os.path.join(src, "roiaware_pool3d_kernel.cu")],
                   build_directory=bdir, verbose=False)
        out = torch.zeros((len(boxes), len(sub)), dtype=torch.int)
        ext.points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(sub), out)
        gold["pib_cpu_out"] = np.packbits(out.numpy().astype(np.uint8), axis=1)
        print("points_in_boxes_cpu (reference build): in-box counts", out.sum(dim=1).tolist())
    except Exception as e:   # noqa: BLE001
        print("reference roiaware_pool3d build failed, pib_cpu_out omitted:", repr(e)[:200])

    np.savez_compressed(os.path.join(OUT, "golden_v1.npz"), **gold)
    print("wrote", os.path.join(OUT, "golden_v1.npz"), {k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
