"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithms on the SEE-VCN object-completion +
voxelization path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
(``see-vcn_b200/``) never does.

Pieces
  liboracle.so (oracle.c)         integer/index work in plain C: points-in-boxes (GPU and CPU
                                  semantics), FPS, kNN, surface select, dynamic + hard voxelization,
                                  MeanVFE, Chamfer
  vcn_forward_ref (this file)     fp32 torch restatement of VCN_VC / VCN_CN forward from a state-dict
  _ref/libref_kernels.so          the reference's OWN .cu files compiled for sm_100a (exact GPU oracle
                                  for points_in_boxes / FPS / gather / group; built by `make ref`)

Pinning: the reference ships no golden vectors for this path.  The oracle is pinned
by *executing reference source*: tests/golden/make_golden.py imports the reference's VCN_VC / VCN_CN
/ MeanVFE / partial_with_KDTree (python) and builds its points_in_boxes_cpu (C++) in this container
and commits input/output vectors; tests/test_oracle.py checks this package against them.  Pieces with
no runnable reference (spconv hard voxelization, torch_scatter mean, open3d crop) are marked
"parity unpinned" where they are defined.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_I = ctypes.c_int


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        L.orc_points_in_boxes_gpu.argtypes = [_I, _I, _I, _f, _f, _i, ctypes.c_void_p]
        L.orc_points_in_boxes_cpu.argtypes = [_I, _I, _f, _f, _i]
        L.orc_fps.argtypes = [_I, _I, _I, _f, _f, _i]
        L.orc_knn.argtypes = [_I, _I, _I, _I, _f, _f, ctypes.c_void_p, _i]
        L.orc_knn_surface_select.argtypes = [_I, _I, _I, _I, _I, _f, _f, _f, _i]
        L.orc_largest_cluster.argtypes = [_I, _I, _I, ctypes.c_double, _I, _f, _f, _i]
        L.orc_dynamic_voxelize.argtypes = [_I, _I, _f, _f, _f, _i, _i, _f, _i]
        L.orc_dynamic_voxelize.restype = _I
        L.orc_hard_voxelize.argtypes = [_I, _I, _f, _f, _f, _i, _I, _I, _f, _i, _i]
        L.orc_hard_voxelize.restype = _I
        L.orc_mean_vfe.argtypes = [_I, _I, _I, _f, _f, _f]
        L.orc_chamfer.argtypes = [_I, _I, _I, _f, _f, _f]
        L.orc_nearest_dist.argtypes = [_I, _I, _f, _f, np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]
        _LIB = L
    return _LIB


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------------------ crop --
def points_in_boxes_gpu(points, boxes, return_slack=False):
    """ref: roiaware_pool3d_kernel.cu:313-336.  points (B,P,3), boxes (B,T,7) -> (B,P) int32.
    ``return_slack``: also the per-point distance from the nearest decision boundary."""
    points, boxes = _c(points), _c(boxes)
    B, P, _ = points.shape
    T = boxes.shape[1]
    out = np.empty((B, P), np.int32)
    slack = np.empty((B, P), np.float32)
    lib().orc_points_in_boxes_gpu(B, T, P, boxes, points, out, slack.ctypes.data_as(ctypes.c_void_p))
    return (out, slack) if return_slack else out


def points_in_boxes_cpu(points, boxes):
    """ref: roiaware_pool3d.cpp:121-168.  points (P,3), boxes (T,7) -> (T,P) int32 0/1."""
    points, boxes = _c(points), _c(boxes)
    out = np.empty((boxes.shape[0], points.shape[0]), np.int32)
    lib().orc_points_in_boxes_cpu(boxes.shape[0], points.shape[0], boxes, points, out)
    return out


def crop_lists(box_idx, num_boxes):
    """Per-box ascending point lists from (P,) box indices: what ``points[idx == k]`` gives the
    reference's callers (e.g. waymo_dataset.py:363-372)."""
    return [np.nonzero(box_idx == k)[0].astype(np.int32) for k in range(num_boxes)]


def resample_points(pts, n_points, rng):
    """ref: ResamplePoints.__call__, data_transforms.py:254-262 (seeded rng instead of np.random).
    Returns (resampled (n,3), choice (n,) indices into the tiled list)."""
    reps = int(np.ceil(n_points / len(pts)))
    tiled = np.tile(pts, (reps, 1))
    choice = rng.permutation(tiled.shape[0])[:n_points]
    return tiled[choice], choice.astype(np.int32)


# ------------------------------------------------------------------------------- FPS --
def furthest_point_sample(xyz, npoint, return_temp=False):
    """ref: sampling_gpu.cu:100-216.  xyz (B,N,3) -> (B,npoint) int32"""
    xyz = _c(xyz)
    B, N, _ = xyz.shape
    idx = np.empty((B, npoint), np.int32)
    temp = np.empty((B, N), np.float32)
    lib().orc_fps(B, N, npoint, xyz, temp, idx)
    return (idx, temp) if return_temp else idx


def gather_operation(features, idx):
    """ref: sampling_gpu.cu:15-31.  features (B,C,N), idx (B,M) -> (B,C,M)"""
    return np.take_along_axis(features, idx[:, None, :].astype(np.int64), axis=2)


def grouping_operation(features, idx):
    """ref: group_points_gpu.cu:53-73.  features (B,C,N), idx (B,P,S) -> (B,C,P,S)"""
    B, Pn, S = idx.shape
    return gather_operation(features, idx.reshape(B, Pn * S)).reshape(B, features.shape[1], Pn, S)


# ------------------------------------------------------------------------------- kNN --
def knn(k, ref, query):
    """cKDTree.query semantics (sampling.py:30-34).  ref (B,R,3), query (B,Q,3) -> dist (B,Q,k), idx (B,Q,k)"""
    ref, query = _c(ref), _c(query)
    B, R, _ = ref.shape
    Q = query.shape[1]
    dist = np.empty((B, Q, k), np.float32)
    idx = np.empty((B, Q, k), np.int32)
    lib().orc_knn(B, R, Q, k, ref, query, dist.ctypes.data_as(ctypes.c_void_p), idx)
    return dist, idx


def knn_gap(ref, query, k):
    """Relative gap between the k-th and (k+1)-th neighbour distance per query (tie-freeness check)."""
    d, _ = knn(min(k + 1, ref.shape[1]), ref, query)
    if d.shape[2] <= k:
        return np.full(d.shape[:2], np.inf, np.float32)
    return (d[:, :, k] - d[:, :, k - 1]) / np.maximum(d[:, :, k], 1e-30)


def get_partial_mesh_batch(partial, complete, k=20, surface_pts=1024):
    """ref: sampling.py:8-41,69-80 -> (B,surface_pts,3) float32, counts (B,)"""
    partial, complete = _c(partial), _c(complete)
    B, NP, _ = partial.shape
    R = complete.shape[1]
    out = np.empty((B, surface_pts, 3), np.float32)
    cnt = np.empty((B,), np.int32)
    lib().orc_knn_surface_select(B, NP, R, k, surface_pts, partial, complete, out, cnt)
    return out, cnt


def get_largest_cluster_batch(pc, eps=0.4, min_points=1, total_pts=1024):
    """ref: sampling.py:83-109 (open3d DBSCAN -> largest cluster) — PARITY UNPINNED.  pc (B,N,3) ->
    (B,total_pts,3), member counts (B,)"""
    pc = _c(pc)
    B, N, _ = pc.shape
    out = np.empty((B, total_pts, 3), np.float32)
    cnt = np.empty((B,), np.int32)
    lib().orc_largest_cluster(B, N, total_pts, float(eps), int(min_points), pc, out, cnt)
    return out, cnt


def distinct_rows(clustered, counts):
    """How many of an object's clustered rows np.unique keeps (SEE_VCN.py:113,244 applies np.unique to the stacked
    clustered clouds): the number of distinct rows among the first counts[o] rows of clustered[o].  (B,) int32."""
    return np.asarray([len(np.unique(clustered[o][: int(counts[o])], axis=0)) if counts[o] > 0 else 0
                       for o in range(len(clustered))], np.int32)


# ------------------------------------------------------------------- mask-based isolation --
def map_pointcloud_to_image(points, calib, img_shape, camera_model="pinhole"):
    """ref: CustomDatasetObjects.map_pointcloud_to_image, datasets/custom_dataset/custom_dataset_objects.py:141-193,
    restated line by line (float64 numpy).  -> dict pc_lidar (K,3), pts_img (K,2) int, fov_inds (N,) bool"""
    points = np.asarray(points, dtype=np.float64)
    IMG_H, IMG_W = int(img_shape[0]), int(img_shape[1])
    cameramat = np.asarray(calib["intrinsic"], np.float64)
    lidar2cam = np.asarray(calib["extrinsic"], np.float64)
    distcoeff = np.zeros(5); dc = np.asarray(calib["distcoeff"], np.float64).reshape(-1); distcoeff[: min(5, dc.size)] = dc[:5]
    pts_3d_hom = np.hstack((points, np.ones((points.shape[0], 1)))).T
    pts_imgframe = (lidar2cam[:3, :] @ pts_3d_hom).T
    with np.errstate(divide="ignore", invalid="ignore"):
        tmpxC = pts_imgframe[:, 0] / pts_imgframe[:, 2]
        tmpyC = pts_imgframe[:, 1] / pts_imgframe[:, 2]
    pre = (pts_imgframe[:, 2] > 0) & (abs(tmpxC) < np.arctan(IMG_W / IMG_H))
    tmpxC, tmpyC, depth = tmpxC[pre], tmpyC[pre], pts_imgframe[:, 2][pre]
    r2 = tmpxC ** 2 + tmpyC ** 2
    if camera_model == "equidistant":
        r1 = np.sqrt(r2)
        a0 = np.arctan(r1)
        a1 = a0 * (1 + distcoeff[0] * (a0 ** 2) + distcoeff[1] * (a0 ** 4) + distcoeff[2] * (a0 ** 6) + distcoeff[3] * (a0 ** 8))
        u = (a1 / r1) * tmpxC
        v = (a1 / r1) * tmpyC
    elif camera_model == "pinhole":
        tmpdist = 1 + distcoeff[0] * r2 + distcoeff[1] * (r2 ** 2) + distcoeff[4] * (r2 ** 3)
        u = tmpxC * tmpdist + 2 * distcoeff[2] * tmpxC * tmpyC + distcoeff[3] * (r2 + 2 * tmpxC ** 2)
        v = tmpyC * tmpdist + distcoeff[2] * (r2 + 2 * tmpyC ** 2) + 2 * distcoeff[3] * tmpxC * tmpyC
    else:
        raise NotImplementedError
    u = cameramat[0, 0] * u + cameramat[0, 2]
    v = cameramat[1, 1] * v + cameramat[1, 2]
    fov = (u > 0) & (u < IMG_W - 1) & (v > 0) & (v < IMG_H - 1)
    combined = np.zeros(pre.shape, dtype=bool)
    combined[pre] = fov
    uv = np.stack([u[fov], v[fov]], axis=1)
    # distance of u, v from the nearest rounding boundary (x.5): where the last float64 bit decides the pixel
    slack = np.abs((uv - np.floor(uv)) - 0.5).min(axis=1) if len(uv) else np.zeros((0,))
    return {"pc_lidar": points[combined].astype(np.float32), "pts_img": np.round(uv, 0).astype(int), "fov_inds": combined,
            "depth": depth[fov], "round_slack": slack}


def get_pts_in_mask(masks, imgfov):
    """ref: get_pts_in_mask, datasets/shared_utils.py:36-106 (binary masks given) -> list of index arrays into the frame."""
    idx = np.nonzero(imgfov["fov_inds"])[0]
    px = imgfov["pts_img"]
    return [idx[np.asarray(m[px[:, 1], px[:, 0]], dtype=bool)] for m in masks]


def cluster_dbscan(xyz, eps, min_points):
    """open3d 0.14 PointCloud::ClusterDBSCAN restated as the sequential expansion it is (PARITY UNPINNED: open3d absent):
    float64 radius search with strict dist < eps (the point itself included), clusters numbered in discovery order,
    a border point joins the first cluster that reaches it.  -> labels (N,) int, -1 = noise."""
    p = np.asarray(xyz, dtype=np.float64)
    n = len(p)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2)
    d2 = (d2[:, :, 0] + d2[:, :, 1]) + d2[:, :, 2]
    nbs = [np.nonzero(d2[i] < eps * eps)[0] for i in range(n)]
    labels = np.full(n, -2, dtype=np.int64)
    cluster = 0
    for idx in range(n):
        if labels[idx] != -2:
            continue
        if len(nbs[idx]) < min_points:
            labels[idx] = -1
            continue
        nxt = list(nbs[idx]); seen = set(nxt) | {idx}
        labels[idx] = cluster
        while nxt:
            nb = nxt.pop(0)
            if labels[nb] == -1:
                labels[nb] = cluster
            if labels[nb] != -2:
                continue
            labels[nb] = cluster
            if len(nbs[nb]) >= min_points:
                for q in nbs[nb]:
                    if q not in seen:
                        seen.add(q); nxt.append(q)
        cluster += 1
    return labels


def isolate_det_pts(points, inst_indices, vres, eps_scaling, min_eps, max_eps, min_cluster=10):
    """ref: SEE_VCN.isolate_det_pts, see/surface_completion/SEE_VCN.py:144-181 -> list over instances of the kept
    cluster's frame indices (None when the instance is dropped), and the eps used."""
    out, eps_used = [], []
    for ind in inst_indices:
        xyz = np.asarray(points, np.float64)[ind]
        sel, eps = None, 0.0
        if xyz.shape[0] > min_cluster:
            dist = np.linalg.norm(xyz.mean(axis=0))
            ring_height = dist * np.tan(vres * np.pi / 180)
            eps = float(np.clip(eps_scaling * ring_height, a_max=max_eps, a_min=min_eps))
            labels = cluster_dbscan(xyz, eps, 3)
            y = np.bincount(labels[labels >= 0])
            if len(y) > 0:
                members = np.argwhere(labels == np.argmax(y)).reshape(-1)
                if len(members) > min_cluster:
                    sel = ind[members]
        out.append(sel); eps_used.append(eps)
    return out, eps_used


# --------------------------------------------------------------------------- splice --
def nearest_dist(points, completed, brute=False):
    """Distance (float64) from every row of points (P,3) to its nearest row of completed (K,3): what
    open3d's compute_point_cloud_distance returns (SEE_VCN.py:258).  KD-tree (scipy) by default, the C
    brute force with brute=True (small cases; the two are checked against each other in tests/test_oracle.py)."""
    points, completed = _c(points), _c(completed)
    if len(completed) == 0:
        return np.full((len(points),), np.inf)
    if brute:
        d = np.empty((len(points),), np.float64)
        lib().orc_nearest_dist(len(points), len(completed), points, completed, d)
        return d
    from scipy.spatial import cKDTree
    return cKDTree(completed.astype(np.float64)).query(points.astype(np.float64), k=1)[0]


def replace_with_completed_pts(points, sc_instances, point_dist_thresh=0.1, brute=False):
    """ref: SEE_VCN.replace_with_completed_pts, see/surface_completion/SEE_VCN.py:247-265 — PARITY UNPINNED
    (open3d absent).  points (P,3), sc_instances (K,3) or None -> (merged (K+kept,3) = vstack(sc_instances,
    surviving originals in order), keep mask (P,) bool)."""
    points = _c(points)
    if sc_instances is None:
        return points, np.ones((len(points),), bool)
    sc_instances = _c(sc_instances)
    keep = ~(nearest_dist(points, sc_instances, brute) < point_dist_thresh)
    return np.vstack((sc_instances, points[keep])), keep


def all_instances(clustered, counts=None):
    """ref: SEE_VCN.py:244 — np.unique(np.vstack(sc_model_ret['clustered']), axis=0): the distinct completed points of
    all objects of a frame, rows in lexicographic order.  counts (O,) limits object o to its first counts[o] rows
    (objects whose every point was noise contribute nothing)."""
    clustered = _c(clustered)
    rows = [clustered[o][: (len(clustered[o]) if counts is None else int(counts[o]))] for o in range(len(clustered))]
    rows = [r for r in rows if len(r)]
    if not rows:
        return np.zeros((0, 3), np.float32)
    return np.unique(np.vstack(rows), axis=0)


# ------------------------------------------------------------------------- voxelize --
def dynamic_voxelize(points, pc_range, voxel_size, grid_size):
    """ref: dynamic_mean_vfe.py:49-76.  points (N,1+C) -> coords (M,4) [b,z,y,x], feats (M,C), counts (M,)
    (torch_scatter absent -> the order-free value its fp32 atomic sums approximate: float64 sum / count rounded to
    fp32; checked against the reference module itself, run with a scatter_mean stub, in tests/test_oracle.py)."""
    points = _c(points)
    N, C1 = points.shape
    C = C1 - 1
    coords = np.empty((max(N, 1), 4), np.int32)
    feats = np.empty((max(N, 1), C), np.float32)
    counts = np.empty((max(N, 1),), np.int32)
    m = lib().orc_dynamic_voxelize(N, C, points, _c(pc_range), _c(voxel_size), _c(grid_size, np.int32), coords, feats,
                                   counts)
    return coords[:m].copy(), feats[:m].copy(), counts[:m].copy()


def hard_voxelize(points, pc_range, voxel_size, grid_size, max_points, max_voxels):
    """spconv v1 loop restated — PARITY UNPINNED (spconv unvendored).  points (N,C) ->
    voxels (M,T,C), coordinates (M,3) zyx, num_points (M,)"""
    points = _c(points)
    N, C = points.shape
    voxels = np.empty((max_voxels, max_points, C), np.float32)
    coords = np.empty((max_voxels, 3), np.int32)
    num = np.empty((max_voxels,), np.int32)
    m = lib().orc_hard_voxelize(N, C, points, _c(pc_range), _c(voxel_size), _c(grid_size, np.int32), max_points,
                                max_voxels, voxels, coords, num)
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy()


def mean_vfe(voxels, num_points):
    """ref: mean_vfe.py:23-29"""
    voxels, num_points = _c(voxels), _c(num_points)
    M, T, C = voxels.shape
    out = np.empty((M, C), np.float32)
    lib().orc_mean_vfe(M, T, C, voxels, num_points, out)
    return out


def chamfer_l2(a, b):
    """ref: chamfer.cu:15-145 + extensions/chamfer_dist/__init__.py:28-44: mean(dist1) + mean(dist2), squared L2.
    a (B,N,3), b (B,M,3) -> scalar per batch (B,)"""
    a, b = _c(a), _c(b)
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = np.empty((B, N), np.float32)
    d2 = np.empty((B, M), np.float32)
    lib().orc_chamfer(B, N, M, a, b, d1)
    lib().orc_chamfer(B, M, N, b, a, d2)
    return d1.mean(axis=1) + d2.mean(axis=1)


# ----------------------------------------------------------------------- VCN forward --
def make_state_dict(model_name="VCN_VC", seed=0, num_coarse=1024):
    """Deterministic random-init weights with the reference's state-dict keys and shapes
    (VCN_VC.py:116-141, VCN_CN.py:118-119), torch default-init scale, BatchNorm running stats
    randomised (mean N(0,0.1), var U(0.5,1.5)) so folding is exercised.
    CPU torch.Generator -> identical on every machine with the same torch build."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, cout, cin, conv):
        bound = 1.0 / np.sqrt(cin)
        w = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
        sd[name + ".weight"] = w[:, :, None].contiguous() if conv else w
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(name, c):
        sd[name + ".weight"] = torch.rand(c, generator=g) * 0.5 + 0.75
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
        sd[name + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    if model_name == "VCN_VC":
        lin("pose_encoder.0", 64, 3, True); lin("pose_encoder.2", 128, 64, True); lin("pose_encoder.4", 1024, 128, True)
        lin("pose_fc.0", 512, 1024, False); lin("pose_fc.2", 9, 512, False)
    lin("encoder.mlp_conv1.0", 128, 3, True); bn("encoder.mlp_conv1.1", 128); lin("encoder.mlp_conv1.3", 256, 128, True)
    lin("encoder.mlp_conv2.0", 512, 512, True); bn("encoder.mlp_conv2.1", 512); lin("encoder.mlp_conv2.3", 1024, 512, True)
    lin("shape_fc.0", 1024, 1024, False); lin("shape_fc.2", 1024, 1024, False); lin("shape_fc.4", 3 * num_coarse, 1024, False)
    if model_name == "VCN_VC":
        lin("final_conv.0", 512, 1029, True); bn("final_conv.1", 512); lin("final_conv.3", 512, 512, True)
        bn("final_conv.4", 512); lin("final_conv.6", 3, 512, True)
    return sd


def _rot_z(points, angle):
    """ref: rotate_points_along_z, utils/transform.py:33-58 (row-vector p . R)"""
    import torch
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    R = torch.stack((c, s, z, -s, c, z, z, z, o), dim=1).view(-1, 3, 3).float()
    return torch.matmul(points, R), R


def vcn_forward_ref(sd, pts, gt_boxes=None, model_name="VCN_VC", dtype=None):
    """fp32 torch restatement of VCN_VC.forward (VCN_VC.py:178-213) / VCN_CN.forward (VCN_CN.py:142-157)
    in eval mode, straight from a state-dict.  pts (B,N,3) -> dict like the reference."""
    import torch
    import torch.nn.functional as F
    pts = torch.as_tensor(pts).float()
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}

    def conv(x, n):   # x (B,C,N)
        return torch.einsum("oc,bcn->bon", sd[n + ".weight"][:, :, 0], x) + sd[n + ".bias"][None, :, None]

    def bnorm(x, n):
        return (x - sd[n + ".running_mean"][None, :, None]) / torch.sqrt(sd[n + ".running_var"][None, :, None] + 1e-5) \
            * sd[n + ".weight"][None, :, None] + sd[n + ".bias"][None, :, None]

    def linear(x, n):
        return x @ sd[n + ".weight"].t() + sd[n + ".bias"]

    def encoder(x):   # FeatureEncoder.forward, VCN_VC.py:95-106
        n = x.shape[2]
        f = conv(F.relu(bnorm(conv(x, "encoder.mlp_conv1.0"), "encoder.mlp_conv1.1")), "encoder.mlp_conv1.3")
        g = f.max(dim=2, keepdim=True)[0]
        f = torch.cat([g.expand(-1, -1, n), f], dim=1)
        f = conv(F.relu(bnorm(conv(f, "encoder.mlp_conv2.0"), "encoder.mlp_conv2.1")), "encoder.mlp_conv2.3")
        return f.max(dim=2)[0]

    def shape_fc(feat):
        h = F.relu(linear(feat, "shape_fc.0"))
        h = F.relu(linear(h, "shape_fc.2"))
        return linear(h, "shape_fc.4")

    B = pts.shape[0]
    ret = {}
    if model_name == "VCN_VC":
        ang = torch.atan2(pts[:, :, 1].mean(dim=1), pts[:, :, 0].mean(dim=1))
        fview, _ = _rot_z(pts, -ang)
        mean = fview.mean(dim=1, keepdim=True)
        x = (fview - mean).permute(0, 2, 1)
        h = F.leaky_relu(conv(x, "pose_encoder.0"))
        h = F.leaky_relu(conv(h, "pose_encoder.2"))
        pose_feat = conv(h, "pose_encoder.4").max(dim=2)[0]
        rel = linear(F.leaky_relu(linear(pose_feat, "pose_fc.0")), "pose_fc.2")
        centre = mean + rel[:, :3].unsqueeze(1)
        xr, yr = rel[:, 3:6], rel[:, 6:9]
        xn = xr / torch.clamp(xr.norm(dim=1, keepdim=True), min=1e-8)
        zn = torch.cross(xn, yr, dim=1)
        zn = zn / torch.clamp(zn.norm(dim=1, keepdim=True), min=1e-8)
        yn = torch.cross(zn, xn, dim=1)
        rot = torch.stack((xn, yn, zn), dim=2)
        pc_cn = torch.matmul(fview - centre, rot.permute(0, 2, 1))
        coarse = shape_fc(encoder(pc_cn.permute(0, 2, 1))).reshape(B, -1, 3)
        coarse_vc = torch.matmul(coarse, rot) + centre
        ret["coarse"], _ = _rot_z(coarse_vc, ang)
        _, Rh = _rot_z(coarse_vc[:, :1], ang)
        ret["reg_rot"] = torch.matmul(rot, Rh)
        ret["reg_centre"] = _rot_z(centre, ang)[0].squeeze(1)
    else:
        gt = torch.as_tensor(gt_boxes).float()
        pc = _rot_z(pts - gt[:, None, :3], -gt[:, 6])[0] / gt[:, 3].view(-1, 1, 1)
        coarse = shape_fc(encoder(pc.permute(0, 2, 1))).reshape(B, -1, 3)
        ret["coarse"] = _rot_z(coarse * gt[:, 3].view(-1, 1, 1), gt[:, 6])[0] + gt[:, None, :3]
    return ret


# ------------------------------------------------------- reference GPU kernels (_ref) --
def ref_kernels():
    """The reference's own CUDA kernels compiled for sm_100a (oracle/_ref, `make -C oracle ref`).
    Returns None when not built.  GPU box only."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libref_kernels.so")
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        P = ctypes.c_void_p
        R.ref_points_in_boxes.argtypes = [_I, _I, _I, P, P, P]
        R.ref_fps.argtypes = [_I, _I, _I, P, P, P]
        R.ref_gather.argtypes = [_I, _I, _I, _I, P, P, P]
        R.ref_group.argtypes = [_I, _I, _I, _I, _I, P, P, P]
        _REF = R
    return _REF
