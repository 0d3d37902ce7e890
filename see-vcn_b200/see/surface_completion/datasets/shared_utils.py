"""Mask-based isolation on the B200 path: LiDAR -> image projection, instance-mask lookup, per-instance DBSCAN.

ref: see/surface_completion/datasets/custom_dataset/custom_dataset_objects.py:141-193 (map_pointcloud_to_image),
     see/surface_completion/datasets/shared_utils.py:36-106 (get_pts_in_mask),
     see/surface_completion/SEE_VCN.py:144-181 (isolate_det_pts).
The reference runs these per frame on the host (numpy float64, pycocotools, open3d); here the frame stays on the device
and three kernels (csrc/isolate.cu) do the work.  The instance masks arrive as binary images: turning the 2D detector's
polygons into masks (pycocotools annToMask, shapely shrink) is outside the path.
"""
import ctypes

import numpy as np
import torch

from .... import _abi


def _darray(vals, n):
    a = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(-1))
    assert a.size >= n, f"need {n} values, got {a.size}"
    return (ctypes.c_double * n)(*a[:n].tolist())


def map_pointcloud_to_image(points, calib, img_shape, camera_model="pinhole"):
    """points (N,3) float32 CUDA; calib {'intrinsic' (3,3), 'extrinsic' (4,4) or (3,4) lidar2cam, 'distcoeff' (>=5,)} host;
    img_shape (H, W[, C]) -> dict of CUDA tensors: pts_img (N,2) int32 pixel (u, v) or (-1,-1), fov_inds (N,) uint8,
    depth (N,) float32.  The reference returns the compacted in-view subsets; here the mask travels with the full
    arrays (``compact_imgfov`` gives the reference's form)."""
    points = points.contiguous()
    _abi.require_cuda(points)
    assert points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 3
    if camera_model not in ("pinhole", "equidistant"):
        raise NotImplementedError(camera_model)
    n = points.shape[0]
    dev = points.device
    H, W = int(img_shape[0]), int(img_shape[1])
    dist = np.zeros(5, np.float64)
    dc = np.asarray(calib["distcoeff"], np.float64).reshape(-1)
    dist[: min(5, dc.size)] = dc[:5]
    uv = torch.empty((n, 2), dtype=torch.int32, device=dev)
    fov = torch.empty((n,), dtype=torch.uint8, device=dev)
    depth = torch.empty((n,), dtype=torch.float32, device=dev)
    with _abi.device_guard(dev):
        _abi.check(_abi.lib().seevcn_project_points(n, _abi.ptr(points), _darray(np.asarray(calib["extrinsic"])[:3, :], 12),
                                                    _darray(calib["intrinsic"], 9), _darray(dist, 5),
                                                    1 if camera_model == "equidistant" else 0, W, H, _abi.ptr(uv), _abi.ptr(fov),
                                                    _abi.ptr(depth), _abi.stream()))
    return {"pc_lidar": points, "pts_img": uv, "fov_inds": fov, "depth": depth, "img_shape": (H, W)}


def compact_imgfov(imgfov):
    """The reference's return value (host numpy): in-view points, their pixels and (u, v, depth) rows."""
    m = imgfov["fov_inds"].bool()
    return {"pc_lidar": imgfov["pc_lidar"][m].cpu().numpy(), "pts_img": imgfov["pts_img"][m].cpu().numpy().astype(np.int64),
            "fov_inds": m.cpu().numpy()}


def get_pts_in_mask(masks, imgfov):
    """masks (I, H, W) uint8 CUDA binary instance masks (largest area first, like get_camera_instances) ->
    lists (I, N) int32 ascending point indices per instance, counts (I,) int32 (CUDA)."""
    masks = masks.contiguous()
    _abi.require_cuda(masks)
    assert masks.dtype == torch.uint8 and masks.dim() == 3
    I, H, W = masks.shape
    assert (H, W) == tuple(imgfov["img_shape"])
    uv, fov = imgfov["pts_img"], imgfov["fov_inds"]
    n = uv.shape[0]
    dev = masks.device
    lists = torch.empty((I, max(n, 1)), dtype=torch.int32, device=dev)
    counts = torch.zeros((I,), dtype=torch.int32, device=dev)
    with _abi.device_guard(dev):
        _abi.check(_abi.lib().seevcn_points_in_masks(n, I, W, H, _abi.ptr(uv), _abi.ptr(fov), _abi.ptr(masks), _abi.ptr(lists),
                                                     _abi.ptr(counts), _abi.stream()))
    return lists, counts


def isolate_det_pts(points, lists, counts, vres, eps_scaling, min_eps, max_eps, min_cluster=10, min_points=3, eps=None):
    """ref: SEE_VCN.isolate_det_pts (SEE_VCN.py:144-181).  points (N,3) CUDA, lists/counts from ``get_pts_in_mask`` ->
    (cluster lists (I, N) int32, cluster counts (I,) int32 [0 = instance dropped], eps (I,) float64), all CUDA.
    ``eps`` fixes the DBSCAN radius instead of the range-adaptive rule."""
    _abi.require_cuda(points, lists, counts)
    I, stride = lists.shape
    dev = points.device
    out_lists = torch.empty_like(lists)
    out_counts = torch.zeros((I,), dtype=torch.int32, device=dev)
    out_eps = torch.zeros((I,), dtype=torch.float64, device=dev)
    # instances too large for shared memory (> 9600 points: a mask that swallowed the road) cluster in this scratch
    ws = _abi.workspace(dev, 256 + 24 * (stride + 64) * min(I, 4), "dbscan")
    with _abi.device_guard(dev):
        _abi.check(_abi.lib().seevcn_dbscan_largest(I, stride, _abi.ptr(points), _abi.ptr(lists), _abi.ptr(counts),
                                                    0 if eps is not None else 1, float(eps or 0.0), float(vres), float(eps_scaling),
                                                    float(min_eps), float(max_eps), int(min_points), int(min_cluster),
                                                    _abi.ptr(out_lists), _abi.ptr(out_counts), _abi.ptr(out_eps), _abi.ptr(ws),
                                                    ws.numel(), _abi.stream()))
    return out_lists, out_counts, out_eps


def resample_instances(points, lists, counts, obj_inst, n_points=1024, seed=0):
    """ResamplePoints (data_transforms.py:247-262) for the kept instances -> (O, n_points, 3) float32 CUDA."""
    _abi.require_cuda(points, lists, counts, obj_inst)
    O = obj_inst.shape[0]
    out = torch.empty((O, n_points, 3), dtype=torch.float32, device=points.device)
    with _abi.device_guard(points.device):
        _abi.check(_abi.lib().seevcn_resample_lists(O, n_points, lists.shape[1], int(seed) & 0xffffffff, _abi.ptr(points),
                                                    _abi.ptr(lists), _abi.ptr(counts), _abi.ptr(obj_inst), _abi.ptr(out), _abi.stream()))
    return out
