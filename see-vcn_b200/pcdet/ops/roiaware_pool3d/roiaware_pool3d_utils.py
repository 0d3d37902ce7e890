"""Drop-in for ``pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils`` (the crop half).

ref: detector3d/pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41
"""
import numpy as np
import torch

from .... import _abi


def points_in_boxes_gpu(points, boxes):
    """
    :param points: (B, M, 3) float32 CUDA
    :param boxes: (B, T, 7) float32 CUDA [x, y, z, dx, dy, dz, heading], (x, y, z) is the box centre
    :return box_idxs_of_pts: (B, M) int32, lowest index of a box containing the point, background = -1
    ref: roiaware_pool3d_utils.py:28-41
    """
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    batch_size, num_points, _ = points.shape
    points = points.contiguous()
    boxes = boxes.contiguous()
    _abi.require_cuda(points, boxes)
    assert points.dtype == torch.float32 and boxes.dtype == torch.float32
    box_idxs_of_pts = torch.empty((batch_size, num_points), dtype=torch.int32, device=points.device)
    with _abi.device_guard(points.device):
        _abi.check(_abi.lib().seevcn_points_in_boxes(batch_size, boxes.shape[1], num_points, _abi.ptr(boxes),
                                                     _abi.ptr(points), _abi.ptr(box_idxs_of_pts), _abi.stream()))
    return box_idxs_of_pts


def points_in_boxes_cpu(points, boxes):
    """
    Args:
        points: (num_points, 3) numpy or CPU tensor
        boxes: (N, 7) [x, y, z, dx, dy, dz, heading], (x, y, z) is the box centre
    Returns:
        point_indices: (N, num_points) int32 0/1 (MARGIN 1e-2, all boxes — no first-box-wins)
    ref: roiaware_pool3d_utils.py:9-25.  "cpu" is the reference's name for host-buffer in/out;
    here the host buffers are staged through pinned memory and the test runs on the GPU.
    cos/sin of the headings are taken on the host (numpy float32 -> libm) so the result equals
    the reference's x86 build bit for bit.
    """
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    is_numpy = isinstance(points, np.ndarray)
    pts = torch.as_tensor(np.ascontiguousarray(points) if is_numpy else points).float().contiguous()
    bxs = torch.as_tensor(np.ascontiguousarray(boxes) if isinstance(boxes, np.ndarray) else boxes).float().contiguous()
    n_box, n_pts = bxs.shape[0], pts.shape[0]
    ang = (-bxs[:, 6]).contiguous()
    trig = torch.stack((torch.cos(ang), torch.sin(ang)), dim=1).contiguous()
    dev = torch.device("cuda", torch.cuda.current_device())
    d_pts = pts.pin_memory().to(dev, non_blocking=True)
    d_box = bxs.pin_memory().to(dev, non_blocking=True)
    d_trig = trig.pin_memory().to(dev, non_blocking=True)
    _abi.require_cuda(d_pts, d_box, d_trig)
    out = torch.empty((n_box, n_pts), dtype=torch.int32, device=dev)
    _abi.check(_abi.lib().seevcn_points_in_boxes_dense_trig(n_box, n_pts, _abi.ptr(d_box), _abi.ptr(d_trig),
                                                            _abi.ptr(d_pts), _abi.ptr(out), _abi.stream()))
    host = out.cpu()
    return host.numpy() if is_numpy else host


def crop_points_in_boxes(points, boxes):
    """Crop with per-box compaction (what SEE_VCN.isolate_gt_pts needs, SEE_VCN.py:61-82).

    points (B, M, 3), boxes (B, T, 7) CUDA ->
      box_idxs_of_pts (B, M) int32, box_counts (B, T) int32, box_offsets (B, T) int32,
      box_points (B, M) int32: ascending point indices grouped by box.
    """
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    points = points.contiguous(); boxes = boxes.contiguous()
    _abi.require_cuda(points, boxes)
    B, M, _ = points.shape
    T = boxes.shape[1]
    L = _abi.lib()
    dev = points.device
    idx = torch.empty((B, M), dtype=torch.int32, device=dev)
    counts = torch.empty((B, T), dtype=torch.int32, device=dev)
    offsets = torch.empty((B, T), dtype=torch.int32, device=dev)
    box_points = torch.empty((B, M), dtype=torch.int32, device=dev)
    ws_bytes = L.seevcn_crop_workspace_bytes(B, T, M)
    ws = _abi.workspace(dev, ws_bytes, "crop")
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_crop_points_in_boxes(B, T, M, _abi.ptr(boxes), _abi.ptr(points), _abi.ptr(idx),
                                                 _abi.ptr(counts), _abi.ptr(offsets), _abi.ptr(box_points),
                                                 _abi.ptr(ws), ws_bytes, _abi.stream()))
    return idx, counts, offsets, box_points


def select_objects(box_counts, min_pts):
    """(frame, box) pairs with at least ``min_pts`` cropped points, frame-major like ``np.argwhere(counts >= min_pts)``
    (SEE_VCN.py:71) -> obj_frame, obj_box (B*T,) int32 CUDA (first num entries valid), num (1,) int32 CUDA."""
    _abi.require_cuda(box_counts)
    B, T = box_counts.shape
    dev = box_counts.device
    obj_frame = torch.empty((B * T,), dtype=torch.int32, device=dev)
    obj_box = torch.empty((B * T,), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    with _abi.device_guard(dev):
        _abi.check(_abi.lib().seevcn_select_objects(B, T, _abi.ptr(box_counts), int(min_pts), _abi.ptr(obj_frame),
                                                    _abi.ptr(obj_box), _abi.ptr(num), _abi.stream()))
    return obj_frame, obj_box, num


def resample_gather(points, box_counts, box_offsets, box_points, obj_frame, obj_box, choice):
    """ResamplePoints on the device (data_transforms.py:247-262) with a host-supplied permutation.

    points (B, M, 3); obj_frame/obj_box (O,) int32; choice (O, n) int32 -> (O, n, 3) float32
    """
    _abi.require_cuda(points, box_counts, box_offsets, box_points, obj_frame, obj_box, choice)
    O, n = choice.shape
    B, M, _ = points.shape
    T = box_counts.shape[1]
    out = torch.empty((O, n, 3), dtype=torch.float32, device=points.device)
    with _abi.device_guard(points.device):
        _abi.check(_abi.lib().seevcn_resample_gather(O, n, T, M, _abi.ptr(points), _abi.ptr(box_counts),
                                                     _abi.ptr(box_offsets), _abi.ptr(box_points), _abi.ptr(obj_frame),
                                                     _abi.ptr(obj_box), _abi.ptr(choice), _abi.ptr(out), _abi.stream()))
    return out


def resample_gather_rng(points, box_counts, box_offsets, box_points, obj_frame, obj_box, n_points, seed):
    """ResamplePoints with the draw made on the device: object o receives the first ``n_points`` entries of a
    pseudo-random permutation of its tiled point list (Feistel network keyed by ``seed`` and frame*T+box).
    points (B, M, 3); obj_frame/obj_box (O,) int32 -> (O, n_points, 3) float32"""
    _abi.require_cuda(points, box_counts, box_offsets, box_points, obj_frame, obj_box)
    O = obj_frame.shape[0]
    B, M, _ = points.shape
    T = box_counts.shape[1]
    out = torch.empty((O, n_points, 3), dtype=torch.float32, device=points.device)
    with _abi.device_guard(points.device):
        _abi.check(_abi.lib().seevcn_resample_gather_rng(O, n_points, T, M, int(seed) & 0xffffffff, _abi.ptr(points),
                                                         _abi.ptr(box_counts), _abi.ptr(box_offsets), _abi.ptr(box_points),
                                                         _abi.ptr(obj_frame), _abi.ptr(obj_box), _abi.ptr(out), _abi.stream()))
    return out


def resample_perm(j, n, seed, frame_box):
    """Host evaluation of the same permutation (tests)."""
    return _abi.lib().seevcn_resample_perm(int(j), int(n), int(seed) & 0xffffffff, int(frame_box))
