"""The splice step of the SEE-VCN frame driver on the B200 path.

ref: see/surface_completion/SEE_VCN.py:247-265 (``SEE_VCN.replace_with_completed_pts``; demo twin
demo/see_vcn_dataset.py:127-135).  The reference measures, on the host, every raw frame point against an open3d
KD-tree of the completed points and drops those closer than ``point_dist_thresh``; here the completed clouds stay on
the device as per-object blocks and one kernel pair (csrc/splice.cu) produces the keep mask, optionally the merged
cloud.
"""
import numpy as np
import torch

from ... import _abi


def all_instances_frames(obj_pts, obj_frame, obj_count, num_frames, out_stride):
    """ref: SEE_VCN.py:113,244 — ``np.unique(np.vstack(clustered), axis=0)`` per frame, on the device.
    obj_pts (O,S,3), obj_frame (O,) int32 non-decreasing, obj_count (O,) int32 or None ->
    (uniq (F, out_stride, 3) f32, ucount (F,) int32): rows in lexicographic (x, y, z) order, duplicates removed."""
    _abi.require_cuda(obj_pts, obj_frame, obj_count)
    O, S, _ = obj_pts.shape
    dev = obj_pts.device
    L = _abi.lib()
    uniq = torch.empty((num_frames, out_stride, 3), dtype=torch.float32, device=dev)
    ucount = torch.empty((num_frames,), dtype=torch.int32, device=dev)
    ws = _abi.workspace(dev, L.seevcn_unique_rows_frames_workspace_bytes(num_frames, O, S, out_stride), "uniq")
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_unique_rows_frames(num_frames, O, S, _abi.ptr(obj_pts), _abi.ptr(obj_count), _abi.ptr(obj_frame), out_stride,
                                               _abi.ptr(uniq), _abi.ptr(ucount), _abi.ptr(ws), ws.numel(), _abi.stream()))
    return uniq, ucount


def splice_frames(frame_pts, obj_pts, obj_frame, obj_count=None, point_dist_thresh=0.1, merged=False, unique=False):
    """frame_pts (F,P,3) f32 CUDA; obj_pts (O,S,3) f32 CUDA completed clouds, obj_frame (O,) int32 non-decreasing,
    obj_count (O,) int32 or None = the distinct rows of every (cyclically tiled) object cloud.

    -> keep (F,P) uint8 CUDA (1 = the raw point survives); with ``merged=True`` also
       (merged (F,P+max rows,3) f32, merged_count (F,) int32, completed_count (F,) int32): per frame
       [completed rows ++ surviving raw points], the reference's return value, rows >= merged_count[f] undefined.
       ``unique=True``: the completed rows are ``np.unique`` of the frame's object rows (lexicographic order, cross-object
       duplicates removed) exactly as SEE_VCN.py:244 builds ``all_instances``; default: distinct rows per object, object order."""
    frame_pts = frame_pts.contiguous()
    _abi.require_cuda(frame_pts)
    assert frame_pts.dtype == torch.float32 and frame_pts.dim() == 3 and frame_pts.shape[2] == 3
    F, P, _ = frame_pts.shape
    dev = frame_pts.device
    if obj_pts is not None and obj_pts.shape[0] > 0:
        obj_pts = obj_pts.contiguous(); obj_frame = obj_frame.contiguous()
        _abi.require_cuda(obj_pts, obj_frame)
        assert obj_pts.dtype == torch.float32 and obj_frame.dtype == torch.int32 and obj_frame.shape[0] == obj_pts.shape[0]
        O, S, _ = obj_pts.shape
        if obj_count is not None:
            obj_count = obj_count.contiguous()
            _abi.require_cuda(obj_count)
            assert obj_count.dtype == torch.int32 and obj_count.shape[0] == O
    else:
        obj_pts = obj_frame = obj_count = None
        O = S = 0
    L = _abi.lib()
    keep = torch.empty((F, P), dtype=torch.uint8, device=dev)
    ws = _abi.workspace(dev, L.seevcn_splice_workspace_bytes(F, P, O), "splice")
    out = m_cnt = c_cnt = None
    stride = 0
    if merged:
        # capacity: every raw point + the rows of the frame with the most objects (host-known upper bound O*S)
        stride = P + O * S
        out = torch.empty((F, stride, 3), dtype=torch.float32, device=dev)
        m_cnt = torch.empty((F,), dtype=torch.int32, device=dev)
        c_cnt = torch.empty((F,), dtype=torch.int32, device=dev)
    rows = rows_cnt = None
    rows_stride = 0
    if merged and unique and O > 0:
        rows_stride = O * S
        rows, rows_cnt = all_instances_frames(obj_pts, obj_frame, obj_count, F, rows_stride)
    with _abi.device_guard(dev):
        _abi.check(L.seevcn_splice(F, P, _abi.ptr(frame_pts), O, S, _abi.ptr(obj_pts), _abi.ptr(obj_count), _abi.ptr(obj_frame),
                                   float(point_dist_thresh), _abi.ptr(keep), stride, _abi.ptr(out), _abi.ptr(m_cnt),
                                   _abi.ptr(c_cnt), _abi.ptr(rows), rows_stride, _abi.ptr(rows_cnt),
                                   _abi.ptr(ws), ws.numel(), _abi.stream()))
    return (keep, out, m_cnt, c_cnt) if merged else keep


def replace_with_completed_pts(original_points, sc_instances, point_dist_thresh=0.1, device=None):
    """Reference-shaped entry (SEE_VCN.py:247-265): host arrays in, host array out.

    original_points (P,3) numpy (the reference takes an open3d PointCloud), sc_instances (N,3) numpy or None
    -> np.vstack((sc_instances, original points farther than point_dist_thresh from every completed point))."""
    pts = np.ascontiguousarray(np.asarray(original_points)[:, :3], dtype=np.float32)
    if sc_instances is None:
        return pts
    sc = np.ascontiguousarray(sc_instances, dtype=np.float32)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    d_pts = torch.from_numpy(pts).to(dev).view(1, -1, 3)
    d_sc = torch.from_numpy(sc).to(dev).view(1, -1, 3)
    frame = torch.zeros((1,), dtype=torch.int32, device=dev)
    _, merged, m_cnt, _ = splice_frames(d_pts, d_sc, frame, None, point_dist_thresh, merged=True)   # sc_instances is already unique
    return merged[0, : int(m_cnt[0])].cpu().numpy()


class DetIsolator:
    """The DET-mode front end of ``SEE_VCN`` (no ground-truth boxes): ``get_det_instances`` + ``isolate_det_pts``
    (SEE_VCN.py:117-181) for one frame and one camera, on the device.

        iso = DetIsolator(vres=0.4, eps_scaling=5, min_eps=0.2, max_eps=1.0)          # cfg.PC_ISOLATION.*
        inst = iso(points_cuda, masks_cuda, calib, img_shape)                          # list of (n_i, 3) CUDA clouds
        clouds = iso.resampled                                                        # (O, 1024, 3) for VCN.forward
    """

    def __init__(self, vres, eps_scaling, min_eps, max_eps, min_cluster=10, min_lidar_pts=30, camera_model="pinhole",
                 resample_num=1024, seed=0):
        self.vres, self.eps_scaling, self.min_eps, self.max_eps = vres, eps_scaling, min_eps, max_eps
        self.min_cluster, self.min_lidar_pts, self.camera_model = min_cluster, min_lidar_pts, camera_model
        self.resample_num, self.seed = resample_num, seed
        self.resampled = None

    @torch.no_grad()
    def __call__(self, points, masks, calib, img_shape):
        from .datasets import shared_utils as su
        imgfov = su.map_pointcloud_to_image(points, calib, img_shape, self.camera_model)
        lists, counts = su.get_pts_in_mask(masks, imgfov)
        clists, ccounts, _ = su.isolate_det_pts(points, lists, counts, self.vres, self.eps_scaling, self.min_eps, self.max_eps,
                                                self.min_cluster)
        h = ccounts.cpu().numpy()                                   # the one host sync: how many instances survive
        keep = np.nonzero(h > self.min_lidar_pts)[0].astype(np.int32)   # SEE_VCN.py:222 (single camera: no merging)
        self.kept = keep
        obj_inst = torch.from_numpy(keep).to(points.device)
        self.resampled = su.resample_instances(points, clists, ccounts, obj_inst, self.resample_num, self.seed)
        return [points[clists[i, : int(h[i])].long()] for i in keep]
