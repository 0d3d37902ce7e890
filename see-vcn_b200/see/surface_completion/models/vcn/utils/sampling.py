"""kNN surface selection on the device.

ref: see/surface_completion/models/vcn/utils/sampling.py:8-41 (partial_with_KDTree), :69-80 (batch)
"""
import torch

from ...... import _abi


def get_partial_mesh_batch(batch_partial, batch_complete, k=20, surface_pts=1024, return_count=False):
    """batch_partial (B, Np, 3), batch_complete (B, R, 3) CUDA float32 -> (B, surface_pts, 3) CUDA.

    For every object: the union of the k nearest completed points of every partial point,
    in ascending index order, tiled cyclically to ``surface_pts`` rows (the reference returns
    the same array as numpy after a host round trip)."""
    batch_partial = batch_partial.contiguous().float()
    batch_complete = batch_complete.contiguous().float()
    _abi.require_cuda(batch_partial, batch_complete)
    B, Np, _ = batch_partial.shape
    R = batch_complete.shape[1]
    out = torch.empty((B, surface_pts, 3), dtype=torch.float32, device=batch_partial.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=batch_partial.device)
    L = _abi.lib()
    ws = _abi.workspace(batch_partial.device, L.seevcn_knn_surface_select_workspace_bytes(B, Np, R), "knn_select")
    with _abi.device_guard(batch_partial.device):
        _abi.check(L.seevcn_knn_surface_select(B, Np, R, k, surface_pts, _abi.ptr(batch_partial),
                                               _abi.ptr(batch_complete), _abi.ptr(out), _abi.ptr(cnt),
                                               _abi.ptr(ws), ws.numel(), _abi.stream()))
    return (out, cnt) if return_count else out


def get_largest_cluster_batch(pc, eps=0.4, min_points=1, total_pts=1024, return_count=False, period=None,
                              return_distinct=False):
    """pc (B, N, 3) CUDA float32 -> (B, total_pts, 3): largest DBSCAN cluster of every object, tiled.

    ref: sampling.py:83-109 (open3d cluster_dbscan per object on the host).  min_points must be 1 or 2
    (the reference always passes 2, models/VCN.py:96): DBSCAN is then connected components of the
    eps-graph, computed per object in one CTA.  ``period`` (B,) int32 CUDA: optional promise that
    pc[b, r] == pc[b, r % period[b]] (the tiling ``get_partial_mesh_batch`` produces, period = its count) — same
    result, only the distinct rows are clustered.  ``return_count``: also the member-row count (np.bincount's size on
    the tiled cloud); ``return_distinct``: also the number of DISTINCT member rows — rows [0, distinct) of the output are
    what ``np.unique(clustered)`` keeps (SEE_VCN.py:113,244) and what the splice / voxelizer take as the object's rows."""
    pc = pc.contiguous().float()
    _abi.require_cuda(pc)
    B, N, _ = pc.shape
    out = torch.empty((B, total_pts, 3), dtype=torch.float32, device=pc.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=pc.device)
    distinct = torch.empty((B,), dtype=torch.int32, device=pc.device)
    with _abi.device_guard(pc.device):
        if period is None:
            _abi.check(_abi.lib().seevcn_largest_cluster(B, N, total_pts, float(eps), int(min_points), _abi.ptr(pc),
                                                         _abi.ptr(out), _abi.ptr(cnt), _abi.ptr(distinct), _abi.stream()))
        else:
            _abi.require_cuda(period)
            assert period.dtype == torch.int32 and period.shape == (B,)
            _abi.check(_abi.lib().seevcn_largest_cluster_periodic(B, N, total_pts, float(eps), int(min_points), _abi.ptr(pc),
                                                                  _abi.ptr(period), _abi.ptr(out), _abi.ptr(cnt),
                                                                  _abi.ptr(distinct), _abi.stream()))
    ret = (out,) + ((cnt,) if return_count else ()) + ((distinct,) if return_distinct else ())
    return ret if len(ret) > 1 else out
