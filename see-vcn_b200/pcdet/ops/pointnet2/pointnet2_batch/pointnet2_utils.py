"""Drop-in for ``pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils`` (forward ops on the
SEE-VCN path) plus the ``knn`` op north_star names.

ref: detector3d/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py:10-197
"""
import torch

from ..... import _abi


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """xyz (B, N, 3) float32 contiguous CUDA -> (B, npoint) int32.  ref: pointnet2_utils.py:12-29"""
    assert xyz.is_contiguous()
    _abi.require_cuda(xyz)
    assert xyz.dtype == torch.float32 and xyz.dim() == 3 and xyz.shape[2] == 3
    B, N, _ = xyz.size()
    output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    temp = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
    with _abi.device_guard(xyz.device):
        _abi.check(_abi.lib().seevcn_furthest_point_sampling(B, N, npoint, _abi.ptr(xyz), _abi.ptr(temp),
                                                             _abi.ptr(output), _abi.stream()))
    return output


farthest_point_sample = furthest_point_sample


def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """features (B, C, N), idx (B, npoint) int32 -> (B, C, npoint).  ref: pointnet2_utils.py:42-60"""
    assert features.is_contiguous()
    assert idx.is_contiguous()
    _abi.require_cuda(features, idx)
    assert features.dtype == torch.float32 and idx.dtype == torch.int32
    B, npoint = idx.size()
    _, C, N = features.size()
    output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
    with _abi.device_guard(features.device):
        _abi.check(_abi.lib().seevcn_gather_points(B, C, N, npoint, _abi.ptr(features), _abi.ptr(idx),
                                                   _abi.ptr(output), _abi.stream()))
    return output


def grouping_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """features (B, C, N), idx (B, npoint, nsample) int32 -> (B, C, npoint, nsample).
    ref: pointnet2_utils.py:159-177"""
    assert features.is_contiguous()
    assert idx.is_contiguous()
    _abi.require_cuda(features, idx)
    assert features.dtype == torch.float32 and idx.dtype == torch.int32
    B, nfeatures, nsample = idx.size()
    _, C, N = features.size()
    output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
    with _abi.device_guard(features.device):
        _abi.check(_abi.lib().seevcn_group_points(B, C, N, nfeatures, nsample, _abi.ptr(features), _abi.ptr(idx),
                                                  _abi.ptr(output), _abi.stream()))
    return output


def knn(k: int, ref: torch.Tensor, query: torch.Tensor):
    """k nearest reference points for every query, ascending distance.

    ref (B, R, 3), query (B, Q, 3) -> dist (B, Q, k) float32 (Euclidean), idx (B, Q, k) int32
    Semantics of cKDTree.query / topk(largest=False): sampling.py:30-34,59-61
    """
    assert ref.is_contiguous() and query.is_contiguous()
    _abi.require_cuda(ref, query)
    B, R, _ = ref.shape
    Q = query.shape[1]
    dist = torch.empty((B, Q, k), dtype=torch.float32, device=ref.device)
    idx = torch.empty((B, Q, k), dtype=torch.int32, device=ref.device)
    with _abi.device_guard(ref.device):
        _abi.check(_abi.lib().seevcn_knn(B, R, Q, k, _abi.ptr(ref), _abi.ptr(query), _abi.ptr(dist), _abi.ptr(idx),
                                         _abi.stream()))
    return dist, idx
