"""seevcn_b200 — B200-native (sm_100a) implementation of SEE-VCN's per-frame
object-completion + voxelization hot path, behind the reference's operator surface.

Layout
  csrc/                         CUDA kernels + the C-ABI (include/seevcn_b200.h)
  _abi.py                       ctypes binding of csrc/libseevcn_b200.so
  pcdet/ops/...                 points_in_boxes_gpu/cpu, furthest_point_sample, gather/grouping, knn
  pcdet/models/backbones_3d/vfe MeanVFE, DynamicMeanVFE
  pcdet/datasets/processor      VoxelGeneratorWrapper
  see/surface_completion/...    VCN_VC / VCN_CN forward, VCN.inference, kNN surface / largest cluster, the splice step
                                (SEE_VCN.replace_with_completed_pts), the .pcd wire format (pcd_io)
  pipeline.py                   frame-level crop -> complete -> select -> cluster -> splice -> voxelize driver,
                                HostStream (pinned host buffers in / out)
  dist.py                       frame sharding over ranks + all-gather of results (static-capacity async form)

There is no CPU fallback: every op raises if the CUDA library is missing or the device is
not sm_100.
"""
__version__ = "0.1.0"
