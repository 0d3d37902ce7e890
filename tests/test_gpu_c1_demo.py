"""BASELINE.json configs[0] (C1): the reference's demo frame (demo/demo_data/pcd/000001.pcd, 26,715 points, committed
as a data fixture) with 10 car boxes drawn on its own points, through the whole path at batch 1:
crop -> VCN completion -> kNN surface -> largest cluster -> splice -> hard voxel generator (KITTI grid) -> MeanVFE,
every stage against the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle
from seevcn_b200.pcdet.datasets.processor.data_processor import VoxelGeneratorWrapper
from seevcn_b200.pcdet.models.backbones_3d.vfe import MeanVFE
from seevcn_b200.see.surface_completion.SEE_VCN import splice_frames
from seevcn_b200.see.surface_completion.pcd_io import read_pcd

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KITTI = ([0.0, -40.0, -3.0, 70.4, 40.0, 1.0], [0.05, 0.05, 0.1], [1408, 1600, 40])   # sc_kitti_dataset.yaml:4, second_iou grid


def demo_boxes(pts, n_boxes=10, seed=3):
    """centres on the frame's own points at range 5-50 m, car sizes N([4.2, 2.0, 1.6], 10 %), heading U(-pi, pi),
    centres >= 6 m apart, each holding >= 60 points"""
    rng = np.random.default_rng(seed)
    rng_xy = np.hypot(pts[:, 0], pts[:, 1])
    cand = np.nonzero((rng_xy > 5) & (rng_xy < 50))[0]
    boxes = []
    for i in rng.permutation(cand):
        c = pts[i].astype(np.float64)
        if any(np.hypot(c[0] - b[0], c[1] - b[1]) < 6.0 for b in boxes):
            continue
        size = np.array([4.2, 2.0, 1.6]) * (1.0 + 0.1 * np.clip(rng.standard_normal(3), -2, 2))
        b = np.concatenate([c, size, [rng.uniform(-np.pi, np.pi)]]).astype(np.float32)
        if (oracle.points_in_boxes_gpu(pts[None], b[None, None])[0] == 0).sum() >= 60:
            boxes.append(b)
        if len(boxes) == n_boxes:
            break
    return np.stack(boxes)


def test_c1_demo_frame_end_to_end(cuda):
    from seevcn_b200.pipeline import CompletionPipeline
    pts = read_pcd(os.path.join(HERE, "golden", "demo_000001.pcd"))
    assert pts.shape == (26715, 3)
    boxes = demo_boxes(pts)
    assert boxes.shape == (10, 7)
    sd = oracle.make_state_dict("VCN_VC", seed=0)
    pipe = CompletionPipeline("VCN_VC", sd, cuda, sel_k=10, cluster_eps=0.3, splice_thresh=0.1, voxel_cfg=KITTI)
    d_pts, d_boxes = torch.from_numpy(pts[None]).to(cuda), torch.from_numpy(boxes[None]).to(cuda)
    out = pipe.run(d_pts, d_boxes, seed=0)

    # crop
    ref_idx, slack = oracle.points_in_boxes_gpu(pts[None], boxes[None], return_slack=True)
    got_idx = out["box_idxs_of_pts"].cpu().numpy()
    np.testing.assert_array_equal(got_idx[slack > 1e-4], ref_idx[slack > 1e-4])
    assert len(out["obj_box"]) == 10                                   # every box holds >= MIN_LIDAR_PTS points
    # completion
    inp = out["input"].cpu().numpy()
    want = oracle.vcn_forward_ref(sd, inp, None, "VCN_VC")["coarse"].numpy()
    cd = oracle.chamfer_l2(out["coarse"].cpu().numpy(), want)
    scale = ((want - want.mean(axis=1, keepdims=True)) ** 2).sum(-1).mean(-1)
    assert (cd / scale < 1e-3).all()
    surf, _ = oracle.get_partial_mesh_batch(inp, out["coarse"].cpu().numpy(), k=10)
    np.testing.assert_array_equal(out["surface"].cpu().numpy(), surf)
    clus, ccnt = oracle.get_largest_cluster_batch(surf, eps=0.3, min_points=2, total_pts=1024)
    np.testing.assert_array_equal(out["clustered"].cpu().numpy(), clus)
    np.testing.assert_array_equal(out["clustered_count"].cpu().numpy(), ccnt)
    dcnt = oracle.distinct_rows(clus, ccnt)                             # what np.unique keeps (SEE_VCN.py:113,244)
    np.testing.assert_array_equal(out["completed_count"].cpu().numpy(), dcnt)
    assert (dcnt < ccnt).any()                                          # tiled clouds: fewer distinct rows than rows
    # splice: the merged frame the reference would save as .pcd (SEE_VCN.py:247-280)
    keep, merged, m_cnt, c_cnt = splice_frames(d_pts, out["clustered"], out["obj_frame_dev"], out["completed_count"], 0.1, merged=True)
    sc = np.concatenate([clus[o][: dcnt[o]] for o in range(len(clus))])
    assert len(sc) == len(oracle.all_instances(clus, ccnt))            # same rows as the reference's np.unique, object order
    want_merged, want_keep = oracle.replace_with_completed_pts(pts, sc, 0.1)
    np.testing.assert_array_equal(keep[0].cpu().numpy().astype(bool), want_keep)
    frame = merged[0, : int(m_cnt[0])].cpu().numpy()
    np.testing.assert_array_equal(frame, want_merged)
    assert int(c_cnt[0]) == dcnt.sum() and 0 < (~want_keep).sum() < len(pts)
    # detector front end of the KITTI configs: hard voxel generator + MeanVFE (data_processor.py:15-60, mean_vfe.py:14-31)
    gen = VoxelGeneratorWrapper(KITTI[1], KITTI[0], 3, 5, 16000)
    v, c, n = gen.generate(frame)
    wv, wc, wn = oracle.hard_voxelize(frame, KITTI[0], KITTI[1], KITTI[2], 5, 16000)
    np.testing.assert_array_equal(c, wc)
    np.testing.assert_array_equal(n, wn)
    np.testing.assert_array_equal(v, wv)
    vfe = MeanVFE(model_cfg={}, num_point_features=3)
    feats = vfe({"voxels": torch.from_numpy(v).to(cuda), "voxel_num_points": torch.from_numpy(n).to(cuda)})["voxel_features"]
    np.testing.assert_allclose(feats.cpu().numpy(), oracle.mean_vfe(wv, wn.astype(np.float32)), rtol=1e-6, atol=1e-6)
    assert len(c) > 3000
    # and the dynamic voxelization the pipeline itself ran, on the same KITTI grid
    vc, vf, vn = oracle.dynamic_voxelize(pipe.voxel_points(out).cpu().numpy(), *KITTI)
    np.testing.assert_array_equal(out["voxel_coords"].cpu().numpy(), vc)
    np.testing.assert_array_equal(out["voxel_num_points"].cpu().numpy(), vn)
    np.testing.assert_allclose(out["voxel_features"].cpu().numpy(), vf, rtol=1e-5, atol=1e-5)
