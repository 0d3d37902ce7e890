"""GPU parity of the mask-based isolation front end (csrc/isolate.cu) against the oracle restatements of
map_pointcloud_to_image (custom_dataset_objects.py:141-193), get_pts_in_mask (shared_utils.py:36-106) and
SEE_VCN.isolate_det_pts (SEE_VCN.py:144-181).  PARITY UNPINNED upstream for the DBSCAN (open3d absent)."""
import numpy as np
import pytest
import torch

import oracle
from seevcn_b200 import synth, _abi
from seevcn_b200.see.surface_completion.datasets import shared_utils as su
from seevcn_b200.see.surface_completion.SEE_VCN import DetIsolator

pytestmark = pytest.mark.gpu

H, W = 720, 1280


def camera(model):
    """front camera on the roof: x forward -> camera z, y left -> -camera x, z up -> -camera y"""
    R = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    yaw = np.deg2rad(2.0)
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    ext = np.eye(4); ext[:3, :3] = R @ Rz; ext[:3, 3] = [0.05, -0.3, 0.1]
    K = np.array([[900.0, 0, 640.5], [0, 905.0, 360.25], [0, 0, 1]])
    dist = np.array([-0.12, 0.03, 0.001, -0.0007, 0.002]) if model == "pinhole" else np.array([0.02, -0.004, 0.001, -0.0002, 0.0])
    return {"intrinsic": K, "extrinsic": ext, "distcoeff": dist}


def instance_masks(pts, boxes, calib, model, rng):
    """binary masks: for every box in view, the bounding rectangle of its projected LiDAR returns, grown a little and
    with an ellipse cut (so masks overlap the background and each other), largest area first."""
    fov = oracle.map_pointcloud_to_image(pts, calib, (H, W), model)
    idx = oracle.points_in_boxes_gpu(pts[None], boxes[None])[0]
    frame_idx = np.nonzero(fov["fov_inds"])[0]
    masks = []
    for k in range(len(boxes)):
        sel = np.nonzero(idx[frame_idx] == k)[0]
        if len(sel) < 15:
            continue
        px = fov["pts_img"][sel]
        u0, v0 = px.min(0) - rng.integers(2, 12, 2); u1, v1 = px.max(0) + rng.integers(2, 12, 2)
        vv, uu = np.mgrid[0:H, 0:W]
        cu, cv, ru, rv = (u0 + u1) / 2, (v0 + v1) / 2, (u1 - u0) / 2 + 1, (v1 - v0) / 2 + 1
        m = (((uu - cu) / ru) ** 2 + ((vv - cv) / rv) ** 2 <= 1.15).astype(np.uint8)
        masks.append(m)
    masks.sort(key=lambda m: -int(m.sum()))
    return np.stack(masks), fov


@pytest.mark.parametrize("model", ["pinhole", "equidistant"])
def test_projection_and_mask_lookup_vs_oracle(model):
    cuda = torch.device("cuda", 0)
    pts, boxes = synth.make_frame(2100, n_boxes=40, box_r_max=45.0)
    calib = camera(model)
    want = oracle.map_pointcloud_to_image(pts, calib, (H, W), model)
    got = su.map_pointcloud_to_image(torch.from_numpy(pts).to(cuda), calib, (H, W), model)
    fov = got["fov_inds"].cpu().numpy().astype(bool)
    # identical up to points whose u or v sits within 1e-9 of a pixel rounding boundary / the image border
    np.testing.assert_array_equal(fov, want["fov_inds"])
    uv = got["pts_img"].cpu().numpy()
    safe = want["round_slack"] > 1e-9
    assert safe.mean() > 0.999999 and 5000 < fov.sum() < len(pts) // 3
    np.testing.assert_array_equal(uv[fov][safe], want["pts_img"][safe])
    assert (uv[~fov] == -1).all()
    np.testing.assert_allclose(got["depth"].cpu().numpy()[fov], want["depth"], rtol=1e-6)
    rng = np.random.default_rng(3)
    masks, _ = instance_masks(pts, boxes, calib, model, rng)
    assert len(masks) >= 4
    lists, counts = su.get_pts_in_mask(torch.from_numpy(masks).to(cuda), got)
    want_lists = oracle.get_pts_in_mask(masks, want)
    for i, wl in enumerate(want_lists):
        assert int(counts[i]) == len(wl)
        np.testing.assert_array_equal(lists[i, : len(wl)].cpu().numpy(), wl)


def test_dbscan_largest_vs_sequential_dbscan():
    """The parallel formulation (core degrees, core-core components, border -> earliest cluster) against the literal
    sequential expansion of open3d's ClusterDBSCAN, min_points = 3 and 1..5, fixed and range-adaptive eps, border points
    between two clusters, instances at and below the min_cluster limits."""
    cuda = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    clouds = []
    for t in range(12):
        n1, n2 = int(rng.integers(30, 400)), int(rng.integers(5, 200))
        c1 = rng.uniform(-30, 30, 3) * [1, 1, 0.05]
        a = c1 + rng.normal(0, [0.8, 0.4, 0.3], (n1, 3))
        b = c1 + [rng.uniform(1.5, 4.0), 0, 0] + rng.normal(0, 0.35, (n2, 3))
        bridge = c1 + np.linspace(0, 1, int(rng.integers(0, 6)))[:, None] * [3.0, 0, 0]          # sparse chain: border points
        noise = c1 + rng.uniform(-15, 15, (int(rng.integers(0, 20)), 3))
        pc = np.concatenate([a, b, bridge, noise])
        clouds.append(pc[rng.permutation(len(pc))].astype(np.float32))
    clouds.append(rng.normal(0, 0.05, (8, 3)).astype(np.float32) + 20)      # too few points
    clouds.append((rng.uniform(-50, 50, (40, 3))).astype(np.float32))       # all noise
    pts = np.concatenate(clouds)
    offs = np.cumsum([0] + [len(c) for c in clouds])
    I, stride = len(clouds), len(pts)
    lists = np.zeros((I, stride), np.int32)
    counts = np.array([len(c) for c in clouds], np.int32)
    for i in range(I):
        lists[i, : counts[i]] = np.arange(offs[i], offs[i + 1])
    d_pts, d_lists, d_counts = (torch.from_numpy(x).to(cuda) for x in (pts, lists, counts))
    inst = [np.arange(offs[i], offs[i + 1]) for i in range(I)]
    # the reference's rule: eps from the range, min_points 3
    cl, cc, ce = su.isolate_det_pts(d_pts, d_lists, d_counts, vres=0.4, eps_scaling=5.0, min_eps=0.2, max_eps=0.9, min_cluster=10)
    want, weps = oracle.isolate_det_pts(pts, inst, 0.4, 5.0, 0.2, 0.9, 10)
    np.testing.assert_allclose(ce.cpu().numpy(), weps, rtol=1e-12)
    kept = 0
    for i in range(I):
        if want[i] is None:
            assert int(cc[i]) == 0
        else:
            kept += 1
            np.testing.assert_array_equal(cl[i, : int(cc[i])].cpu().numpy(), want[i])
    assert 8 <= kept <= 12
    # fixed eps, other min_points: labels of the largest cluster against the sequential algorithm
    for mp, eps in ((1, 0.3), (2, 0.3), (3, 0.25), (4, 0.5), (5, 0.6)):
        with _abi.device_guard(cuda):
            out_l = torch.empty_like(d_lists); out_c = torch.zeros((I,), dtype=torch.int32, device=cuda)
            _abi.check(_abi.lib().seevcn_dbscan_largest(I, stride, _abi.ptr(d_pts), _abi.ptr(d_lists), _abi.ptr(d_counts), 0, eps, 0.0,
                                                        0.0, 0.0, 0.0, mp, 0, _abi.ptr(out_l), _abi.ptr(out_c), None, None, 0, _abi.stream()))
        for i in range(I):
            labels = oracle.cluster_dbscan(pts[inst[i]], eps, mp)
            y = np.bincount(labels[labels >= 0])
            if len(y) == 0:
                assert int(out_c[i]) == 0
                continue
            members = inst[i][labels == np.argmax(y)]
            np.testing.assert_array_equal(out_l[i, : int(out_c[i])].cpu().numpy(), members)


def test_det_isolator_end_to_end():
    """DET mode on a synthetic frame: projection -> masks -> DBSCAN -> resampled clouds ready for VCN.forward."""
    cuda = torch.device("cuda", 0)
    pts, boxes = synth.make_frame(2101, n_boxes=40, box_r_max=40.0)
    calib = camera("pinhole")
    masks, fov = instance_masks(pts, boxes, calib, "pinhole", np.random.default_rng(5))
    iso = DetIsolator(vres=0.4, eps_scaling=5.0, min_eps=0.2, max_eps=1.0, min_cluster=10, min_lidar_pts=30)
    inst = iso(torch.from_numpy(pts).to(cuda), torch.from_numpy(masks).to(cuda), calib, (H, W))
    want, _ = oracle.isolate_det_pts(pts, oracle.get_pts_in_mask(masks, fov), 0.4, 5.0, 0.2, 1.0, 10)
    want = [w for w in want if w is not None and len(w) > 30]
    assert len(inst) == len(want) >= 3
    assert max(len(w) for w in want) > 9600        # one mask swallowed the road: clustered in the global workspace
    L = _abi.lib()
    for k, (g, w) in enumerate(zip(inst, want)):
        np.testing.assert_array_equal(g.cpu().numpy(), pts[w])
        # resampled rows = the keyed permutation of the tiled cluster (ResamplePoints semantics)
        cnt = len(w)
        reps = -(-1024 // cnt)
        choice = np.array([L.seevcn_resample_perm(j, reps * cnt, 0, int(iso.kept[k])) % cnt for j in range(1024)])
        np.testing.assert_array_equal(iso.resampled[k].cpu().numpy(), pts[w][choice])
