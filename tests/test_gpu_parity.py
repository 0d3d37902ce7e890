"""GPU parity tests: the CUDA path (through the python wrappers -> ctypes -> C-ABI) against the
oracle on the same seeded inputs, against the committed goldens, against the reference's own
kernels compiled for sm_100a (oracle/_ref, when present), and size-independent properties at
BASELINE.json's full sizes.  Run with `pytest -m gpu`."""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle
from seevcn_b200 import synth
from seevcn_b200.pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils as roi
from seevcn_b200.pcdet.ops.pointnet2.pointnet2_batch import pointnet2_utils as pn2
from seevcn_b200.pcdet.models.backbones_3d.vfe import MeanVFE, DynamicMeanVFE
from seevcn_b200.pcdet.models.backbones_3d.vfe.dynamic_mean_vfe import dynamic_voxelize
from seevcn_b200.pcdet.datasets.processor.data_processor import VoxelGeneratorWrapper
from seevcn_b200.see.surface_completion.models.vcn.utils.sampling import get_partial_mesh_batch, get_largest_cluster_batch
from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS

pytestmark = pytest.mark.gpu


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------ stage 1: crop --
@pytest.mark.parametrize("shape", [(1, 20000, 10), (2, 4097, 3), (3, 1000, 300), (1, 5, 1)])
def test_points_in_boxes_vs_oracle(cuda, shape):
    B, P, T = shape
    pts = np.stack([synth.make_frame(1000 + b, n_beams=16, n_az=P // 16 + 1, n_boxes=min(T, 40))[0][:P] for b in range(B)])
    boxes = np.stack([synth.make_boxes(np.random.default_rng(50 + b), T, r_max=60 + T) for b in range(B)])
    if T > 40:   # many boxes: draw them on top of the points so they are hit
        sel = np.linspace(0, P - 1, T).astype(np.int64)
        boxes[:, :, :3] = pts[:, sel] + 0.2
    out = roi.points_in_boxes_gpu(dev(pts, cuda), dev(boxes, cuda)).cpu().numpy()
    ref, slack = oracle.points_in_boxes_gpu(pts, boxes, return_slack=True)
    decided = slack > 1e-4      # within 1e-4 m of an x/y face libm and CUDA cosf/sinf may round differently
    np.testing.assert_array_equal(out[decided], ref[decided])
    assert decided.mean() > 0.99
    assert (out >= 0).sum() > 0 or T == 1


def test_points_in_boxes_bit_exact_vs_reference_kernel(cuda):
    """Full C2 size against the reference's own kernel compiled for sm_100a: every point, no tolerance."""
    R = oracle.ref_kernels()
    if R is None:
        pytest.skip("oracle/_ref/libref_kernels.so not built (needs /root/reference at build time)")
    pts, boxes = synth.make_stream(2)
    boxes = np.concatenate([boxes, np.zeros((2, 6, 7), np.float32)], axis=1)   # zero-padded rows (dataset.py:193-198)
    d_pts, d_box = dev(pts, cuda), dev(boxes, cuda)
    out = roi.points_in_boxes_gpu(d_pts, d_box)
    ref = torch.full_like(out, -1)
    assert R.ref_points_in_boxes(2, boxes.shape[1], pts.shape[1], _ptr(d_box), _ptr(d_pts), _ptr(ref)) == 0
    assert torch.equal(out, ref)
    assert (out >= 0).sum().item() > 20000


def test_points_in_boxes_cpu_surface_matches_reference_build(cuda, golden):
    """host buffers in/out; bit-exact against the reference's x86 build (golden)"""
    out = roi.points_in_boxes_cpu(golden["pib_points"], golden["pib_boxes"])
    assert isinstance(out, np.ndarray) and out.dtype == np.int32
    np.testing.assert_array_equal(out, oracle.points_in_boxes_cpu(golden["pib_points"], golden["pib_boxes"]))
    if "pib_cpu_out" in golden:
        ref = np.unpackbits(golden["pib_cpu_out"], axis=1)[:, : out.shape[1]].astype(np.int32)
        np.testing.assert_array_equal(out, ref)


def test_crop_compaction_lists(cuda):
    pts, boxes = synth.make_stream(3, n_beams=32, n_az=1090, n_boxes=20)
    idx, counts, offsets, lists = roi.crop_points_in_boxes(dev(pts, cuda), dev(boxes, cuda))
    idx, counts, offsets, lists = (t.cpu().numpy() for t in (idx, counts, offsets, lists))
    for b in range(3):
        want = oracle.crop_lists(idx[b], 20)     # ascending point indices per box, like points[idx == k]
        assert counts[b].tolist() == [len(w) for w in want]
        assert offsets[b].tolist() == np.concatenate([[0], np.cumsum(counts[b])[:-1]]).tolist()
        for k in range(20):
            np.testing.assert_array_equal(lists[b, offsets[b, k]: offsets[b, k] + counts[b, k]], want[k])
    assert counts.sum() > 1000


def test_crop_empty_and_ragged(cuda):
    pts = torch.zeros((1, 0, 3), device=cuda)
    boxes = torch.zeros((1, 4, 7), device=cuda)
    assert roi.points_in_boxes_gpu(pts, boxes).shape == (1, 0)
    pts = torch.rand((2, 77, 3), device=cuda)
    assert (roi.points_in_boxes_gpu(pts, torch.zeros((2, 0, 7), device=cuda)) == -1).all()


def test_resample_gather(cuda):
    pts, boxes = synth.make_stream(2, n_beams=32, n_az=1090, n_boxes=10)
    d_pts = dev(pts, cuda)
    idx, counts, offsets, lists = roi.crop_points_in_boxes(d_pts, dev(boxes, cuda))
    cnt = counts.cpu().numpy()
    objs = [(f, k) for f in range(2) for k in range(10) if cnt[f, k] >= 30]
    rng = np.random.default_rng(0)
    idx_h = idx.cpu().numpy()
    want, choice = [], []
    for f, k in objs:
        r, c = oracle.resample_points(pts[f][idx_h[f] == k], 1024, rng)
        want.append(r); choice.append(c)
    got = roi.resample_gather(d_pts, counts, offsets, lists, dev(np.array([o[0] for o in objs], np.int32), cuda),
                              dev(np.array([o[1] for o in objs], np.int32), cuda), dev(np.stack(choice), cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), np.stack(want))


def test_resample_gather_device_rng(cuda):
    """device-side draw: every object's cloud is the first 1024 entries of a permutation of its tiled list"""
    pts, boxes = synth.make_stream(2, n_beams=32, n_az=1090, n_boxes=10)
    d_pts = dev(pts, cuda)
    idx, counts, offsets, lists = roi.crop_points_in_boxes(d_pts, dev(boxes, cuda))
    cnt, off, lst = counts.cpu().numpy(), offsets.cpu().numpy(), lists.cpu().numpy()
    objs = [(f, k) for f in range(2) for k in range(10) if cnt[f, k] >= 5]
    of = dev(np.array([o[0] for o in objs], np.int32), cuda); ob = dev(np.array([o[1] for o in objs], np.int32), cuda)
    got = roi.resample_gather_rng(d_pts, counts, offsets, lists, of, ob, 1024, seed=11).cpu().numpy()
    again = roi.resample_gather_rng(d_pts, counts, offsets, lists, of, ob, 1024, seed=11).cpu().numpy()
    other = roi.resample_gather_rng(d_pts, counts, offsets, lists, of, ob, 1024, seed=12).cpu().numpy()
    np.testing.assert_array_equal(got, again)
    assert not np.array_equal(got, other)
    for o, (f, k) in enumerate(objs):
        c = int(cnt[f, k]); reps = -(-1024 // c)
        choice = np.array([roi.resample_perm(j, reps * c, 11, f * 10 + k) for j in range(1024)])
        assert len(np.unique(choice)) == 1024 and choice.max() < reps * c          # a permutation prefix
        src = lst[f, off[f, k] + choice % c]
        np.testing.assert_array_equal(got[o], pts[f][src])
        if c <= 1024:   # ResamplePoints property: every original point survives when the cloud is tiled up
            assert len(np.unique(choice % c)) >= min(c, 1024 - c + 1) or reps > 1


# ------------------------------------------------------------------------- stage 3: FPS --
@pytest.mark.parametrize("shape", [(3, 1024, 256), (2, 1000, 64), (2, 16384, 1024), (2, 40, 40), (1, 7, 3), (1, 20000, 128)])
def test_fps_vs_oracle_bit_exact(cuda, shape):
    B, N, M = shape
    _, dense, _ = synth.make_object_clouds(21, B, 64, N)
    got = pn2.furthest_point_sample(dev(dense, cuda), M).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.furthest_point_sample(dense, M))


def test_fps_duplicates_follow_reference_tie_rule(cuda):
    part, _, _ = synth.make_object_clouds(3, 2, 200, 0)
    tiled = np.tile(part, (1, 6, 1))[:, :1024]          # resampled clouds contain exact duplicates
    got = pn2.furthest_point_sample(dev(tiled, cuda), 300).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.furthest_point_sample(tiled, 300))


def test_fps_gather_group_vs_reference_kernels(cuda):
    R = oracle.ref_kernels()
    if R is None:
        pytest.skip("oracle/_ref/libref_kernels.so not built")
    _, dense, _ = synth.make_object_clouds(33, 4, 64, 16384)
    xyz = dev(dense, cuda)
    got = pn2.furthest_point_sample(xyz, 1024)
    ref = torch.empty_like(got)
    temp = torch.full((4, 16384), 1e10, device=cuda)
    assert R.ref_fps(4, 16384, 1024, _ptr(xyz), _ptr(temp), _ptr(ref)) == 0
    assert torch.equal(got, ref)
    feats = xyz.transpose(1, 2).contiguous()
    g = pn2.gather_operation(feats, got)
    gref = torch.empty_like(g)
    assert R.ref_gather(4, 3, 16384, 1024, _ptr(feats), _ptr(got), _ptr(gref)) == 0
    assert torch.equal(g, gref)
    _, nn_idx = pn2.knn(16, xyz[:, :2048].contiguous(), g.transpose(1, 2).contiguous())
    feats2 = feats[:, :, :2048].contiguous()
    grp = pn2.grouping_operation(feats2, nn_idx)
    gref = torch.empty_like(grp)
    assert R.ref_group(4, 3, 2048, 1024, 16, _ptr(feats2), _ptr(nn_idx), _ptr(gref)) == 0
    assert torch.equal(grp, gref)


def test_fps_ties_bit_exact_vs_reference_kernel(cuda):
    """clouds full of exact duplicates (what ResamplePoints produces): the tie rule of the reference's
    shared-memory tree must be reproduced, not just its arithmetic"""
    R = oracle.ref_kernels()
    if R is None:
        pytest.skip("oracle/_ref/libref_kernels.so not built")
    for n, m, src in ((1024, 400, 150), (1000, 300, 77), (4096, 512, 300), (48, 48, 5), (16384, 256, 3000), (2048, 300, 500),
                      (64, 64, 9)):
        part, _, _ = synth.make_object_clouds(90 + n, 3, src, 0)
        tiled = np.tile(part, (1, n // src + 1, 1))[:, :n].copy()
        xyz = dev(tiled, cuda)
        got = pn2.furthest_point_sample(xyz, m)
        ref = torch.empty_like(got)
        temp = torch.full((3, n), 1e10, device=cuda)
        assert R.ref_fps(3, n, m, _ptr(xyz), _ptr(temp), _ptr(ref)) == 0
        assert torch.equal(got, ref), (n, m)
        np.testing.assert_array_equal(got.cpu().numpy(), oracle.furthest_point_sample(tiled, m))


def test_gather_group_vs_oracle(cuda):
    rng = np.random.default_rng(5)
    feats = rng.standard_normal((3, 7, 500)).astype(np.float32)
    idx = rng.integers(0, 500, (3, 64)).astype(np.int32)
    np.testing.assert_array_equal(pn2.gather_operation(dev(feats, cuda), dev(idx, cuda)).cpu().numpy(),
                                  oracle.gather_operation(feats, idx))
    idx3 = rng.integers(0, 500, (3, 33, 9)).astype(np.int32)
    np.testing.assert_array_equal(pn2.grouping_operation(dev(feats, cuda), dev(idx3, cuda)).cpu().numpy(),
                                  oracle.grouping_operation(feats, idx3))


# ------------------------------------------------------------------------- stage 4: kNN --
@pytest.mark.parametrize("cfg", [(2, 1024, 1024, 10), (2, 1024, 1024, 30), (1, 5000, 300, 16), (1, 64, 10, 64), (2, 16384, 256, 20)])
def test_knn_vs_oracle(cuda, cfg):
    B, R, Q, k = cfg
    part, dense, _ = synth.make_object_clouds(41, B, Q, R)
    gap = oracle.knn_gap(dense, part, k)
    dist, idx = pn2.knn(k, dev(dense, cuda), dev(part, cuda))
    rd, ri = oracle.knn(k, dense, part)
    ok = gap > 1e-5                     # tie-free queries: k / k+1 gap above fp32 resolution
    assert ok.mean() > 0.95
    idx = idx.cpu().numpy()
    # inside the k-set neighbours closer than fp32 resolution may swap places: compare as sets per query ...
    np.testing.assert_array_equal(np.sort(idx[ok], axis=-1), np.sort(ri[ok], axis=-1))
    # ... and exactly where consecutive distances are separated
    sep = np.all(np.diff(rd, axis=-1) > 1e-5 * np.maximum(rd[..., 1:], 1e-6), axis=-1) & ok
    np.testing.assert_array_equal(idx[sep], ri[sep])
    np.testing.assert_allclose(dist.cpu().numpy(), rd, rtol=1e-5, atol=1e-6)


def test_surface_select_vs_oracle_and_golden(cuda, golden):
    for k in (10, 30):
        out, cnt = get_partial_mesh_batch(dev(golden["vcn_input"], cuda), dev(golden["VCN_VC.coarse"], cuda), k=k,
                                          return_count=True)
        want, wcnt = oracle.get_partial_mesh_batch(golden["vcn_input"], golden["VCN_VC.coarse"], k=k)
        np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
        np.testing.assert_array_equal(out.cpu().numpy(), want)
        for b in range(3):   # same SET as the reference's cKDTree run (its row order is CPython set order)
            np.testing.assert_array_equal(np.unique(out[b].cpu().numpy(), axis=0), np.unique(golden[f"surface_k{k}"][b], axis=0))


def test_surface_select_with_duplicate_and_padded_objects(cuda):
    part, dense, _ = synth.make_object_clouds(43, 3, 150, 1024)
    tiled = np.tile(part, (1, 7, 1))[:, :1024]
    tiled[2] = 0   # zero-padded object (models/VCN.py:58)
    out, cnt = get_partial_mesh_batch(dev(tiled, cuda), dev(dense, cuda), k=20, return_count=True)
    want, wcnt = oracle.get_partial_mesh_batch(tiled, dense, k=20)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    assert wcnt[2] == 20


@pytest.mark.parametrize("cfg", [(2, 2048, 16384, 20), (2, 300, 5000, 16), (3, 1024, 4096, 64), (2, 37, 40, 5), (1, 1, 1, 1)])
def test_surface_select_shapes_vs_oracle(cuda, cfg):
    """C4 (2048 partial points against a 16,384-point completed cloud: sorted cloud read from global memory)
    and ragged sizes; the sweep along the longest axis must give exactly the brute-force union."""
    B, NP, R, k = cfg
    part, dense, _ = synth.make_object_clouds(47, B, NP, R)
    out, cnt = get_partial_mesh_batch(dev(part, cuda), dev(dense, cuda), k=k, surface_pts=1500, return_count=True)
    want, wcnt = oracle.get_partial_mesh_batch(part, dense, k=k, surface_pts=1500)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_surface_select_degenerate_axis_and_exact_ties(cuda):
    """A flat cloud (two axes constant) and a lattice full of equidistant neighbours: ties go to the lower index."""
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(4), indexing="ij"), -1).reshape(1, -1, 3).astype(np.float32)
    g = g[:, rng.permutation(g.shape[1])]
    flat = np.zeros((1, 1024, 3), np.float32); flat[0, :, 1] = rng.permutation(1024) * 0.25
    dense = np.concatenate([g, flat])
    part = np.concatenate([g[:, :512] + np.float32(0.5), flat[:, ::2] + np.float32(0.125)])
    out, cnt = get_partial_mesh_batch(dev(part, cuda), dev(dense, cuda), k=7, return_count=True)
    want, wcnt = oracle.get_partial_mesh_batch(part, dense, k=7)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_largest_cluster_vs_oracle(cuda):
    _, dense, _ = synth.make_object_clouds(51, 4, 64, 1024)
    pc = dense.copy()
    pc[:, 600:1000] += 30.0                                             # second, smaller blob
    pc[:, 1000:] = 99.0 + 5.0 * np.arange(24, dtype=np.float32)[None, :, None]   # isolated points = noise
    pc[3, :512] = pc[3, 512:]                                            # exact duplicates (tiled surfaces)
    for eps, mp in ((0.4, 2), (0.2, 2), (0.3, 1)):
        out, cnt = get_largest_cluster_batch(dev(pc, cuda), eps=eps, min_points=mp, total_pts=1024, return_count=True)
        want, wcnt = oracle.get_largest_cluster_batch(pc, eps=eps, min_points=mp, total_pts=1024)
        np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
        np.testing.assert_array_equal(out.cpu().numpy(), want)
    assert wcnt.max() < 1024 and wcnt.min() >= 1


@pytest.mark.parametrize("scale", [1.0, 0.05])
def test_largest_cluster_periodic_vs_oracle(cuda, scale):
    """tiled clouds (the kNN surface selection's output) with the period hint: same result as clustering all rows.
    scale 0.05 squeezes the cloud so that every pair is adjacent (what a random-init VCN produces)."""
    _, dense, _ = synth.make_object_clouds(52, 6, 64, 1024)
    dense = (dense - dense.mean(axis=1, keepdims=True)) * scale + dense.mean(axis=1, keepdims=True)
    periods = np.array([1, 7, 178, 512, 1000, 1024], dtype=np.int32)
    pc = np.stack([np.tile(dense[b, :m], (1024 // m + 1, 1))[:1024] for b, m in enumerate(periods)]).astype(np.float32)
    pc[2, 100:130] += 40.0       # a second blob and some isolated rows inside the period
    pc[2] = np.tile(pc[2, :178], (6, 1))[:1024]
    pc[4, 990:1000] = 500.0 + 3.0 * np.arange(10, dtype=np.float32)[:, None]
    pc[4] = np.tile(pc[4, :1000], (2, 1))[:1024]
    for eps, mp, total in ((0.3, 2, 1024), (0.2, 1, 1024), (0.3, 2, 2500)):
        out, cnt, dis = get_largest_cluster_batch(dev(pc, cuda), eps=eps, min_points=mp, total_pts=total, return_count=True,
                                                  period=dev(periods, cuda), return_distinct=True)
        plain, pcnt = get_largest_cluster_batch(dev(pc, cuda), eps=eps, min_points=mp, total_pts=total, return_count=True)
        want, wcnt = oracle.get_largest_cluster_batch(pc, eps=eps, min_points=mp, total_pts=total)
        np.testing.assert_array_equal(pcnt.cpu().numpy(), wcnt)
        np.testing.assert_array_equal(plain.cpu().numpy(), want)
        np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
        np.testing.assert_array_equal(out.cpu().numpy(), want)
        # distinct member rows = what np.unique(clustered) keeps (SEE_VCN.py:113,244); with period << 1024 far fewer than cnt
        wdis = oracle.distinct_rows(want, np.minimum(wcnt, total))
        np.testing.assert_array_equal(dis.cpu().numpy(), wdis)
        assert (wdis[:3] < wcnt[:3]).all()
        for b in range(len(pc)):   # and they are exactly the leading rows of the output
            lead = out[b, : int(dis[b])].cpu().numpy()
            assert len(np.unique(lead, axis=0)) == len(lead)
            np.testing.assert_array_equal(np.unique(lead, axis=0), np.unique(want[b][: min(int(wcnt[b]), total)], axis=0))


def test_vcn_inference_wrapper(cuda, golden):
    """host numpy in / out through the reference-shaped VCN.inference"""
    from seevcn_b200.see.surface_completion.models.VCN import VCN
    cfg = {"MODEL": "VCN_VC", "NORM_WITH_GT": False, "SEL_K_NEAREST": 10, "CLUSTER_EPS": 0.3, "BATCH_SIZE_LIMIT": 32}
    sd = oracle.make_state_dict("VCN_VC", seed=0)
    vcn = VCN(cfg, gpu_id=0, state_dict=sd, precision="fp32")
    clouds = [golden["vcn_input"][0][:300], golden["vcn_input"][1], golden["vcn_input"][2][:57]]
    ret = vcn.inference(clouds, batch_size_limit=32, k=10, eps=0.3, rng=np.random.default_rng(4))
    assert ret["input"].shape == (3, 1024, 3) and ret["coarse"].shape == (3, 1024, 3)
    want = oracle.vcn_forward_ref(sd, ret["input"], None, "VCN_VC")["coarse"].numpy()
    np.testing.assert_allclose(ret["coarse"], want, atol=5e-4)
    surf, _ = oracle.get_partial_mesh_batch(ret["input"], ret["coarse"], k=10)
    np.testing.assert_array_equal(ret["surface"], surf)
    clus, _ = oracle.get_largest_cluster_batch(surf, eps=0.3, min_points=2, total_pts=1024)
    np.testing.assert_array_equal(ret["clustered"], clus)


# --------------------------------------------------------------------- stages 2+5: VCN --
def rel_chamfer(a, b, inp):
    """Chamfer(a, b) relative to the Chamfer scale of the cloud (Chamfer(b, centroid))"""
    cd = oracle.chamfer_l2(a, b)
    scale = ((b - b.mean(axis=1, keepdims=True)) ** 2).sum(-1).mean(-1)
    return cd / scale


@pytest.mark.parametrize("cfg", [(256, 64, 128, 256, 0), (1024, 128, 1024, 1024, 2), (3000, 512, 512, 1000, 1),
                                 (37, 1024, 3072, 1, 0), (2048 + 77, 256, 256, 2048 + 77, 1), (5, 100, 130, 2, 0)])
def test_tc_linear_layer_bf16(cuda, cfg):
    """the tcgen05 GEMM on its own against a float64 product of the bf16-rounded operands"""
    from seevcn_b200 import _abi
    rows, cin, cout, rpo, act = cfg
    g = torch.Generator().manual_seed(rows + cin)
    X = torch.randn(rows, cin, generator=g); W = torch.randn(cout, cin, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g)
    nobj = (rows + rpo - 1) // rpo
    ob = torch.randn(nobj, cout, generator=g)
    L = _abi.lib()
    ws = torch.empty(L.seevcn_linear_bf16_workspace_bytes(rows, cin, cout), dtype=torch.uint8, device=cuda)
    Y = torch.empty(rows, cout, device=cuda)
    cm = torch.full((nobj, cout), float("-inf"), device=cuda)
    dX, dW, db, dob = X.to(cuda), W.to(cuda), bias.to(cuda), ob.to(cuda)
    _abi.check(L.seevcn_linear_bf16(rows, cin, cout, _abi.ptr(dX), _abi.ptr(dW), _abi.ptr(db), _abi.ptr(dob), rpo, act,
                                    _abi.ptr(Y), _abi.ptr(cm), _abi.ptr(ws), ws.numel(), _abi.stream()))
    torch.cuda.synchronize()
    ref = X.bfloat16().double() @ W.bfloat16().double().t() + bias.double() + ob.double().repeat_interleave(rpo, 0)[:rows]
    ref = torch.relu(ref) if act == 1 else torch.where(ref > 0, ref, 0.01 * ref) if act == 2 else ref
    np.testing.assert_allclose(Y.cpu().double().numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)
    want_max = torch.stack([ref[o * rpo: min((o + 1) * rpo, rows)].max(0)[0] for o in range(nobj)])
    np.testing.assert_allclose(cm.cpu().double().numpy(), want_max.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["VCN_VC", "VCN_CN"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vcn_forward_vs_golden(cuda, golden, name, precision):
    sd = oracle.make_state_dict(name, seed=0)
    model = MODELS.build({"NAME": name}, precision=precision)
    model.load_state_dict(sd)
    model.to(cuda).eval()
    ret = model({"input": dev(golden["vcn_input"], cuda), "gt_boxes": dev(golden["vcn_gt_boxes"], cuda)})
    coarse = ret["coarse"].cpu().numpy()
    want = golden[f"{name}.coarse"]
    rc = rel_chamfer(coarse, want, golden["vcn_input"])
    assert (rc < 1e-3).all(), rc                       # north_star tolerance: 1e-3 relative Chamfer
    if precision == "fp32":
        np.testing.assert_allclose(coarse, want, rtol=0, atol=5e-4)
        if name == "VCN_VC":
            np.testing.assert_allclose(ret["reg_rot"].cpu().numpy(), golden["VCN_VC.reg_rot"], atol=1e-4)
            np.testing.assert_allclose(ret["reg_centre"].cpu().numpy(), golden["VCN_VC.reg_centre"], atol=1e-4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vcn_forward_many_objects_vs_oracle(cuda, precision):
    """crosses the internal object-chunk boundary; ragged last chunk"""
    part, _, _ = synth.make_object_clouds(77, 37, 1024, 0)
    sd = oracle.make_state_dict("VCN_VC", seed=5)
    model = MODELS.build({"NAME": "VCN_VC"}, precision=precision)
    model.load_state_dict(sd)
    model.to(cuda).eval()
    got = model({"input": dev(part, cuda)})["coarse"].cpu().numpy()
    want = oracle.vcn_forward_ref(sd, part, None, "VCN_VC")["coarse"].numpy()
    assert (rel_chamfer(got, want, part) < 1e-3).all()


@pytest.mark.parametrize("name,nobj,n", [("VCN_VC", 5, 1024), ("VCN_VC", 150, 1024), ("VCN_CN", 3, 1024), ("VCN_VC", 3, 1000),
                                         ("VCN_VC", 2, 2048)])
def test_vcn_fused_chains_vs_layerwise_and_oracle(cuda, name, nobj, n):
    """the fused tcgen05 chains (activations in TMEM, A operand from tensor memory) against the layer-by-layer
    tcgen05 path and the fp32 restatement; ragged tiles (n = 1000), several objects per CTA, chunk boundary (150)"""
    from seevcn_b200 import _abi
    part, _, boxes = synth.make_object_clouds(91, nobj, n, 0)
    sd = oracle.make_state_dict(name, seed=5)
    model = MODELS.build({"NAME": name}, precision="bf16")
    model.load_state_dict(sd)
    model.to(cuda).eval()
    in_dict = {"input": dev(part, cuda)}
    if name == "VCN_CN":
        in_dict["gt_boxes"] = dev(boxes[:, :7].astype(np.float32), cuda)
    fused = model(in_dict)["coarse"].cpu().numpy()
    model.precision = "bf16_layerwise"                       # same weights, one tcgen05 GEMM launch per layer
    layer = model(in_dict)["coarse"].cpu().numpy()
    model.precision = "bf16"
    want = oracle.vcn_forward_ref(sd, part, boxes[:, :7].astype(np.float32) if name == "VCN_CN" else None, name)["coarse"].numpy()
    assert np.isfinite(fused).all()
    assert (rel_chamfer(fused, want, part) < 1e-3).all(), rel_chamfer(fused, want, part).max()
    assert (rel_chamfer(fused, layer, part) < 1e-3).all()
    # same bf16 roundings in both paths up to accumulation order: point-wise agreement, not just Chamfer
    scale = np.abs(want - want.mean(axis=1, keepdims=True)).max()
    assert np.abs(fused - layer).max() < 2e-2 * scale, (np.abs(fused - layer).max(), scale)


def test_vcn_forward_beyond_one_pass(cuda):
    """more objects than one internal pass holds (512): the passes are independent, so the result equals the forward of
    the two halves, and the fp32 restatement on a sample of objects from both passes"""
    part, _, _ = synth.make_object_clouds(93, 530, 1024, 0)
    sd = oracle.make_state_dict("VCN_VC", seed=5)
    model = MODELS.build({"NAME": "VCN_VC"}, precision="bf16")
    model.load_state_dict(sd)
    model.to(cuda).eval()
    x = dev(part, cuda)
    full = model({"input": x})["coarse"].cpu().numpy()
    halves = np.concatenate([model({"input": x[:300].contiguous()})["coarse"].cpu().numpy(),
                             model({"input": x[300:].contiguous()})["coarse"].cpu().numpy()])
    assert np.isfinite(full).all()
    assert (rel_chamfer(full, halves, part) < 1e-4).all()
    sel = np.array([0, 1, 255, 511, 512, 513, 529])
    want = oracle.vcn_forward_ref(sd, part[sel], None, "VCN_VC")["coarse"].numpy()
    assert (rel_chamfer(full[sel], want, part[sel]) < 1e-3).all()


def test_vcn_forward_dense_decoder_c4_shape(cuda):
    """BASELINE.json configs[3] shape: 2048 input points per object, 16,384-point decoder (shape_fc.4: 1024 -> 49,152),
    then the 16,384-reference kNN surface selection on the output.  The reference class ties the encoder width to
    number_coarse (VCN_VC.py:130) and so only runs with 1024; the decoder is widened here the way §8(a6) counts it."""
    nc, n, nobj = 16384, 2048, 3
    part, _, _ = synth.make_object_clouds(97, nobj, n, 0)
    sd = oracle.make_state_dict("VCN_VC", seed=7, num_coarse=nc)
    model = MODELS.build({"NAME": "VCN_VC"}, precision="bf16")
    model.number_coarse = nc
    model.shape_fc[4] = torch.nn.Linear(1024, 3 * nc)
    model.load_state_dict(sd)
    model.to(cuda).eval()
    x = dev(part, cuda)
    got = model({"input": x})["coarse"]
    assert tuple(got.shape) == (nobj, nc, 3)
    want = oracle.vcn_forward_ref(sd, part, None, "VCN_VC")["coarse"].numpy()
    assert (rel_chamfer(got.cpu().numpy(), want, part) < 1e-3).all()
    surf, cnt = get_partial_mesh_batch(x, got, k=20, surface_pts=n, return_count=True)
    wsurf, wcnt = oracle.get_partial_mesh_batch(part, got.cpu().numpy(), k=20, surface_pts=n)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wcnt)
    np.testing.assert_array_equal(surf.cpu().numpy(), wsurf)


def test_data_processor_range_mask_shuffle_voxels(cuda):
    """pcdet DataProcessor on the device (data_processor.py:78-143): x/y range mask (inclusive), seeded shuffle, hard voxels."""
    from seevcn_b200.pcdet.datasets.processor.data_processor import DataProcessor, mask_points_by_range
    from seevcn_b200 import _abi
    pts, _ = synth.make_frame(1003)
    pts = np.concatenate([pts, np.array([[75.2, -75.2, 0.0], [75.2000046, 0.0, 0.0], [-75.2, 75.2, 1.0], [80.0, 0.0, 0.0]], np.float32)])
    cfgs = [{"NAME": "mask_points_and_boxes_outside_range", "REMOVE_OUTSIDE_BOXES": True},
            {"NAME": "shuffle_points", "SHUFFLE_ENABLED": {"train": True, "test": False}},
            {"NAME": "transform_points_to_voxels", "VOXEL_SIZE": WAYMO[1], "MAX_POINTS_PER_VOXEL": 5,
             "MAX_NUMBER_OF_VOXELS": {"train": 80000, "test": 90000}}]
    mask = mask_points_by_range(pts, WAYMO[0])
    assert 0 < (~mask).sum() < len(pts) // 4 and mask[-4] and not mask[-3] and mask[-2]
    for training in (False, True):
        dp = DataProcessor(cfgs, WAYMO[0], training=training, num_point_features=3, seed=9)
        assert dp.grid_size.tolist() == WAYMO[2]
        out = dp.forward({"points": dev(pts, cuda), "use_lead_xyz": True})
        want = pts[mask]
        if training:
            perm = np.array([_abi.lib().seevcn_shuffle_perm(j, len(want), 9) for j in range(len(want))])
            assert sorted(perm.tolist()) == list(range(len(want)))
            want = want[perm]
        np.testing.assert_array_equal(out["points"].cpu().numpy(), want)
        wv, wc, wn = oracle.hard_voxelize(want, WAYMO[0], WAYMO[1], WAYMO[2], 5, 80000 if training else 90000)
        np.testing.assert_array_equal(out["voxel_coords"].cpu().numpy(), wc)
        np.testing.assert_array_equal(out["voxel_num_points"].cpu().numpy(), wn)
        np.testing.assert_array_equal(out["voxels"].cpu().numpy(), wv)


def test_hard_voxelize_frames_vs_oracle_per_frame(cuda):
    """The batched hard voxelizer (F frames, padded outputs, device-side row counts) against the oracle frame by frame;
    MeanVFE with the int32 counts on the padded slots."""
    rng = np.random.default_rng(4)
    F, S, MV, T = 3, 40000, 20000, 5
    pts, _ = synth.make_stream(F, n_beams=32, n_az=1250, n_boxes=8, first_seed=40)
    pts = pts[:, rng.permutation(pts.shape[1])]
    counts = np.array([40000, 25000, 31111], np.int32)
    gen = VoxelGeneratorWrapper(WAYMO[1], WAYMO[0], 3, T, MV)
    v, c, n, nv = gen.generate_frames_device(dev(pts, cuda), dev(counts, cuda))
    assert tuple(v.shape) == (F, MV, T, 3) and tuple(c.shape) == (F, MV, 4)
    feats = MeanVFE(model_cfg={}, num_point_features=3)({"voxels": v.view(F * MV, T, 3), "voxel_num_points": n.view(-1)})["voxel_features"]
    feats = feats.view(F, MV, 3).cpu().numpy()
    v, c, n, nv = v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy(), nv.cpu().numpy()
    capped = 0
    for f in range(F):
        wv, wc, wn = oracle.hard_voxelize(pts[f, : counts[f]], WAYMO[0], WAYMO[1], WAYMO[2], T, MV)
        m = int(nv[f])
        assert m == len(wc)
        capped += m == MV
        np.testing.assert_array_equal(c[f, :m, 1:], wc)
        assert (c[f, :m, 0] == f).all()
        np.testing.assert_array_equal(n[f, :m], wn)
        np.testing.assert_array_equal(v[f, :m], wv)
        assert (n[f, m:] == 0).all()
        np.testing.assert_allclose(feats[f, :m], oracle.mean_vfe(wv, wn.astype(np.float32)), rtol=1e-6, atol=1e-6)
    assert capped >= 1            # the per-frame voxel cap was hit at least once


# ------------------------------------------------------------------ stage 6: voxelization --
WAYMO = ([-75.2, -75.2, -2, 75.2, 75.2, 4], [0.1, 0.1, 0.15], [1504, 1504, 40])


def test_mean_vfe_vs_golden(cuda, golden):
    vfe = MeanVFE(model_cfg={}, num_point_features=3)
    out = vfe({"voxels": dev(golden["meanvfe_voxels"], cuda), "voxel_num_points": dev(golden["meanvfe_num"], cuda)})
    np.testing.assert_allclose(out["voxel_features"].cpu().numpy(), golden["meanvfe_out"], rtol=1e-6, atol=1e-6)
    assert vfe.get_output_feature_dim() == 3


@pytest.mark.parametrize("nframes", [1, 3])
def test_dynamic_voxelize_vs_oracle(cuda, nframes):
    pts, _ = synth.make_stream(nframes)
    points = np.concatenate([np.concatenate([np.full((pts.shape[1], 1), b, np.float32), pts[b]], axis=1) for b in range(nframes)])
    want_c, want_f, want_n = oracle.dynamic_voxelize(points, *WAYMO)
    vfe = DynamicMeanVFE({}, 3, WAYMO[1], WAYMO[2], WAYMO[0])
    out = vfe({"points": dev(points, cuda), "batch_size": nframes})
    np.testing.assert_array_equal(out["voxel_coords"].cpu().numpy(), want_c)          # bit-exact, reference order
    np.testing.assert_array_equal(out["voxel_num_points"].cpu().numpy(), want_n)
    np.testing.assert_allclose(out["voxel_features"].cpu().numpy(), want_f, rtol=1e-5, atol=1e-5)
    assert len(want_c) > 50000
    # sort=False is accepted and ignored: rows always come out in the reference's order; batch_size derived from the rows
    c, f, n = dynamic_voxelize(dev(points, cuda), *WAYMO, sort=False)
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    np.testing.assert_array_equal(n.cpu().numpy(), want_n)
    # bit-reproducible: integer accumulation does not depend on the order the points arrive in
    np.testing.assert_array_equal(f.cpu().numpy(), out["voxel_features"].cpu().numpy())
    perm = np.random.default_rng(0).permutation(len(points))
    c2, f2, n2 = dynamic_voxelize(dev(points[perm], cuda), *WAYMO, batch_size=nframes)
    np.testing.assert_array_equal(c2.cpu().numpy(), want_c)
    np.testing.assert_array_equal(f2.cpu().numpy(), f.cpu().numpy())


def test_dynamic_voxelize_vs_reference_module_golden(cuda):
    """a9 pinned by the reference's own code: tests/golden/make_golden_dynvfe.py executed DynamicMeanVFE.forward
    (dynamic_mean_vfe.py:37-76; scatter_mean stubbed by index_add_/bincount) and committed its outputs."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_dynvfe_v1.npz"))
    KITTI = ([0.0, -40.0, -3.0, 70.4, 40.0, 1.0], [0.1, 0.1, 0.15], [704, 800, 27])
    for tag, cfg, c in (("A", WAYMO, 3), ("B", KITTI, 5)):
        vfe = DynamicMeanVFE({}, c, cfg[1], cfg[2], cfg[0])
        assert vfe.get_output_feature_dim() == c
        out = vfe({"points": dev(g[f"{tag}_points"], cuda), "batch_size": 2})
        np.testing.assert_array_equal(out["voxel_coords"].cpu().numpy(), g[f"{tag}_voxel_coords"])
        np.testing.assert_allclose(out["voxel_features"].cpu().numpy(), g[f"{tag}_voxel_features"], rtol=1e-5, atol=1e-6)


def test_dynamic_voxelize_dense_voxels_vs_float64_mean(cuda):
    """Thousands of points in one voxel (what a collapsed completed cloud produces): the mean equals the float64 mean
    at 1e-6 and is identical run to run; more voxels in a bucket than one accumulation pass holds; voxel cap."""
    rng = np.random.default_rng(5)
    blobs = []
    for centre, n in (((61.234, -33.71, 1.03), 6000), ((-70.01, 74.9, -1.9), 4500), ((0.01, 0.02, 0.03), 5000)):
        lo = np.floor((np.array(centre) - WAYMO[0][:3]) / WAYMO[1]) * WAYMO[1] + WAYMO[0][:3]
        blobs.append(lo + rng.uniform(0.001, 0.999, (n, 3)) * WAYMO[1])
    # 700 occupied voxels in a y run at fixed x (small buckets next to each other)
    run = np.stack([np.full(2100, 10.03), -20.0 + 0.1 * (np.arange(2100) % 700) + 0.05, np.full(2100, 0.51)], axis=1)
    run += rng.uniform(-0.02, 0.02, run.shape)
    # a wall: 6000 points over ~800 voxels of two neighbouring buckets (x fixed, 2 m of y, all of z): big buckets with
    # several accumulation passes of 64 rows
    wall = np.stack([rng.uniform(20.0, 20.1, 6000), rng.uniform(5.0, 7.0, 6000), rng.uniform(-2.0, 4.0, 6000)], axis=1)
    xyz = np.concatenate(blobs + [run, wall]).astype(np.float32)
    points = np.concatenate([np.zeros((len(xyz), 1), np.float32), xyz], axis=1)
    points = points[rng.permutation(len(points))]
    want_c, want_f, want_n = oracle.dynamic_voxelize(points, *WAYMO)          # float64 sums
    assert want_n.max() >= 4096
    c, f, n = dynamic_voxelize(dev(points, cuda), *WAYMO, batch_size=1)
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    np.testing.assert_array_equal(n.cpu().numpy(), want_n)
    np.testing.assert_allclose(f.cpu().numpy(), want_f, rtol=1e-6, atol=1e-7)
    for _ in range(3):
        c2, f2, n2 = dynamic_voxelize(dev(points, cuda), *WAYMO, batch_size=1)
        np.testing.assert_array_equal(f2.cpu().numpy(), f.cpu().numpy())
    # capacity: rows beyond max_voxels are dropped, the leading rows are the same
    c3, f3, n3 = dynamic_voxelize(dev(points, cuda), *WAYMO, batch_size=1, max_voxels=100)
    np.testing.assert_array_equal(c3.cpu().numpy(), want_c[:100])
    np.testing.assert_array_equal(f3.cpu().numpy(), f.cpu().numpy()[:100])


def test_dynamic_voxelize_frames_matches_concatenated_matrix(cuda):
    """The two-source form (frames + object clouds, no [b,x,y,z] matrix) against the oracle on the concatenation."""
    from seevcn_b200.pcdet.models.backbones_3d.vfe.dynamic_mean_vfe import dynamic_voxelize_frames
    from seevcn_b200.pipeline import WAYMO_VOXEL_CFG
    rng = np.random.default_rng(11)
    F, P, O, S = 3, 5000, 7, 256
    frames = (rng.standard_normal((F, P, 3)) * [30, 30, 1.5]).astype(np.float32)
    objs = (rng.standard_normal((O, S, 3)) * [2, 1, 0.7] + rng.standard_normal((O, 1, 3)) * [20, 20, 0.3]).astype(np.float32)
    obj_frame = rng.integers(0, F, O).astype(np.int32)
    rows = [np.concatenate([np.full((P, 1), f, np.float32), frames[f]], 1) for f in range(F)]
    rows += [np.concatenate([np.full((S, 1), obj_frame[o], np.float32), objs[o]], 1) for o in range(O)]
    wc, wf, wn = oracle.dynamic_voxelize(np.concatenate(rows), *WAYMO_VOXEL_CFG)
    for objs_d, of_d in ((dev(objs, cuda), dev(obj_frame, cuda)), (None, None)):
        coords, feats, counts, num = dynamic_voxelize_frames(dev(frames, cuda), objs_d, of_d, *WAYMO_VOXEL_CFG)
        m = int(num.item())
        if objs_d is None:
            wc, wf, wn = oracle.dynamic_voxelize(np.concatenate(rows[:F]), *WAYMO_VOXEL_CFG)
        assert m == len(wc)
        np.testing.assert_array_equal(coords[:m].cpu().numpy(), wc)
        np.testing.assert_array_equal(counts[:m].cpu().numpy(), wn)
        np.testing.assert_allclose(feats[:m].cpu().numpy(), wf, rtol=1e-5, atol=1e-5)


def test_pipeline_stream_equals_single_runs(cuda):
    """run_stream (crop look-ahead, deferred voxel count) and HostStream give what run() gives batch by batch."""
    from seevcn_b200.pipeline import CompletionPipeline, HostStream
    pipe = CompletionPipeline("VCN_VC", oracle.make_state_dict("VCN_VC", seed=0), cuda, sel_k=10, cluster_eps=0.3)
    batches = [synth.make_stream(2, n_beams=32, n_az=1090, n_boxes=10, first_seed=300 + 2 * i) for i in range(4)]
    dbat = [(dev(p, cuda), dev(b, cuda)) for p, b in batches]
    singles = [pipe.run(p, b, seed=0) for p, b in dbat]
    streamed = list(pipe.run_stream(iter(dbat), seed=0))
    pipe2 = CompletionPipeline("VCN_VC", oracle.make_state_dict("VCN_VC", seed=0), cuda, sel_k=10, cluster_eps=0.3, streams=2)
    two = list(pipe2.run_stream(iter(dbat), seed=0))             # consecutive batches on two CUDA streams
    torch.cuda.synchronize()
    for a, b in zip(singles, two):
        for key in ("clustered", "voxel_coords", "voxel_num_points", "voxel_features"):
            np.testing.assert_array_equal(a[key].cpu().numpy(), b[key].cpu().numpy())
    hs = HostStream(pipe, 2, batches[0][0].shape[1], batches[0][1].shape[1])
    pinned = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(b).pin_memory()) for p, b in batches]
    hosted = [{k: v.clone() for k, v in res.items()} for res in hs.run(iter(pinned), seed=0)]
    assert len(streamed) == len(hosted) == 4
    for a, b, c in zip(singles, streamed, hosted):
        assert a["input"].shape[0] > 0
        for key in ("clustered", "voxel_coords", "voxel_num_points"):
            np.testing.assert_array_equal(a[key].cpu().numpy(), b[key].cpu().numpy())
            np.testing.assert_array_equal(a[key].cpu().numpy(), c[key].numpy())
        np.testing.assert_allclose(a["voxel_features"].cpu().numpy(), c["voxel_features"].numpy(), rtol=1e-5, atol=1e-5)


def test_dynamic_voxelize_extra_features_and_overflow_batch(cuda):
    rng = np.random.default_rng(9)
    n = 20000
    pts = np.concatenate([rng.integers(20, 30, (n, 1)).astype(np.float32), rng.uniform(-80, 80, (n, 2)).astype(np.float32),
                          rng.uniform(-3, 5, (n, 1)).astype(np.float32), rng.standard_normal((n, 2)).astype(np.float32)], axis=1)
    want_c, want_f, want_n = oracle.dynamic_voxelize(pts, *WAYMO)   # batch >= 24 overflows the reference's int32 key
    c, f, cnt = dynamic_voxelize(dev(pts, cuda), *WAYMO)
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    np.testing.assert_array_equal(cnt.cpu().numpy(), want_n)
    np.testing.assert_allclose(f.cpu().numpy(), want_f, rtol=1e-5, atol=1e-5)


def test_dynamic_voxelize_properties_full_size(cuda):
    """C5-size: counts sum to the in-range points, means lie inside their voxel, voxels are unique."""
    pts, _ = synth.make_stream(8)
    points = np.concatenate([np.concatenate([np.full((pts.shape[1], 1), b, np.float32), pts[b]], axis=1) for b in range(8)])
    c, f, n = dynamic_voxelize(dev(points, cuda), *WAYMO, batch_size=8)
    c, f, n = c.cpu().numpy(), f.cpu().numpy(), n.cpu().numpy()
    lo, vs, gs = np.array(WAYMO[0][:3], np.float32), np.array(WAYMO[1], np.float32), np.array(WAYMO[2])
    pc = np.floor((points[:, 1:4] - lo) / vs)
    inside = ((pc >= 0) & (pc < gs)).all(axis=1)
    assert n.sum() == inside.sum()
    assert len(np.unique(c, axis=0)) == len(c)
    vox_lo = lo + c[:, [3, 2, 1]] * vs
    assert ((f >= vox_lo - 1e-3) & (f <= vox_lo + vs + 1e-3)).all()


def test_hard_voxelize_vs_oracle(cuda):
    pts, _ = synth.make_frame(1002)
    pts = pts[np.random.default_rng(1).permutation(len(pts))]     # shuffled like data_processor.py:93-103
    for max_voxels in (90000, 5000):
        gen = VoxelGeneratorWrapper(WAYMO[1], WAYMO[0], 3, 5, max_voxels)
        assert gen.grid_size == WAYMO[2]
        v, c, n = gen.generate(pts)
        wv, wc, wn = oracle.hard_voxelize(pts, WAYMO[0], WAYMO[1], WAYMO[2], 5, max_voxels)
        np.testing.assert_array_equal(c, wc)
        np.testing.assert_array_equal(n, wn)
        np.testing.assert_array_equal(v, wv)
        assert len(c) == min(max_voxels, len(c)) and len(c) > 1000


def test_chamfer_kernel(cuda):
    from seevcn_b200 import _abi
    rng = np.random.default_rng(2)
    a = rng.standard_normal((2, 700, 3)).astype(np.float32)
    b = rng.standard_normal((2, 1300, 3)).astype(np.float32)
    da, db = dev(a, cuda), dev(b, cuda)
    d1 = torch.empty((2, 700), device=cuda); d2 = torch.empty((2, 1300), device=cuda)
    _abi.check(_abi.lib().seevcn_chamfer(2, 700, 1300, _abi.ptr(da), _abi.ptr(db), _abi.ptr(d1), _abi.ptr(d2), _abi.stream()))
    got = (d1.mean(1) + d2.mean(1)).cpu().numpy()
    np.testing.assert_allclose(got, oracle.chamfer_l2(a, b), rtol=1e-4)


def test_select_objects_matches_argwhere(cuda):
    """seevcn_select_objects = np.argwhere(counts >= MIN_LIDAR_PTS) in frame-major order (SEE_VCN.py:71)."""
    rng = np.random.default_rng(17)
    for B, T in ((1, 1), (3, 50), (8, 300), (2, 0)):
        counts = rng.integers(0, 60, (B, T)).astype(np.int32)
        of, ob, num = roi.select_objects(dev(counts, cuda), 30)
        want = np.argwhere(counts >= 30)
        n = int(num.item())
        assert n == len(want)
        np.testing.assert_array_equal(of[:n].cpu().numpy(), want[:, 0])
        np.testing.assert_array_equal(ob[:n].cpu().numpy(), want[:, 1])
