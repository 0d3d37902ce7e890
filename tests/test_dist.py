"""world_size-2 gloo test of the frame sharding / all-gather-v host logic (CPU)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seevcn_b200.dist import shard_range, all_gather_v, all_gather_padded, rebase_batch_index, FrameGather


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(5, rank, world)                       # 5 frames over 2 ranks -> 3 + 2
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3).contiguous()   # (n_r, 4, 3)
    full, counts = all_gather_v(local)
    coords = torch.tensor([[0, 1, 2, 3], [hi - lo - 1, 4, 5, 6]], dtype=torch.int32)
    allc, _ = all_gather_v(rebase_batch_index(coords, lo))
    pg = all_gather_padded(local, 4, async_op=True)            # streaming form: fixed capacity, counts as a tensor
    parts = pg.parts()
    # the per-batch collection of completed clouds + voxel tensors (gloo runs the all-gather fallback)
    fg = FrameGather(torch.device("cpu"), world, rank, max_obj=4, rows_per_obj=4, frames=2, max_rows=64)
    gens = []
    for k in range(3):
        m = 3 + rank + k
        out = {"clustered": local + k, "voxel_coords": torch.arange(m * 4, dtype=torch.int32).view(m, 4) * (rank + 1),
               "voxel_features": torch.full((m, 3), float(rank + k)), "voxel_num_points": torch.full((m,), rank + 7, dtype=torch.int32)}
        gens.append(fg.push(out, frame_offset=lo))
    got = fg.parts(gens[-1])                                   # generation of the last push (k = 2)
    if rank == 0:
        ret["fg_kind"] = fg.kind
        ret["fg_obj"] = [int(p["clustered"].shape[0]) for p in got]
        ret["fg_clu"] = torch.cat([p["clustered"] for p in got]).clone()
        ret["fg_m"] = [int(p["voxel_coords"].shape[0]) for p in got]
        ret["fg_b0"] = [int(p["voxel_coords"][0, 0]) for p in got]
        ret["fg_num"] = [int(p["voxel_num_points"][0]) for p in got]
        ret["fg_feat"] = [float(p["voxel_features"][0, 0]) for p in got]
    if rank == 0:
        ret["full"] = full.clone(); ret["counts"] = counts; ret["coords"] = allc.clone()
        ret["padded_counts"] = pg.counts.tolist(); ret["padded"] = torch.cat(parts).clone(); ret["padded_shape"] = tuple(pg.out.shape)
    dist.destroy_process_group()


def test_all_gather_v_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["counts"] == [3, 2]
    assert ret["full"][:, 0, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert ret["coords"][:, 0].tolist() == [0, 2, 3, 4]
    assert ret["padded_counts"] == [3, 2] and ret["padded_shape"] == (2, 4, 4, 3)
    assert torch.equal(ret["padded"], ret["full"])
    assert ret["fg_kind"] == "gloo" and ret["fg_obj"] == [3, 2] and ret["fg_m"] == [5, 6]
    assert torch.equal(ret["fg_clu"], ret["full"] + 2)
    assert ret["fg_b0"] == [0, 3] and ret["fg_num"] == [7, 8] and ret["fg_feat"] == [2.0, 3.0]
