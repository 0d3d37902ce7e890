"""world_size-2 gloo test of the frame sharding / all-gather-v host logic (CPU)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seevcn_b200.dist import shard_range, all_gather_v, all_gather_padded, rebase_batch_index


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
