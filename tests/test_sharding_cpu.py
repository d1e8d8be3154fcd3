"""CPU, world_size 2 over gloo: the window-sharding exchange reassembles every window's record on every rank."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _make_pred(w, T, H, W):
    g = torch.Generator().manual_seed(100 + w)
    return {"pts3d": torch.randn(T, H, W, 3, generator=g), "conf": torch.rand(T, H, W, 1, generator=g),
            "inverse_depthmap": torch.rand(T, H, W, 1, generator=g), "traj": torch.randn(T, 4, 4, generator=g)}


def _worker(rank, world, port, n_windows, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geo4d_b200 import sharding
    T, H, W = 16, 6, 8
    mine = sharding.windows_for_rank(n_windows, rank, world)
    local = {w: _make_pred(w, T, H, W) for w in mine}
    allp = sharding.gather_predictions(local, n_windows, T, H, W)
    ok = len(allp) == n_windows
    for w in range(n_windows):
        ref = _make_pred(w, T, H, W)
        for k in ref:
            ok = ok and torch.equal(allp[w][k], ref[k])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_predictions_world2():
    for n_windows, port in ((2, 29611), (5, 29612)):
        with mp.Manager() as m:
            ret = m.dict()
            mp.spawn(_worker, args=(2, port, n_windows, ret), nprocs=2, join=True)
            assert dict(ret) == {0: True, 1: True}


def test_assignment_covers_all_windows_once():
    from geo4d_b200 import sharding
    for n, world in ((32, 8), (6, 4), (1, 1), (3, 8)):
        seen = sorted(w for r in range(world) for w in sharding.windows_for_rank(n, r, world))
        assert seen == list(range(n))


def _worker_groups(rank, world, port, n_groups, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geo4d_b200 import sharding
    per = -(-n_groups // world)
    mine = sharding.windows_for_rank(n_groups, rank, world)
    rec = torch.zeros(per, 3, dtype=torch.float64)
    for j, g in enumerate(mine):                       # what a rank would compute for its windows
        rec[j] = torch.tensor([g + 0.25, -g - 0.5, 1.0 / (g + 1)], dtype=torch.float64)
    full = sharding.gather_group_records(rec, n_groups)
    want = torch.tensor([[g + 0.25, -g - 0.5, 1.0 / (g + 1)] for g in range(n_groups)], dtype=torch.float64)
    ret[rank] = bool(torch.equal(full, want))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_group_records_world2():
    """the per-window LAD results of the replicated alignment: fitted round-robin, identical table everywhere"""
    for n_groups, port in ((2, 29621), (5, 29622)):
        with mp.Manager() as m:
            ret = m.dict()
            mp.spawn(_worker_groups, args=(2, port, n_groups, ret), nprocs=2, join=True)
            assert dict(ret) == {0: True, 1: True}


def test_partition_images_is_contiguous_balanced_and_complete():
    """image ranges of the sharded alignment loop: contiguous, cover every image once, balanced by (1 + windows
    observing the image) -- including more ranks than images and a single rank"""
    from geo4d_b200 import sharding
    for edges, world in (([1] * 8 + [2] * 56 + [1] * 8, 8), ([1] * 16, 1), ([1] * 16, 2), ([2, 2, 2], 8),
                         ([1] * 8 + [2] * 8 + [3] * 14 + [2] * 4 + [1] * 16, 4), ([1] * 72, 16)):
        lo = sharding.partition_images(edges, world)
        assert len(lo) == world + 1 and lo[0] == 0 and lo[-1] == len(edges)
        assert all(b >= a for a, b in zip(lo, lo[1:]))
        cost = [sum(1 + e for e in edges[lo[r]:lo[r + 1]]) for r in range(world)]
        if len(edges) >= 4 * world:
            assert max(cost) <= 1.5 * sum(cost) / world + 4, (edges, world, cost)


def test_flag_epochs_grow_monotonically():
    from geo4d_b200 import sharding
    a = sharding.reserve_flags(500)
    b = sharding.reserve_flags(500)
    c = sharding.reserve_flags(60)
    assert b >= a + 500 and c >= b + 500
