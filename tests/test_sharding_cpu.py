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
