"""world_size-2 gloo tests (CPU) of the N>1 host path: task partition + per-round grid exchange give the same grid as
a single rank."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffuman4d_b200.sharding import exchange_grid_updates, frame_shard, shard_tasks


def test_shard_tasks_balanced():
    sizes = [len(shard_tasks(44, r, 8)) for r in range(8)]
    assert sizes == [6, 6, 6, 6, 5, 5, 5, 5]
    assert sorted(sum((shard_tasks(44, r, 8) for r in range(8)), [])) == list(range(44))
    assert shard_tasks(3, 5, 8) == [] and shard_tasks(16, 1, 8) == [2, 3]
    with pytest.raises(ValueError):
        shard_tasks(4, 8, 8)


def test_frame_shard():
    assert frame_shard(16, 3, 8) == (6, 8) and frame_shard(24, 7, 8) == (21, 24)
    with pytest.raises(ValueError):
        frame_shard(12, 0, 8)


def _fake_denoise(task, lat):
    return lat * 0.5 + task


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_spa, n_tem = 5, 3            # grid; one round of "temporal" tasks = one task per camera
        grid = {(s, t): torch.full((4, 2, 2), float(s * 10 + t)) for s in range(n_spa) for t in range(n_tem)}
        ti = {k: 0 for k in grid}
        mine = shard_tasks(n_spa, rank, world)
        keys, lats, tis = [], [], []
        for s in mine:
            for t in range(n_tem):
                keys.append((s, t))
                lats.append(_fake_denoise(s, grid[(s, t)]))
                tis.append(6)
        lat = torch.stack(lats) if lats else torch.zeros(0, 4, 2, 2)
        upd = exchange_grid_updates(keys, lat, torch.tensor(tis, dtype=torch.int64))
        for k, (v, i) in upd.items():
            grid[k], ti[k] = v, i
        q.put((rank, {k: v.tolist() for k, v in grid.items()}, ti))   # plain lists: no shared-memory handles
    finally:
        dist.destroy_process_group()


def test_two_rank_round_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = {(s, t): _fake_denoise(s, torch.full((4, 2, 2), float(s * 10 + t))) for s in range(5) for t in range(3)}
    for _, grid, ti in res:
        assert set(grid) == set(ref)
        for k in ref:
            assert torch.equal(torch.tensor(grid[k]), ref[k]) and ti[k] == 6
