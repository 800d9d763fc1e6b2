"""CPU-only, world_size-2 gloo: the track sharding + pose all-gather logic of dist.py."""
import importlib, os, socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

D = importlib.import_module('iros20-6d-pose-tracking_b200.dist')


def test_shard_tracks_balanced_and_grouped():
    wid = np.array([i % 21 for i in range(512)])
    shards = D.shard_tracks(wid, 8)
    assert sorted(np.concatenate(shards).tolist()) == list(range(512))
    assert all(len(s) == 64 for s in shards)
    # grouped by weight id: each rank sees at most ceil(21/8)+1 distinct sets, ids non-decreasing
    for s in shards:
        ids = wid[s]
        assert np.all(np.diff(ids) >= 0) and len(set(ids.tolist())) <= 4
    shards = D.shard_tracks(np.zeros(10, int), 4)
    assert [len(s) for s in shards] == [3, 3, 2, 2]
    assert [len(s) for s in D.shard_tracks(np.zeros(2, int), 4)] == [1, 1, 0, 0]


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_tracks, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    wid = np.array([i % 3 for i in range(n_tracks)])
    shards = D.shard_tracks(wid, world)
    # every rank "tracks" its slice: pose[i] = i * ones
    local = torch.stack([torch.full((4, 4), float(i), dtype=torch.float64) for i in shards[rank]]) if len(shards[rank]) else torch.zeros(0, 4, 4, dtype=torch.float64)
    full = D.all_gather_poses(local, shards, rank, world)
    ok = all(bool((full[i] == float(i)).all()) for i in range(n_tracks))
    q.put((rank, ok, tuple(full.shape)))
    dist.destroy_process_group()


def test_all_gather_poses_gloo_world2():
    ctx = mp.get_context('spawn')
    for n_tracks in (7, 8):                       # uneven and even shards
        q = ctx.Queue(); port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_tracks, q)) for r in range(2)]
        [p.start() for p in ps]
        res = [q.get(timeout=120) for _ in ps]
        [p.join(timeout=60) for p in ps]
        assert all(ok for _, ok, _ in res) and all(shape == (n_tracks, 4, 4) for _, _, shape in res)


def test_all_gather_single_rank_is_identity_permutation():
    wid = np.array([2, 0, 1, 0])
    shards = D.shard_tracks(wid, 1)
    local = torch.arange(4, dtype=torch.float64).view(4, 1, 1).expand(4, 4, 4).contiguous()[torch.as_tensor(shards[0])]
    full = D.all_gather_poses(local, shards, 0, 1)
    assert [float(full[i, 0, 0]) for i in range(4)] == [0.0, 1.0, 2.0, 3.0]


def test_gather_plan_identity_and_permutation():
    """GatherPlan: consecutive equal shards need no permutation (the gathered buffer IS the result); anything else gets the
    inverse permutation once, and a plan is reusable call after call (ShardedTracker keeps one)."""
    wid = np.repeat(np.arange(4), 4)                               # already grouped and sorted: shards are consecutive blocks
    plan = D.GatherPlan(D.shard_tracks(wid, 4), 2, 4)
    assert plan.identity and plan.perm is None and plan.per == 4 and plan.n_total == 16
    wid = np.array([i % 5 for i in range(13)])                     # weight_id = i mod G, uneven shards
    shards = D.shard_tracks(wid, 3)
    plan = D.GatherPlan(shards, 0, 3)
    assert not plan.identity and plan.per == 5 and plan.perm.shape == (13,)
    # what the ranks would send, rank-major and padded, and what comes out in original order
    gathered = torch.zeros(3 * plan.per, 4, 4, dtype=torch.float64)
    for r, idx in enumerate(shards):
        for j, t in enumerate(idx):
            gathered[r * plan.per + j] = float(t)
    out = gathered.index_select(0, plan.perm)
    assert [float(out[t, 0, 0]) for t in range(13)] == [float(t) for t in range(13)]
    one = D.GatherPlan(D.shard_tracks(wid, 1), 0, 1)               # a single rank still has to undo the sort by weight id
    local = torch.stack([torch.full((4, 4), float(t), dtype=torch.float64) for t in one.shards[0]])
    for _ in range(2):
        full = one.gather(local)
        assert [float(full[t, 0, 0]) for t in range(13)] == [float(t) for t in range(13)]
