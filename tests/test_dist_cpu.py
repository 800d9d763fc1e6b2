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
