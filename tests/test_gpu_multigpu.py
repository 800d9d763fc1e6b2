"""Sharded multi-GPU run == single-GPU run, per track, bit for bit (SURVEY.md section 4 tier 3 / section 8e): tracks are
partitioned over ranks with no collective on the data path, so rank r's poses must be exactly what one GPU computes for
the same tracks.  Needs two GPUs (skipped otherwise; `gpurun --gpus 2`)."""
import importlib, os, socket, sys
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N_TRACKS = 24
TN, RN = 0.03, 5 * np.pi / 180


def _inputs(synth):
    rgb, depth = synth.raw_frame(31)
    poses = synth.raw_poses(N_TRACKS, seed=31)
    rgbA, depthA = synth.rendered_views(N_TRACKS, poses, seed=31)
    wids = (np.arange(N_TRACKS) % 2).astype(np.int32)          # interleaved ids: shard_tracks has to regroup them
    return rgb, depth, poses, rgbA, depthA, wids


def _engine(pkg, synth, device, max_batch):
    eng = pkg.Engine(max_batch=max_batch, device=device)
    mean, std = synth.default_mean_std()
    for wid in (0, 1):
        eng.load_state_dict(synth.make_state_dict(wid), wid); eng.set_stats(mean, std, wid)
    return eng


def _worker(rank, world, port, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    pkg = importlib.import_module('iros20-6d-pose-tracking_b200')
    dmod = importlib.import_module('iros20-6d-pose-tracking_b200.dist')
    synth = pkg.synth
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world, device_id=dev)
    rgb, depth, poses, rgbA, depthA, wids = _inputs(synth)
    eng = _engine(pkg, synth, rank, N_TRACKS)
    trk = dmod.ShardedTracker(eng, wids, synth.CAMERA_K, 200.0, TN, RN, rank, world, 'bf16x3')
    mine = trk.mine
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = []
    for _ in range(2):                                           # two steps: the overlapped gather of step 1 is waited for by step 2
        local, gathered = trk.step(t(rgb), t(depth), t(poses[mine]), t(rgbA[mine]), t(depthA[mine]), gather=True)
        trk.wait_gather()
        torch.cuda.synchronize(dev)
        res.append((local.cpu().numpy(), gathered.cpu().numpy()))
    q.put((rank, mine, res))
    dist.destroy_process_group()
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sharded_tracks_equal_single_gpu_bit_for_bit(pkg, synth):
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0)); port = sk.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    got = {}
    for _ in ps:
        rank, mine, res = q.get(timeout=300)
        got[rank] = (mine, res)
    for p in ps: p.join(60)
    # the same 24 tracks on ONE GPU, in the original order, per-track weight ids in one launch sequence
    rgb, depth, poses, rgbA, depthA, wids = _inputs(synth)
    eng = _engine(pkg, synth, 0, N_TRACKS)
    try:
        dev = eng.device
        order = np.argsort(wids, kind='stable')                  # the library wants equal ids contiguous for speed, any order is correct
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        single, _, _ = eng.track_batch(t(rgb), t(depth), synth.CAMERA_K, t(poses[order]), torch.full((N_TRACKS,), 200.0, dtype=torch.float64, device=dev),
                                       t(rgbA[order]), t(depthA[order]), TN, RN, weight_ids_host=wids[order], precision='bf16x3')
        ref = np.empty((N_TRACKS, 4, 4)); ref[order] = single.cpu().numpy()
    finally:
        eng.close()
    for rank in range(world):
        mine, res = got[rank]
        for local, gathered in res:
            assert np.array_equal(local, ref[mine]), 'rank %d: sharded poses differ from the single-GPU run' % rank
            assert np.array_equal(gathered, ref), 'rank %d: gathered pose set differs' % rank
