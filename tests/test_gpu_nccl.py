"""se3tn_allgather_poses over a raw NCCL communicator.  Needs two GPUs: skipped on a single-GPU box (the
torch.distributed path of the same exchange is covered on CPU with gloo in test_dist_cpu.py and by bench.py --gpus N)."""
import ctypes as C, glob, importlib, os, sys
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _find_nccl():
    import nvidia  # noqa: torch's bundled NCCL wheel
    for base in list(getattr(nvidia, '__path__', [])):
        hits = glob.glob(os.path.join(base, 'nccl', 'lib', 'libnccl.so*'))
        if hits:
            return sorted(hits)[0]
    return 'libnccl.so.2'


class _UniqueId(C.Structure):
    _fields_ = [('internal', C.c_byte * 128)]          # opaque; c_char would truncate at the first NUL when read back


def _worker(rank, world, uid_bytes, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    pkg = importlib.import_module('iros20-6d-pose-tracking_b200')
    torch.cuda.set_device(rank)
    nccl = C.CDLL(_find_nccl(), mode=C.RTLD_GLOBAL)
    uid = _UniqueId(); C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    eng = pkg.Engine(max_batch=4, device=rank)
    n = 3
    local = torch.from_numpy(pkg.synth.raw_poses(n, seed=100 + rank)).cuda(rank)
    out = eng.allgather_poses_nccl(comm.value, local, world_size=world)
    torch.cuda.synchronize(rank)
    q.put((rank, out.cpu().numpy()))
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_allgather_poses_over_raw_nccl_comm(synth):
    import torch.multiprocessing as mp
    nccl = C.CDLL(_find_nccl(), mode=C.RTLD_GLOBAL)
    uid = _UniqueId()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    world = 2
    ps = [ctx.Process(target=_worker, args=(r, world, C.string_at(C.byref(uid), 128), q)) for r in range(world)]
    for p in ps: p.start()
    res = dict(q.get(timeout=180) for _ in ps)
    for p in ps: p.join(60)
    want = np.concatenate([synth.raw_poses(3, seed=100 + r) for r in range(world)])
    for r in range(world):
        assert np.array_equal(res[r], want)
