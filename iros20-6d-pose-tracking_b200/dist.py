"""Multi-GPU: tracks (object instances) are independent (reference predict.py:217-296 has no
cross-track data flow), so the track set is partitioned across ranks -- one process per GPU --
with NO collective on the data path.  The only exchange step is an all-gather of the updated
4x4 poses (128 B per track) when every rank wants the full pose set (NCCL over NVLink on GPUs;
the same code runs on gloo/CPU tensors in the unit tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_tracks(weight_ids, world_size):
    """Partition track indices over ranks: sort by weight id (one checkpoint per object class, F17),
    then cut into `world_size` contiguous, equally sized (+-1) slices so each rank touches as few
    weight sets as possible and all ranks carry the same number of tracks.
    Returns a list (len world_size) of int64 index arrays into the original track order."""
    weight_ids = np.asarray(weight_ids)
    order = np.argsort(weight_ids, kind='stable')
    n = len(order)
    base, extra = divmod(n, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append(order[start:start + cnt].astype(np.int64))
        start += cnt
    return out


def padded_count(n_tracks, world_size):
    return (n_tracks + world_size - 1) // world_size


class GatherPlan:
    """The pose exchange for one partition of the track set, prepared once: shard sizes, padding, and -- when the shards are
    not simply consecutive blocks of the original order (e.g. weight_id = i mod G) -- the permutation
    that puts the rank-major gathered buffer back into ORIGINAL track order, resident on the device.  gather() is then one
    all_gather_into_tensor plus at most one index_select; nothing is copied from the host per call (a per-call index
    upload is a synchronous copy: it blocks the host until the step in front of it has finished, and the next step's launch
    with it)."""
    def __init__(self, shards, rank, world_size, device=None):
        self.shards, self.rank, self.world_size = shards, rank, world_size
        self.sizes = [len(s) for s in shards]
        self.n_total = int(sum(self.sizes))
        self.per = max(self.sizes) if self.sizes else 0
        self.identity = (all(z == self.per for z in self.sizes)
                         and np.array_equal(np.concatenate(shards) if shards else np.zeros(0, np.int64), np.arange(self.n_total)))
        self.perm = None
        if not self.identity:
            pos = np.empty(self.n_total, dtype=np.int64)          # original track t sits at gathered[pos[t]]
            for r, idx in enumerate(shards):
                pos[idx] = r * self.per + np.arange(len(idx))
            self.perm = torch.from_numpy(pos)
            if device is not None:
                self.perm = self.perm.to(device)

    def gather(self, local_poses, group=None):
        """local_poses: (len(shards[rank]),4,4) on this rank's device -> (n_tracks,4,4) in original track order on every rank."""
        dev = local_poses.device
        if self.perm is not None and self.perm.device != dev:
            self.perm = self.perm.to(dev)
        if self.identity:
            # equal consecutive shards in original order: the gathered buffer IS the result (one NCCL call, no other kernel)
            out = torch.empty(self.n_total, 4, 4, dtype=local_poses.dtype, device=dev)
            if self.world_size == 1:
                out.copy_(local_poses)
            else:
                dist.all_gather_into_tensor(out.view(-1), local_poses.contiguous().view(-1), group=group)
            return out
        buf = local_poses.contiguous()
        if buf.shape[0] != self.per:                               # uneven shards are padded to the largest one
            buf = torch.zeros(self.per, 4, 4, dtype=local_poses.dtype, device=dev)
            buf[:local_poses.shape[0]] = local_poses
        gathered = torch.empty(self.world_size * self.per, 4, 4, dtype=local_poses.dtype, device=dev)
        if self.world_size == 1:
            gathered.copy_(buf)
        else:
            dist.all_gather_into_tensor(gathered.view(-1), buf.view(-1), group=group)
        return gathered.index_select(0, self.perm)


def all_gather_poses(local_poses, shards, rank, world_size, group=None):
    """local_poses: (len(shards[rank]),4,4) float64 on this rank's device.  Returns the full
    (n_tracks,4,4) tensor in ORIGINAL track order on every rank.  Uneven shards are padded to the
    largest shard so a single fixed-size all_gather_into_tensor suffices.  One-off convenience: callers that exchange
    poses every frame keep a GatherPlan (ShardedTracker does)."""
    return GatherPlan(shards, rank, world_size, local_poses.device).gather(local_poses, group=group)


class ShardedTracker:
    """Runs this rank's slice of a multi-object track set through an Engine and (optionally) gathers
    all poses.  `engine` needs every weight set referenced by this rank's slice loaded.

    overlap_gather (default): the pose all-gather is the only exchange step and nothing on this rank's data path needs
    its result (frame k+1's crop uses the rank's OWN updated poses), so it is issued on a side stream behind an event
    and the next step's kernels do not wait for it -- a 4 KB-per-rank NCCL call costs ~25 us of latency that would
    otherwise sit on the critical path of a 0.8 ms step.  Nor does a step wait for the PREVIOUS step's gather: an all-gather
    completes only when every rank has reached it, so waiting for it at the start of each step would put the ranks in lock
    step (one slow step anywhere stalls all of them); a step only waits for the gather issued TWO steps earlier, the last
    reader of the output set it is about to overwrite.  The gathered tensor of step k may be read on the compute stream
    after wait_gather(), which waits for every gather still in flight."""
    def __init__(self, engine, weight_ids, K, object_width, trans_normalizer, rot_normalizer,
                 rank=0, world_size=1, precision='bf16x3', overlap_gather=True):
        self.engine = engine
        self.rank, self.world_size = rank, world_size
        self.weight_ids = np.asarray(weight_ids, dtype=np.int32)
        self.shards = shard_tracks(self.weight_ids, world_size)
        self.mine = self.shards[rank]
        self.plan = GatherPlan(self.shards, rank, world_size, engine.device)
        # device-side index of this rank's tracks (indexing a CUDA tensor with the numpy array would stage a synchronous
        # host->device copy on every call and serialise host and GPU)
        self.mine_dev = torch.as_tensor(np.ascontiguousarray(self.mine), dtype=torch.long).to(engine.device)
        self.K = K
        self.tn, self.rn = trans_normalizer, rot_normalizer
        self.precision = precision
        dev = engine.device
        self.local_wids_host = np.ascontiguousarray(self.weight_ids[self.mine])      # sorted => contiguous runs
        self.local_wids_dev = torch.from_numpy(self.local_wids_host).to(dev)
        ow = np.broadcast_to(np.asarray(object_width, dtype=np.float64), self.weight_ids.shape)
        self.local_ow = torch.from_numpy(np.ascontiguousarray(ow[self.mine])).to(dev)
        self.overlap_gather = bool(overlap_gather) and world_size > 1
        self._comm_stream = torch.cuda.Stream(device=dev) if self.overlap_gather else None
        self._pending = []                         # events of the all-gathers still in flight, oldest first
        self._readers = [None, None]               # per rotating output set: the gather that last read it
        # two rotating sets of output tensors: stable device addresses keep the step's CUDA graph key stable (libse3tn replays one
        # graph per distinct set of pointers); a returned pose tensor stays valid until the step after the next one
        n = len(self.mine)
        self._outs = [(torch.empty(n, 4, 4, dtype=torch.float64, device=dev), torch.empty(n, 3, dtype=torch.float32, device=dev),
                       torch.empty(n, 3, dtype=torch.float32, device=dev)) for _ in range(2)]
        self._flip = 0

    def wait_gather(self):
        """Make the current stream wait for every overlapped all-gather still in flight (no-op when none is)."""
        cur = torch.cuda.current_stream(self.engine.device)
        for ev in self._pending:
            cur.wait_event(ev)
        self._pending = []

    def step(self, frame_rgb, frame_depth, local_poses, local_rgbA, local_depthA, gather=True):
        # two rotating output sets: the one written now was last read by the gather of the step before the previous one
        self._flip ^= 1
        if self._readers[self._flip] is not None:
            torch.cuda.current_stream(self.engine.device).wait_event(self._readers[self._flip])
            self._readers[self._flip] = None
        o = self._outs[self._flip]
        out, _, _ = self.engine.track_batch(frame_rgb, frame_depth, self.K, local_poses, self.local_ow,
                                            local_rgbA, local_depthA, self.tn, self.rn,
                                            weight_ids_host=self.local_wids_host, weight_ids_dev=self.local_wids_dev,
                                            precision=self.precision, out_poses=o[0], out_trans=o[1], out_rot=o[2])
        if not gather:
            return out, None
        if not self.overlap_gather:
            return out, self.plan.gather(out)
        cur = torch.cuda.current_stream(self.engine.device)
        ready = torch.cuda.Event(); ready.record(cur)
        self._comm_stream.wait_event(ready)
        with torch.cuda.stream(self._comm_stream):
            gathered = self.plan.gather(out)
            done = torch.cuda.Event(); done.record(self._comm_stream)
            self._readers[self._flip] = done
            self._pending = self._pending[-1:] + [done]          # anything older is ordered before these on the side stream
        gathered.record_stream(cur)                # ... and `gathered` will be read here after wait_gather()
        return out, gathered
