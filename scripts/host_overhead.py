"""Host-side cost of one Tracker.on_track_batch call (pinned host tensors in): wall time until the call returns, GPU not waited for."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
nb = 64
mean, std = synth.default_mean_std()
info = {'resolution': 176, 'boundingbox': 10, 'object_width': 200.0,
        'camera': {'focalX': synth.CAMERA_K[0, 0], 'focalY': synth.CAMERA_K[1, 1], 'centerX': synth.CAMERA_K[0, 2], 'centerY': synth.CAMERA_K[1, 2], 'height': 480, 'width': 640}}
trk = pkg.Tracker(info, mean, std, {'state_dict': synth.make_state_dict(0)}, model_path=None, max_batch=nb)
rgb, depth = synth.raw_frame(0); poses = synth.raw_poses(nb, seed=0); rgbA, depthA = synth.rendered_views(nb, poses, seed=0)
h = [torch.from_numpy(x).pin_memory() for x in (poses, rgb, depth, rgbA, depthA)]
for _ in range(5): trk.on_track_batch(*h)
torch.cuda.synchronize()
import cProfile, pstats
K = 200
t0 = time.perf_counter()
for _ in range(K): trk.on_track_batch(*h)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host %.1f us/call, total %.1f us/call' % ((t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(K): trk.on_track_batch(*h)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
