"""A/B of the frame upload variants for the one-object numpy calling pattern (wall clock per frame)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
sd = synth.make_state_dict(0); mean, std = synth.default_mean_std(); K = synth.CAMERA_K
info = {'resolution': 176, 'boundingbox': 10, 'object_width': 200.0, 'camera': {'focalX': K[0, 0], 'focalY': K[1, 1], 'centerX': K[0, 2], 'centerY': K[1, 2], 'height': 480, 'width': 640}}
trk = pkg.Tracker(info, mean, std, {'state_dict': sd}, model_path=None, max_batch=4)
rgb, depth = synth.raw_frame(0); poses = synth.raw_poses(4, seed=0); rgbA, depthA = synth.rendered_views(4, poses, seed=0)
for rep in range(2):
    for mode in ('off', 'pageable', 'pinned'):
        os.environ['SE3TN_WINDOW_UPLOAD'] = mode
        for _ in range(20): trk.on_track(poses[0], rgb, depth, rgbA=rgbA[0], depthA=depthA[0])
        t0 = time.perf_counter(); n = 300
        for _ in range(n): trk.on_track(poses[0], rgb, depth, rgbA=rgbA[0], depthA=depthA[0])
        print('upload %-8s: %.3f ms per frame' % (mode, (time.perf_counter() - t0) / n * 1e3))
import cProfile, pstats
os.environ['SE3TN_WINDOW_UPLOAD'] = 'off'
pr = cProfile.Profile(); pr.enable()
for _ in range(200): trk.on_track(poses[0], rgb, depth, rgbA=rgbA[0], depthA=depthA[0])
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
