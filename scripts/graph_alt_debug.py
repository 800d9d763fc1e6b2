"""Debug: per-step device times when the precision changes under graph replay (bench.py's alt_precisions leg)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
dist_mod = importlib.import_module('iros20-6d-pose-tracking_b200.dist')
nb = 64
eng = pkg.Engine(max_batch=nb); eng.load_state_dict(synth.make_state_dict(0), 0)
mean, std = synth.default_mean_std(); eng.set_stats(mean, std, 0)
sets = []
for k in range(16):
    rgb, depth = synth.raw_frame(k); poses = synth.raw_poses(nb, seed=k); rgbA, depthA = synth.rendered_views(nb, poses, seed=k)
    sets.append([torch.from_numpy(x).cuda() for x in (rgb, depth, poses, rgbA, depthA)])
trk = dist_mod.ShardedTracker(eng, np.zeros(nb, np.int32), synth.CAMERA_K, 200.0, 0.03, 5 * np.pi / 180, 0, 1, 'bf16x3')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(prec, n):
    trk.precision = prec
    out = []
    for k in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); d = sets[k % 16]; trk.step(d[0], d[1], d[2], d[3], d[4], gather=False); e1.record()
        host = (time.perf_counter() - t0) * 1e3
        out.append((e0, e1, host, eng.last_step_was_graph()))
    torch.cuda.synchronize()
    print(prec, ' '.join('%.2f/%.2f%s' % (a.elapsed_time(b), h, 'g' if g else 'p') for a, b, h, g in out))
run('bf16x3', 25); run('tf32', 23); run('bf16', 23); run('tf32', 23); run('bf16x3', 23)
