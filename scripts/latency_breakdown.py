"""One object, one frame at a time (the reference's calling pattern, predict.py:416): where do the microseconds go?
   (1) per-kernel device times at n = 1 (library profiling mode: plain launches bracketed by events)
   (2) device time of the graph-replayed step
   (3) wall clock of Tracker.on_track with numpy in / numpy out, and of its host-side pieces
Run on a GPU box:  python scripts/latency_breakdown.py"""
import importlib, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200')
synth = importlib.import_module('iros20-6d-pose-tracking_b200.synth')


def main():
    dev = torch.device('cuda:0')
    eng = pkg.Engine(max_batch=8, device=0)
    sd = synth.make_state_dict(0)
    mean, std = synth.default_mean_std()
    eng.load_state_dict(sd, 0); eng.set_stats(mean, std, 0)
    rgb, depth = synth.raw_frame(3)
    TN, RN = 0.03, 5 * np.pi / 180
    for n in (1, 2, 4):
        poses = synth.raw_poses(n, seed=3)
        rgbA, depthA = synth.rendered_views(n, poses, seed=3)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d = dict(rgb=t(rgb), depth=t(depth), poses=t(poses), rgbA=t(rgbA), depthA=t(depthA))
        ow = torch.full((n,), 200.0, dtype=torch.float64, device=dev)
        run = lambda: eng.track_batch(d['rgb'], d['depth'], synth.CAMERA_K, d['poses'], ow, d['rgbA'], d['depthA'], TN, RN)
        for _ in range(5): run()
        eng.set_profiling(True)
        acc = []
        for _ in range(20):
            run(); acc.append(eng.get_profile())
        eng.set_profiling(False)
        p = np.mean(acc, 0) * 1e3
        print('n=%d per-kernel us (plain launches): preprocess %.1f | stems %.1f %.1f | 64-ch %s | trunk %.1f | head+pose %.1f | sum %.1f'
              % (n, p[17], p[0], p[1], ' '.join('%.1f' % v for v in p[2:8]), p[8], p[16], p[:21].sum()))
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): run()
        e1.record(); torch.cuda.synchronize()
        print('n=%d graph-replayed step, back to back: %.1f us/step (graph: %s)' % (n, e0.elapsed_time(e1) / 200 * 1e3, eng.last_step_was_graph()))

    info = {'resolution': 176, 'boundingbox': 10, 'object_width': 200.0,
            'camera': {'focalX': synth.CAMERA_K[0, 0], 'focalY': synth.CAMERA_K[1, 1], 'centerX': synth.CAMERA_K[0, 2],
                       'centerY': synth.CAMERA_K[1, 2], 'height': 480, 'width': 640}}
    trk = pkg.Tracker(info, mean, std, {'state_dict': sd}, model_path=None, engine=eng)
    poses = synth.raw_poses(1, seed=3)
    rgbA, depthA = synth.rendered_views(1, poses, seed=3)
    p1, a1, d1 = poses[0], rgbA[0], depthA[0]
    for _ in range(20): trk.on_track(p1, rgb, depth, rgbA=a1, depthA=d1)
    reps = 300
    t0 = time.perf_counter()
    for _ in range(reps): trk.on_track(p1, rgb, depth, rgbA=a1, depthA=d1)
    print('Tracker.on_track numpy->numpy: %.1f us/frame' % ((time.perf_counter() - t0) / reps * 1e6))
    # host pieces: the same call with CUDA tensors in (no uploads, no read-back, no sync) = pure launch overhead
    dp, dr, dd, da, dda = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (poses, rgb, depth, rgbA, depthA))
    for _ in range(20): trk.on_track_batch(dp, dr, dd, da, dda)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): trk.on_track_batch(dp, dr, dd, da, dda)
    host = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()
    print('on_track_batch with CUDA tensors (host time to enqueue, no sync): %.1f us/call; incl. drain %.1f us/call' % (host, (time.perf_counter() - t0) / reps * 1e6))
    try:
        import cProfile, pstats, io
        pr = cProfile.Profile(); pr.enable()
        for _ in range(200): trk.on_track(p1, rgb, depth, rgbA=a1, depthA=d1)
        pr.disable()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22)
        print('\n'.join(l[:150] for l in s.getvalue().splitlines()[:45]))
    except Exception as e:                                      # the profile is a convenience
        print('cProfile failed:', e)
    eng.close()


if __name__ == '__main__':
    main()
