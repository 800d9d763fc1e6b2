"""Device-side timeline of one step (SE3TN_TRACE=1): where each conv launch spends its time, with PDL on and no events between kernels.
   python scripts/trace_timeline.py [precision] [n]"""
import importlib, os, sys
os.environ['SE3TN_TRACE'] = '1'
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = pkg.Engine(max_batch=nb); eng.load_state_dict(synth.make_state_dict(0), 0)
mean, std = synth.default_mean_std(); eng.set_stats(mean, std, 0)
rgb, depth = synth.raw_frame(0); poses = synth.raw_poses(nb, seed=0); rgbA, depthA = synth.rendered_views(nb, poses, seed=0)
d = [torch.from_numpy(x).cuda() for x in (rgb, depth, poses, rgbA, depthA)]
ow = torch.full((nb,), 200.0, dtype=torch.float64, device='cuda')
def step():
    return eng.track_batch(d[0], d[1], synth.CAMERA_K, d[2], ow, d[3], d[4], 0.03, 5 * np.pi / 180, precision=prec)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
tr = eng.get_trace().astype(np.int64)
print('%s n=%d step %.1f us' % (prec, nb, e0.elapsed_time(e1) * 1e3))
names = ['stemA', 'stemB', 'A2c1', 'A2c2', 'B2c1', 'B2c2', 'B3c1', 'B3c2', 'trunk', '-', '-', '-', '-', '-']   # slot 8 = conv_trunk_kernel (six layers, one launch)
t_first = None
prev_end = None
print('layer  ctas | start(first,last)  end(first,last) | per-CTA medians: setup  wgt  firstA  mma_span  acc0  epi_tail  exit | span  gap_prev')
for l in range(9):                          # slots 9..13 hold the trunk's per-unit stamps (scripts/trunk_units.py)
    t = tr[l]; used = t[:, 0] > 0
    t = t[used]
    if not len(t): continue
    s0 = t[:, 0]; ex = (t[:, 7] & ~0xff)
    if t_first is None: t_first = s0.min()
    med = lambda a: float(np.median(a)) / 1e3
    print('%-6s %4d | %7.1f %7.1f   %7.1f %7.1f | %5.1f %5.1f %5.1f %7.1f %6.1f %6.1f %5.1f | %6.1f %6.1f' % (
        names[l], len(t), (s0.min() - t_first) / 1e3, (s0.max() - t_first) / 1e3, (ex.min() - t_first) / 1e3, (ex.max() - t_first) / 1e3,
        med(t[:, 1] - s0), med(t[:, 2] - s0), med(t[:, 3] - s0), med(t[:, 4] - t[:, 3]), med(t[:, 5] - s0), med(t[:, 6] - t[:, 4]), med(ex - t[:, 6]),
        (ex.max() - s0.min()) / 1e3, ((s0.min() - prev_end) / 1e3) if prev_end is not None else 0.0))
    prev_end = ex.max()
print('conv stack first entry -> last exit: %.1f us' % ((prev_end - t_first) / 1e3))
