"""Per-unit timeline of the trunk launch at a small batch (SE3TN_TRACE=1): for every layer, when its units' dependencies were met, when
their MMAs ran and when their epilogues finished.   python scripts/trunk_units.py [precision] [n]"""
import importlib, os, sys
os.environ['SE3TN_TRACE'] = '1'; os.environ['SE3TN_GRAPH'] = '0'
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = pkg.Engine(max_batch=max(nb, 4)); eng.load_state_dict(synth.make_state_dict(0), 0)
A, B = synth.tensor_pairs(nb, seed=1); A = A.cuda(); B = B.cuda()
for _ in range(4): eng.forward(A, B, precision=prec)
torch.cuda.synchronize()
tr = eng.get_trace().astype(np.int64)
trunk = tr[8]                               # per-CTA stamps of the trunk launch
t0 = trunk[trunk[:, 0] > 0, 0].min()
units = tr[9:14].reshape(-1)[:2048 * 5].reshape(2048, 5)
used = units[:, 4] > 0
idx = np.nonzero(used)[0]
print('%s n=%d: %d work units; times in us from the first trunk CTA entry' % (prec, nb, len(idx)))
# layer boundaries: infer from unit counts (n * units_per_image * ksplit per layer) -- print in groups of equal size
per_layer = len(idx) // 6
for l in range(6):
    u = units[idx[l * per_layer:(l + 1) * per_layer]]
    f = lambda a: '%7.1f..%7.1f' % ((a.min() - t0) / 1e3, (a.max() - t0) / 1e3)
    print('layer %d: dep met %s | first A %s | last MMA commit %s | acc seen %s | epilogue done %s | mma %.1f us, epi %.1f us (medians)' % (
        l, f(u[:, 0]), f(u[:, 1]), f(u[:, 2]), f(u[:, 3]), f(u[:, 4]), np.median(u[:, 2] - u[:, 1]) / 1e3, np.median(u[:, 4] - u[:, 3]) / 1e3))
