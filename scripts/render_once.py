"""Render 64 tracks of the 20k-face synthetic model a few times (profiling target for ncu)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
n = 64
eng = pkg.Engine(max_batch=n)
eng.set_mesh(synth.mesh(int(sys.argv[1]) if len(sys.argv) > 1 else 5, seed=0), 0)
poses = torch.from_numpy(synth.raw_poses(n, seed=0)).cuda(); ow = torch.full((n,), 200.0, dtype=torch.float64, device='cuda')
for _ in range(4):
    rgb, dep = eng.render(synth.CAMERA_K, poses, ow)
torch.cuda.synchronize()
print('foreground share', float((dep.to(torch.int32) > 0).float().mean()))
