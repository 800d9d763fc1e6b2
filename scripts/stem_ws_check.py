"""Weights-stationary stem (SE3TN_STEM_WS) vs the resident-weight stem: pooled stem outputs P1A / P1B must agree to fp32
accumulation-order noise (both form hi*w_hi + lo*w_hi + hi*w_lo in fp32), final 6-vectors within the gate; then time both."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
pkg = importlib.import_module('iros20-6d-pose-tracking_b200'); synth = pkg.synth
nb = 64
sd = synth.make_state_dict(0)
A, B = synth.tensor_pairs(nb, seed=3); Ad, Bd = A.cuda(), B.cuda()

def decode(buf, n):      # P1 buffers: (n, 44*44, 64 ch) as [32 hi | 32 lo] chunks -> float32
    raw = buf.view(torch.uint8).view(n, 44 * 44, 2, 2, 32, 2).contiguous()         # chunk, (hi|lo), 32 x bf16
    w = raw.view(torch.bfloat16).view(n, 44 * 44, 2, 2, 32).float()
    return (w[:, :, :, 0] + w[:, :, :, 1]).reshape(n, 44 * 44, 64)

res = {}
for mode in ('0', '1', '3'):
    os.environ['SE3TN_STEM_WS'] = mode
    eng = pkg.Engine(max_batch=nb); eng.load_state_dict(sd, 0)
    try:
        t, r, _ = eng.forward(Ad, Bd, precision='bf16x3')
        torch.cuda.synchronize()
        p1a = decode(eng.debug_buffer(4, nb).clone(), nb); p1b = decode(eng.debug_buffer(5, nb).clone(), nb)
        for _ in range(3): eng.forward(Ad, Bd, precision='bf16x3')
        eng.set_profiling(True); acc = []
        for _ in range(10):
            eng.forward(Ad, Bd, precision='bf16x3'); acc.append(eng.get_profile())
        eng.set_profiling(False)
        m = np.mean(np.stack(acc), 0)
        res[mode] = (torch.cat((t, r), 1).cpu(), p1a.cpu(), p1b.cpu(), m[:2])
        print('mode %s: stems %.4f %.4f ms; 6-vector[0] %s' % (mode, m[0], m[1], res[mode][0][0].numpy()))
    except Exception as e:
        print('mode %s failed: %s' % (mode, e))
    eng.close()
if '0' in res:
    import se3_oracle as O
    ref = O.forward(sd, A[:8], B[:8]); ref6 = torch.cat((ref['trans'], ref['rot']), 1)
    for mode in ('1',):
        if mode not in res: continue
        d6 = (res[mode][0] - res['0'][0]).abs().max().item()
        da = (res[mode][1] - res['0'][1]).abs().max().item(); db = (res[mode][2] - res['0'][2]).abs().max().item()
        scale = res['0'][1].abs().max().item()
        err = ((res[mode][0][:8] - ref6).abs() / (1e-4 + 1e-3 * ref6.abs())).max().item()
        print('mode %s vs resident stem: max |d P1A| %.3e, |d P1B| %.3e (values up to %.2f), max |d 6-vector| %.3e, err/tol vs oracle %.3f' % (mode, da, db, scale, d6, err))
