"""CPU-only host logic: the weight packer (BN fold + OIHW -> K-major rows) checked by rebuilding
plain conv weights from the packed blob and running them through torch CPU against the oracle."""
import importlib
import numpy as np
import torch
import torch.nn.functional as F
import se3_oracle as O

W = importlib.import_module('iros20-6d-pose-tracking_b200.weights')

# (rows, ktot, kind) in blob order -- mirrors include/se3tn.h
LAYERS = [(64, 224, 'stem')] * 2 + [(64, 576, 64)] * 6 + [(256, 1152, 128), (256, 2304, 256), (256, 2304, 256),
          (1024, 2304, 256), (1024, 4608, 512), (1024, 4608, 512)]


def unpack(blob):
    off, out = 0, []
    for rows, ktot, kind in LAYERS:
        w = torch.from_numpy(blob[off:off + rows * ktot].reshape(rows, ktot)); off += rows * ktot
        b = torch.from_numpy(blob[off:off + rows]); off += rows
        if kind == 'stem':
            w4 = w.reshape(rows, 7, 8, 4)
            assert float(w4[:, :, 7, :].abs().max()) == 0.0          # the padded 8th filter column
            w4 = w4[:, :, :7, :].permute(0, 3, 1, 2).contiguous()    # (Co,4,7,7)
        else:
            w4 = w.reshape(rows, 3, 3, kind).permute(0, 3, 1, 2).contiguous()
        out.append((w4, b))
    fcw = torch.from_numpy(blob[off:off + 6 * 512].reshape(6, 512)); off += 6 * 512
    fcb = torch.from_numpy(blob[off:off + 6]); off += 6
    assert off == blob.size == W.BLOB_FLOATS
    return out, fcw, fcb


def run_packed(layers, fcw, fcb, A, B):
    def conv(x, i, stride, pad, groups=1):
        return F.conv2d(x, layers[i][0], layers[i][1], stride=stride, padding=pad, groups=groups)
    a = F.max_pool2d(F.selu(conv(A, 0, 2, 3)), 3, 2, 1)
    b = F.max_pool2d(F.selu(conv(B, 1, 2, 3)), 3, 2, 1)
    a = F.relu(conv(F.relu(conv(a, 2, 1, 1)), 3, 1, 1) + a)
    t = F.relu(conv(F.relu(conv(b, 4, 1, 1)), 5, 1, 1) + b)
    b = F.relu(conv(F.relu(conv(t, 6, 1, 1)), 7, 1, 1) + t)
    ab = F.selu(conv(torch.cat((a, b), 1), 8, 2, 1))
    ab = F.relu(conv(F.relu(conv(ab, 9, 1, 1)), 10, 1, 1) + ab)
    h = F.selu(conv(ab, 11, 2, 1))                                   # (N,1024,11,11): trans | rot
    h = F.relu(conv(F.relu(conv(h, 12, 1, 1, groups=2)), 13, 1, 1, groups=2) + h)
    p = h.mean((2, 3))
    trans = torch.tanh(F.linear(p[:, :512], fcw[:3], fcb[:3]))
    rot = torch.tanh(F.linear(p[:, 512:], fcw[3:], fcb[3:]))
    return trans, rot, ab


def test_packed_blob_reproduces_reference_forward(synth):
    sd = synth.make_state_dict(0)
    blob = W.pack_state_dict(sd)
    assert blob.dtype == np.float32 and blob.flags['C_CONTIGUOUS']
    layers, fcw, fcb = unpack(blob)
    A, B = synth.tensor_pairs(1, seed=3)
    with torch.no_grad():
        trans, rot, feat = run_packed(layers, fcw, fcb, A, B)
        ref = O.forward(sd, A, B)
    # BN folding moves roundings around: ~1e-6 relative, nowhere near the 1e-4/1e-3 parity gate
    assert torch.allclose(trans, ref['trans'], rtol=1e-4, atol=2e-6)
    assert torch.allclose(rot, ref['rot'], rtol=1e-4, atol=2e-6)
    assert torch.allclose(feat, ref['feature'], rtol=1e-3, atol=1e-4)


def test_packer_rejects_foreign_state_dict():
    import pytest
    with pytest.raises((KeyError, ValueError)):
        W.pack_state_dict({'foo.weight': torch.zeros(3)})
