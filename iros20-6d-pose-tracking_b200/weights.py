"""Weight packer: reference-format state_dict (123 keys, reference problems.py:142-151 /
predict.py:151-155) -> the flat float32 blob se3tn_load_weights takes (layout in include/se3tn.h).

Eval-mode BatchNorm (eps 1e-5, reference network_modules.py:54,64,98,101) is folded into the
preceding conv in float64 and rounded once to float32:
    w' = w * gamma / sqrt(var + eps)        b' = (b - mean) * gamma / sqrt(var + eps) + beta
Conv weights go from OIHW to K-major rows W[co][tap*Cin + c] (tap = r*3+s); the 7x7 stem becomes
W[co][r*32 + s*4 + c] with a zero 8th column so one filter row is one 128-byte K chunk.
"""
import numpy as np
import torch

BN_EPS = 1e-5
BLOB_FLOATS = 13528326


def _fold(sd, conv, bn):
    w = sd[conv + '.weight'].detach().cpu().double()
    b = sd[conv + '.bias'].detach().cpu().double()
    g = sd[bn + '.weight'].detach().cpu().double()
    beta = sd[bn + '.bias'].detach().cpu().double()
    mu = sd[bn + '.running_mean'].detach().cpu().double()
    var = sd[bn + '.running_var'].detach().cpu().double()
    s = g / torch.sqrt(var + BN_EPS)
    return w * s[:, None, None, None], (b - mu) * s + beta


def _rows3x3(w):                       # (Co,Ci,3,3) -> (Co, 9*Ci), k = (r*3+s)*Ci + c
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _rows_stem(w):                     # (Co,4,7,7) -> (Co, 7*32), k = r*32 + s*4 + c, s=7 zero
    co = w.shape[0]
    t = torch.zeros(co, 7, 8, 4, dtype=w.dtype)
    t[:, :, :7, :] = w.permute(0, 2, 3, 1)
    return t.reshape(co, 7 * 32)


def pack_state_dict(sd):
    """-> np.float32 array of BLOB_FLOATS values, C-contiguous."""
    parts = []

    def add(wrows, b):
        parts.append(wrows.float().contiguous().reshape(-1))
        parts.append(b.float().contiguous().reshape(-1))

    for name in ('convA1', 'convB1'):
        w, b = _fold(sd, name + '.0', name + '.1')
        add(_rows_stem(w), b)
    for name in ('convA2', 'convB2', 'convB3'):
        for cv, bn in (('conv1', 'bn1'), ('conv2', 'bn2')):
            w, b = _fold(sd, '%s.%s' % (name, cv), '%s.%s' % (name, bn))
            add(_rows3x3(w), b)
    w, b = _fold(sd, 'convAB1.0', 'convAB1.1'); add(_rows3x3(w), b)
    for cv, bn in (('conv1', 'bn1'), ('conv2', 'bn2')):
        w, b = _fold(sd, 'convAB2.' + cv, 'convAB2.' + bn); add(_rows3x3(w), b)
    wt, bt = _fold(sd, 'trans_conv1.0', 'trans_conv1.1')
    wr, br = _fold(sd, 'rot_conv1.0', 'rot_conv1.1')
    add(torch.cat([_rows3x3(wt), _rows3x3(wr)], 0), torch.cat([bt, br]))
    for cv, bn in (('conv1', 'bn1'), ('conv2', 'bn2')):
        wt, bt = _fold(sd, 'trans_conv2.' + cv, 'trans_conv2.' + bn)
        wr, br = _fold(sd, 'rot_conv2.' + cv, 'rot_conv2.' + bn)
        add(torch.cat([_rows3x3(wt), _rows3x3(wr)], 0), torch.cat([bt, br]))
    parts.append(torch.cat([sd['trans_out.0.weight'].detach().cpu(), sd['rot_out.0.weight'].detach().cpu()], 0).float().reshape(-1))
    parts.append(torch.cat([sd['trans_out.0.bias'].detach().cpu(), sd['rot_out.0.bias'].detach().cpu()], 0).float().reshape(-1))
    blob = torch.cat(parts).numpy()
    if blob.size != BLOB_FLOATS:
        raise ValueError('packed %d floats, expected %d (is this a Se3TrackNet state_dict?)' % (blob.size, BLOB_FLOATS))
    return np.ascontiguousarray(blob, dtype=np.float32)
