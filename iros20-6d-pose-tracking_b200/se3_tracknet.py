"""Drop-in for the reference's se3_tracknet.py: same class name, constructor, load_state_dict /
cuda / eval / __call__ surface and output dict -- but forward() runs the hand-written sm_100a
kernels of libse3tn through the C ABI instead of torch.nn -> cuDNN.

Reference: se3_tracknet.py:52-112 (Se3TrackNet), network_modules.py:59-66,86-120.
Inference only: the training loss (se3_tracknet.py:114-121) and autograd are out of scope, so
train(True) raises.
"""
import torch
from .engine import Engine


class Se3TrackNet(torch.nn.Module):
    def __init__(self, image_size=174, max_batch=64, precision='bf16x3', engine=None, weight_id=0):
        super().__init__()
        self.rot_dim = 3
        self.image_size = image_size          # unused by the reference too (fully convolutional, F5)
        self.max_batch = max_batch
        self.precision = precision
        self.weight_id = weight_id
        self._engine = engine
        self._sd = None
        self._loaded = False

    # -- nn.Module surface the reference's callers use (predict.py:153-158) ---------------------
    def load_state_dict(self, state_dict, strict=True):
        missing = [k for k in ('convA1.0.weight', 'trans_out.0.weight', 'rot_out.0.bias') if k not in state_dict]
        if missing:
            raise RuntimeError('not a Se3TrackNet state_dict, missing keys: %s' % missing)
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self._loaded = False
        if self._engine is not None:
            self._upload()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *args, **kwargs):
        return dict(self._sd) if self._sd is not None else {}

    def cuda(self, device=None):
        if self._engine is None:
            self._engine = Engine(max_batch=self.max_batch, device=device)
        if self._sd is not None and not self._loaded:
            self._upload()
        return self

    def to(self, *args, **kwargs):
        dev = args[0] if args else kwargs.get('device')
        if dev is not None and torch.device(dev).type == 'cuda':
            return self.cuda(torch.device(dev).index)
        raise RuntimeError('Se3TrackNet (B200) only lives on a CUDA device; there is no CPU path')

    def train(self, mode=True):
        if mode:
            raise NotImplementedError('this is the inference hot path only (training is out of scope)')
        return super().train(False)

    @property
    def engine(self):
        if self._engine is None:
            self.cuda()
        return self._engine

    def _upload(self):
        self._engine.load_state_dict(self._sd, self.weight_id)
        self._loaded = True

    # -- se3_tracknet.py:81-112 -------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, A, B, return_feature=True):
        if self._sd is None:
            raise RuntimeError('load_state_dict() must be called before forward()')
        eng = self.engine
        if not self._loaded:
            self._upload()
        A = A.to(eng.device, torch.float32).contiguous()
        B = B.to(eng.device, torch.float32).contiguous()
        trans, rot, feat = eng.forward(A, B, weight_id=self.weight_id, precision=self.precision,
                                       want_feature=return_feature)
        out = {'trans': trans, 'rot': rot}
        if return_feature:
            out['feature'] = feat
        return out
