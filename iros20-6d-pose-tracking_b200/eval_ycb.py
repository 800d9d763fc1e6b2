"""Drop-in for the scoring function of the reference's eval_ycb.py: VOCap (reference eval_ycb.py:45-64), the
area under the accuracy-vs-threshold curve below 0.1 m, computed on the GPU (sort + one reduction).
The directory walking / YCB-Video file layout of eval_one_class (eval_ycb.py:67-119) needs the dataset and is
out of scope; it only calls Utils.add / Utils.adi / VOCap, which all exist here."""
import numpy as np
import torch
from . import Utils as U


def VOCap(rec):
    eng = U._eng()
    errs = torch.from_numpy(np.ascontiguousarray(np.asarray(rec, dtype=np.float64).reshape(-1))).to(eng.device)
    return eng.vocap(errs)
