"""Generate tests/golden/decode_case.npz from the UNMODIFIED reference decode (run in the build container only).

    python tests/golden/make_golden_decode.py

`ObjectEncoder.__init__` needs a dataset on disk, `decode3d` / `decode2d` (reference vfa/data/encoder.py:234-305) read only
`maxpool`, `topk`, `grid_size`, `world_size`, `classname`, `dataset.classAverage.get_mean` and `dataset.base.__name__`: an
instance is made with `__new__` and exactly those attributes.  matplotlib is stubbed (imported at encoder.py:10, unused here).
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
for _n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec'):
    sys.modules.setdefault(_n, types.ModuleType(_n))
sys.path.insert(0, '/root/reference')
import torch.nn as nn                                  # noqa: E402
from vfa.data.encoder import ObjectEncoder             # noqa: E402  (the reference)

from decode_case_inputs import CLS_THRESH, DIM_MEAN, GRID_SIZE, TOPK, WORLD_SIZE, case_pred   # noqa: E402

out = {}
for seed in (0, 1):
    pred = case_pred(seed)
    for name, base in (('3d', 'MultiviewC'), ('2d', 'MultiviewX'), ('2dw', 'Wildtrack')):
        enc = ObjectEncoder.__new__(ObjectEncoder)
        enc.maxpool = nn.MaxPool2d(kernel_size=5, padding=2, stride=1)
        enc.topk = TOPK
        enc.grid_size, enc.world_size = np.array(GRID_SIZE), np.array(WORLD_SIZE)
        enc.classname = ['person']
        enc.dataset = SimpleNamespace(classAverage=SimpleNamespace(get_mean=lambda c: DIM_MEAN),
                                      base=SimpleNamespace(__name__=base))
        with torch.no_grad():
            d = enc.decode3d(pred, CLS_THRESH) if name == '3d' else enc.decode2d(pred, CLS_THRESH)
        for k, v in d.items():
            out[f's{seed}/{name}/{k}'] = v.numpy()
np.savez_compressed(os.path.join(HERE, 'decode_case.npz'), **out)
print({k: v.shape for k, v in out.items()})
