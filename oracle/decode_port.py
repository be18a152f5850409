"""Torch-CPU port of the reference's decode tail  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement of the tensor part of reference `ObjectEncoder.decode3d` / `decode2d` (reference
vfa/data/encoder.py:230-305) with the encoder's attributes passed as arguments (`topk`, `grid_size`, `world_size`, the
class-average dimensions): sigmoid, 5 x 5 max-pool NMS (`self.maxpool`, encoder.py:48, :230-232), top-k, gathers, orientation
argmax.  Returns the gathered top-k lists BEFORE the `conf > cls_thresh` mask (plus the top-k cell indices), and the
reference's output dictionary after it.

Pinning: tests/golden/make_golden_decode.py calls the UNMODIFIED reference `decode3d` / `decode2d` (an `ObjectEncoder`
instance with the five attributes they read set by hand -- constructing one needs a dataset on disk) on seeded head maps and
stores its outputs in tests/golden/decode_case.npz; tests/test_oracle_golden.py holds this port to them exactly.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def topk_lists(pred, topk, grid_size, world_size, dim_mean=None):
    """pred: heatmap [1,1,L,W], loc_offset [1,L,W,2], dim_offset [1,L,W,3] / rotation [1,L,W,A] (or absent).
    -> dict of [1, topk] tensors: conf, cy, cx, (h, w, l, orient_idx), index."""
    heatmap, tytx = pred['heatmap'], pred['loc_offset']
    dtype = heatmap.dtype
    heatmap = torch.sigmoid(heatmap)
    mask = torch.eq(F.max_pool2d(heatmap, kernel_size=5, padding=2, stride=1), heatmap).to(dtype)      # encoder.py:230-232
    heatmap = (mask * heatmap).flatten(start_dim=2).transpose(1, 2)
    conf, _ = torch.max(heatmap, dim=-1)                                                               # :240
    L, W = pred['heatmap'].shape[2:]
    grid_y, grid_x = torch.meshgrid(torch.arange(L, dtype=dtype), torch.arange(W, dtype=dtype), indexing='ij')
    tytx = torch.sigmoid(tytx)
    cy = (grid_y[None] + tytx[..., 0]).flatten(start_dim=1) / grid_size[0] * world_size[0]             # :246
    cx = (grid_x[None] + tytx[..., 1]).flatten(start_dim=1) / grid_size[1] * world_size[1]
    lists = {'conf': conf, 'cy': cy, 'cx': cx}
    if pred.get('dim_offset') is not None:
        t = pred['dim_offset']
        for i, k in enumerate(('h', 'w', 'l')):                                                        # :250-252
            lists[k] = torch.exp(t[..., i]).flatten(start_dim=1) * dim_mean[i]
    if pred.get('rotation') is not None:
        _, idx = torch.max(torch.sigmoid(pred['rotation']), dim=-1)                                    # :254-256
        lists['orient_idx'] = idx.flatten(start_dim=1)
    _, index = torch.topk(conf, k=topk, dim=1)                                                         # :259
    out = {k: torch.gather(v, dim=1, index=index) for k, v in lists.items()}
    out['index'] = index
    return out


def decode3d(pred, cls_thresh, topk, grid_size, world_size, dim_mean):
    o = topk_lists(pred, topk, grid_size, world_size, dim_mean)
    m = o['conf'] > cls_thresh                                                                         # :264
    return {'conf': o['conf'][m],
            'location': torch.stack([o['cx'][m], o['cy'][m], torch.zeros_like(o['cy'][m])], dim=-1),
            'dimension': torch.stack([o['h'][m], o['w'][m], o['l'][m]], dim=-1),
            'rotation': torch.deg2rad(o['orient_idx'][m].to(torch.float32))}


def decode2d(pred, cls_thresh, topk, grid_size, world_size, wildtrack=False):
    o = topk_lists({'heatmap': pred['heatmap'], 'loc_offset': pred['loc_offset']}, topk, grid_size, world_size)
    m = o['conf'] > cls_thresh
    a, b = (o['cy'][m], o['cx'][m]) if wildtrack else (o['cx'][m], o['cy'][m])                        # :296-299
    return {'conf': o['conf'][m], 'location': torch.stack([a, b, torch.zeros_like(a)], dim=-1)}
