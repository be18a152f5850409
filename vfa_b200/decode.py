"""Decode tail of the detector on the GPU: the heads' maps -> detections.

`decode_topk` replaces the tensor part of reference `ObjectEncoder.decode3d` / `decode2d` (reference
vfa/data/encoder.py:230-305: sigmoid, 5 x 5 max-pool NMS, top-k, gathers, orientation argmax -- ~15 full-map ATen launches)
with two small kernels behind `vfa_decode_topk` (include/vfa_b200.h); `decode3d` / `decode2d` return the reference's
dictionaries (`conf`, `location`, `dimension`, `rotation`) after its `conf > cls_thresh` mask.  The reference is
structurally batch-1; here every frame of the batch is decoded in the same launch.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .vfa_op import _require_cuda, _stream


def _head(t: torch.Tensor | None, L: int, W: int, channels_last_dim: bool):
    """(pointer, strides) of a head given as [B, L, W, C] (the reference's permuted views) or [B, C, L, W]."""
    if t is None:
        return None, (0, 0, 0), None
    if not channels_last_dim:
        t = t.permute(0, 2, 3, 1)
    if t.dtype != torch.float32:
        raise TypeError('head maps must be float32')
    if t.shape[1] != L or t.shape[2] != W:
        raise ValueError(f'head map is {tuple(t.shape)}, heatmap is {L} x {W}')
    if t.stride(1) != W * t.stride(2):                   # rows not evenly spaced: one copy
        t = t.contiguous()
    return t.data_ptr(), (t.stride(0), t.stride(3), t.stride(2)), t


def decode_topk(pred: dict, topk: int, grid_size, world_size, dim_mean=(1.0, 1.0, 1.0), heads_last: bool = True):
    """pred: 'heatmap' [B,1,L,W] logits, 'loc_offset' [B,L,W,2], optional 'dim_offset' [B,L,W,3] and 'rotation'
    [B,L,W,A] (heads_last=False: heads are [B,C,L,W]).  Returns (vals [B,topk,7] = conf, cy, cx, h, w, l, orientation bin;
    cells [B,topk] int32, -1 where a frame has fewer NMS survivors), sorted by confidence."""
    heat = pred['heatmap']
    _require_cuda(heat, 'heatmap')
    if heat.dim() != 4 or heat.shape[1] != 1:
        raise ValueError(f'heatmap must be [B,1,L,W], got {tuple(heat.shape)}')
    heat = heat.detach().to(torch.float32).contiguous()
    B, _, L, W = heat.shape
    d = _lib.Decode()
    d.batch, d.grid_l, d.grid_w, d.topk = B, L, W, int(topk)
    keep = []
    d.heatmap = heat.data_ptr()
    for name, field, stride in (('loc_offset', 'loc_offset', d.loc_stride), ('dim_offset', 'dim_offset', d.dim_stride),
                                ('rotation', 'rotation', d.rot_stride)):
        t = pred.get(name)
        ptr, st, ref = _head(None if t is None else t.detach(), L, W, heads_last)
        setattr(d, field, ptr)
        for i in range(3):
            stride[i] = st[i]
        keep.append(ref)
    if d.loc_offset is None:
        raise ValueError("pred['loc_offset'] is required")
    d.n_angles = 0 if pred.get('rotation') is None else (keep[2].shape[3])
    for i in range(2):
        d.grid_size[i], d.world_size[i] = float(grid_size[i]), float(world_size[i])
    for i in range(3):
        d.dim_mean[i] = float(dim_mean[i])
    dev = heat.device
    vals = torch.empty(B, int(topk), 7, dtype=torch.float32, device=dev)
    cells = torch.empty(B, int(topk), dtype=torch.int32, device=dev)
    L_ = _lib.lib()
    ws = torch.empty(L_.vfa_decode_workspace_bytes(B), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L_.vfa_decode_topk(C.byref(d), vals.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return vals, cells


def decode3d(pred: dict, cls_thresh: float, topk: int, grid_size, world_size, dim_mean):
    """Reference `ObjectEncoder.decode3d` (encoder.py:234-273) for frame 0 of the batch (the reference is batch-1)."""
    vals, _ = decode_topk(pred, topk, grid_size, world_size, dim_mean)
    v = vals[0]
    m = v[:, 0] > cls_thresh
    v = v[m]
    return {'conf': v[:, 0],
            'location': torch.stack([v[:, 2], v[:, 1], torch.zeros_like(v[:, 1])], dim=-1),          # x y z
            'dimension': v[:, 3:6],
            'rotation': torch.deg2rad(v[:, 6])}


def decode2d(pred: dict, cls_thresh: float, topk: int, grid_size, world_size, wildtrack: bool = False):
    """Reference `ObjectEncoder.decode2d` (encoder.py:275-305); Wildtrack swaps the two ground axes (:296-299)."""
    vals, _ = decode_topk({'heatmap': pred['heatmap'], 'loc_offset': pred['loc_offset']}, topk, grid_size, world_size)
    v = vals[0]
    v = v[v[:, 0] > cls_thresh]
    first, second = (v[:, 1], v[:, 2]) if wildtrack else (v[:, 2], v[:, 1])
    return {'conf': v[:, 0], 'location': torch.stack([first, second, torch.zeros_like(first)], dim=-1)}


__all__ = ['decode_topk', 'decode3d', 'decode2d']
_ = math
