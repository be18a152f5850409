"""Caller-side glue for the reference network.

`aggregate_cameras` replaces the per-camera / per-scale loop of reference `VFANet.forward`
(reference vfa/model/vfanet.py:64-82) for any module that carries the reference's attribute names
(`lat8/16/32`, `bn8/16/32`, `vfa8/16/32` -- the latter being `vfa_b200.VFA` modules): the three lateral
1x1 conv + GroupNorm + ReLU stages run ONCE over all cameras (exact: GroupNorm statistics are per sample,
reference vfanet.py:72-74) in channels-last memory format, and one fused kernel launch aggregates all
cameras and scales.  INTEGRATION.md shows the two-line edit of the reference's forward.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .vfa_op import aggregate, build_table


def head_flags(flags: int, channels: int) -> int:
    """Flags of the aggregation when its consumer is a convolutional head (reference vfanet.py:131-139): with the C = 256
    feature-side forward the BEV map is emitted [B, L, W, C] and returned as a [B, C, L, W] tensor in torch.channels_last
    -- cuDNN reads it without a permute, the pooling epilogue writes 1 KB rows instead of a stride-L*W scatter, and the
    backward reads the heads' channels-last cotangent in place."""
    if channels == 256 and not (int(flags) & (_lib.FLAG_GRID_SIDE | _lib.FLAG_FORCE_SIMT)):
        return int(flags) | _lib.FLAG_OUT_NHWC
    return int(flags)


def lateral_features(model, feats8, feats16, feats32):
    """relu(GN(conv1x1(.))) for all cameras at once, emitted channels-last (zero-copy hand-off to the gather)."""
    outs = []
    for conv, norm, f in ((model.lat8, model.bn8, feats8), (model.lat16, model.bn16, feats16),
                          (model.lat32, model.bn32, feats32)):
        x = f.contiguous(memory_format=torch.channels_last)
        outs.append(F.relu(norm(conv(x))))
    return outs


def aggregate_cameras(model, feats8, feats16, feats32, calibs, grid, crange=(-1.0, 0.95), batch: int = 1):
    """feats* [B*V, C_s, fH_s, fW_s] backbone maps (camera-major within a frame, as reference utils.py:43 collates),
    calibs [V,3,4], grid [1,L,W,3] or [L,W,3]  ->  ortho [B, 256, L, W] = sum over cameras and scales (vfanet.py:79-82)."""
    lats = lateral_features(model, feats8, feats16, feats32)
    N = lats[0].shape[0]
    if N % batch != 0:
        raise ValueError(f'{N} feature maps do not split into {batch} frames')
    V = N // batch
    calibs = calibs.reshape(-1, 3, 4)
    if calibs.shape[0] != V:
        raise ValueError(f'{calibs.shape[0]} calibrations for {V} cameras')
    L, W = grid.shape[-3], grid.shape[-2]
    vfas = (model.vfa8, model.vfa16, model.vfa32)
    geom = vfas[0].geometry((L, W), crange)
    table = build_table(geom, calibs, grid)
    feats = [x.reshape(batch, V, *x.shape[1:]) for x in lats]
    return aggregate(feats, table, [m.collapse.weight for m in vfas], [m.collapse.bias for m in vfas],
                     flags=head_flags(vfas[0].flags, lats[0].shape[1]))


class MultiScaleVFA(torch.nn.Module):
    """The whole aggregation stage of the reference network as ONE module with the batched multi-view signature of
    BASELINE.json's north_star:  forward(feats, calibs, grid) with feats = list of per-scale `[B, V, C, fH, fW]` lateral
    maps, calibs `[V, 3, 4]`, grid `[L, W, 3]` or `[1, L, W, 3]`  ->  `[B, C, L, W]`.

    It owns three `VFA` modules under the reference's attribute names (`vfa8`, `vfa16`, `vfa32`, reference
    vfa/model/vfanet.py:30-32), so the `vfa{8,16,32}.*` entries of a `VFANet` checkpoint load with
    `load_state_dict(..., strict=False)` filtering, and its result equals the reference's loop
    `sum_cam (vfa8(lat8[cam]) + vfa16(lat16[cam]) + vfa32(lat32[cam]))` (vfanet.py:64-82) for every frame of the batch.
    """

    def __init__(self, channel, grid_height, cube_size, args, flags: int = 0):
        super().__init__()
        from .vfa_op import VFA
        self.vfa8 = VFA(channel, grid_height, cube_size, 1 / 8., args)
        self.vfa16 = VFA(channel, grid_height, cube_size, 1 / 16., args)
        self.vfa32 = VFA(channel, grid_height, cube_size, 1 / 32., args)
        self.flags = int(flags)

    def forward(self, feats, calibs, grid, crange=(-1.0, 0.95), channels_last: bool | None = None):
        vfas = (self.vfa8, self.vfa16, self.vfa32)[:len(feats)]
        if len(feats) != 3:
            raise ValueError('MultiScaleVFA aggregates the three FPN scales (stride 8, 16, 32) of the reference network')
        L, W = grid.shape[-3], grid.shape[-2]
        table = build_table(vfas[0].geometry((L, W), crange), calibs.reshape(-1, 3, 4), grid)
        return aggregate(list(feats), table, [m.collapse.weight for m in vfas], [m.collapse.bias for m in vfas],
                         flags=self.flags, channels_last=channels_last)
