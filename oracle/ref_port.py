"""Torch-CPU port of the reference aggregation  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement of reference vfa/model/vfa_op.py:61-125 (one `VFA.forward`) and of the per-view /
per-scale loop of reference vfa/model/vfanet.py:64-82, using the same torch operators the reference uses
(cumsum o cumsum, four bilinear `grid_sample`s, `linear`, `relu`), so that it

  * is the floating-point half of the oracle: run in float64 on the bit-exact fp32 boxes it is the "hybrid
    oracle" of SURVEY.md section 8(c), and autograd through it is the gradient oracle;
  * is the CPU baseline that bench.py times beside the GPU (`cpu_baseline.kind = "port"` and
    `--impl reference`): in float32 it performs the reference's own operator sequence on all host threads.

The reference is Python and cannot travel to the GPU box; tests/golden/make_golden.py pins this port against
the unmodified reference (imported from /root/reference in the build container) and commits the vectors.

The boxes are formed with explicit elementwise fp32 operations in the scalar order of SURVEY.md appendix A.2
(not with torch.matmul), so they do not depend on which BLAS kernel a given host CPU selects.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPSILON = 1e-6               # reference vfa_op.py:14
MAXIMUM_AREA_RATIO = 0.3     # reference vfa_op.py:15

_CONVERT = {'MultiviewC': ('div', 1.0, (0.0, 0.0, 0.0)),
            'MultiviewX': ('div', 40.0, (0.0, 0.0, 0.0)),
            'Wildtrack': ('affine', 2.5, (300.0, 900.0, 0.0))}


def cube_offsets(cube_size):
    l, w, h = (float(v) for v in cube_size)
    return torch.tensor([[-l / 2, -w / 2, 0], [l / 2, -w / 2, 0], [l / 2, w / 2, 0], [-l / 2, w / 2, 0],
                         [-l / 2, -w / 2, h], [l / 2, -w / 2, h], [l / 2, w / 2, h], [-l / 2, w / 2, h]],
                        dtype=torch.float32)


def boxes_fp32(calib, grid, grid_height, cube_size, dataset, image_size, crange=(-1.0, 0.95)):
    """fp32 [nl, L*W, 4] clamped normalised boxes (reference vfa_op.py:64-88, utils.py:50-59)."""
    calib = calib.to(torch.float32).reshape(3, 4)
    grid = grid.to(torch.float32).reshape(grid.shape[-3], grid.shape[-2], 3)
    L, W, _ = grid.shape
    zs = torch.arange(0, int(grid_height), int(cube_size[2])).to(torch.float32)
    zvec = torch.stack([torch.zeros_like(zs), torch.zeros_like(zs), zs], dim=-1)
    pts = (grid[None, :, :, None, :] + zvec[:, None, None, None, :]) + cube_offsets(cube_size)[None, None, None]
    kind, scale, off = _CONVERT[dataset]
    if kind == 'div':
        pts = pts / scale
    else:
        pts = torch.stack([pts[..., 0] * scale - off[0], pts[..., 1] * scale - off[1], pts[..., 2] * scale], -1)
    X, Y, Z = pts.unbind(-1)
    h = [((calib[r, 0] * X + calib[r, 1] * Y) + calib[r, 2] * Z) + calib[r, 3] for r in range(3)]
    u, v = h[0] / h[2], h[1] / h[2]
    nx = ((2 * u) / float(image_size[1]) - 1).clamp(crange[0], crange[1])
    ny = ((2 * v) / float(image_size[0]) - 1).clamp(crange[0], crange[1])
    box = torch.stack([nx.min(-1)[0], ny.min(-1)[0], nx.max(-1)[0], ny.max(-1)[0]], dim=-1)
    return box.reshape(len(zs), L * W, 4)


def pooled_voxels(feature, boxes):
    """[LW, C*nl] pre-collapse matrix in feature.dtype from given boxes (reference vfa_op.py:104-120)."""
    C, fh, fw = feature.shape[-3:]
    feature = feature.reshape(1, C, fh, fw)
    b = boxes.to(feature.dtype)[None]                                        # [1, nl, LW, 4]
    area = ((b[..., 2:] - b[..., :2]).prod(dim=-1) * fh * fw + EPSILON).unsqueeze(1)
    visible = torch.logical_and(area > EPSILON, area < (fh * fw * MAXIMUM_AREA_RATIO))
    integral = torch.cumsum(torch.cumsum(feature, dim=-1), dim=-2)
    lt = F.grid_sample(integral, b[..., [0, 1]], align_corners=False)
    rb = F.grid_sample(integral, b[..., [2, 3]], align_corners=False)
    rt = F.grid_sample(integral, b[..., [2, 1]], align_corners=False)
    lb = F.grid_sample(integral, b[..., [0, 3]], align_corners=False)
    vox = (lt + rb - rt - lb) / area
    vox = vox * visible
    return vox.permute(0, 3, 1, 2).flatten(0, 1).flatten(1, 2)


def preactivation(feature, calib, grid, weight, bias, grid_height, cube_size, dataset, image_size,
                  crange=(-1.0, 0.95), boxes=None):
    """The linear part of one (view, scale), before the ReLU of reference vfa_op.py:124: [1, C, L, W]."""
    L, W = grid.shape[-3], grid.shape[-2]
    if boxes is None:
        boxes = boxes_fp32(calib, grid, grid_height, cube_size, dataset, image_size, crange)
    vox = pooled_voxels(feature, boxes)
    return F.linear(vox, weight, bias).view(1, L, W, -1).permute(0, 3, 1, 2)


def vfa_forward(feature, calib, grid, weight, bias, grid_height, cube_size, dataset, image_size,
                crange=(-1.0, 0.95), boxes=None):
    """One (view, scale): -> [1, C, L, W] in feature.dtype.  With feature/weight/bias in float64 this is the
    hybrid oracle (boxes stay the bit-exact fp32 ones)."""
    L, W = grid.shape[-3], grid.shape[-2]
    if boxes is None:
        boxes = boxes_fp32(calib, grid, grid_height, cube_size, dataset, image_size, crange)
    vox = pooled_voxels(feature, boxes)
    out = F.linear(vox, weight, bias).view(1, L, W, -1)
    return F.relu(out.permute(0, 3, 1, 2))


def aggregate(feats, calibs, grid, params, grid_height, cube_size, dataset, image_size, cache_boxes=False):
    """feats: three [B,V,C,fH,fW]; calibs [V,3,4]; grid [L,W,3]; params: three (weight, bias) -> [B,C,L,W].
    Frames are looped because the reference is structurally batch-1 (reference train.py:57-59).  With
    cache_boxes=False (default, used for baseline timing) the boxes are re-derived for every (view, scale,
    frame) exactly as the reference does; cache_boxes=True only saves oracle time."""
    B, V = feats[0].shape[:2]
    boxes = [boxes_fp32(calibs[v], grid, grid_height, cube_size, dataset, image_size) if cache_boxes else None
             for v in range(V)]
    frames = []
    for b in range(B):
        ortho = 0
        for v in range(V):
            per_view = 0
            for s in range(len(feats)):
                per_view = per_view + vfa_forward(feats[s][b, v], calibs[v], grid, params[s][0], params[s][1],
                                                  grid_height, cube_size, dataset, image_size, boxes=boxes[v])
            ortho = ortho + per_view
        frames.append(ortho)
    return torch.cat(frames, dim=0)
