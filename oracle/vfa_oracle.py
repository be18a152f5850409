"""CPU oracle of VFA's voxelized feature aggregation  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package; the product (vfa_b200/) never does and fails loudly without its CUDA library.

This file is the *bit-exact* half of the oracle: a numpy restatement, in scalar fp32 operation order, of
everything the reference computes with integer / bit-pattern-level meaning on this path:

  boxes      reference vfa/model/vfa_op.py:64-88 + vfa/utils.py:50-59   (corner build, convert, project,
             normalise, clamp, min/max over the 8 cuboid corners)
  area       reference vfa/model/vfa_op.py:104-105
  visible    reference vfa/model/vfa_op.py:106
  taps       ATen grid_sampler unnormalisation floor(((c+1)*S-1)/2)  (torch ATen/native/GridSampler.h:27-35,
             the arithmetic behind F.grid_sample at reference vfa_op.py:112-115; torch is an un-vendored,
             un-pinned dependency of the reference -- torch 2.11.0 is the version of record)

and a float64 restatement of the pooled voxel features, both as the reference computes them (integral image
sampled bilinearly at the four box corners, vfa_op.py:110-118, :172-173) and in the algebraically equal direct
form, used to cross-check the torch port in oracle/ref_port.py.

Pinning: tests/golden/make_golden.py imports the unmodified reference from /root/reference, runs it on seeded
inputs and stores its outputs in tests/golden/*.npz; tests/test_oracle_golden.py holds this file to those
vectors bit for bit (boxes, area, visible) and to 1e-12 (fp64 features).

Every numpy expression below is a single separately-rounded IEEE fp32 operation per array op (numpy never
contracts a*b+c into an FMA), which is what torch's CPU kernels do for the reference.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPSILON = 1e-6               # reference vfa_op.py:14
MAXIMUM_AREA_RATIO = 0.3     # reference vfa_op.py:15

CONVERT = {
    # dataset -> (kind, scale, offsets)   reference vfa_op.py:23-35
    'MultiviewC': ('div', 1.0, (0.0, 0.0, 0.0)),
    'MultiviewX': ('div', 40.0, (0.0, 0.0, 0.0)),
    'Wildtrack': ('affine', 2.5, (300.0, 900.0, 0.0)),
}


def make_grid(world_size, cube_lw, dataset, grid_offset=(0, 0, 0)):
    """fp32 [L,W,3] cell origins (reference vfa/utils.py:16-37)."""
    length, width = world_size[::-1] if dataset == 'Wildtrack' else world_size
    xs = np.arange(0., width, cube_lw[0], dtype=F32) + F32(grid_offset[0])
    ys = np.arange(0., length, cube_lw[1], dtype=F32) + F32(grid_offset[1])
    if dataset == 'Wildtrack':
        xx, yy = np.meshgrid(xs, ys, indexing='ij')
    else:
        yy, xx = np.meshgrid(ys, xs, indexing='ij')
    return np.stack([xx, yy, np.full_like(xx, F32(grid_offset[2]))], axis=-1).astype(F32)


def cube_offsets(cube_size):
    """fp32 [8,3] corner offsets in the reference's corner order (vfa_op.py:127-133)."""
    l, w, h = (float(v) for v in cube_size)
    x = [-l / 2, l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2]
    y = [-w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2, w / 2]
    z = [0, 0, 0, 0, h, h, h, h]
    return np.stack([x, y, z], axis=1).astype(F32)


def layer_heights(grid_height, cube_h):
    """int64 [nl] layer base heights (vfa_op.py:50)."""
    return np.arange(0, int(grid_height), int(cube_h), dtype=np.int64)


def to_world(p, dataset):
    """grid units -> world units, fp32, same operation per dataset as reference vfa_op.py:23-44."""
    kind, scale, off = CONVERT[dataset]
    if kind == 'div':
        return p / F32(scale)
    out = np.empty_like(p)
    for a in range(3):
        out[..., a] = p[..., a] * F32(scale)
        if a < 2:                                  # the reference subtracts only on x and y (vfa_op.py:32-34)
            out[..., a] = out[..., a] - F32(off[a])
    return out


def project_boxes(calib, grid, grid_height, cube_size, dataset, image_size, crange=(-1.0, 0.95)):
    """Clamped normalised bounding boxes of every voxel's 8 projected corners.

    calib [3,4] fp32, grid [L,W,3] fp32  ->  boxes fp32 [nl, L*W, 4] = (left, top, right, bottom).
    """
    calib = np.asarray(calib, F32)
    grid = np.asarray(grid, F32)
    L, W, _ = grid.shape
    zs = layer_heights(grid_height, cube_size[2]).astype(F32)            # int64 -> fp32 promotion
    offs = cube_offsets(cube_size)
    img_w, img_h = F32(image_size[1]), F32(image_size[0])                 # image_size[::-1], vfa_op.py:75
    lo, hi = F32(crange[0]), F32(crange[1])
    with np.errstate(all='ignore'):
        # (grid + z_n) + off_k ; z_n is added to all three components as (0, 0, z_n)   vfa_op.py:64-66
        base = grid[None, :, :, None, :] + np.stack(
            [np.zeros_like(zs), np.zeros_like(zs), zs], axis=-1)[:, None, None, None, :]
        pts = base + offs[None, None, None, :, :]                         # [nl, L, W, 8, 3]
        pts = to_world(pts, dataset)
        X, Y, Z = pts[..., 0], pts[..., 1], pts[..., 2]
        h = []
        for r in range(3):                                                # utils.py:57, left-to-right, no FMA
            acc = calib[r, 0] * X
            acc = acc + calib[r, 1] * Y
            acc = acc + calib[r, 2] * Z
            h.append(acc + calib[r, 3])
        u = h[0] / h[2]                                                   # utils.py:59 (no depth test)
        v = h[1] / h[2]
        nx = np.clip((F32(2.0) * u) / img_w - F32(1.0), lo, hi)           # vfa_op.py:76 ; NaN propagates
        ny = np.clip((F32(2.0) * v) / img_h - F32(1.0), lo, hi)
        boxes = np.stack([nx.min(-1), ny.min(-1), nx.max(-1), ny.max(-1)], axis=-1)   # vfa_op.py:81-86
    return boxes.reshape(len(zs), L * W, 4).astype(F32)


def area_visible(boxes, fh, fw):
    """fp32 area [nl, LW] and bool visible [nl, LW] for one feature scale (vfa_op.py:104-106)."""
    boxes = np.asarray(boxes, F32)
    with np.errstate(all='ignore'):
        wh = boxes[..., 2:] - boxes[..., :2]
        area = (wh[..., 0] * wh[..., 1]) * F32(fh) * F32(fw) + F32(EPSILON)
        # `area < fh*fw*0.3` compares an fp32 tensor with a Python double; torch casts the scalar to fp32
        visible = (area > F32(EPSILON)) & (area < F32(fh * fw * MAXIMUM_AREA_RATIO))
    return area.astype(F32), visible


def tap_index(coord, size):
    """int32 floor of the unnormalised sampling coordinate, fp32 (ATen GridSampler.h:27-35)."""
    coord = np.asarray(coord, F32)
    with np.errstate(all='ignore'):
        ix = ((coord + F32(1.0)) * F32(size) - F32(1.0)) / F32(2.0)
        fl = np.floor(ix)
    return np.where(np.isfinite(fl), fl, -1).astype(np.int32)


# ----------------------------------------------------------------------------------------------------------
# float64 pooled features on given (fp32) boxes -- the "hybrid oracle" of SURVEY.md section 8(c).
# ----------------------------------------------------------------------------------------------------------

def _unnormalise(c, size):
    return ((c + 1.0) * size - 1.0) / 2.0


def _bilinear_zero_pad(img, nx, ny):
    """F.grid_sample(img[C,H,W], (nx, ny)), bilinear, zeros padding, align_corners=False -> [C, N]."""
    C, H, W = img.shape
    ix = _unnormalise(nx.astype(np.float64), W)
    iy = _unnormalise(ny.astype(np.float64), H)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    out = np.zeros((C, nx.shape[0]), np.float64)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            wgt = (1.0 - np.abs(ix - xi)) * (1.0 - np.abs(iy - yi))
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            xi_c = np.clip(xi, 0, W - 1).astype(np.int64)
            yi_c = np.clip(yi, 0, H - 1).astype(np.int64)
            out += img[:, yi_c, xi_c] * (wgt * ok)[None, :]
    return out


def vox_features_integral(feature, boxes, visible=None):
    """fp64 [LW, C*nl] pre-collapse matrix exactly as the reference forms it (vfa_op.py:104-120):
    integral image, four bilinear samples, / area, * visible, column index c*nl + n."""
    f = np.asarray(feature, np.float64)
    C, H, W = f.shape
    nl, LW, _ = boxes.shape
    integ = np.cumsum(np.cumsum(f, axis=-1), axis=-2)                       # vfa_op.py:172-173
    b = boxes.astype(np.float64).reshape(nl * LW, 4)
    bad = ~np.isfinite(b).all(axis=1)
    b = np.where(bad[:, None], 0.0, b)
    lt = _bilinear_zero_pad(integ, b[:, 0], b[:, 1])
    rb = _bilinear_zero_pad(integ, b[:, 2], b[:, 3])
    rt = _bilinear_zero_pad(integ, b[:, 2], b[:, 1])
    lb = _bilinear_zero_pad(integ, b[:, 0], b[:, 3])
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) * H * W + EPSILON
    if visible is None:
        visible = area_visible(boxes, H, W)[1]
    vis = visible.reshape(nl * LW) & ~bad
    vox = (lt + rb - rt - lb) / area[None, :] * vis[None, :]               # [C, nl*LW]
    return vox.reshape(C, nl, LW).transpose(2, 0, 1).reshape(LW, C * nl)


def coverage_weights(lo, hi, size, max_taps):
    """Direct-form separable weights.  For normalised fp32 box edges lo <= hi returns
    (first tap index int64 [N], weights fp64 [N, max_taps]) with
    w_j = clamp(xr-(j-1),0,1) - clamp(xl-(j-1),0,1), j = first .. first+max_taps-1, zeroed outside [0,size)."""
    xl = _unnormalise(lo.astype(np.float64), size)
    xr = _unnormalise(hi.astype(np.float64), size)
    first = np.floor(xl).astype(np.int64)
    j = first[:, None] + np.arange(max_taps)[None, :]
    w = np.clip(xr[:, None] - (j - 1), 0.0, 1.0) - np.clip(xl[:, None] - (j - 1), 0.0, 1.0)
    w = np.where((j >= 0) & (j < size), w, 0.0)
    return first, w


def vox_features_direct(feature, boxes, visible=None):
    """Same quantity as vox_features_integral by the cancellation-free direct sum (valid for the reference's
    default clamp range whenever fW, fH > 20)."""
    f = np.asarray(feature, np.float64)
    C, H, W = f.shape
    nl, LW, _ = boxes.shape
    if visible is None:
        visible = area_visible(boxes, H, W)[1]
    b = boxes.reshape(nl * LW, 4)
    vis = visible.reshape(-1) & np.isfinite(b).all(axis=1)
    idx = np.nonzero(vis)[0]
    out = np.zeros((C, nl * LW), np.float64)
    if idx.size:
        bb = b[idx]
        span_x = int(np.max(np.floor(_unnormalise(bb[:, 2].astype(np.float64), W))
                            - np.floor(_unnormalise(bb[:, 0].astype(np.float64), W)))) + 2
        span_y = int(np.max(np.floor(_unnormalise(bb[:, 3].astype(np.float64), H))
                            - np.floor(_unnormalise(bb[:, 1].astype(np.float64), H)))) + 2
        x0, wx = coverage_weights(bb[:, 0], bb[:, 2], W, span_x)
        y0, wy = coverage_weights(bb[:, 1], bb[:, 3], H, span_y)
        bd = bb.astype(np.float64)
        area = (bd[:, 2] - bd[:, 0]) * (bd[:, 3] - bd[:, 1]) * H * W + EPSILON
        acc = np.zeros((C, idx.size), np.float64)
        for a in range(span_y):
            yi = np.clip(y0 + a, 0, H - 1)
            for c in range(span_x):
                wgt = wy[:, a] * wx[:, c]
                if not wgt.any():
                    continue
                xi = np.clip(x0 + c, 0, W - 1)
                acc += f[:, yi, xi] * wgt[None, :]
        out[:, idx] = acc / area[None, :]
    return out.reshape(C, nl, LW).transpose(2, 0, 1).reshape(LW, C * nl)


def collapse_relu(vox, weight, bias):
    """relu(vox @ W^T + b) -> [C, LW]  (vfa_op.py:123-124), fp64."""
    y = vox.astype(np.float64) @ np.asarray(weight, np.float64).T + np.asarray(bias, np.float64)[None, :]
    return np.maximum(y, 0.0).T


def vfa_forward(feature, calib, grid, weight, bias, grid_height, cube_size, dataset, image_size,
                crange=(-1.0, 0.95)):
    """One reference `VFA.forward` on fp32 boxes with fp64 features: [C, L, W] float64."""
    L, W = grid.shape[:2]
    boxes = project_boxes(calib, grid, grid_height, cube_size, dataset, image_size, crange)
    vox = vox_features_integral(feature, boxes)
    return collapse_relu(vox, weight, bias).reshape(-1, L, W)


def aggregate(feats, calibs, grid, params, grid_height, cube_size, dataset, image_size):
    """The loop of reference vfa/model/vfanet.py:64-82 without the laterals.
    feats: three arrays [V,C,fH,fW]; calibs [V,3,4]; params: three (weight, bias).  -> [C, L, W] float64."""
    out = 0.0
    for v in range(calibs.shape[0]):
        boxes = project_boxes(calibs[v], grid, grid_height, cube_size, dataset, image_size)
        for s in range(3):
            vox = vox_features_integral(feats[s][v], boxes)
            out = out + collapse_relu(vox, *params[s])
    return out.reshape(-1, grid.shape[0], grid.shape[1])
