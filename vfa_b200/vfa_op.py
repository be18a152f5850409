"""Host-side mirror of the reference aggregation module, backed by libvfa_b200.so.

`VFA` keeps the constructor, forward signature, buffers and parameter names of the reference module (reference
vfa/model/vfa_op.py:46-125) so `vfa.model.vfanet.VFANet` (reference vfa/model/vfanet.py:30-32, :76-78) and released
checkpoints work unchanged.  `aggregate` is the fused multi-view / multi-scale entry that replaces the loop of
reference vfa/model/vfanet.py:64-82 (minus the lateral convs) with one kernel launch per frame batch.

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic runs in hand-written sm_100a
kernels behind the C ABI.  Nothing in this file computes on the CPU or falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .geometry import convert_descriptor

__all__ = ['VFA', 'ProjectionTable', 'build_table', 'aggregate', 'aggregate_forward_raw', 'prepare_weights',
           'make_shape', 'to_channels_last', 'last_kernel_path']


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: vfa_b200 has no CPU path (got device {t.device})')


def make_geometry(n_layers, cube_size, layer_z, grid_lw, dataset, image_size, crange=(-1.0, 0.95)) -> _lib.Geometry:
    kind, scale, offset = convert_descriptor(dataset)
    g = _lib.Geometry()
    g.n_layers = int(n_layers)
    g.grid_l, g.grid_w = int(grid_lw[0]), int(grid_lw[1])
    g.convert_kind = kind
    g.convert_scale = float(scale)
    for a in range(3):
        g.convert_offset[a] = float(offset[a])
        g.cube[a] = float(cube_size[a])
    if n_layers > _lib.VFA_MAX_LAYERS:
        raise ValueError(f'{n_layers} height layers > supported maximum {_lib.VFA_MAX_LAYERS}')
    for n in range(int(n_layers)):
        g.layer_z[n] = float(layer_z[n])
    g.image_h, g.image_w = float(image_size[0]), float(image_size[1])
    g.clamp_lo, g.clamp_hi = float(crange[0]), float(crange[1])
    return g


class ProjectionTable:
    """Boxes [V, nl, L*W, 4] fp32 for a set of cameras (bit-identical to the reference's `box_corners`)."""

    def __init__(self, geom: _lib.Geometry, boxes: torch.Tensor):
        self.geom = geom
        self.boxes = boxes

    @property
    def n_views(self):
        return self.boxes.shape[0]

    def scale_table(self, feat_h: int, feat_w: int):
        """(area fp32, visible bool, taps int32[...,4]) of one feature scale -- the parity-checked quantities."""
        n = self.boxes.numel() // 4
        area = torch.empty(self.boxes.shape[:-1], dtype=torch.float32, device=self.boxes.device)
        vis = torch.empty(self.boxes.shape[:-1], dtype=torch.uint8, device=self.boxes.device)
        taps = torch.empty(self.boxes.shape, dtype=torch.int32, device=self.boxes.device)
        with torch.cuda.device(self.boxes.device):
            _lib.check(_lib.lib().vfa_table_scale(self.boxes.data_ptr(), n, int(feat_h), int(feat_w), area.data_ptr(),
                                                  vis.data_ptr(), taps.data_ptr(), _stream()))
        return area, vis.bool(), taps


def build_table(geom: _lib.Geometry, calibs: torch.Tensor, grid: torch.Tensor) -> ProjectionTable:
    """calibs [V,3,4] (or [3,4]) fp32 CUDA, grid [L,W,3] / [1,L,W,3] fp32 CUDA."""
    _require_cuda(calibs, 'calibs')
    _require_cuda(grid, 'grid')
    calibs = calibs.detach().reshape(-1, 3, 4).to(torch.float32).contiguous()
    grid = grid.detach().reshape(-1, 3).to(torch.float32).contiguous()
    if grid.shape[0] != geom.grid_l * geom.grid_w:
        raise ValueError(f'grid has {grid.shape[0]} cells, geometry says {geom.grid_l}x{geom.grid_w}')
    V = calibs.shape[0]
    boxes = torch.empty(V, geom.n_layers, geom.grid_l * geom.grid_w, 4, dtype=torch.float32, device=calibs.device)
    with torch.cuda.device(calibs.device):
        _lib.check(_lib.lib().vfa_table_build(C.byref(geom), V, calibs.data_ptr(), grid.data_ptr(), boxes.data_ptr(),
                                              _stream()))
    return ProjectionTable(geom, boxes)


class _ChannelsLast(torch.autograd.Function):
    """[N, C, H, W] contiguous -> [N, H, W, C] contiguous with the library's tiled transpose (and back for grads)."""

    @staticmethod
    def forward(ctx, x):
        N, Cc, H, W = x.shape
        out = torch.empty(N, H, W, Cc, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().vfa_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), N, Cc, H * W, _stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N, H, W, Cc = g.shape
        out = torch.empty(N, Cc, H, W, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().vfa_nhwc_to_nchw(g.data_ptr(), out.data_ptr(), N, Cc, H * W, _stream()))
        return out


def to_channels_last(x: torch.Tensor) -> torch.Tensor:
    """[..., C, H, W] fp32 CUDA -> contiguous [..., H, W, C].  Zero-copy when the tensor already has
    channels-last strides (e.g. produced by a cuDNN conv in torch.channels_last memory format)."""
    _require_cuda(x, 'feature')
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f'features must be float32 or bfloat16, got {x.dtype}')
    lead = x.shape[:-3]
    Cc, H, W = x.shape[-3:]
    perm = x.movedim(-3, -1)
    if perm.is_contiguous():
        return perm
    if x.dtype == torch.bfloat16:
        raise ValueError('bfloat16 features must already be channels-last ([...,C,H,W] with C innermost in memory), '
                         'e.g. the output of a conv run in torch.channels_last under autocast')
    x4 = x.reshape(-1, Cc, H, W).contiguous()
    return _ChannelsLast.apply(x4).reshape(*lead, H, W, Cc)


def make_shape(feats_cl, n_layers: int) -> _lib.Shape:
    """vfa_shape_t of a list of channels-last feature tensors [B,V,fH,fW,C]."""
    B, V, _, _, Cc = feats_cl[0].shape
    shape = _lib.Shape()
    shape.batch, shape.n_views, shape.channels, shape.n_scales = B, V, Cc, len(feats_cl)
    for s, f in enumerate(feats_cl):
        if f.dim() != 5 or f.shape[0] != B or f.shape[1] != V or f.shape[4] != Cc:
            raise ValueError('feature tensors disagree on batch / views / channels')
        if f.dtype not in (torch.float32, torch.bfloat16) or f.dtype != feats_cl[0].dtype or not f.is_contiguous():
            raise ValueError('features must be contiguous float32 (or all bfloat16) [B,V,fH,fW,C]')
        shape.feat_h[s], shape.feat_w[s] = f.shape[2], f.shape[3]
    return shape


def workspace_for(geom, shape, flags, device, direction: str = 'forward') -> torch.Tensor:
    """Scratch for one direction ('forward' / 'backward') or both ('both') of the kernel family `flags` selects."""
    size_flag = {'forward': _lib.FLAG_WS_FORWARD, 'backward': _lib.FLAG_WS_BACKWARD, 'both': 0}[direction]
    n = _lib.lib().vfa_aggregate_workspace_bytes(C.byref(geom), C.byref(shape), int(flags) | size_flag)
    return torch.empty(max(n, 256), dtype=torch.uint8, device=device)


def prepare_weights(geom, shape, weights, flags=0, workspace=None) -> torch.Tensor:
    """Re-lay the collapse weights for the kernel family `flags` selects; returns the workspace holding them
    (pass it back to aggregate_forward_raw(..., prepared=True)).  Inference callers do this once."""
    weights = [w.detach().contiguous() for w in weights]
    dev = weights[0].device
    ws = workspace if workspace is not None else workspace_for(geom, shape, flags, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vfa_prepare_weights(C.byref(geom), C.byref(shape),
                                                  _lib.ptr_array([w.data_ptr() for w in weights]), ws.data_ptr(),
                                                  ws.numel(), int(flags), _stream()))
    return ws


def aggregate_forward_raw(feats_cl, table, weights, biases, flags=0, out=None, workspace=None, prepared=False,
                          relu_mask=None, table_prepared=False, out_ptr=None):
    """No-autograd forward on channels-last features [B,V,fH,fW,C]: one C-ABI call on the current stream.
    relu_mask: optional int32 tensor [B,V,S,ceil(C/32),L*W] receiving the ReLU pass bits for the backward.
    table_prepared: `workspace` was used by the previous call with the same table, shapes and flags (static cameras): the
    tap records, coverage and texel lists in it are reused instead of rebuilt (VFA_FLAG_TABLE_PREPARED).
    out_ptr: raw device address written instead of `out` (a peer GPU's buffer or the multicast address of a symmetric
    allocation, with FLAG_OUT_ACCUMULATE / FLAG_OUT_MULTICAST); returns None in that case."""
    geom = table.geom
    shape = make_shape(feats_cl, geom.n_layers)
    dev = feats_cl[0].device
    if out is None and out_ptr is None:
        if int(flags) & _lib.FLAG_OUT_NHWC:      # [B, L, W, C]: what a channels-last head reads without a permute
            out = torch.empty(shape.batch, geom.grid_l, geom.grid_w, shape.channels, dtype=torch.float32, device=dev)
        else:
            out = torch.empty(shape.batch, shape.channels, geom.grid_l, geom.grid_w, dtype=torch.float32, device=dev)
    if workspace is None and int(flags) & (_lib.FLAG_TABLE_PREPARED | _lib.FLAG_WEIGHTS_PREPARED):
        raise ValueError('FLAG_TABLE_PREPARED / FLAG_WEIGHTS_PREPARED describe the contents of a caller-owned workspace: '
                         'pass `workspace=` (a freshly allocated one holds garbage)')
    ws = workspace if workspace is not None else workspace_for(geom, shape, flags, dev)
    f = int(flags) | (_lib.FLAG_WEIGHTS_PREPARED if prepared else 0)
    if table_prepared:
        if workspace is None:
            raise ValueError('table_prepared=True needs the workspace of the previous call')
        f |= _lib.FLAG_TABLE_PREPARED
    if feats_cl[0].dtype == torch.bfloat16:
        f |= _lib.FLAG_BF16_FEATURES
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vfa_aggregate_fwd(C.byref(geom), C.byref(shape), table.boxes.data_ptr(),
                                                _lib.ptr_array([t.data_ptr() for t in feats_cl]),
                                                _lib.ptr_array([t.data_ptr() for t in weights]),
                                                _lib.ptr_array([t.data_ptr() for t in biases]),
                                                out_ptr if out_ptr is not None else out.data_ptr(),
                                                relu_mask.data_ptr() if relu_mask is not None else None,
                                                ws.data_ptr(), ws.numel(), f, _stream()))
    return out


class _AggregateFn(torch.autograd.Function):
    """out [B,C,L,W] = sum_v sum_s relu(collapse_s(pooled voxels)); tensors = feats(S) + weights(S) + biases(S)."""

    @staticmethod
    def forward(ctx, geom, boxes, flags, n_scales, *tensors):
        S = n_scales
        feats = [t.contiguous() for t in tensors[:S]]
        weights = [t.contiguous() for t in tensors[S:2 * S]]
        biases = [t.contiguous() for t in tensors[2 * S:3 * S]]
        shape = make_shape(feats, geom.n_layers)
        B, V, Cc = shape.batch, shape.n_views, shape.channels
        for s in range(S):
            if tuple(weights[s].shape) != (Cc, Cc * geom.n_layers) or tuple(biases[s].shape) != (Cc,):
                raise ValueError(f'collapse parameters of scale {s} have shapes {tuple(weights[s].shape)}, '
                                 f'{tuple(biases[s].shape)}; expected ({Cc}, {Cc * geom.n_layers}) and ({Cc},)')
        if boxes.shape[0] != V:
            raise ValueError(f'table holds {boxes.shape[0]} views, features hold {V}')
        dev = feats[0].device
        need_grad = any(ctx.needs_input_grad[4:])
        mask = None
        if need_grad:
            mask = torch.empty(B, V, S, (Cc + 31) // 32, geom.grid_l * geom.grid_w, dtype=torch.int32, device=dev)
        out = aggregate_forward_raw(feats, ProjectionTable(geom, boxes), weights, biases, flags, relu_mask=mask)
        ctx.geom, ctx.shape, ctx.flags, ctx.S = geom, shape, flags, S
        if need_grad:
            ctx.save_for_backward(boxes, mask, *feats, *weights)
        if flags & _lib.FLAG_OUT_NHWC:                # memory is [B,L,W,C]: a [B,C,L,W] tensor in torch.channels_last
            return out.permute(0, 3, 1, 2)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        S = ctx.S
        saved = ctx.saved_tensors
        boxes, mask, feats, weights = saved[0], saved[1], saved[2:2 + S], saved[2 + S:2 + 2 * S]
        if ctx.flags & _lib.FLAG_OUT_NHWC:            # zero-copy when the consumer ran in channels_last (cuDNN heads)
            grad_out = grad_out.permute(0, 2, 3, 1).contiguous()
        else:
            grad_out = grad_out.contiguous()
        dev = grad_out.device
        L = _lib.lib()
        need_f = [ctx.needs_input_grad[4 + s] for s in range(S)]
        need_w = [ctx.needs_input_grad[4 + S + s] for s in range(S)]
        need_b = [ctx.needs_input_grad[4 + 2 * S + s] for s in range(S)]
        gf = [torch.zeros_like(feats[s]) if need_f[s] else None for s in range(S)]
        want_wb = [need_w[s] or need_b[s] for s in range(S)]
        gw = [torch.empty_like(weights[s]) if want_wb[s] else None for s in range(S)]
        gb = [torch.empty(weights[s].shape[0], dtype=torch.float32, device=dev) if want_wb[s] else None
              for s in range(S)]
        ws = workspace_for(ctx.geom, ctx.shape, ctx.flags, dev, direction='backward')
        with torch.cuda.device(dev):
            _lib.check(L.vfa_aggregate_bwd(C.byref(ctx.geom), C.byref(ctx.shape), boxes.data_ptr(),
                                           _lib.ptr_array([t.data_ptr() for t in feats]),
                                           _lib.ptr_array([t.data_ptr() for t in weights]),
                                           mask.data_ptr(), grad_out.data_ptr(),
                                           _lib.ptr_array([t.data_ptr() if t is not None else None for t in gf]),
                                           _lib.ptr_array([t.data_ptr() if t is not None else None for t in gw]),
                                           _lib.ptr_array([t.data_ptr() if t is not None else None for t in gb]),
                                           ws.data_ptr(), ws.numel(), ctx.flags, _stream()))
        grads = [g if n else None for g, n in zip(gf, need_f)]
        grads += [g if n else None for g, n in zip(gw, need_w)]
        grads += [g if n else None for g, n in zip(gb, need_b)]
        return (None, None, None, None, *grads)


def aggregate(feats, table: ProjectionTable, weights, biases, flags: int = 0, channels_last: bool | None = None):
    """Fused aggregation of a frame batch.

    feats    list of S tensors [B,V,C,fH,fW] fp32 CUDA (any strides; channels-last strides are consumed zero-copy),
             or [B,V,fH,fW,C] when channels_last=True
    table    ProjectionTable for the V cameras (build_table)
    weights  list of S tensors [C, C*nl] (`collapse.weight`, column order c*nl+n), biases list of S tensors [C]
    returns  [B,C,L,W] fp32 = sum over views and scales of relu(collapse(pooled voxels)), autograd-connected.
    """
    S = len(feats)
    if not (1 <= S <= _lib.VFA_MAX_SCALES) or len(weights) != S or len(biases) != S:
        raise ValueError('need 1..3 feature scales with one (weight, bias) pair each')
    cl = []
    for f in feats:
        if f.dim() != 5:
            raise ValueError(f'features must be 5-D [B,V,C,H,W], got {tuple(f.shape)}')
        cl.append(f if channels_last else to_channels_last(f))
    for t in list(weights) + list(biases):
        _require_cuda(t, 'collapse parameter')
    if cl[0].dtype == torch.bfloat16 or (int(flags) & _lib.FLAG_BF16_MMA):
        # bf16 feature storage / bf16 tensor-core variant: inference only (the backward kernels are fp32)
        if any(t.requires_grad for t in list(cl) + list(weights) + list(biases)) and torch.is_grad_enabled():
            raise RuntimeError('bfloat16 feature maps / FLAG_BF16_MMA are forward-only; run under torch.no_grad() or use '
                               'the float32 path')
        return aggregate_forward_raw([t.contiguous() for t in cl], table, [w.detach() for w in weights],
                                     [b.detach() for b in biases], flags)
    # the PREPARED bits describe a caller-owned workspace; the autograd path allocates its own
    flags = int(flags) & ~(_lib.FLAG_TABLE_PREPARED | _lib.FLAG_WEIGHTS_PREPARED)
    return _AggregateFn.apply(table.geom, table.boxes, flags, S, *cl, *weights, *biases)


def last_kernel_path() -> str:
    """Kernel family the most recent aggregate() on this thread dispatched to ('simt_fp32', 'umma_tf32x3', ...)."""
    return _lib.last_path()


class VFA(nn.Module):
    """Drop-in for the reference `VFA` module (reference vfa/model/vfa_op.py:46-125).

    Same constructor arguments, same buffers (`z_corners` int64 [nl,1,1,3], `corners_offset` fp32 [1,1,1,1,8,3]) and
    parameters (`collapse.weight` [C, C*nl], `collapse.bias` [C]); `args` is any object with `.data` in
    {'MultiviewC','MultiviewX','Wildtrack'} and `.image_size = (H, W)`.

    Smallest input: with the default `crange = (-1, 0.95)` every feature map needs MORE THAN 20 texels along each axis
    (the reference's 720 x 1280 configs give 23 x 40 at stride 32); smaller maps or a `crange[1]` nearer 1 raise
    `VFAError` (UNSUPPORTED) -- there the reference samples its integral image outside the map and reads zero padding,
    which the direct pooling form does not reproduce.
    """

    def __init__(self, channel, grid_height=160, cube_size=(25, 25, 32), feat_scale=1, args=None):
        super().__init__()
        cube_size = [float(v) for v in np.asarray(cube_size).tolist()]
        self.cube_size = tuple(cube_size)
        self.cube_height = cube_size[2]
        z_corners = torch.arange(0, grid_height, int(cube_size[2]) if float(cube_size[2]).is_integer() else cube_size[2])
        z_corners = F.pad(z_corners.view(-1, 1, 1, 1), [2, 0])
        l, w, h = cube_size
        offs = torch.tensor([[-l / 2, -w / 2, 0], [l / 2, -w / 2, 0], [l / 2, w / 2, 0], [-l / 2, w / 2, 0],
                             [-l / 2, -w / 2, h], [l / 2, -w / 2, h], [l / 2, w / 2, h], [-l / 2, w / 2, h]],
                            dtype=torch.float32).view(1, 1, 1, 1, 8, 3)
        self.register_buffer('z_corners', z_corners)
        self.register_buffer('corners_offset', offs)
        self.feat_scale = feat_scale          # accepted and unused, like the reference (vfa_op.py:57, :74)
        self.args = args
        self.channel = channel
        self.num_grid_layer = z_corners.shape[0]
        self.collapse = nn.Linear(channel * self.num_grid_layer, channel)
        self.flags = 0

    def geometry(self, grid_lw, crange=(-1.0, 0.95)) -> _lib.Geometry:
        if self.args is None:
            raise ValueError('VFA needs `args` with .data and .image_size (reference vfa_op.py:38-43, :75)')
        # the layer heights live in a (device) buffer; read them back once, not per call (a .tolist() is a stream sync)
        key = (self.z_corners.data_ptr(), self.z_corners._version)
        if getattr(self, '_layer_z_key', None) != key:
            self._layer_z, self._layer_z_key = self.z_corners[:, 0, 0, 2].tolist(), key
        layer_z = self._layer_z
        return make_geometry(self.num_grid_layer, self.cube_size, layer_z, grid_lw, self.args.data, self.args.image_size,
                             crange)

    def forward(self, feature, calib, grid, crange=(-1, 0.95), visualize=False):
        """feature [1,C,fH,fW], calib [3,4], grid [1,L,W,3] -> [1,C,L,W]   (reference vfa_op.py:61-125)."""
        if visualize:
            raise NotImplementedError('visualize=True is a matplotlib debugging aid of the reference and is not provided')
        if feature.dim() != 4 or feature.shape[0] != 1:
            raise ValueError(f'feature must be [1,C,fH,fW] (the reference is batch-1, vfa_op.py:64), got {tuple(feature.shape)}')
        L, W = grid.shape[-3], grid.shape[-2]
        geom = self.geometry((L, W), crange)
        table = build_table(geom, calib, grid)
        return aggregate([feature.unsqueeze(0)], table, [self.collapse.weight], [self.collapse.bias], self.flags)
