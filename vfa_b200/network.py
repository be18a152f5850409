"""The reference detector around the aggregation path, for BASELINE config 5.

`VFANet` has the constructor, the `forward(images, calibs, grid)` contract, the returned dictionary and the
state-dict keys of the reference network (reference vfa/model/vfanet.py:14-56, :58-63, :131-150;
backbone: vfa/model/resnet.py:26-55, :100-147), so a released checkpoint loads key for key and the reference's
`train.py:249-256` / `evaluate.py:62-70` can construct it in place of their own.  What differs is the data path
between the backbone and the heads: instead of `7 cameras x 3 scales` Python iterations with a lateral conv, a
GroupNorm and an aggregation call each (reference vfanet.py:64-82), the laterals run once over all cameras in
channels-last memory and ONE fused launch sequence aggregates all cameras, scales and frames
(`vfa_b200.vfanet.aggregate_cameras`).  Frames are batched: `images` may hold `batch * V` views.

The backbone, `fuse` and the heads are stock cuDNN layers -- they are not part of the hand-written path and are here
only so that the full training step of config 5 can be measured and checked end to end.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .vfa_op import VFA
from .vfanet import aggregate_cameras

_STAGES = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}     # reference resnet.py:150-171
# ImageNet checkpoints the reference initialises the trunk from when `pretrained=True` (reference resnet.py:6-12)
_IMAGENET = {'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth',
             'resnet34': 'https://download.pytorch.org/models/resnet34-333f7ec4.pth'}
_GN_GROUPS = 16                                                     # every norm of the backbone (resnet.py:33, :36)


class _Residual(nn.Module):
    """Two 3x3 convs with GroupNorm; a strided 1x1 projection on the skip when the shape changes.
    Attribute names (conv1/bn1/conv2/bn2/downsample) are the checkpoint's (reference resnet.py:26-55)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.GroupNorm(_GN_GROUPS, cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.GroupNorm(_GN_GROUPS, cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.GroupNorm(_GN_GROUPS, cout))

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Trunk(nn.Module):
    """GroupNorm ResNet trunk returning the stride-8/16/32 maps (reference resnet.py:100-147)."""

    def __init__(self, depths):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.GroupNorm(_GN_GROUPS, 64)
        widths, cin = (64, 128, 256, 512), 64
        for i, (w, d) in enumerate(zip(widths, depths)):
            blocks = [_Residual(cin if j == 0 else w, w, (1 if i == 0 else 2) if j == 0 else 1) for j in range(d)]
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
            cin = w
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        f8 = self.layer2(self.layer1(x))
        f16 = self.layer3(f8)
        return f8, f16, self.layer4(f16)


def _head(cout):
    return nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.GroupNorm(16, 256), nn.ReLU(True),
                         nn.Conv2d(256, cout, 3, padding=1, bias=False))


class VFANet(nn.Module):
    """Drop-in for reference `vfa.model.vfanet.VFANet` with a batched, fused aggregation stage."""

    def __init__(self, args, base='resnet18', grid_height=160, cube_size=(25, 25, 32), angle_range=360, mode='3D',
                 pretrained=False, flags: int = 0):
        super().__init__()
        if base not in _STAGES:
            raise ValueError(f'Unrecognized model, expect `resnet18` or `resnet34`, got {base}.')
        if mode not in ('2D', '3D'):
            raise ValueError(f'mode error, expect `2D` or `3D`, got {mode}')
        self.mode = mode
        self.base = _Trunk(_STAGES[base])
        if pretrained:
            self.load_imagenet_trunk(base)
        for s in (8, 16, 32):
            vfa = VFA(channel=256, grid_height=grid_height, cube_size=cube_size, feat_scale=1. / s, args=args)
            vfa.flags = int(flags)
            setattr(self, f'vfa{s}', vfa)
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]))
        for s, cin in ((8, 128), (16, 256), (32, 512)):
            setattr(self, f'lat{s}', nn.Conv2d(cin, 256, 1))
        for s in (8, 16, 32):
            setattr(self, f'bn{s}', nn.GroupNorm(16, 256))
        self.fuse = nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(True),
                                  nn.Conv2d(256, 256, 3, padding=2, dilation=2), nn.BatchNorm2d(256), nn.ReLU(True))
        self.map_classifier = nn.Sequential(nn.Conv2d(256, 1, 3, padding=4, dilation=4, bias=False))
        self.tytx_pred = _head(2)
        if mode == '3D':
            self.orient_pred = nn.Sequential(nn.Conv2d(256, angle_range, 3, padding=4, dilation=4, bias=False))
            self.thtwtl_pred = _head(3)

    def load_imagenet_trunk(self, base: str, state: dict | None = None):
        """`pretrained=True` of the reference (resnet.py:150-180): take from the torchvision ImageNet checkpoint every
        entry whose name the trunk knows -- the convolutions, and the BatchNorm scales / shifts, which land in the
        GroupNorm layers of the same names; running statistics have no counterpart and are dropped.  Needs network
        access (torch.hub download) unless `state` is given."""
        if state is None:
            state = torch.hub.load_state_dict_from_url(_IMAGENET[base], progress=False)
        own = self.base.state_dict()
        own.update({k: v for k, v in state.items() if k in own and v.shape == own[k].shape})
        self.base.load_state_dict(own)

    def bev_features(self, images, calibs, grid, batch: int = 1):
        """images [batch*V, 3, iH, iW] -> ortho [batch, 256, L, W]  (reference vfanet.py:61-82)."""
        x = (images - self.mean.view(3, 1, 1)) / self.std.view(3, 1, 1)
        f8, f16, f32 = self.base(x.contiguous(memory_format=torch.channels_last))
        return aggregate_cameras(self, f8, f16, f32, calibs, grid, (-1.0, 0.95), batch=batch)

    def forward(self, images, calibs, grid, visualize=False, visualize_ortho=False, batch: int = 1):
        if visualize or visualize_ortho:
            raise NotImplementedError('the matplotlib views of reference vfanet.py:84-125 are not part of the drop-in')
        ortho = self.bev_features(images, calibs, grid, batch)
        fused = self.fuse(ortho)                               # reference vfanet.py:131-134
        pred = {'heatmap': self.map_classifier(fused),
                'loc_offset': self.tytx_pred(ortho).permute(0, 2, 3, 1)}
        if self.mode == '3D':
            pred['dim_offset'] = self.thtwtl_pred(ortho).permute(0, 2, 3, 1)
            pred['rotation'] = self.orient_pred(fused).permute(0, 2, 3, 1)
        return pred


def procedural_state(template: dict, seed: int = 0) -> dict:
    """A state dict with the template's keys / shapes whose values depend only on (key, shape, seed) -- not on the
    order modules were constructed in -- so the reference network (build container) and this one (GPU box) can be given
    identical weights without shipping a checkpoint.  Weights ~ N(0, 1/fan_in), norm scales near 1, biases small."""
    import zlib
    out = {}
    for k in sorted(template):
        t = template[k]
        if not t.is_floating_point() or k in ('mean', 'std') or k.endswith('corners_offset'):
            out[k] = t.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) & 0x7fffffff)
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if k.endswith('running_var'):
            v = 1.0 + 0.1 * r.abs()
        elif k.endswith('running_mean'):
            v = 0.05 * r
        elif t.dim() == 1 and k.endswith('weight'):
            v = 1.0 + 0.1 * r
        elif t.dim() == 1:
            v = 0.05 * r
        else:
            fan_in = t[0].numel()
            v = r * (2.0 / fan_in) ** 0.5
        out[k] = v.to(t.dtype)
    return out
