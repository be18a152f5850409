"""Seeded synthetic inputs of the aggregation path.

The same generators feed the oracle, the parity tests, the CPU baseline and the GPU bench, so every arm sees
identical tensors.  Cameras form a deterministic look-at ring; `P = K [R|t]` is built in fp64 and cast to
fp32 exactly like the reference's data path (reference vfa/data/dataset.py:64 -> vfa/utils.py:44).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .geometry import Geometry

# world-unit rig parameters per dataset: centre (x, y), ring radius, camera height, focal length (pixels)
RIGS = {
    'MultiviewC': dict(centre=(1950.0, 1950.0), radius=2600.0, height=600.0, focal=900.0),
    'MultiviewX': dict(centre=(12.5, 8.0), radius=17.0, height=4.0, focal=1400.0),
    'Wildtrack': dict(centre=(300.0, 900.0), radius=2200.0, height=350.0, focal=1500.0),
}


def look_at(eye, target, focal, image_size) -> np.ndarray:
    """fp64 3x4 projection K[R|t] of a pinhole camera at `eye` looking at `target` (world z is up)."""
    eye = np.asarray(eye, np.float64)
    target = np.asarray(target, np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    t = -R @ eye
    H, W = image_size
    K = np.array([[focal, 0.0, W / 2.0], [0.0, focal, H / 2.0], [0.0, 0.0, 1.0]])
    return K @ np.concatenate([R, t[:, None]], axis=1)


def ring_calibs(geom: Geometry, n_views: int | None = None, in_field: bool = False) -> torch.Tensor:
    """fp32 [V,3,4] look-at ring.  `in_field=True` appends one camera standing inside the field (its
    behind-camera voxels exercise the reference's missing depth test)."""
    rig = RIGS[geom.name]
    V = geom.n_views if n_views is None else n_views
    rs = np.random.RandomState(0)
    cx, cy = rig['centre']
    out = []
    for v in range(V):
        ang = 2.0 * math.pi * v / V + 0.1
        eye = (cx + rig['radius'] * math.cos(ang), cy + rig['radius'] * math.sin(ang), rig['height'])
        jit = rs.uniform(-0.1, 0.1, size=2) * rig['radius']
        out.append(look_at(eye, (cx + jit[0], cy + jit[1], 0.0), rig['focal'], geom.image_size))
    if in_field:
        # low and nearly level: ground voxels behind it mirror into the upper image half and pass `visible`
        eye = (cx + 0.13 * rig['radius'], cy - 0.07 * rig['radius'], 0.35 * rig['height'])
        out.append(look_at(eye, (cx - 0.4 * rig['radius'], cy + 0.3 * rig['radius'], 0.30 * rig['height']),
                           rig['focal'], geom.image_size))
    return torch.from_numpy(np.stack(out)).to(torch.float32)


def features(geom: Geometry, batch: int = 1, n_views: int | None = None, channels: int | None = None,
             seed: int = 0, sizes=None, device='cpu', pin: bool = False) -> list:
    """Three fp32 tensors [B,V,C,fH,fW] = relu(randn) (post-ReLU laterals are non-negative, reference
    vfa/model/vfanet.py:72-74)."""
    V = geom.n_views if n_views is None else n_views
    C = geom.channels if channels is None else channels
    g = torch.Generator(device='cpu').manual_seed(seed)
    out = []
    for (h, w) in (sizes or geom.feature_sizes()):
        t = torch.randn(batch, V, C, h, w, generator=g).relu_()
        if pin:
            t = t.pin_memory()
        out.append(t.to(device))
    return out


def collapse_params(geom: Geometry, channels: int | None = None, seed: int = 0):
    """Three (weight [C, C*nl], bias [C]) pairs with nn.Linear's default init under a fixed seed."""
    C = geom.channels if channels is None else channels
    K = C * geom.n_layers
    g = torch.Generator(device='cpu').manual_seed(seed + 1000)
    bound = 1.0 / math.sqrt(K)
    out = []
    for _ in range(3):
        w = (torch.rand(C, K, generator=g) * 2 - 1) * bound
        b = (torch.rand(C, generator=g) * 2 - 1) * bound
        out.append((w, b))
    return out
