"""The fused camera/scale entry against the reference's own loop structure (reference vfanet.py:64-82), with the
drop-in VFA modules called once per (camera, scale) as the reference does, and against the float64 port."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ref_port                        # noqa: E402
import vfa_b200                                     # noqa: E402
from vfa_b200 import geometry, synthetic, vfanet    # noqa: E402


class _Stem(nn.Module):
    """The attributes of reference VFANet that the loop touches (vfanet.py:30-43), at reduced width."""
    def __init__(self, g, ch=32):
        super().__init__()
        args = SimpleNamespace(data=g.name, image_size=g.image_size)
        self.vfa8 = vfa_b200.VFA(ch, g.grid_height, g.cube_size, 1 / 8., args)
        self.vfa16 = vfa_b200.VFA(ch, g.grid_height, g.cube_size, 1 / 16., args)
        self.vfa32 = vfa_b200.VFA(ch, g.grid_height, g.cube_size, 1 / 32., args)
        self.lat8, self.lat16, self.lat32 = nn.Conv2d(16, ch, 1), nn.Conv2d(24, ch, 1), nn.Conv2d(40, ch, 1)
        self.bn8, self.bn16, self.bn32 = nn.GroupNorm(8, ch), nn.GroupNorm(8, ch), nn.GroupNorm(8, ch)


def test_fused_cameras_equal_reference_loop():
    g = geometry.MULTIVIEWC
    torch.manual_seed(0)
    m = _Stem(g).cuda()
    V = 3
    grid = geometry.grid_for(g)[None].cuda()
    calibs = synthetic.ring_calibs(g, n_views=V).cuda()
    sizes = g.feature_sizes()
    f8 = torch.randn(V, 16, *sizes[0], device='cuda')
    f16 = torch.randn(V, 24, *sizes[1], device='cuda')
    f32 = torch.randn(V, 40, *sizes[2], device='cuda')
    with torch.no_grad():
        fused = vfanet.aggregate_cameras(m, f8, f16, f32, calibs, grid)
        # the reference's loop, one drop-in VFA.forward per (camera, scale)   (vfanet.py:64-82).  The laterals are
        # shared: cuDNN picks different (TF32) conv kernels for batch 1 and batch V, which is not what is under test.
        lats = vfanet.lateral_features(m, f8, f16, f32)
        ortho = 0
        for cam in range(V):
            ortho = ortho + (m.vfa8(lats[0][[cam]], calibs[cam], grid) + m.vfa16(lats[1][[cam]], calibs[cam], grid)
                             + m.vfa32(lats[2][[cam]], calibs[cam], grid))
        # float64 port on the same lateral features
        params = [(v.collapse.weight.detach().cpu(), v.collapse.bias.detach().cpu()) for v in (m.vfa8, m.vfa16, m.vfa32)]
        want = ref_port.aggregate([x.cpu().double()[None] for x in lats], calibs.cpu(), grid[0].cpu(),
                                  [(w.double(), b.double()) for w, b in params], g.grid_height, g.cube_size, g.name,
                                  g.image_size, cache_boxes=True).numpy()
    assert fused.shape == (1, 32, 156, 156)
    torch.testing.assert_close(fused, ortho, rtol=1e-5, atol=2e-6)
    err = np.abs(fused.cpu().numpy().astype(np.float64) - want)
    assert (err <= 1e-6 + 1e-5 * np.abs(want)).all(), f'worst abs err {err.max():.2e}'


def test_fused_cameras_train_step_updates_collapse_and_laterals():
    g = geometry.WILDTRACK
    torch.manual_seed(1)
    m = _Stem(g, ch=16).cuda()
    V, B = 2, 2
    grid = geometry.grid_for(g)[::4, ::4].contiguous().cuda()
    calibs = synthetic.ring_calibs(g, n_views=V).cuda()
    sizes = g.feature_sizes()
    f8 = torch.randn(B * V, 16, *sizes[0], device='cuda')
    f16 = torch.randn(B * V, 24, *sizes[1], device='cuda')
    f32 = torch.randn(B * V, 40, *sizes[2], device='cuda')
    out = vfanet.aggregate_cameras(m, f8, f16, f32, calibs, grid, batch=B)
    assert out.shape == (B, 16, 30, 90)
    out.square().mean().backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
        assert float(p.grad.abs().max()) > 0, name


def test_multiscale_module_equals_module_loop():
    """MultiScaleVFA(feats [B,V,C,H,W] x 3, calibs, grid) == the reference's loop over cameras and scales of drop-in VFA
    modules (vfanet.py:64-82), per frame of the batch, and its gradients reach the collapse parameters."""
    from types import SimpleNamespace
    from vfa_b200 import geometry, synthetic
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::2, ::2].contiguous().cuda()
    V, B = 3, 2
    calibs = synthetic.ring_calibs(g, n_views=V).cuda()
    feats = [f.cuda() for f in synthetic.features(g, batch=B, n_views=V, seed=31)]
    m = vfa_b200.MultiScaleVFA(256, g.grid_height, g.cube_size, SimpleNamespace(data=g.name, image_size=g.image_size)).cuda()
    out = m(feats, calibs, grid)
    assert out.shape == (B, 256) + tuple(grid.shape[:2])
    with torch.no_grad():
        for b in range(B):
            want = 0
            for v in range(V):
                want = want + sum(mod(f[b, v:v + 1], calibs[v], grid[None])
                                  for mod, f in zip((m.vfa8, m.vfa16, m.vfa32), feats))
            torch.testing.assert_close(out[b:b + 1], want, rtol=1e-5, atol=2e-6)
    out.sum().backward()
    for mod in (m.vfa8, m.vfa16, m.vfa32):
        assert mod.collapse.weight.grad is not None and float(mod.collapse.weight.grad.abs().max()) > 0
