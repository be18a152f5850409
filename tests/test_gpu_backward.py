"""GPU gradient parity: the CUDA backward (through the C ABI and torch.autograd) against autograd through the
float64 reference (golden vectors from the unmodified reference, and the oracle port on larger cases)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_port                      # noqa: E402
import vfa_b200                                   # noqa: E402
from vfa_b200 import geometry, synthetic          # noqa: E402

NAMES = ['MultiviewC', 'MultiviewX', 'Wildtrack']
# fp32 gradients: sums of up to ~1e5 products accumulated with atomics; compare relative to the gradient's scale
REL = 2e-5


def _close(got, want, what):
    got = np.asarray(got, np.float64)
    scale = np.abs(want).max() + 1e-30
    err = np.abs(got - want).max() / scale
    print(f'{what}: max err / max|grad| = {err:.2e}')
    assert err < REL, f'{what}: relative error {err:.3e}'


@pytest.mark.parametrize('name', NAMES)
def test_backward_matches_reference_autograd_golden(golden, name):
    g = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    args = SimpleNamespace(data=name, image_size=g.image_size)
    checked = 0
    for v in range(calibs.shape[0]):
        for s in range(3):
            key = f'{name}/dfeat_{v}_{s}'
            if key not in golden:
                continue
            m = vfa_b200.VFA(6, g.grid_height, g.cube_size, 1.0, args).cuda()
            with torch.no_grad():
                m.collapse.weight.copy_(torch.from_numpy(golden[f'{name}/weight{s}']))
                m.collapse.bias.copy_(torch.from_numpy(golden[f'{name}/bias{s}']))
            feat = torch.from_numpy(golden[f'{name}/feat{s}'])[None].cuda().requires_grad_(True)
            out = m(feat, torch.from_numpy(calibs[v]).cuda(), torch.from_numpy(grid)[None].cuda())
            out.backward(torch.from_numpy(golden[f'{name}/gout_{v}_{s}']).float()[None].cuda())
            _close(feat.grad[0].cpu().numpy(), golden[key], f'{name} v{v} s{s} dFeature')
            _close(m.collapse.weight.grad.cpu().numpy(), golden[f'{name}/dweight_{v}_{s}'], f'{name} v{v} s{s} dWeight')
            _close(m.collapse.bias.grad.cpu().numpy(), golden[f'{name}/dbias_{v}_{s}'], f'{name} v{v} s{s} dBias')
            checked += 1
    assert checked >= 3


def _port_grads(name, feats, calibs, grid, params, gout):
    g = geometry.GEOMETRIES[name]
    f = [t.double().requires_grad_(True) for t in feats]
    p = [(w.double().requires_grad_(True), b.double().requires_grad_(True)) for w, b in params]
    out = ref_port.aggregate(f, calibs, grid, p, g.grid_height, g.cube_size, name, g.image_size, cache_boxes=True)
    out.backward(gout.double())
    return out.detach(), [t.grad for t in f], [w.grad for w, _ in p], [b.grad for _, b in p]


@pytest.mark.parametrize('name,channels', [('MultiviewC', 16), ('Wildtrack', 8)])
def test_fused_backward_matches_port(name, channels):
    """Batched, multi-view, 3-scale fused entry on the full grid; NCHW inputs (exercises the transpose backward)."""
    g = geometry.GEOMETRIES[name]
    V, B = 2, 2
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=V - 1, in_field=True)
    feats = synthetic.features(g, batch=B, n_views=V, channels=channels, seed=21)
    params = synthetic.collapse_params(g, channels=channels, seed=21)
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(B, channels, *grid.shape[:2], generator=gen)
    _, gf, gw, gb = _port_grads(name, feats, calibs, grid, params, gout)

    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], name, g.image_size)
    table = vfa_b200.build_table(geom, calibs.cuda(), grid.cuda())
    fc = [f.cuda().requires_grad_(True) for f in feats]
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    out = vfa_b200.aggregate(fc, table, ws, bs)
    out.backward(gout.cuda())
    for s in range(3):
        _close(fc[s].grad.cpu().numpy(), gf[s].numpy(), f'{name} scale {s} dFeature')
        _close(ws[s].grad.cpu().numpy(), gw[s].numpy(), f'{name} scale {s} dWeight')
        _close(bs[s].grad.cpu().numpy(), gb[s].numpy(), f'{name} scale {s} dBias')


@pytest.mark.parametrize('path,bwd', [('fside_tf32x3', 'gather'), ('umma_tf32x3', 'gather'), ('fside_tf32x3', 'overflow'),
                                      ('fside_tf32x3', 'scatter')])
def test_full_width_backward_matches_port(path, bwd, libenv):
    """C = 256: tcgen05 forward (either formulation; writes the ReLU mask) + CUDA backward on a strided sub-grid vs
    the float64 port.

    ReLU makes the gradient discontinuous where a pre-activation is ~0, and 3xTF32 vs float64 may disagree on the
    sign of a few of the 2.7 M pre-activations; so the oracle gradient is taken through the SAME pass mask the
    forward kernel recorded (checked to agree with the float64 sign wherever |pre-activation| > 1e-5)."""
    name = 'MultiviewC'
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)[::3, ::3].contiguous()
    LW = grid.shape[0] * grid.shape[1]
    calibs = synthetic.ring_calibs(g, n_views=2)
    feats = synthetic.features(g, batch=1, n_views=2, seed=8, sizes=g.feature_sizes()[1:])
    params = synthetic.collapse_params(g, seed=8)[:2]
    gen = torch.Generator().manual_seed(6)
    gout = torch.randn(1, 256, *grid.shape[:2], generator=gen)

    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], name, g.image_size)
    table = vfa_b200.build_table(geom, calibs.cuda(), grid.cuda())
    fc = [t.cuda().requires_grad_(True) for t in feats]
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    # backward variants: CSR gather + tcgen05 dFeature product (default); the same with a CSR too small for most rows
    # (they are completed by the atomic overflow kernel); the older scatter + cuBLAS path
    if bwd == 'overflow':
        libenv.setenv('VFA_BWD_CSR_PER_BOX', '1')
    if bwd == 'scatter':
        libenv.setenv('VFA_BWD_SCATTER', '1')
    flags = vfa_b200.FLAG_GRID_SIDE if path == 'umma_tf32x3' else 0
    out = vfa_b200.aggregate(fc, table, ws, bs, flags=flags)
    assert vfa_b200.last_kernel_path() == path
    out.backward(gout.cuda())
    # the pass mask of the same forward
    mask = torch.empty(1, 2, 2, 8, LW, dtype=torch.int32, device='cuda')
    cl = [vfa_b200.to_channels_last(t.detach()) for t in fc]
    vfa_b200.aggregate_forward_raw(cl, table, [w.detach() for w in ws], [b.detach() for b in bs], flags=flags,
                                   relu_mask=mask)
    bits = ((mask.cpu().long().unsqueeze(4) >> torch.arange(32).view(1, 1, 1, 1, 32, 1)) & 1).bool()
    bits = bits.reshape(1, 2, 2, 256, *grid.shape[:2])

    f = [t.double().requires_grad_(True) for t in feats]
    p = [(w.double().requires_grad_(True), b.double().requires_grad_(True)) for w, b in params]
    out64 = 0
    for v in range(2):
        for s in range(2):
            pre = ref_port.preactivation(f[s][0, v], calibs[v], grid, p[s][0], p[s][1], g.grid_height, g.cube_size, name,
                                         g.image_size)
            sure = pre.detach().abs() > 1e-5
            assert bool(((pre.detach() > 0) == bits[:, v, s])[sure].all()), 'recorded ReLU mask disagrees with float64'
            out64 = out64 + pre * bits[:, v, s]
    out64.backward(gout.double())
    np.testing.assert_allclose(out.detach().cpu().numpy(), out64.detach().numpy(), rtol=1e-5, atol=2e-6)
    for s in range(2):
        _close(fc[s].grad.cpu().numpy(), f[s].grad.numpy(), f'scale {s} dFeature')
        _close(ws[s].grad.cpu().numpy(), p[s][0].grad.numpy(), f'scale {s} dWeight')
        _close(bs[s].grad.cpu().numpy(), p[s][1].grad.numpy(), f'scale {s} dBias')


def test_grad_flags_are_respected():
    """Only tensors that require grad receive one (frozen collapse weights / inference features)."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::4, ::4].contiguous()
    calibs = synthetic.ring_calibs(g, n_views=1)
    feats = synthetic.features(g, batch=1, n_views=1, channels=8, seed=3)
    params = synthetic.collapse_params(g, channels=8, seed=3)
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    table = vfa_b200.build_table(geom, calibs.cuda(), grid.cuda())
    fc = [f.cuda().requires_grad_(s == 0) for s, f in enumerate(feats)]
    ws = [w.cuda().requires_grad_(s == 1) for s, (w, _) in enumerate(params)]
    bs = [b.cuda() for _, b in params]
    vfa_b200.aggregate(fc, table, ws, bs).sum().backward()
    assert fc[0].grad is not None and fc[1].grad is None and fc[2].grad is None
    assert ws[1].grad is not None and ws[0].grad is None and ws[2].grad is None


def test_full_size_backward_variants_agree(libenv):
    """BASELINE-size problem (MultiviewC, 7 views, 3 scales, C = 256, 2 frames): the gather-form backward (CSR + tcgen05
    dFeature / dWeight) against the scatter + SGEMM backward, gradients of a random cotangent."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g)
    feats = synthetic.features(g, batch=2, seed=21)
    params = synthetic.collapse_params(g, seed=21)
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    table = vfa_b200.build_table(geom, calibs.cuda(), grid.cuda())
    gout = torch.randn(2, 256, *grid.shape[:2], generator=torch.Generator().manual_seed(3)).cuda()
    grads = {}
    for tag in ('gather', 'scatter'):
        if tag == 'scatter':
            libenv.setenv('VFA_BWD_SCATTER', '1')
        fc = [t.cuda().requires_grad_(True) for t in feats]
        ws = [w.cuda().requires_grad_(True) for w, _ in params]
        bs = [b.cuda().requires_grad_(True) for _, b in params]
        vfa_b200.aggregate(fc, table, ws, bs).backward(gout)
        grads[tag] = [t.grad for t in fc + ws + bs]
    libenv.delenv('VFA_BWD_SCATTER', raising=False)
    names = [f'dFeature{s}' for s in range(3)] + [f'dWeight{s}' for s in range(3)] + [f'dBias{s}' for s in range(3)]
    for name, a, b in zip(names, grads['gather'], grads['scatter']):
        _close(a.cpu().numpy(), b.cpu().numpy().astype(np.float64), f'full size {name}')
