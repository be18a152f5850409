"""GPU parity of the BENCHED configuration: the default C = 256 path (row-compacted tcgen05 3xTF32 GEMM + staged-tile
pooling) on whole frames -- every view, all three scales (stride 8 included), full BEV grid -- directly against the
float64 port of the reference, forward and backward.

The oracle is per-cell independent, so it is evaluated on a strided subset of the cells (grid[r0::sr, c0::sc]) while the
CUDA path runs the complete grid; the compared values are the CUDA outputs at exactly those cells.  For the backward the
cotangent is non-zero on the subset only, which makes every gradient a function of the subset cells alone.
Tolerance: 1e-5 relative + 1e-6 absolute on the forward (north_star); 2e-5 of max|grad| on the gradients.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_port                      # noqa: E402
import vfa_b200                                   # noqa: E402
from vfa_b200 import geometry, synthetic          # noqa: E402

NAMES = ['MultiviewC', 'MultiviewX', 'Wildtrack']
RTOL, ATOL = 1e-5, 1e-6
SUBSET = {'MultiviewC': (1, 5, 2, 6), 'MultiviewX': (0, 6, 3, 9), 'Wildtrack': (2, 5, 1, 11)}    # r0, sr, c0, sc


def _geom(name, grid_lw):
    g = geometry.GEOMETRIES[name]
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    return vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid_lw, name, g.image_size)


def _port(name, feats, calibs, grid, params):
    g = geometry.GEOMETRIES[name]
    return ref_port.aggregate([t.double() for t in feats], calibs, grid, [(w.double(), b.double()) for w, b in params],
                              g.grid_height, g.cube_size, name, g.image_size, cache_boxes=True)


@pytest.mark.parametrize('name', NAMES)
def test_whole_frame_default_path_vs_float64_port(name):
    """Benched workload shape: all ring cameras x 3 scales x C = 256, batch 2, full grid, default flags."""
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g)
    feats = synthetic.features(g, batch=2, seed=31)
    params = synthetic.collapse_params(g, seed=31)
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    out = vfa_b200.aggregate([f.cuda() for f in feats], table, [w.cuda() for w, _ in params], [b.cuda() for _, b in params])
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
    r0, sr, c0, sc = SUBSET[name]
    sub = grid[r0::sr, c0::sc].contiguous()
    got = out[1, :, r0::sr, c0::sc].cpu().numpy().astype(np.float64)          # frame 1 of the batch
    want = _port(name, [f[1:2] for f in feats], calibs, sub, params)[0].numpy()
    assert got.shape == want.shape
    err = np.abs(got - want)
    bad = err > ATOL + RTOL * np.abs(want)
    print(f'{name}: whole frame, {want.size} outputs of {sub.shape[0]}x{sub.shape[1]} cells: outside tol {bad.mean():.2e}, '
          f'worst abs {err.max():.2e}, max|out| {np.abs(want).max():.3f}')
    assert not bad.any(), f'{bad.sum()} of {bad.size} outside tolerance (worst {err.max():.3e})'


@pytest.mark.parametrize('scale', [0, 1, 2])
def test_every_scale_single_call_vs_float64_port(scale):
    """One (view, scale) at C = 256 for EACH of the three feature scales (stride 8 = 90 x 160 included), in-field camera
    (very large boxes and behind-camera ghosts), full grid on the GPU, every 3rd cell in the oracle."""
    name = 'MultiviewC'
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=1, in_field=True)
    sizes = [g.feature_sizes()[scale]]
    params = synthetic.collapse_params(g, seed=40 + scale)[:1]
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    feats = synthetic.features(g, batch=1, n_views=2, seed=40 + scale, sizes=sizes)
    out = vfa_b200.aggregate([feats[0].cuda()], table, [params[0][0].cuda()], [params[0][1].cuda()])
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
    sub = grid[::3, 1::3].contiguous()
    want = _port(name, feats, calibs, sub, params)[0].numpy()
    got = out[0, :, ::3, 1::3].cpu().numpy().astype(np.float64)
    err = np.abs(got - want)
    bad = err > ATOL + RTOL * np.abs(want)
    print(f'scale {scale} {sizes[0]}: outside tol {bad.mean():.2e}, worst abs {err.max():.2e}')
    assert not bad.any(), f'{bad.sum()} of {bad.size} outside tolerance (worst {err.max():.3e})'


@pytest.mark.parametrize('name', ['MultiviewC', 'Wildtrack'])
def test_whole_frame_backward_vs_float64_port(name):
    """C = 256 forward + backward of the default path at all three scales on the FULL grid against float64 autograd
    through the port.  No quantity of the CUDA run enters the oracle: the cotangent is zeroed wherever any (view, scale)
    pre-activation of the float64 oracle is within 1e-5 of the ReLU kink (where 3xTF32 and float64 may legitimately
    disagree on the sign), and each side then uses its own ReLU mask."""
    g = geometry.GEOMETRIES[name]
    V = 2
    grid = geometry.grid_for(g)
    L, W = grid.shape[:2]
    calibs = synthetic.ring_calibs(g, n_views=V)
    feats = synthetic.features(g, batch=1, n_views=V, seed=51)
    params = synthetic.collapse_params(g, seed=51)
    r0, sr, c0, sc = SUBSET[name]
    sub = grid[r0::sr, c0::sc].contiguous()

    f64 = [t.double().requires_grad_(True) for t in feats]
    p64 = [(w.double().requires_grad_(True), b.double().requires_grad_(True)) for w, b in params]
    out64, near_kink = 0, torch.zeros(1, 256, *sub.shape[:2], dtype=torch.bool)
    for v in range(V):
        for s in range(3):
            pre = ref_port.preactivation(f64[s][0, v], calibs[v], sub, p64[s][0], p64[s][1], g.grid_height, g.cube_size,
                                         name, g.image_size)
            near_kink |= pre.detach().abs() <= 1e-5
            out64 = out64 + torch.relu(pre)
    gen = torch.Generator().manual_seed(9)
    gsub = torch.randn(1, 256, *sub.shape[:2], generator=gen) * (~near_kink)
    out64.backward(gsub.double())
    print(f'{name}: {int(near_kink.sum())} of {near_kink.numel()} outputs masked out of the cotangent (near the ReLU kink)')

    gout = torch.zeros(1, 256, L, W)
    gout[:, :, r0::sr, c0::sc] = gsub
    table = vfa_b200.build_table(_geom(name, (L, W)), calibs.cuda(), grid.cuda())
    fc = [t.cuda().requires_grad_(True) for t in feats]
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    out = vfa_b200.aggregate(fc, table, ws, bs)
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
    out.backward(gout.cuda())
    got = out.detach()[:, :, r0::sr, c0::sc].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, out64.detach().numpy(), rtol=RTOL, atol=ATOL)
    for s in range(3):
        for what, a, b in (('dFeature', fc[s].grad, f64[s].grad), ('dWeight', ws[s].grad, p64[s][0].grad),
                           ('dBias', bs[s].grad, p64[s][1].grad)):
            want = b.numpy()
            err = np.abs(a.cpu().numpy().astype(np.float64) - want).max() / (np.abs(want).max() + 1e-30)
            print(f'{name} scale {s} {what}: max err / max|grad| = {err:.2e}')
            assert err < 2e-5, f'{name} scale {s} {what}: {err:.3e}'


def test_tile_pooling_overflow_nhwc_and_determinism(libenv):
    """Staged-tile pooling (pool_tile_kernel, the default): bit-reproducible run to run; with its pools forced to overflow
    (VFA_POOL_TILE_CAP = 1 %: nearly every tile is left to the walking kernel) and with the quads' list kernel
    (VFA_POOL_TILE = 0) the sums only change their order (fp32 rounding); the [B, L, W, C] output (VFA_FLAG_OUT_NHWC) holds
    the same bits as [B, C, L, W]; so does every tile order of the scheduler (VFA_TILE_ORDER); all on a ragged grid with an in-field camera, inference and training (mask) variants."""
    g = geometry.WILDTRACK
    grid = geometry.grid_for(g)[3:80, 5:132].contiguous()             # 77 x 127 cells: partial tiles on both sides
    calibs = synthetic.ring_calibs(g, n_views=3, in_field=True)
    V = calibs.shape[0]
    params = synthetic.collapse_params(g, seed=8)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    feats = [f.cuda() for f in synthetic.features(g, batch=2, n_views=V, seed=8)]
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    for k in ('VFA_POOL_TILE_CAP', 'VFA_POOL_TILE', 'VFA_TILE_ORDER'):
        libenv.delenv(k, raising=False)
    ref = vfa_b200.aggregate(feats, table, ws, bs)
    assert torch.equal(vfa_b200.aggregate(feats, table, ws, bs), ref)
    # the order in which the scheduler deals the tiles (heavy first: rank sort; the bitonic sort for > 1024 tiles; identity
    # for > 8192) changes no bit -- every tile is pooled exactly once, by one CTA
    for mode in ('1', '2'):
        libenv.setenv('VFA_TILE_ORDER', mode)
        assert torch.equal(vfa_b200.aggregate(feats, table, ws, bs), ref), f'VFA_TILE_ORDER={mode}'
    libenv.delenv('VFA_TILE_ORDER', raising=False)
    simt = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT)
    err = (ref - simt).abs()
    assert bool((err <= ATOL + RTOL * simt.abs()).all()), f'worst {float(err.max()):.2e}'
    cl = [vfa_b200.to_channels_last(f) for f in feats]
    nhwc = torch.empty(2, grid.shape[0], grid.shape[1], 256, device='cuda')
    vfa_b200.aggregate_forward_raw(cl, table, ws, bs, flags=vfa_b200.FLAG_OUT_NHWC, out=nhwc)
    assert torch.equal(nhwc.permute(0, 3, 1, 2), ref)

    def train_run():
        f = [t.detach().clone().requires_grad_(True) for t in feats]
        out = vfa_b200.aggregate(f, table, ws, bs)
        out.backward(torch.ones_like(out))
        return out.detach(), f[0].grad

    out_t, grad_t = train_run()
    assert torch.equal(out_t, ref)
    for key, val in (('VFA_POOL_TILE_CAP', '1'), ('VFA_POOL_TILE', '0')):
        libenv.setenv(key, val)
        other = vfa_b200.aggregate(feats, table, ws, bs)
        err = (other - ref).abs()
        assert bool((err <= 1e-6 + 2e-6 * ref.abs()).all()), f'{key}={val}: worst {float(err.max()):.2e}'
        vfa_b200.aggregate_forward_raw(cl, table, ws, bs, flags=vfa_b200.FLAG_OUT_NHWC, out=nhwc)
        assert torch.equal(nhwc.permute(0, 3, 1, 2), other)
        out_c, grad_c = train_run()
        assert torch.equal(out_c, other)
        # a pre-activation within rounding of zero may flip its ReLU bit when the summation order changes: a handful of
        # gradient elements move, everything else agrees to rounding
        moved = ((grad_c - grad_t).abs() > 2e-5 * float(grad_t.abs().max())).float().mean()
        assert float(moved) < 1e-4, f'{key}={val}: {float(moved):.2e} of dFeature elements differ'
        libenv.delenv(key, raising=False)


def test_channels_last_output_through_autograd():
    """VFA_FLAG_OUT_NHWC through the autograd entry (what VFANet hands its heads): the result is a [B, C, L, W] tensor in
    torch.channels_last with the bits of the [B, C, L, W]-contiguous result, and the backward reads a channels-last
    cotangent in place -- same gradients (atomics reorder the sums: compared to rounding)."""
    g = geometry.MULTIVIEWX
    grid = geometry.grid_for(g)[::2, ::2].contiguous()
    calibs = synthetic.ring_calibs(g, n_views=2)
    feats = synthetic.features(g, batch=2, n_views=2, seed=61)
    params = synthetic.collapse_params(g, seed=61)
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    gout = torch.randn(2, 256, *grid.shape[:2], generator=torch.Generator().manual_seed(7)).cuda()
    res = {}
    for tag, flags in (('nchw', 0), ('nhwc', vfa_b200.FLAG_OUT_NHWC)):
        f = [t.cuda().requires_grad_(True) for t in feats]
        w = [x.cuda().requires_grad_(True) for x, _ in params]
        b = [x.cuda().requires_grad_(True) for _, x in params]
        out = vfa_b200.aggregate(f, table, w, b, flags=flags)
        go = gout.contiguous(memory_format=torch.channels_last) if tag == 'nhwc' else gout
        out.backward(go)
        res[tag] = (out.detach(), [t.grad for t in f + w + b])
    assert res['nhwc'][0].shape == res['nchw'][0].shape
    assert res['nhwc'][0].is_contiguous(memory_format=torch.channels_last) and not res['nhwc'][0].is_contiguous()
    assert torch.equal(res['nhwc'][0], res['nchw'][0])
    for a, b_ in zip(res['nhwc'][1], res['nchw'][1]):
        assert float((a - b_).abs().max()) <= 1e-5 * float(b_.abs().max())


@pytest.mark.parametrize('storage', ['f32', 'bf16'])
def test_bf16_mma_variant(storage):
    """VFA_FLAG_BF16_MMA (north_star: "bf16 features within a stated tolerance"): bf16 operands in ONE tcgen05 kind::f16
    pass with fp32 accumulation, Y stored in bf16, pooling / bias / ReLU / sums in fp32.  Stated tolerance against the
    float64 port of the reference on the SAME fp32 inputs: 4e-3 of the output scale (max |out|) on every element (measured 1.6e-3),
    mean |error| below 5e-4 of it (measured 2e-4) -- operands and Y carry 8 bits of mantissa (relative rounding 2^-9 = 2e-3), the
    contraction and the sum over layers average the rounding errors down.  Whole frame, all views and scales, batch 2."""
    name = 'MultiviewC'
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g)
    feats = synthetic.features(g, batch=2, seed=71)
    params = synthetic.collapse_params(g, seed=71)
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    cl = [vfa_b200.to_channels_last(f.cuda()) for f in feats]
    if storage == 'bf16':
        cl = [t.to(torch.bfloat16) for t in cl]
    with torch.no_grad():
        out = vfa_b200.aggregate(cl, table, ws, bs, flags=vfa_b200.FLAG_BF16_MMA, channels_last=True)
        assert vfa_b200.last_kernel_path() == ('fside_bf16mma_bf16feat' if storage == 'bf16' else 'fside_bf16mma')
        again = vfa_b200.aggregate(cl, table, ws, bs, flags=vfa_b200.FLAG_BF16_MMA, channels_last=True)
        assert torch.equal(out, again)                                   # deterministic
        full = vfa_b200.aggregate([f.cuda() for f in feats], table, ws, bs)
    r0, sr, c0, sc = SUBSET[name]
    want = _port(name, [f[1:2] for f in feats], calibs, grid[r0::sr, c0::sc].contiguous(), params)[0].numpy()
    got = out[1, :, r0::sr, c0::sc].cpu().numpy().astype(np.float64)
    scale = np.abs(want).max()
    err = np.abs(got - want)
    print(f'bf16 MMA ({storage} features): max |err| / max|out| = {err.max() / scale:.2e}, mean = {err.mean() / scale:.2e}; '
          f'vs the fp32 path: {float((out - full).abs().max()) / scale:.2e}')
    assert err.max() <= 4e-3 * scale and err.mean() <= 5e-4 * scale
    with pytest.raises(RuntimeError, match='forward-only'):
        vfa_b200.aggregate([t.float().requires_grad_(True) for t in cl], table, ws, bs, flags=vfa_b200.FLAG_BF16_MMA,
                           channels_last=True)


@pytest.mark.parametrize('name', NAMES)
def test_random_rigs_tile_pooling_vs_float64_port(name):
    """Randomised cameras (outside the field, inside it and low -- voxels behind the camera, boxes tens of texels wide --,
    looking along a grid axis), random ragged sub-grids, batch 2, C = 256, three scales: the default path (compacted GEMM +
    staged-tile pooling, including tiles whose chunk lists overflow into the walking kernel) against the float64 port on
    every cell.  The rigs are the generator of tests/test_gpu_parity.py::test_table_bit_exact_random_rigs."""
    g = geometry.GEOMETRIES[name]
    full = geometry.grid_for(g)
    L, W = full.shape[:2]
    rng = np.random.RandomState(4321 + NAMES.index(name))
    from oracle import vfa_oracle as onp
    world = onp.to_world(full.numpy().reshape(-1, 3).astype(np.float64), name).reshape(L, W, 3)
    lo, hi = world.reshape(-1, 3).min(0), world.reshape(-1, 3).max(0)
    span = float(max(hi[0] - lo[0], hi[1] - lo[1]))
    cams = []
    for k in range(3):
        if k == 0:        # outside the field, elevated
            ang = rng.uniform(0, 2 * np.pi)
            eye = np.array([(lo[0] + hi[0]) / 2 + np.cos(ang) * span * rng.uniform(0.6, 1.0),
                            (lo[1] + hi[1]) / 2 + np.sin(ang) * span * rng.uniform(0.6, 1.0), span * rng.uniform(0.1, 0.3)])
            target = np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1]), 0.0])
        elif k == 1:      # inside the field, low
            eye = np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1]), span * rng.uniform(0.01, 0.05)])
            target = np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1]), 0.0])
        else:             # above a cell origin, looking along +x
            i, j = rng.randint(L), rng.randint(W)
            eye = np.array([world[i, j, 0], world[i, j, 1], span * 0.03])
            target = eye + np.array([1.0, 0.0, 0.0])
        cams.append(synthetic.look_at(eye, target, rng.uniform(0.6, 1.4) * g.image_size[1] * 0.7, g.image_size).astype(np.float32))
    calibs = torch.from_numpy(np.stack(cams))
    i0, j0 = rng.randint(0, L // 2), rng.randint(0, W // 2)
    grid = full[i0:i0 + rng.randint(17, 30), j0:j0 + rng.randint(17, 30)].contiguous()        # partial tiles on both sides
    feats = synthetic.features(g, batch=2, n_views=3, seed=81)
    params = synthetic.collapse_params(g, seed=81)
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    out = vfa_b200.aggregate([f.cuda() for f in feats], table, [w.cuda() for w, _ in params], [b.cuda() for _, b in params])
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
    want = _port(name, feats, calibs, grid, params).numpy()
    got = out.cpu().numpy().astype(np.float64)
    finite = np.isfinite(want)                 # a 0/0 projection makes the reference's row NaN (documented deviation)
    err = np.abs(got - want)[finite]
    bad = err > ATOL + RTOL * np.abs(want[finite])
    print(f'{name}: grid {tuple(grid.shape[:2])}, {finite.mean():.4f} finite, outside tol {bad.mean():.2e}, worst abs {err.max():.2e}, '
          f'max|out| {np.abs(want[finite]).max():.2f}')
    assert not bad.any(), f'{bad.sum()} of {bad.size} outside tolerance (worst {err.max():.3e})'
