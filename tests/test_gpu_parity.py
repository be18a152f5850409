"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Bit-exact: boxes, area, visible, tap indices.  Floating point: fused fp32 features within 1e-5 relative +
1e-6 absolute of the float64 hybrid oracle (BASELINE.json north_star).
"""
import hashlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_port                      # noqa: E402
from oracle import vfa_oracle as onp             # noqa: E402
import vfa_b200                                   # noqa: E402
from vfa_b200 import geometry, synthetic          # noqa: E402

NAMES = ['MultiviewC', 'MultiviewX', 'Wildtrack']
SMALL_SIZES = [(45, 80), (30, 52), (23, 40)]
RTOL, ATOL = 1e-5, 1e-6       # north_star tolerance for fused fp32 features


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _geom(name, grid_lw, crange=(-1.0, 0.95)):
    g = geometry.GEOMETRIES[name]
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    return vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid_lw, name, g.image_size, crange)


def _within(got, want, rtol=RTOL, atol=ATOL):
    err = np.abs(got - want)
    ok = err <= atol + rtol * np.abs(want)
    return ok, err


@pytest.mark.parametrize('name', NAMES)
def test_table_bit_exact_small(golden, name):
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    L, W = grid.shape[:2]
    table = vfa_b200.build_table(_geom(name, (L, W)), torch.from_numpy(calibs).cuda(), torch.from_numpy(grid).cuda())
    boxes = table.boxes.cpu().numpy()
    for v in range(calibs.shape[0]):
        assert np.array_equal(_bits(boxes[v]), _bits(golden[f'{name}/boxes{v}'])), (name, v)
    for s, (fh, fw) in enumerate(SMALL_SIZES):
        area, vis, taps = table.scale_table(fh, fw)
        for v in range(calibs.shape[0]):
            assert np.array_equal(_bits(area[v].cpu().numpy()), _bits(golden[f'{name}/area{v}_{s}']))
            assert np.array_equal(vis[v].cpu().numpy(), golden[f'{name}/visible{v}_{s}'])
            b = golden[f'{name}/boxes{v}']
            want = np.stack([onp.tap_index(b[..., 0], fw), onp.tap_index(b[..., 1], fh),
                             onp.tap_index(b[..., 2], fw), onp.tap_index(b[..., 3], fh)], -1)
            assert np.array_equal(taps[v].cpu().numpy(), want)


@pytest.mark.parametrize('name', NAMES)
def test_table_full_size_digests(digests, name):
    """Config-of-record grids, ring + in-field camera: identical bits to the reference at full size."""
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, in_field=True)
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    boxes = table.boxes.cpu().numpy()
    for v in range(calibs.shape[0]):
        assert _digest(_bits(boxes[v])) == digests[f'{name}/boxes{v}'], (name, v)
    for s, (fh, fw) in enumerate(g.feature_sizes()):
        _, vis, taps = table.scale_table(fh, fw)
        vis, taps = vis.cpu().numpy(), taps.cpu().numpy()
        for v in range(calibs.shape[0]):
            assert _digest(vis[v].astype(np.uint8)) == digests[f'{name}/visible{v}_{s}']
            assert _digest(taps[v]) == digests[f'{name}/taps{v}_{s}']


@pytest.mark.parametrize('name', NAMES)
def test_forward_matches_hybrid_oracle_golden(golden, name):
    """One reference `VFA.forward` per (camera, scale) of the golden set, through the drop-in module."""
    g = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    args = SimpleNamespace(data=name, image_size=g.image_size)
    n_checked, n_bad, worst = 0, 0, 0.0
    ref32_bad = 0
    for s in range(3):
        m = vfa_b200.VFA(6, g.grid_height, g.cube_size, 1.0, args).cuda()
        with torch.no_grad():
            m.collapse.weight.copy_(torch.from_numpy(golden[f'{name}/weight{s}']))
            m.collapse.bias.copy_(torch.from_numpy(golden[f'{name}/bias{s}']))
        feat = torch.from_numpy(golden[f'{name}/feat{s}'])[None].cuda()
        for v in range(calibs.shape[0]):
            key = f'{name}/out64_{v}_{s}'
            if key not in golden:
                continue
            with torch.no_grad():
                out = m(feat, torch.from_numpy(calibs[v]).cuda(), torch.from_numpy(grid)[None].cuda())
            assert out.shape == (1, 6) + grid.shape[:2]
            want = golden[key]
            ok, err = _within(out[0].cpu().numpy().astype(np.float64), want)
            n_checked += ok.size
            n_bad += int((~ok).sum())
            worst = max(worst, float(err.max()))
            ok32, _ = _within(golden[f'{name}/out32_{v}_{s}'].astype(np.float64), want)
            ref32_bad += int((~ok32).sum())
    print(f'{name}: {n_bad}/{n_checked} outside 1e-5/1e-6 (worst abs {worst:.2e}); the reference fp32 itself: {ref32_bad}')
    assert n_bad == 0, f'{n_bad}/{n_checked} elements outside tolerance, worst abs err {worst:.3e}'


def _port_frame(name, feats, calibs, grid, params, dtype=torch.float64):
    g = geometry.GEOMETRIES[name]
    f = [t.to(dtype) for t in feats]
    p = [(w.to(dtype), b.to(dtype)) for w, b in params]
    return ref_port.aggregate(f, calibs, grid, p, g.grid_height, g.cube_size, name, g.image_size, cache_boxes=True)


@pytest.mark.parametrize('name', NAMES)
def test_fused_aggregate_matches_port(name):
    """Fused multi-view / multi-scale / batched entry vs the float64 port of the reference loop, full-size grid,
    real feature-map sizes, reduced channel count so the oracle finishes in seconds."""
    g = geometry.GEOMETRIES[name]
    Cc, V, B = 16, 3, 2
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=V - 1, in_field=True)
    feats = synthetic.features(g, batch=B, n_views=V, channels=Cc, seed=3)
    params = synthetic.collapse_params(g, channels=Cc, seed=3)
    want = _port_frame(name, feats, calibs, grid, params).numpy()
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    out = vfa_b200.aggregate([f.cuda() for f in feats], table, [w.cuda() for w, _ in params],
                             [b.cuda() for _, b in params])
    assert out.shape == want.shape
    ok, err = _within(out.cpu().numpy().astype(np.float64), want)
    frac = 1.0 - ok.mean()
    print(f'{name}: outside tol {frac:.2e}, worst abs {err.max():.2e}, path {vfa_b200.last_kernel_path()}')
    assert frac == 0.0, f'{frac:.3e} of elements outside tolerance (worst {err.max():.3e})'
    # channels-last input is consumed zero-copy and gives identical bits
    cl = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in feats]
    out2 = vfa_b200.aggregate(cl, table, [w.cuda() for w, _ in params], [b.cuda() for _, b in params], channels_last=True)
    assert torch.equal(out, out2)


PATH_FLAGS = {'simt_fp32': vfa_b200.FLAG_FORCE_SIMT,
              'fside_tf32x3': vfa_b200.FLAG_FORCE_UMMA,                              # default for C = 256
              'umma_tf32x3': vfa_b200.FLAG_FORCE_UMMA | vfa_b200.FLAG_GRID_SIDE}


@pytest.mark.parametrize('name,path', [('MultiviewC', 'fside_tf32x3'), ('MultiviewC', 'umma_tf32x3'),
                                       ('MultiviewC', 'simt_fp32'), ('MultiviewX', 'fside_tf32x3'),
                                       ('MultiviewX', 'umma_tf32x3'), ('Wildtrack', 'fside_tf32x3'),
                                       ('Wildtrack', 'umma_tf32x3')])
def test_full_width_single_call_matches_port(name, path):
    """C = 256 (the real channel count), full grid, one (view, scale) -- the three kernel families (feature-side
    tcgen05, grid-side fused tcgen05, generic fp32) vs the float64 port."""
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=1)
    feats = synthetic.features(g, batch=1, n_views=1, seed=5, sizes=[g.feature_sizes()[1]])
    params = synthetic.collapse_params(g, seed=5)[:1]
    want = ref_port.vfa_forward(feats[0][0, 0].double(), calibs[0], grid, params[0][0].double(), params[0][1].double(),
                                g.grid_height, g.cube_size, name, g.image_size).numpy()
    args = SimpleNamespace(data=name, image_size=g.image_size)
    m = vfa_b200.VFA(256, g.grid_height, g.cube_size, 1 / 16., args).cuda()
    m.flags = PATH_FLAGS[path]
    with torch.no_grad():
        m.collapse.weight.copy_(params[0][0])
        m.collapse.bias.copy_(params[0][1])
        out = m(feats[0][0].cuda(), calibs[0].cuda(), grid[None].cuda())
    ok, err = _within(out.cpu().numpy().astype(np.float64), want)
    frac = 1.0 - ok.mean()
    assert vfa_b200.last_kernel_path() == path
    print(f'C=256 single call {name} {path}: outside tol {frac:.2e}, worst abs {err.max():.2e}')
    assert frac == 0.0, f'{frac:.3e} of elements outside tolerance (worst {err.max():.3e})'


@pytest.mark.parametrize('name', NAMES)
def test_tensor_core_path_matches_simt_path_full_size(name):
    """Whole frame (all views, 3 scales, C=256, batch 2 -> no view split; batch 1 -> view-split + atomics): both
    tcgen05 3xTF32 formulations against the fp32 FFMA kernel, inside the north_star tolerance of each other."""
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=g.n_views - 1, in_field=True)      # includes the ghost-producing camera
    feats = [f.cuda() for f in synthetic.features(g, batch=2, seed=11)]
    params = synthetic.collapse_params(g, seed=11)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    simt = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT)
    assert vfa_b200.last_kernel_path() == 'simt_fp32'
    for path in ('fside_tf32x3', 'umma_tf32x3'):
        got = vfa_b200.aggregate(feats, table, ws, bs, flags=PATH_FLAGS[path])
        assert vfa_b200.last_kernel_path() == path
        ok, err = _within(got.cpu().numpy().astype(np.float64), simt.cpu().numpy().astype(np.float64))
        print(f'{name}: {path} vs simt outside tol {1 - ok.mean():.2e}, worst abs {err.max():.2e}')
        assert ok.all()
        one = vfa_b200.aggregate([f[:1] for f in feats], table, ws, bs, flags=PATH_FLAGS[path])
        ok, err = _within(one.cpu().numpy().astype(np.float64), simt[:1].cpu().numpy().astype(np.float64))
        assert ok.all()


def test_full_size_properties():
    """BASELINE-size problem (MultiviewC, 7 views, 3 scales, C=256): properties that need no oracle."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g)
    feats = [f.cuda() for f in synthetic.features(g, batch=2, seed=1)]
    params = synthetic.collapse_params(g, seed=1)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    out = vfa_b200.aggregate(feats, table, ws, bs)
    assert out.shape == (2, 256, 156, 156) and bool(torch.isfinite(out).all()) and float(out.min()) >= 0.0
    # frames are independent: batch == per-frame calls (a single frame is view-split with atomic accumulation,
    # so the summation order differs; fp32 rounding only)
    one = vfa_b200.aggregate([f[1:2] for f in feats], table, ws, bs)
    torch.testing.assert_close(one[0], out[1], rtol=1e-5, atol=1e-5)
    # summing over views is order-free up to fp32 rounding: permuting cameras permutes nothing else
    perm = [3, 0, 6, 2, 5, 1, 4]
    tperm = vfa_b200.ProjectionTable(table.geom, table.boxes[perm].contiguous())
    outp = vfa_b200.aggregate([f[:1, perm] for f in feats], tperm, ws, bs)
    torch.testing.assert_close(outp[0], out[0], rtol=1e-5, atol=1e-5)
    # positive homogeneity with zero bias: relu(W (a x)) = a relu(W x)
    zb = [torch.zeros_like(b) for b in bs]
    o1 = vfa_b200.aggregate([f[:1] for f in feats], table, ws, zb)
    o2 = vfa_b200.aggregate([f[:1] * 4.0 for f in feats], table, ws, zb)
    torch.testing.assert_close(o2, o1 * 4.0, rtol=1e-5, atol=1e-5)
    # a camera that sees nothing contributes relu(bias) per scale: all-zero features -> sum_v sum_s relu(b_s)
    zf = [torch.zeros_like(f[:1]) for f in feats]
    oz = vfa_b200.aggregate(zf, table, ws, bs)
    want = sum(torch.relu(b) for b in bs) * calibs.shape[0]
    torch.testing.assert_close(oz[0], want[:, None, None].expand_as(oz[0]), rtol=1e-6, atol=1e-6)


def test_module_state_dict_and_errors():
    g = geometry.MULTIVIEWC
    args = SimpleNamespace(data=g.name, image_size=g.image_size)
    m = vfa_b200.VFA(256, g.grid_height, np.array(g.cube_size), 1 / 8., args)
    sd = m.state_dict()
    assert {k: (tuple(v.shape), v.dtype) for k, v in sd.items()} == {
        'z_corners': ((5, 1, 1, 3), torch.int64), 'corners_offset': ((1, 1, 1, 1, 8, 3), torch.float32),
        'collapse.weight': ((256, 1280), torch.float32), 'collapse.bias': ((256,), torch.float32)}
    grid = geometry.grid_for(g)[None]
    feat = torch.zeros(1, 256, 90, 160)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(feat, torch.zeros(3, 4), grid)
    m = m.cuda()
    calib = synthetic.ring_calibs(g, n_views=1)[0].cuda()
    with pytest.raises(NotImplementedError):
        m(feat.cuda(), calib, grid.cuda(), visualize=True)
    with pytest.raises(vfa_b200.VFAError, match='crange'):
        m(feat.cuda(), calib, grid.cuda(), crange=(-1, 1.0))
    out = m(feat.cuda(), calib, grid.cuda())
    assert out.shape == (1, 256, 156, 156)


def test_streaming_aggregator_matches_direct_call():
    """Pinned-host in, pinned-host out, copies overlapped with compute: same bits as the direct call."""
    g = geometry.WILDTRACK
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g).cuda()
    params = synthetic.collapse_params(g, seed=2)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs, grid.cuda())
    batches = [[f.pin_memory() for f in synthetic.features(g, batch=4, seed=30 + i)] for i in range(3)]
    agg = vfa_b200.StreamingAggregator(table, ws, bs, [tuple(f.shape) for f in batches[0]], depth=2)
    tickets, got = [], []
    for i, hb in enumerate(batches):
        tickets.append(agg.submit(hb))
        if i >= 1:                                   # read results one step behind, as a pipeline would
            got.append(agg.result(tickets[i - 1]).clone())
    got.append(agg.result(tickets[-1]).clone())
    agg.drain()
    with pytest.raises(ValueError):
        agg.result(0)
    for hb, out in zip(batches, got):
        want = vfa_b200.aggregate([f.cuda() for f in hb], table, ws, bs)
        assert torch.equal(out.cuda(), want)


def test_bf16_feature_storage():
    """bf16 feature maps (VFA_FLAG_BF16_FEATURES): the kernel widens bf16 -> fp32 exactly, so against the float64 oracle
    fed the SAME bf16-rounded features the fp32 tolerance holds; against the fp32-feature oracle the stated tolerance is
    the input quantisation, 2^-9 relative per feature value -> 4e-3 of the output scale (measured ~1e-3)."""
    name = 'MultiviewC'
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=2)
    feats = synthetic.features(g, batch=1, n_views=2, seed=9)
    params = synthetic.collapse_params(g, seed=9)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    f16 = [f.cuda().permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16) for f in feats]       # [B,V,H,W,C] bf16
    out = vfa_b200.aggregate(f16, table, ws, bs, channels_last=True)
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3_bf16feat' and out.dtype == torch.float32
    # (a) same quantised inputs, fp32 path: identical arithmetic after the widening -> fp32 tolerance
    fq = [t.float() for t in f16]
    ref = vfa_b200.aggregate(fq, table, ws, bs, channels_last=True)
    ok, err = _within(out.cpu().numpy().astype(np.float64), ref.cpu().numpy().astype(np.float64))
    assert ok.all(), f'worst {err.max():.2e}'
    # (b) stated tolerance against unquantised fp32 features
    full = vfa_b200.aggregate([f.cuda() for f in feats], table, ws, bs)
    scale = float(full.abs().max())
    rel = float((out - full).abs().max()) / scale
    print(f'bf16 feature storage: max |err| / max|out| = {rel:.2e}')
    assert rel < 4e-3
    with pytest.raises(RuntimeError, match='forward-only'):
        vfa_b200.aggregate(f16, table, [w.requires_grad_(True) for w in ws], bs, channels_last=True)


def test_feature_side_chunking_and_ragged_grids(libenv):
    """Feature-side forward (default, C = 256): frame chunks (Y budget forced down to one frame per chunk), grids whose
    sides are not multiples of the 2 x 2 quads / 4 x 8 CTA tiles, a texel-row count that is not a multiple of the
    256-row GEMM tile -- all bit-identical to the unchunked run and inside the tolerance of the fp32 FFMA kernel."""
    g = geometry.MULTIVIEWC
    full = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, n_views=2, in_field=True)
    params = synthetic.collapse_params(g, seed=4)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    feats = [f.cuda() for f in synthetic.features(g, batch=3, n_views=3, seed=4)]
    for sl in ((slice(0, 45, 2), slice(1, 60, 3)),            # 23 x 20 cells
               (slice(0, 7), slice(0, 9)),                    # 7 x 9: one partial CTA tile, odd quads on both sides
               (slice(100, 101), slice(0, 156))):             # a single row
        grid = full[sl].contiguous()
        table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
        libenv.delenv('VFA_FSIDE_Y_BUDGET_MB', raising=False)
        ref = vfa_b200.aggregate(feats, table, ws, bs)
        assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
        simt = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT)
        ok, err = _within(ref.cpu().numpy().astype(np.float64), simt.cpu().numpy().astype(np.float64))
        assert ok.all(), f'grid {tuple(grid.shape[:2])}: worst {err.max():.2e}'
        # Y of one MultiviewC frame with 3 views is 291 MB: 300 MB -> one frame per chunk, 3 chunks
        libenv.setenv('VFA_FSIDE_Y_BUDGET_MB', '300')
        chunked = vfa_b200.aggregate(feats, table, ws, bs)
        assert torch.equal(chunked, ref)
    libenv.delenv('VFA_FSIDE_Y_BUDGET_MB', raising=False)


def test_list_pooling_equals_walking_pooling_bit_for_bit(libenv):
    """The list pooling (VFA_POOL_TILE=0) reads precomputed texel lists (pool_list_kernel); quads whose list does not fit its slot are
    pooled by the walking kernel (pool_quad_kernel, completion pass).  Both apply the same weights in the same order, so
    any split of the quads between them gives the same bits: default slots (every quad listed), slots of 1 and 3 entries
    per iteration (most / some quads walked), for inference and for the training variant that writes the ReLU mask, on a
    rig with an in-field camera (very large boxes) and a ragged grid."""
    g = geometry.WILDTRACK
    grid = geometry.grid_for(g)[3:80, 5:132].contiguous()             # 77 x 127 cells: odd on both sides
    calibs = synthetic.ring_calibs(g, n_views=3, in_field=True)
    V = calibs.shape[0]
    params = synthetic.collapse_params(g, seed=8)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    feats = [f.cuda() for f in synthetic.features(g, batch=2, n_views=V, seed=8)]
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    libenv.delenv('VFA_POOL_LIST_CAP', raising=False)
    libenv.setenv('VFA_POOL_TILE', '0')              # the list kernel (the staged-tile pooling is the default)
    ref = vfa_b200.aggregate(feats, table, ws, bs)
    assert vfa_b200.last_kernel_path() == 'fside_tf32x3'
    simt = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT)
    ok, err = _within(ref.cpu().numpy().astype(np.float64), simt.cpu().numpy().astype(np.float64))
    assert ok.all(), f'worst {err.max():.2e}'
    gout = torch.randn(ref.shape, generator=torch.Generator(device='cuda').manual_seed(2), device='cuda')

    def train_run():
        f = [t.detach().clone().requires_grad_(True) for t in feats]
        out = vfa_b200.aggregate(f, table, ws, bs)
        out.backward(gout)
        return out.detach(), f[0].grad

    out_t, grad_t = train_run()
    assert torch.equal(out_t, ref)
    for cap in ('1', '3'):
        libenv.setenv('VFA_POOL_LIST_CAP', cap)
        assert torch.equal(vfa_b200.aggregate(feats, table, ws, bs), ref), f'slot of {cap} entries per iteration'
        out_c, grad_c = train_run()
        assert torch.equal(out_c, ref)
        # same ReLU mask -> same masked gradient (the backward's atomics reorder sums: compare to rounding)
        assert float((grad_c - grad_t).abs().max()) <= 1e-5 * float(grad_t.abs().max())
    libenv.delenv('VFA_POOL_LIST_CAP', raising=False)


def test_relu_mask_agrees_between_formulations():
    """The ReLU pass bits the backward consumes: feature-side and grid-side forwards may only disagree where the
    pre-activation is within rounding of zero."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::2, ::2].contiguous()
    LW = grid.shape[0] * grid.shape[1]
    calibs = synthetic.ring_calibs(g, n_views=2)
    feats = [vfa_b200.to_channels_last(f.cuda()) for f in synthetic.features(g, batch=2, n_views=2, seed=12)]
    params = synthetic.collapse_params(g, seed=12)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs.cuda(), grid.cuda())
    masks = []
    for flags in (0, vfa_b200.FLAG_GRID_SIDE):
        m = torch.zeros(2, 2, 3, 8, LW, dtype=torch.int32, device='cuda')
        vfa_b200.aggregate_forward_raw(feats, table, ws, bs, flags=flags, relu_mask=m)
        masks.append(m)
    diff = (masks[0] ^ masks[1]).cpu().numpy().view(np.uint32)
    flipped = int(np.unpackbits(diff.view(np.uint8)).sum())
    total = diff.size * 32
    print(f'mask bits that differ: {flipped} of {total}')
    assert flipped <= total * 1e-4


@pytest.mark.parametrize('name', NAMES)
def test_table_bit_exact_random_rigs(name):
    """Randomised cameras (outside, inside and just above the field; looking across, down, and along a grid axis so that
    corners land behind the camera and on the principal plane), random sub-grids: boxes, area, visibility and tap
    indices of the CUDA table kernels are bit-identical to the numpy restatement that the golden vectors pin to the
    reference."""
    g = geometry.GEOMETRIES[name]
    full = geometry.grid_for(g).numpy()
    L, W = full.shape[:2]
    rng = np.random.RandomState(1234 + NAMES.index(name))
    # field extent in the dataset's world units (the cameras of synthetic.ring_calibs live in the same frame)
    ring = synthetic.ring_calibs(g).numpy()
    world = onp.to_world(full.reshape(-1, 3).astype(np.float64) + 0.0, name).reshape(L, W, 3)
    lo, hi = world.reshape(-1, 3).min(0), world.reshape(-1, 3).max(0)
    span = float(max(hi[0] - lo[0], hi[1] - lo[1]))
    cams = []
    for k in range(10):
        where = k % 3
        if where == 0:      # outside the field, elevated
            ang = rng.uniform(0, 2 * np.pi)
            eye = np.array([(lo[0] + hi[0]) / 2 + np.cos(ang) * span * rng.uniform(0.6, 1.2),
                            (lo[1] + hi[1]) / 2 + np.sin(ang) * span * rng.uniform(0.6, 1.2), span * rng.uniform(0.05, 0.4)])
        elif where == 1:    # inside the field, low: voxels behind the camera
            eye = np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1]), span * rng.uniform(0.005, 0.05)])
        else:               # exactly above a grid cell origin, looking straight along +x: corners on the principal plane
            i, j = rng.randint(L), rng.randint(W)
            eye = np.array([world[i, j, 0], world[i, j, 1], span * 0.02])
        if where == 2:
            target = eye + np.array([1.0, 0.0, 0.0])
        else:
            target = np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1]), 0.0])
        cams.append(synthetic.look_at(eye, target, rng.uniform(0.5, 1.6) * g.image_size[1] * 0.7, g.image_size).astype(np.float32))
    calibs = np.stack(cams)
    assert calibs.shape[1:] == ring.shape[1:]
    i0, j0 = rng.randint(0, L // 2), rng.randint(0, W // 2)
    grid = np.ascontiguousarray(full[i0:i0 + rng.randint(8, 40):1, j0:j0 + rng.randint(8, 40):1])
    table = vfa_b200.build_table(_geom(name, grid.shape[:2]), torch.from_numpy(calibs).cuda(), torch.from_numpy(grid).cuda())
    boxes = table.boxes.cpu().numpy()
    n_nan = 0
    for v in range(len(cams)):
        want = onp.project_boxes(calibs[v], grid, g.grid_height, g.cube_size, name, g.image_size)
        assert np.array_equal(_bits(boxes[v]), _bits(want)), (name, v)
        n_nan += int(np.isnan(want).sum())
    for fh, fw in g.feature_sizes():
        area, vis, taps = table.scale_table(fh, fw)
        for v in range(len(cams)):
            b = boxes[v]
            a_want, v_want = onp.area_visible(b, fh, fw)
            assert np.array_equal(_bits(area[v].cpu().numpy()), _bits(a_want))
            assert np.array_equal(vis[v].cpu().numpy(), v_want)
            t_want = np.stack([onp.tap_index(b[..., 0], fw), onp.tap_index(b[..., 1], fh),
                               onp.tap_index(b[..., 2], fw), onp.tap_index(b[..., 3], fh)], -1)
            assert np.array_equal(taps[v].cpu().numpy(), t_want)
    print(f'{name}: 10 random cameras on a {grid.shape[0]} x {grid.shape[1]} sub-grid, {n_nan} NaN box edges, all bit-exact')


def test_cameras_that_see_nothing():
    """Every box invisible (cameras looking past the field): no texel is covered, the compacted GEMM has zero units,
    and every cell gets relu(bias) per view and scale -- both C = 256 formulations and the fp32 kernel."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::4, ::4].contiguous()
    world_c = np.array([1950.0, 1950.0, 0.0])
    cams = []
    for k in range(2):
        # beside the field, looking ALONG its edge: every corner projects beyond the left / right image border (or, behind
        # the camera, wraps to the other side), so boxes clamp to zero width or exceed the area cap -- none is visible
        eye = world_c + np.array([6000.0 * (1 if k == 0 else -1), 0.0, 400.0])
        cams.append(synthetic.look_at(eye, eye + np.array([0.0, 1.0, 0.0]), 900.0, g.image_size).astype(np.float32))
    calibs = torch.from_numpy(np.stack(cams)).cuda()
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs, grid.cuda())
    for fh, fw in g.feature_sizes():
        _, vis, _ = table.scale_table(fh, fw)
        assert not bool(vis.any())
    feats = [f.cuda() for f in synthetic.features(g, batch=2, n_views=2, seed=17)]
    params = synthetic.collapse_params(g, seed=17)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    want = sum(torch.relu(b) for b in bs) * 2
    for path, flags in PATH_FLAGS.items():
        out = vfa_b200.aggregate(feats, table, ws, bs, flags=flags)
        assert vfa_b200.last_kernel_path() == path
        torch.testing.assert_close(out, want[None, :, None, None].expand_as(out), rtol=0, atol=1e-6)


def test_cameras_that_see_nothing_backward():
    """Same rig through the backward: no CSR entries, every tile skipped -> dFeature = dWeight = 0, dBias = the masked
    cotangent summed over frames, views and cells."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::4, ::4].contiguous()
    world_c = np.array([1950.0, 1950.0, 0.0])
    cams = [synthetic.look_at(world_c + np.array([6000.0 * sgn, 0.0, 400.0]),
                              world_c + np.array([6000.0 * sgn, 1.0, 400.0]), 900.0, g.image_size).astype(np.float32)
            for sgn in (1, -1)]
    calibs = torch.from_numpy(np.stack(cams)).cuda()
    table = vfa_b200.build_table(_geom(g.name, grid.shape[:2]), calibs, grid.cuda())
    feats = [f.cuda().requires_grad_(True) for f in synthetic.features(g, batch=2, n_views=2, seed=18)]
    params = synthetic.collapse_params(g, seed=18)
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    gout = torch.randn(2, 256, *grid.shape[:2], generator=torch.Generator().manual_seed(5)).cuda()
    vfa_b200.aggregate(feats, table, ws, bs).backward(gout)
    for f, w in zip(feats, ws):
        assert float(f.grad.abs().max()) == 0.0 and float(w.grad.abs().max()) == 0.0
    for b in bs:
        want = (gout.sum(dim=(0, 2, 3)) * 2) * (b.detach() > 0)           # 2 views; relu passes where bias > 0
        torch.testing.assert_close(b.grad, want, rtol=1e-5, atol=1e-4)


def test_graphed_aggregator_replays_the_eager_call_bit_for_bit():
    """GraphedAggregator: table + forward captured once in a CUDA graph; replays with new features and new calibrations
    (the table is rebuilt inside the graph) equal the eager call sequence exactly."""
    g = geometry.MULTIVIEWC
    grid = geometry.grid_for(g)[::3, ::3].contiguous().cuda()
    calibs = synthetic.ring_calibs(g, n_views=3).cuda()
    params = synthetic.collapse_params(g, seed=6)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=2, n_views=3, seed=6)]
    geom = _geom(g.name, grid.shape[:2])
    ga = vfa_b200.GraphedAggregator(geom, [tuple(f.shape) for f in feats], ws, bs)
    ga.load(feats, calibs, grid)
    ga.capture()
    other = [torch.flip(f, dims=[3]).contiguous() for f in feats]
    for f, c in ((feats, calibs), (other, calibs.roll(1, 0).contiguous()), (feats, calibs)):
        want = vfa_b200.aggregate(f, vfa_b200.build_table(geom, c, grid), ws, bs, channels_last=True)
        got = ga(f, c)
        assert torch.equal(got, want)


def test_table_prepared_flag_reuses_the_workspace_bit_for_bit():
    """VFA_FLAG_TABLE_PREPARED (static cameras): the second call with the same table, shapes, flags and workspace skips
    the tap records / coverage / row lists / texel lists and must give the bits of a full call, on new features too."""
    g = geometry.MULTIVIEWX
    grid = geometry.grid_for(g)[::2, ::2].contiguous().cuda()
    calibs = synthetic.ring_calibs(g, n_views=3).cuda()
    params = synthetic.collapse_params(g, seed=12)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=2, n_views=3, seed=12)]
    other = [torch.flip(f, dims=[2]).contiguous() for f in feats]
    geom = _geom(g.name, grid.shape[:2])
    table = vfa_b200.build_table(geom, calibs, grid)
    shape = vfa_b200.make_shape(feats, geom.n_layers)
    wsp = vfa_b200.workspace_for(geom, shape, 0, feats[0].device)
    vfa_b200.prepare_weights(geom, shape, ws, 0, workspace=wsp)
    full = [vfa_b200.aggregate_forward_raw(f, table, ws, bs, 0, workspace=wsp, prepared=True).clone() for f in (feats, other)]
    assert not torch.equal(full[0], full[1])
    for f, want in ((other, full[1]), (feats, full[0])):
        got = vfa_b200.aggregate_forward_raw(f, table, ws, bs, 0, workspace=wsp, prepared=True, table_prepared=True)
        assert torch.equal(got, want)
    with pytest.raises(ValueError, match='workspace'):
        vfa_b200.aggregate_forward_raw(feats, table, ws, bs, 0, table_prepared=True)
