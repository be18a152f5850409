"""World-size-2 / 3 gloo tests (CPU) of the multi-GPU host logic: slab partition + all-gather reassembly + gradient
all-reduce, camera sharding + all-reduce of the partial maps, and data-parallel collapse-gradient reduction.  The per-rank compute is the oracle port here (the CUDA
kernels need a GPU); on GPUs the same functions run with vfa_b200.distributed.cuda_compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vfa_b200 import distributed as vd
from vfa_b200 import geometry, synthetic


def test_slab_bounds_cover_rows_exactly():
    for L in (1, 7, 120, 156, 160):
        for world in (1, 2, 3, 4, 8):
            rows = [vd.slab_bounds(L, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _port_compute(name):
    from oracle import ref_port
    g = geometry.GEOMETRIES[name]

    def compute(feats, calibs, grid_slab, weights, biases):
        return ref_port.aggregate(feats, calibs, grid_slab, list(zip(weights, biases)), g.grid_height, g.cube_size, name,
                                  g.image_size, cache_boxes=True)
    return compute


def _worker(rank, world, port, name, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = geometry.GEOMETRIES[name]
        grid = geometry.grid_for(g)[::12, ::9].contiguous()            # 13 rows: uneven split 7 + 6
        calibs = synthetic.ring_calibs(g, n_views=2)
        sizes = [(30, 52), (23, 40)]
        feats = [f.double() for f in synthetic.features(g, batch=2, n_views=2, channels=4, seed=5, sizes=sizes)]
        params = synthetic.collapse_params(g, channels=4, seed=5)[:2]
        # rank 1 starts with garbage features: broadcast must fix them
        if rank == 1:
            feats = [torch.full_like(f, 123.0) for f in feats]
        vd.broadcast_features(feats, src=0)
        feats = [f.requires_grad_(True) for f in feats]
        ws = [w.double().requires_grad_(True) for w, _ in params]
        bs = [b.double().requires_grad_(True) for _, b in params]
        compute = _port_compute(name)
        full = vd.aggregate_slab(feats, calibs, grid, ws, bs, compute)
        gen = torch.Generator().manual_seed(3)
        gout = torch.randn(full.shape, generator=gen, dtype=torch.float64)
        full.backward(gout)
        # single-process truth
        f2 = [f.detach().clone().requires_grad_(True) for f in feats]
        w2 = [w.detach().clone().requires_grad_(True) for w in ws]
        b2 = [b.detach().clone().requires_grad_(True) for b in bs]
        want = compute(f2, calibs, grid, w2, b2)
        want.backward(gout)
        ok = torch.allclose(full, want, rtol=1e-12, atol=1e-12)
        for a, b in zip(feats + ws + bs, f2 + w2 + b2):
            ok = ok and torch.allclose(a.grad, b.grad, rtol=1e-10, atol=1e-12)
        # data-parallel gradient reduction: each rank holds rank+1 -> sum 3
        p = torch.nn.Parameter(torch.zeros(5))
        p.grad = torch.full((5,), float(rank + 1))
        q = torch.nn.Parameter(torch.zeros(2, 2))
        q.grad = torch.full((2, 2), 10.0 * (rank + 1))
        vd.allreduce_collapse_grads([p, q])
        ok = ok and bool((p.grad == 3).all()) and bool((q.grad == 30).all())
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['MultiviewC', 'Wildtrack'])
def test_slab_sharding_equals_single_process_gloo(name):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def _views_worker(rank, world, port, name, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = geometry.GEOMETRIES[name]
        V, B, C = 2, 3, 4                                              # world 3 > V: the last rank owns no camera
        grid = geometry.grid_for(g)[::12, ::9].contiguous()
        calibs = synthetic.ring_calibs(g, n_views=V)
        sizes = [(30, 52), (23, 40)]
        feats = [f.double() for f in synthetic.features(g, batch=B, n_views=V, channels=C, seed=9, sizes=sizes)]
        params = synthetic.collapse_params(g, channels=C, seed=9)[:2]
        v0, v1 = vd.view_bounds(V, world, rank)
        local = [f[:, v0:v1].clone().requires_grad_(True) for f in feats]
        ws = [w.double().requires_grad_(True) for w, _ in params]
        bs = [b.double().requires_grad_(True) for _, b in params]
        compute = _port_compute(name)
        full = vd.aggregate_views(local, calibs[v0:v1], grid, ws, bs, compute, out_channels=C, frames_per_chunk=2)
        gout = torch.randn(full.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
        full.backward(gout)
        f2 = [f.detach().clone().requires_grad_(True) for f in feats]
        w2 = [w.detach().clone().requires_grad_(True) for w in ws]
        b2 = [b.detach().clone().requires_grad_(True) for b in bs]
        want = compute(f2, calibs, grid, w2, b2)
        want.backward(gout)
        ok = tuple(full.shape) == (B, C) + tuple(grid.shape[:2]) and torch.allclose(full, want, rtol=1e-12, atol=1e-12)
        for a, b in zip(local, f2):                                      # dFeature: complete on the owner, no exchange
            if v1 > v0:
                ok = ok and torch.allclose(a.grad, b.grad[:, v0:v1], rtol=1e-10, atol=1e-12)
            else:
                ok = ok and a.grad is None                               # a rank without cameras has no features
        for a, b in zip(ws + bs, w2 + b2):                               # dWeight / dBias: summed over ranks
            ok = ok and a.grad is not None and torch.allclose(a.grad, b.grad, rtol=1e-10, atol=1e-12)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_camera_sharding_equals_single_process_gloo(world):
    assert [vd.view_bounds(7, 8, r) for r in (0, 6, 7)] == [(0, 1), (6, 7), (7, 7)]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_views_worker, args=(world, _free_port(), 'MultiviewX', results), nprocs=world, join=True)
    assert dict(results) == {r: True for r in range(world)}


def test_camera_sharding_without_a_process_group_is_the_plain_call():
    g = geometry.GEOMETRIES['MultiviewC']
    grid = geometry.grid_for(g)[::12, ::12].contiguous()
    calibs = synthetic.ring_calibs(g, n_views=2)
    feats = [f.double() for f in synthetic.features(g, batch=3, n_views=2, channels=4, seed=2, sizes=[(30, 52)])]
    w, b = synthetic.collapse_params(g, channels=4, seed=2)[0]
    compute = _port_compute('MultiviewC')
    got = vd.aggregate_views(feats, calibs, grid, [w.double()], [b.double()], compute, out_channels=4, frames_per_chunk=2)
    assert torch.equal(got, compute(feats, calibs, grid, [w.double()], [b.double()]))


def test_row_bands_partition_the_grid_like_the_kernel_routes_rows():
    """FusedViewAggregator's reduce-scatter: every BEV row has exactly one owner, bands are whole 8-row tiles, and the
    owner the pooling kernel computes (min(cy // band_rows, world - 1), csrc/vfa_pool.cuh owner_base) is the rank whose
    [r0, r1) holds the row -- for grids that do and do not divide evenly, and for more ranks than tile rows."""
    from vfa_b200.distributed import row_band
    for L in (156, 160, 120, 39, 7, 8, 9):
        for world in (1, 2, 3, 4, 8, 16):
            bands = [row_band(L, world, r) for r in range(world)]
            band = bands[0][2]
            assert band % 8 == 0 and all(b[2] == band for b in bands)
            owner = [-1] * L
            for r, (r0, r1, _) in enumerate(bands):
                assert 0 <= r0 <= r1 <= L
                for cy in range(r0, r1):
                    assert owner[cy] == -1
                    owner[cy] = r
            assert all(o >= 0 for o in owner)
            assert owner == [min(cy // band, world - 1) for cy in range(L)]
