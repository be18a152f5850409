"""Two-GPU checks (skipped on a 1-GPU box): BEV row-slab sharding and camera sharding over NCCL reproduce the single-GPU
result and gradients; data-parallel ranks reduce the collapse gradients."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist
    import vfa_b200
    from vfa_b200 import distributed as vd
    from vfa_b200 import geometry, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        g = geometry.MULTIVIEWC
        zs = list(range(0, g.grid_height, g.cube_size[2]))
        geom_of = lambda lw: vfa_b200.make_geometry(len(zs), g.cube_size, zs, lw, g.name, g.image_size)  # noqa: E731
        grid = geometry.grid_for(g).to(dev)
        calibs = synthetic.ring_calibs(g, n_views=3).to(dev)
        feats = [f.to(dev) for f in synthetic.features(g, batch=2, n_views=3, seed=13)]
        if rank != 0:
            feats = [torch.zeros_like(f) for f in feats]
        vd.broadcast_features(feats, src=0)
        params = synthetic.collapse_params(g, seed=13)
        ws = [w.to(dev).requires_grad_(True) for w, _ in params]
        bs = [b.to(dev).requires_grad_(True) for _, b in params]
        feats = [f.requires_grad_(True) for f in feats]
        full = vd.aggregate_slab(feats, calibs, grid, ws, bs, vd.cuda_compute(geom_of))
        gout = torch.randn(full.shape, generator=torch.Generator(device=dev).manual_seed(1), device=dev)
        full.backward(gout)
        # single-GPU truth on every rank
        f2 = [f.detach().clone().requires_grad_(True) for f in feats]
        w2 = [w.detach().clone().requires_grad_(True) for w in ws]
        b2 = [b.detach().clone().requires_grad_(True) for b in bs]
        table = vfa_b200.build_table(geom_of(grid.shape[:2]), calibs, grid)
        want = vfa_b200.aggregate(f2, table, w2, b2)
        want.backward(gout)
        ok = torch.allclose(full, want, rtol=1e-5, atol=1e-5)
        for a, b in zip(feats + ws + bs, f2 + w2 + b2):
            scale = float(b.grad.abs().max()) + 1e-30
            ok = ok and float((a.grad - b.grad).abs().max()) / scale < 1e-4
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_slab_sharding_two_gpus_matches_single_gpu():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def _views_worker(rank, world, port, results):
    import torch.distributed as dist
    import vfa_b200
    from vfa_b200 import distributed as vd
    from vfa_b200 import geometry, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        g = geometry.WILDTRACK
        V, B = 3, 3
        zs = list(range(0, g.grid_height, g.cube_size[2]))
        grid = geometry.grid_for(g).to(dev)
        cgeom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
        calibs = synthetic.ring_calibs(g, n_views=V).to(dev)
        feats = [f.to(dev) for f in synthetic.features(g, batch=B, n_views=V, seed=21)]
        params = synthetic.collapse_params(g, seed=21)
        ws = [w.to(dev).requires_grad_(True) for w, _ in params]
        bs = [b.to(dev).requires_grad_(True) for _, b in params]
        v0, v1 = vd.view_bounds(V, world, rank)                          # 2 + 1 cameras
        local = [f[:, v0:v1].contiguous().requires_grad_(True) for f in feats]

        def compute(f, c, grid_, w, b_):
            return vfa_b200.aggregate(f, vfa_b200.build_table(cgeom, c, grid_), w, b_)
        full = vd.aggregate_views(local, calibs[v0:v1], grid, ws, bs, compute, out_channels=256, frames_per_chunk=1)
        gout = torch.randn(full.shape, generator=torch.Generator(device=dev).manual_seed(1), device=dev)
        full.backward(gout)
        f2 = [f.detach().clone().requires_grad_(True) for f in feats]
        w2 = [w.detach().clone().requires_grad_(True) for w in ws]
        b2 = [b.detach().clone().requires_grad_(True) for b in bs]
        want = compute(f2, calibs, grid, w2, b2)
        want.backward(gout)
        ok = torch.allclose(full, want, rtol=1e-5, atol=1e-5)
        pairs = [(a.grad, b.grad[:, v0:v1]) for a, b in zip(local, f2)] + [(a.grad, b.grad) for a, b in zip(ws + bs, w2 + b2)]
        for a, b in pairs:
            ok = ok and float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30) < 1e-4
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_camera_sharding_two_gpus_matches_single_gpu():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_views_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def _fused_views_worker(rank, world, port, results):
    import torch.distributed as dist
    import vfa_b200
    from vfa_b200 import distributed as vd
    from vfa_b200 import geometry, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        g = geometry.MULTIVIEWC
        V, B = 5, 2
        zs = list(range(0, g.grid_height, g.cube_size[2]))
        grid = geometry.grid_for(g).to(dev)
        cgeom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
        calibs = synthetic.ring_calibs(g, n_views=V).to(dev)
        params = synthetic.collapse_params(g, seed=23)
        ws, bs = [w.to(dev) for w, _ in params], [b.to(dev) for _, b in params]
        v0, v1 = vd.view_bounds(V, world, rank)                          # 3 + 2 cameras
        aggs = [vd.FusedViewAggregator(cgeom, B, mode=m) for m in ('reduce_scatter', 'multicast_red')]
        table_local = vfa_b200.build_table(cgeom, calibs[v0:v1], grid)
        table_all = vfa_b200.build_table(cgeom, calibs, grid)
        ok = True
        for step in range(3):                                            # both buffer slots, and a reuse
            feats = [vfa_b200.to_channels_last(f.to(dev)) for f in synthetic.features(g, batch=B, n_views=V, seed=40 + step)]
            local = [f[:, v0:v1].contiguous() for f in feats]
            want = vfa_b200.aggregate_forward_raw(feats, table_all, ws, bs)
            for agg in aggs:
                got = agg(local, table_local, ws, bs).clone()
                err = (got - want).abs()
                ok = ok and bool((err <= 2e-6 + 2e-6 * want.abs()).all())     # the same terms, summed in another order
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_fused_multicast_camera_sharding_two_gpus_matches_single_gpu():
    """Camera shards reduced inside the pooling kernel over NVLink -- fused reduce-scatter into the owners' bands (red.add on
    peer memory) + multicast all-gather, and multimem.red on the multicast address: same map as one GPU aggregating all
    cameras, on every rank, for three consecutive steps."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_fused_views_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_one_process_drives_two_devices():
    """Kernel attributes, resident-cluster counts and the cuBLAS handle are cached per device: forward + backward on
    cuda:1 after cuda:0 in ONE process (nn.DataParallel-style use) give the same results (C = 256 tcgen05 path with its >
    48 KB dynamic shared memory, and the generic C = 16 path whose backward calls cuBLAS)."""
    import vfa_b200
    from vfa_b200 import geometry, synthetic
    g = geometry.MULTIVIEWC
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    grid = geometry.grid_for(g)[::2, ::2].contiguous()
    cgeom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    calibs = synthetic.ring_calibs(g, n_views=2)
    for channels in (256, 16):
        feats = synthetic.features(g, batch=2, n_views=2, channels=channels, seed=3)
        params = synthetic.collapse_params(g, channels=channels, seed=3)
        gout = torch.randn(2, channels, *grid.shape[:2], generator=torch.Generator().manual_seed(4))
        res = []
        for d in (0, 1):
            dev = torch.device('cuda', d)
            table = vfa_b200.build_table(cgeom, calibs.to(dev), grid.to(dev))
            f = [t.to(dev).requires_grad_(True) for t in feats]
            w = [x.to(dev).requires_grad_(True) for x, _ in params]
            b = [x.to(dev).requires_grad_(True) for _, x in params]
            out = vfa_b200.aggregate(f, table, w, b)
            out.backward(gout.to(dev))
            torch.cuda.synchronize(dev)
            res.append([out.detach().cpu()] + [t.grad.cpu() for t in f + w + b])
        assert torch.equal(res[0][0], res[1][0])                                  # the forward is deterministic
        for a, b_ in zip(res[0][1:], res[1][1:]):                                 # the backward sums with atomics
            assert float((a - b_).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-12
