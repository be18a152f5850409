"""Scratch check of the feature-side forward against the grid-side fused kernel + timing of both:
python scripts/check_fside.py [workload ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic


def timed(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = int(os.environ.get('B', 4))
for name in (sys.argv[1:] or ['MultiviewC', 'MultiviewX', 'Wildtrack']):
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g).cuda()
    feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, seed=0)]
    params = synthetic.collapse_params(g, seed=0)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    table = vfa_b200.build_table(geom, calibs, grid.cuda())
    run = lambda flags: vfa_b200.aggregate(feats, table, ws, bs, flags=flags, channels_last=True)
    os.environ.pop('VFA_UMMA_VARIANT', None)
    ref = run(vfa_b200.FLAG_GRID_SIDE); pg = vfa_b200.last_kernel_path()
    new = run(0); pf = vfa_b200.last_kernel_path()
    torch.cuda.synchronize()
    d = (new - ref).abs()
    tol = 1e-6 + 1e-5 * ref.abs()
    print(f'{name} B={B}: {pf} vs {pg}: max abs diff {float(d.max()):.3e}, mean {float(d.mean()):.3e}, '
          f'outside tol {float((d > tol).float().mean()):.2e}, ref mean {float(ref.mean()):.5f} new mean {float(new.mean()):.5f}',
          flush=True)
    tg = timed(lambda: run(vfa_b200.FLAG_GRID_SIDE))
    tf = timed(lambda: run(0))
    os.environ['VFA_UMMA_VARIANT'] = '64'
    t_gemm = timed(lambda: run(0))
    os.environ['VFA_UMMA_VARIANT'] = '128'
    t_pool = timed(lambda: run(0))
    os.environ.pop('VFA_UMMA_VARIANT', None)
    print(f'{name} B={B}: grid-side {tg:.3f} ms ({B / tg * 1e3:.1f} fps) | feature-side {tf:.3f} ms ({B / tf * 1e3:.1f} fps) '
          f'[gemm-only {t_gemm:.3f}, pool-only {t_pool:.3f}]', flush=True)
