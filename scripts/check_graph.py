"""Scratch check + timing of the CUDA-graph path: python scripts/check_graph.py [workload] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else 'MultiviewC'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = geometry.GEOMETRIES[name]
grid = geometry.grid_for(g).cuda()
calibs = synthetic.ring_calibs(g).cuda()
feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, seed=0)]
params = synthetic.collapse_params(g, seed=0)
ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
shape = vfa_b200.make_shape(feats, geom.n_layers)
wsp = vfa_b200.workspace_for(geom, shape, 0, feats[0].device)
vfa_b200.prepare_weights(geom, shape, ws, 0, workspace=wsp)
out = torch.empty(B, 256, *grid.shape[:2], device='cuda')


def eager(f, c):
    table = vfa_b200.build_table(geom, c, grid)
    return vfa_b200.aggregate_forward_raw(f, table, ws, bs, 0, out=out, workspace=wsp, prepared=True)


ga = vfa_b200.GraphedAggregator(geom, [tuple(f.shape) for f in feats], ws, bs)
ga.load(feats, calibs, grid)
ga.capture()
feats2 = [torch.flip(f, dims=[2]).contiguous() for f in feats]
calibs2 = calibs.roll(1, 0).contiguous()
ok = True
for f, c in ((feats, calibs), (feats2, calibs2), (feats, calibs)):
    want = eager(f, c).clone()
    got = ga(f, c)
    torch.cuda.synchronize()
    ok = ok and torch.equal(got, want)
print('graph == eager:', ok)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f'{name} B={B}: eager {timed(lambda: eager(feats, calibs)):.3f} ms/call, graph replay {timed(lambda: ga.graph.replay()):.3f} ms/call '
      f'(wall clock, {B} frame(s) per call)')
