"""Scratch check + timing of VFA_FLAG_TABLE_PREPARED (static cameras): python scripts/check_static.py [workload] [B]"""
import os, sys, time
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
feats2 = [torch.flip(f, dims=[2]).contiguous() for f in feats]
params = synthetic.collapse_params(g, seed=0)
ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
shape = vfa_b200.make_shape(feats, geom.n_layers)
wsp = vfa_b200.workspace_for(geom, shape, 0, feats[0].device)
vfa_b200.prepare_weights(geom, shape, ws, 0, workspace=wsp)
table = vfa_b200.build_table(geom, calibs, grid)
full = lambda f: vfa_b200.aggregate_forward_raw(f, table, ws, bs, 0, workspace=wsp, prepared=True)
fast = lambda f: vfa_b200.aggregate_forward_raw(f, table, ws, bs, 0, workspace=wsp, prepared=True, table_prepared=True)
a1, a2 = full(feats).clone(), full(feats2).clone()
b2, b1 = fast(feats2).clone(), fast(feats).clone()
torch.cuda.synchronize()
print('table_prepared == full:', torch.equal(a1, b1) and torch.equal(a2, b2), 'different inputs differ:', not torch.equal(a1, a2))


def timed(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f'{name} B={B}: full call {timed(lambda: full(feats)):.3f} ms, table prepared {timed(lambda: fast(feats)):.3f} ms')
