"""Scratch: tcgen05 path vs SIMT path vs float64 port on C=256 problems (run under `timeout`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic
from oracle import ref_port

def geom_of(g, grid):
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    return vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)

def report(tag, got, want):
    err = (got.double() - want.double()).abs()
    tol = 1e-6 + 1e-5 * want.double().abs()
    print(f'{tag}: max_abs={err.max().item():.3e} mean_abs={err.mean().item():.3e} outside_tol={(err > tol).double().mean().item():.3e}', flush=True)

name = sys.argv[1] if len(sys.argv) > 1 else 'MultiviewC'
g = geometry.GEOMETRIES[name]
grid = geometry.grid_for(g)
# 1) one (view, scale), vs float64 port
calibs = synthetic.ring_calibs(g, n_views=1)
feats = synthetic.features(g, batch=1, n_views=1, seed=5, sizes=[g.feature_sizes()[1]])
params = synthetic.collapse_params(g, seed=5)[:1]
table = vfa_b200.build_table(geom_of(g, grid), calibs.cuda(), grid.cuda())
fc = [f.cuda() for f in feats]; ws = [w.cuda() for w, _ in params]; bs = [b.cuda() for _, b in params]
simt = vfa_b200.aggregate(fc, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT); torch.cuda.synchronize()
print('simt done', vfa_b200.last_kernel_path(), flush=True)
umma = vfa_b200.aggregate(fc, table, ws, bs, flags=vfa_b200.FLAG_FORCE_UMMA); torch.cuda.synchronize()
print('umma done', vfa_b200.last_kernel_path(), flush=True)
want = ref_port.vfa_forward(feats[0][0, 0].double(), calibs[0], grid, params[0][0].double(), params[0][1].double(),
                            g.grid_height, g.cube_size, name, g.image_size)
report('single simt vs f64', simt.cpu(), want)
report('single umma vs f64', umma.cpu(), want)
report('single umma vs simt', umma.cpu(), simt.cpu())
# 2) full frame batch 2, all views/scales: umma vs simt
calibs = synthetic.ring_calibs(g)
feats = [f.cuda() for f in synthetic.features(g, batch=2, seed=1)]
params = synthetic.collapse_params(g, seed=1)
ws = [w.cuda() for w, _ in params]; bs = [b.cuda() for _, b in params]
table = vfa_b200.build_table(geom_of(g, grid), calibs.cuda(), grid.cuda())
simt = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_SIMT); torch.cuda.synchronize()
umma = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_FORCE_UMMA); torch.cuda.synchronize()
report('full B=2 umma vs simt', umma.cpu(), simt.cpu())
one = vfa_b200.aggregate([f[:1] for f in feats], table, ws, bs, flags=vfa_b200.FLAG_FORCE_UMMA); torch.cuda.synchronize()
report('full B=1 (view-split, atomics) umma vs simt', one.cpu(), simt[:1].cpu())
