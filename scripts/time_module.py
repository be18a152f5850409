"""Scratch timing: the reference-style loop over the drop-in VFA modules (21 calls per MultiviewC frame) vs the fused
entry at several batch sizes.  python scripts/time_module.py [workload]"""
import os, sys
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic


def timed(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


name = sys.argv[1] if len(sys.argv) > 1 else 'MultiviewC'
g = geometry.GEOMETRIES[name]
grid = geometry.grid_for(g).cuda()
calibs = synthetic.ring_calibs(g).cuda()
params = synthetic.collapse_params(g, seed=0)
args = SimpleNamespace(data=g.name, image_size=g.image_size)
mods = []
for (w, b), scale in zip(params, (1 / 8., 1 / 16., 1 / 32.)):
    m = vfa_b200.VFA(256, g.grid_height, g.cube_size, scale, args).cuda()
    with torch.no_grad():
        m.collapse.weight.copy_(w)
        m.collapse.bias.copy_(b)
    mods.append(m)
feats = [f.cuda() for f in synthetic.features(g, batch=1, seed=0)]          # [1,V,C,h,w]


def loop():
    with torch.no_grad():
        ortho = 0
        for v in range(g.n_views):
            ortho = ortho + sum(m(f[0, v:v + 1], calibs[v], grid[None]) for m, f in zip(mods, feats))
    return ortho


t = timed(loop)
print(f'{name}: reference-style loop over {3 * g.n_views} drop-in VFA.forward calls: {t:.2f} ms/frame', flush=True)
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
for B in (1, 2, 4, 16):
    fb = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, seed=0)]

    def fused():
        table = vfa_b200.build_table(geom, calibs, grid)
        return vfa_b200.aggregate(fb, table, ws, bs, channels_last=True)
    t = timed(fused)
    print(f'{name}: fused entry B={B}: {t:.2f} ms/step = {t / B:.3f} ms/frame ({B / t * 1e3:.0f} frames/s)', flush=True)
    del fb
