"""Scratch: device time of one forward+backward step (CUDA events), MultiviewC / Wildtrack, B = 4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

for name in (sys.argv[1:] or ['MultiviewC', 'Wildtrack']):
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g).cuda()
    calibs = synthetic.ring_calibs(g).cuda()
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    table = vfa_b200.build_table(geom, calibs, grid)
    feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous().requires_grad_(True) for f in synthetic.features(g, batch=4, seed=0)]
    params = synthetic.collapse_params(g, seed=0)
    ws, bs = [w.cuda().requires_grad_(True) for w, _ in params], [b.cuda().requires_grad_(True) for _, b in params]
    gout = torch.randn(4, 256, *grid.shape[:2], device='cuda')

    def step():
        for t in feats + ws + bs:
            t.grad = None
        vfa_b200.aggregate(feats, table, ws, bs, channels_last=True).backward(gout)
    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name}: fwd+bwd {e0.elapsed_time(e1) / 5:.3f} ms/step', flush=True)
