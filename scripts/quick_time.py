"""Scratch timing of the fused forward (not the bench): python scripts/quick_time.py [workload] [B] [flags]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else 'MultiviewC'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = geometry.GEOMETRIES[name]
grid = geometry.grid_for(g)
calibs = synthetic.ring_calibs(g).cuda()
feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, seed=0)]
params = synthetic.collapse_params(g, seed=0)
ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
table = vfa_b200.build_table(geom, calibs, grid.cuda())
for _ in range(2):
    out = vfa_b200.aggregate(feats, table, ws, bs, flags=flags, channels_last=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    out = vfa_b200.aggregate(feats, table, ws, bs, flags=flags, channels_last=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f'var={os.environ.get("VFA_UMMA_VARIANT","0")} {name} B={B} flags={flags} path={vfa_b200.last_kernel_path()}: {ms:.3f} ms/step, {B / ms * 1e3:.1f} frames/s, mean={float(out.mean()):.5f}')
