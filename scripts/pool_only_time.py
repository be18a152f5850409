"""Scratch: duration of the pooling-only call (VFA_UMMA_VARIANT=128|256) with and without a busy GPU queue ahead of the
start event (is bench.py's per-kernel figure inflated by host launch latency?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

g = geometry.GEOMETRIES['MultiviewC']
B = 4
grid = geometry.grid_for(g)
calibs = synthetic.ring_calibs(g).cuda()
feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, seed=0)]
params = synthetic.collapse_params(g, seed=0)
ws_, bs_ = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
table = vfa_b200.build_table(geom, calibs, grid.cuda())
out = torch.empty(B, 256, grid.shape[0], grid.shape[1], device='cuda')
shape = vfa_b200.make_shape(feats, geom.n_layers)
ws = vfa_b200.workspace_for(geom, shape, 0, feats[0].device)
vfa_b200.prepare_weights(geom, shape, ws_, 0, ws)
for _ in range(2):
    vfa_b200.aggregate_forward_raw(feats, table, ws_, bs_, 0, out=out, workspace=ws, prepared=True)
for bits, name in ((64 | 256, 'gemm only'), (128 | 256, 'pool only'), (64 | 256, 'gemm only'), (128 | 256, 'pool only')):
    os.environ['VFA_UMMA_VARIANT'] = str(bits)
    vfa_b200.reload_env()
    for busy in (False, True):
        n = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(-2, n):
            if busy:
                torch.cuda._sleep(3_000_000)
            if i >= 0:
                evs[i][0].record()
            vfa_b200.aggregate_forward_raw(feats, table, ws_, bs_, 0, out=out, workspace=ws, prepared=True)
            if i >= 0:
                evs[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        print(f'{name}, queue {"busy" if busy else "empty"} at the start event: median {ts[n // 2]:.3f} ms, min {ts[0]:.3f}')

# per-iteration series of a longer pool-only loop (does the clock drop under sustained load?)
os.environ['VFA_UMMA_VARIANT'] = str(128 | 256)
vfa_b200.reload_env()
n = 60
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for i in range(n):
    evs[i][0].record()
    vfa_b200.aggregate_forward_raw(feats, table, ws_, bs_, 0, out=out, workspace=ws, prepared=True)
    evs[i][1].record()
torch.cuda.synchronize()
print('pool-only series (ms):', ' '.join(f'{a.elapsed_time(b):.2f}' for a, b in evs))
