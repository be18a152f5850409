"""Step time of the camera-sharded fused aggregation, both collectives, with a per-phase breakdown (torchrun, N ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vfa_b200
from vfa_b200 import distributed as vd, geometry, synthetic

rank, lr, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
g = geometry.MULTIVIEWC; B = 4
zs = list(range(0, g.grid_height, g.cube_size[2]))
grid = geometry.grid_for(g).to(dev)
cgeom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
v0, v1 = vd.view_bounds(g.n_views, world, rank)
calibs = synthetic.ring_calibs(g).to(dev)[v0:v1].contiguous()
params = synthetic.collapse_params(g, seed=0)
ws, bs = [w.to(dev) for w, _ in params], [b.to(dev) for _, b in params]
gen = torch.Generator(device=dev).manual_seed(rank)
feats = [torch.randn(B, v1 - v0, h, w, 256, generator=gen, device=dev).relu_() for h, w in g.feature_sizes()]

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

out_local = torch.empty(B, grid.shape[0], grid.shape[1], 256, device=dev)
def local_only():
    if v1 > v0:
        tb = vfa_b200.build_table(cgeom, calibs, grid)
        vfa_b200.aggregate_forward_raw(feats, tb, ws, bs, vfa_b200.FLAG_OUT_NHWC, out=out_local)
res = {'local compute (store to own HBM)': timed(local_only)}
for mode in ('reduce_scatter', 'multicast_red'):
    agg = vd.FusedViewAggregator(cgeom, B, mode=mode)
    def step():
        tb = vfa_b200.build_table(cgeom, calibs, grid) if v1 > v0 else None
        return agg(feats, tb, ws, bs)
    res[mode] = timed(step)
    res[mode + ' barrier only'] = timed(lambda: agg.hdl.barrier(channel=0))
part = torch.zeros(B, 256, grid.shape[0], grid.shape[1], device=dev)
res['nccl all_reduce of the map alone'] = timed(lambda: dist.all_reduce(part))
if rank == 0:
    print(f'N={world} cameras/rank={v1 - v0}:', {k: round(v, 3) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
