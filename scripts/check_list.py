"""Scratch check of the list-based pooling kernel: the all-lists run, a run whose list capacity forces most quads through
the walking completion pass (VFA_POOL_LIST_CAP=1) and a run with no list at all must agree bit for bit:
python scripts/check_list.py [workload ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

B = int(os.environ.get('B', 2))
for name in (sys.argv[1:] or ['MultiviewC', 'MultiviewX', 'Wildtrack']):
    g = geometry.GEOMETRIES[name]
    grid = geometry.grid_for(g)
    calibs = synthetic.ring_calibs(g, in_field=True).cuda()
    feats = [f.cuda().permute(0, 1, 3, 4, 2).contiguous() for f in synthetic.features(g, batch=B, n_views=calibs.shape[0], seed=0)]
    params = synthetic.collapse_params(g, seed=0)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    table = vfa_b200.build_table(geom, calibs, grid.cuda())
    run = lambda: vfa_b200.aggregate(feats, table, ws, bs, channels_last=True)
    outs = {}
    for cap in ('', '1', '3', '64'):
        os.environ.pop('VFA_POOL_LIST_CAP', None)
        if cap:
            os.environ['VFA_POOL_LIST_CAP'] = cap
        outs[cap] = run().clone()
    os.environ.pop('VFA_POOL_LIST_CAP', None)
    f2 = [f.detach().clone().requires_grad_(True) for f in feats]
    o2 = vfa_b200.aggregate(f2, table, ws, bs, channels_last=True)                # training variant (writes the ReLU mask)
    torch.cuda.synchronize()
    print(f'{name}: training variant differs in {int((outs[""] != o2.detach()).sum())} elements')
    ref = vfa_b200.aggregate(feats, table, ws, bs, flags=vfa_b200.FLAG_GRID_SIDE, channels_last=True)
    d = (outs[''] - ref).abs()
    ne = {c: int((outs[''] != outs[c]).sum()) for c in ('1', '3', '64')}
    print(f'{name} B={B} V={calibs.shape[0]}: elements differing from the default run: slot cap 1: {ne["1"]}, cap 3: {ne["3"]}, '
          f'cap 64: {ne["64"]}; vs grid-side max abs {float(d.max()):.3e}, '
          f'outside tol {float((d > 1e-6 + 1e-5 * ref.abs()).float().mean()):.2e}', flush=True)
