"""Scratch: dWeight of the tcgen05 kernel vs the cuBLAS path on a small C=256 case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

g = geometry.MULTIVIEWC
grid = geometry.grid_for(g)[::3, ::3].contiguous().cuda()
calibs = synthetic.ring_calibs(g, n_views=2).cuda()
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
table = vfa_b200.build_table(geom, calibs, grid)
feats0 = synthetic.features(g, batch=1, n_views=2, seed=8, sizes=g.feature_sizes()[1:])
params = synthetic.collapse_params(g, seed=8)[:2]
gout = torch.randn(1, 256, *grid.shape[:2], generator=torch.Generator().manual_seed(6)).cuda()
res = {}
for tag, env in (('tc', None), ('cublas', '1')):
    if env: os.environ['VFA_BWD_CUBLAS_DW'] = env
    else: os.environ.pop('VFA_BWD_CUBLAS_DW', None)
    vfa_b200.reload_env()
    fc = [t.cuda().requires_grad_(True) for t in feats0]
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    out = vfa_b200.aggregate(fc, table, ws, bs)
    out.backward(gout)
    torch.cuda.synchronize()
    res[tag] = [w.grad.clone() for w in ws]
for s in range(2):
    a, b = res['tc'][s], res['cublas'][s]          # [256, 1280] (o, c*nl+n)
    nl = a.shape[1] // 256
    print(f'scale {s}: |tc| max {float(a.abs().max()):.4e} |cublas| max {float(b.abs().max()):.4e} '
          f'max diff {float((a - b).abs().max()):.4e} nonzero tc {int((a != 0).sum())} / {a.numel()}')
    A = a.view(256, 256, nl).permute(2, 0, 1)      # [n][o][c]
    Bm = b.view(256, 256, nl).permute(2, 0, 1)
    for n in range(nl):
        d = (A[n] - Bm[n]).abs().max()
        dt = (A[n].t() - Bm[n]).abs().max()
        print(f'  layer {n}: diff {float(d):.3e}  diff if transposed {float(dt):.3e}  ratio mean {float((A[n].abs().mean() / (Bm[n].abs().mean() + 1e-30))):.3f}')
    # row / column block structure of the error
    E = (A - Bm).abs()
    print('  err by o-block of 32:', [f'{float(E[:, i*32:(i+1)*32].max()):.1e}' for i in range(8)])
    print('  err by c-block of 32:', [f'{float(E[:, :, i*32:(i+1)*32].max()):.1e}' for i in range(8)])
