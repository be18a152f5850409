"""Scratch: error anatomy of the tcgen05 path (signed error, K dependence, variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic
from oracle import ref_port

g = geometry.MULTIVIEWC
grid = geometry.grid_for(g)
calibs = synthetic.ring_calibs(g, n_views=1)
sizes = [g.feature_sizes()[1]]
for grid_h in (32, 64, 160, 320, 512):
    zs = list(range(0, grid_h, g.cube_size[2]))
    nl = len(zs)
    geom = vfa_b200.make_geometry(nl, g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
    feats = synthetic.features(g, batch=1, n_views=1, seed=5, sizes=sizes)
    gen = torch.Generator().manual_seed(9)
    K = 256 * nl
    w = (torch.rand(256, K, generator=gen) * 2 - 1) / K ** 0.5
    b = (torch.rand(256, generator=gen) * 2 - 1) / K ** 0.5
    table = vfa_b200.build_table(geom, calibs.cuda(), grid.cuda())
    want = ref_port.vfa_forward(feats[0][0, 0].double(), calibs[0], grid, w.double(), b.double(), grid_h, g.cube_size,
                                g.name, g.image_size)
    # also pre-activation sign: only look at active outputs
    act = want > 0
    for tag, flags in (('simt', vfa_b200.FLAG_FORCE_SIMT), ('umma', vfa_b200.FLAG_FORCE_UMMA)):
        got = vfa_b200.aggregate([f.cuda() for f in feats], table, [w.cuda()], [b.cuda()], flags=flags).cpu().double()
        e = (got - want)[act]
        tol = 1e-6 + 1e-5 * want[act].abs()
        print(f'nl={nl} K={K} {tag} var={os.environ.get("VFA_UMMA_VARIANT","0")}: mean_signed={e.mean().item():+.3e} mean_abs={e.abs().mean().item():.3e} '
              f'rms={e.pow(2).mean().sqrt().item():.3e} max={e.abs().max().item():.3e} outside={(e.abs() > tol).double().mean().item():.2e} '
              f'mean_out={want[act].mean().item():.3f}', flush=True)
