"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / synccheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vfa_b200
from vfa_b200 import geometry, synthetic

g = geometry.MULTIVIEWC
grid = geometry.grid_for(g)[::6, ::5].contiguous().cuda()          # 26 x 32 cells: edge tiles + an odd tile count
calibs = synthetic.ring_calibs(g, n_views=2, in_field=True).cuda()
zs = list(range(0, g.grid_height, g.cube_size[2]))
geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, grid.shape[:2], g.name, g.image_size)
table = vfa_b200.build_table(geom, calibs, grid)
for C, flags in ((256, 0), (256, vfa_b200.FLAG_GRID_SIDE), (256, vfa_b200.FLAG_FORCE_SIMT), (16, 0)):
    feats = [f.cuda().requires_grad_(True) for f in synthetic.features(g, batch=2, n_views=3, channels=C, seed=1)]
    params = synthetic.collapse_params(g, channels=C, seed=1)
    ws = [w.cuda().requires_grad_(True) for w, _ in params]
    bs = [b.cuda().requires_grad_(True) for _, b in params]
    out = vfa_b200.aggregate(feats, table, ws, bs, flags=flags)
    out.sum().backward()
    torch.cuda.synchronize()
    print(C, flags, vfa_b200.last_kernel_path(), float(out.mean()), float(feats[0].grad.abs().mean()))
f16 = [f.detach().permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16) for f in
       [t.cuda() for t in synthetic.features(g, batch=1, n_views=3, seed=2)]]
params = synthetic.collapse_params(g, seed=2)
o = vfa_b200.aggregate(f16, table, [w.cuda() for w, _ in params], [b.cuda() for _, b in params], channels_last=True)
torch.cuda.synchronize()
print('bf16', vfa_b200.last_kernel_path(), float(o.mean()))
# round 2: staged-tile pooling without the mask (inference), channels-last output, single frame, forced overflow tiles are
# covered by the tests; the bf16 tensor-core variant; the decode tail
with torch.no_grad():
    feats = [f.cuda() for f in synthetic.features(g, batch=2, n_views=3, seed=3)]
    params = synthetic.collapse_params(g, seed=3)
    ws, bs = [w.cuda() for w, _ in params], [b.cuda() for _, b in params]
    for B_, fl in ((2, 0), (2, vfa_b200.FLAG_OUT_NHWC), (1, 0), (2, vfa_b200.FLAG_BF16_MMA)):
        o = vfa_b200.aggregate([f[:B_] for f in feats], table, ws, bs, flags=fl)
        torch.cuda.synchronize()
        print('inference', B_, fl, vfa_b200.last_kernel_path(), float(o.float().mean()))
    L_, W_ = grid.shape[:2]
    gen = torch.Generator(device='cuda').manual_seed(5)
    pred = {'heatmap': torch.randn(2, 1, L_, W_, device='cuda', generator=gen),
            'loc_offset': torch.randn(2, L_, W_, 2, device='cuda', generator=gen),
            'dim_offset': torch.randn(2, L_, W_, 3, device='cuda', generator=gen),
            'rotation': torch.randn(2, L_, W_, 36, device='cuda', generator=gen)}
    from vfa_b200 import decode
    vals, cells = decode.decode_topk(pred, 50, (L_, W_), (3900.0, 3900.0), (1.7, 0.6, 0.6))
    torch.cuda.synchronize()
    print('decode', tuple(vals.shape), int((cells >= 0).sum()))
