"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/vfa/model/vfa_op.py (matplotlib stubbed: it is only used by the `visualize=True`
branch, reference vfa_op.py:7-8, :90-101) and records, for seeded inputs on all three dataset geometries:

  * the reference's own boxes, captured from the `grid` arguments of its F.grid_sample calls
    (reference vfa_op.py:112-113), as fp32 bit patterns;
  * its fp32 output, and its float64 output when fed those fp32 boxes ("hybrid oracle": the reference module
    in .double() with torch.cat at vfa_op.py:81 answered by the captured fp32 boxes);
  * autograd gradients of the float64 hybrid run w.r.t. feature / collapse.weight / collapse.bias;
  * sha256 digests of boxes / visibility / tap indices on the FULL-size grids.

/root/reference does not exist on the GPU box, so nothing but this script reads it; the tests read the .npz.
"""
import contextlib
import copy
import hashlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for _n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec'):
    sys.modules.setdefault(_n, types.ModuleType(_n))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import vfa.model.vfa_op as ref_op                     # noqa: E402  (the reference)
from vfa.utils import make_grid as ref_make_grid      # noqa: E402
from types import SimpleNamespace                     # noqa: E402

from vfa_b200 import geometry, synthetic             # noqa: E402
from oracle import vfa_oracle as onp                  # noqa: E402
from oracle import ref_port                           # noqa: E402

C_SMALL = 6
SMALL_SIZES = [(45, 80), (30, 52), (23, 40)]
STRIDE = {'MultiviewC': 8, 'MultiviewX': 9, 'Wildtrack': 7}


@contextlib.contextmanager
def capture_grid_sample(store):
    orig = ref_op.F.grid_sample

    def wrapper(inp, grid, *a, **k):
        store.append(grid.detach().clone())
        return orig(inp, grid, *a, **k)
    ref_op.F.grid_sample = wrapper
    try:
        yield
    finally:
        ref_op.F.grid_sample = orig


class _TorchProxy:
    """Stands in for the `torch` global of the reference module; answers torch.cat (vfa_op.py:81) with given
    boxes and forwards everything else."""
    def __init__(self, boxes):
        self._boxes = boxes

    def cat(self, tensors, dim=0):
        assert dim == -1 and len(tensors) == 4
        want = tuple(tensors[0].shape[:-1]) + (4,)
        return self._boxes.reshape(want)

    def __getattr__(self, name):
        return getattr(torch, name)


@contextlib.contextmanager
def boxes_injected(boxes):
    orig = ref_op.torch
    ref_op.torch = _TorchProxy(boxes)
    try:
        yield
    finally:
        ref_op.torch = orig


def ref_module(geom, channels, seed):
    args = SimpleNamespace(data=geom.name, image_size=geom.image_size)
    torch.manual_seed(seed)
    return ref_op.VFA(channels, geom.grid_height, geom.cube_size, 1.0, args)


def run_ref_fp32(m, feature, calib, grid):
    """-> (out [1,C,L,W] fp32, boxes [nl, LW, 4] fp32) from the unmodified reference forward."""
    store = []
    with capture_grid_sample(store), torch.no_grad():
        out = m(feature, calib, grid[None])
    lt, rb = store[0][0], store[1][0]                    # [nl, LW, 2] each
    return out, torch.cat([lt, rb], dim=-1).contiguous()


def run_ref_hybrid(m, feature, calib, grid, boxes, grad_out=None):
    """Reference forward in float64 on given fp32 boxes; optional autograd gradients."""
    m64 = copy.deepcopy(m).double()
    f64 = feature.double().clone().requires_grad_(grad_out is not None)
    L, W = grid.shape[:2]
    with boxes_injected(boxes.double().reshape(1, boxes.shape[0], L, W, 4)):
        out = m64(f64, calib.double(), grid[None].double())
        grads = None
        if grad_out is not None:
            out.backward(grad_out.double())
            grads = (f64.grad.clone(), m64.collapse.weight.grad.clone(), m64.collapse.bias.grad.clone())
    return out.detach(), grads


def principal_plane_calib(geom):
    """A projection whose third row makes h2 == 0 exactly on the top corners of layer 0 (x/0 -> +-inf)."""
    probe = onp.to_world(np.array([[0, 0, geom.cube_size[2]]], np.float32), geom.name)
    z0 = float(probe[0, 2])
    H, W = geom.image_size
    return torch.tensor([[900., 0., W / 2, 37.0], [0., 900., H / 2, -11.0], [0., 0., 1., -z0]])


def small_cases():
    out = {}
    for name, geom in geometry.GEOMETRIES.items():
        full = ref_make_grid(geom.world_size, cube_LW=list(geom.cube_size[:2]), dataset=name)
        st = STRIDE[name]
        grid = full[1::st, 2::st].contiguous()
        calibs = synthetic.ring_calibs(geom, n_views=2, in_field=True)
        calibs = torch.cat([calibs, principal_plane_calib(geom)[None]], dim=0)
        g = torch.Generator().manual_seed(7)
        feats = [torch.randn(1, C_SMALL, h, w, generator=g).relu_() for (h, w) in SMALL_SIZES]
        mods = [ref_module(geom, C_SMALL, 11 + s) for s in range(3)]
        out[f'{name}/grid'] = grid.numpy()
        out[f'{name}/calibs'] = calibs.numpy()
        for s in range(3):
            out[f'{name}/feat{s}'] = feats[s][0].numpy()
            out[f'{name}/weight{s}'] = mods[s].collapse.weight.detach().numpy()
            out[f'{name}/bias{s}'] = mods[s].collapse.bias.detach().numpy()
        for v in range(calibs.shape[0]):
            for s in range(3):
                o32, boxes = run_ref_fp32(mods[s], feats[s], calibs[v], grid)
                if s == 0:
                    out[f'{name}/boxes{v}'] = boxes.numpy()
                else:                                    # boxes do not depend on the scale
                    assert np.array_equal(out[f'{name}/boxes{v}'].view(np.uint32), boxes.numpy().view(np.uint32))
                fh, fw = SMALL_SIZES[s]
                b = boxes[None]
                area = ((b[..., 2:] - b[..., :2]).prod(dim=-1) * fh * fw + ref_op.EPSILON).unsqueeze(1)
                vis = torch.logical_and(area > ref_op.EPSILON, area < (fh * fw * ref_op.MAXIMUM_AREA_RATIO))
                out[f'{name}/area{v}_{s}'] = area[0, 0].numpy()
                out[f'{name}/visible{v}_{s}'] = vis[0, 0].numpy()
                finite = bool(torch.isfinite(boxes).all())
                out[f'{name}/out32_{v}_{s}'] = o32[0].numpy()
                if not finite:
                    continue
                want_grad = (s == v % 3)
                gout = None
                if want_grad:
                    gg = torch.Generator().manual_seed(100 + 10 * v + s)
                    gout = torch.randn(o32.shape, generator=gg, dtype=torch.float64)
                    out[f'{name}/gout_{v}_{s}'] = gout[0].numpy()
                o64, grads = run_ref_hybrid(mods[s], feats[s], calibs[v], grid, boxes, gout)
                out[f'{name}/out64_{v}_{s}'] = o64[0].numpy()
                if want_grad:
                    out[f'{name}/dfeat_{v}_{s}'] = grads[0][0].numpy()
                    out[f'{name}/dweight_{v}_{s}'] = grads[1].numpy()
                    out[f'{name}/dbias_{v}_{s}'] = grads[2].numpy()
                # the oracle restatements must already agree here (also enforced by tests on the .npz)
                pb = ref_port.boxes_fp32(calibs[v], grid, geom.grid_height, geom.cube_size, name, geom.image_size)
                assert np.array_equal(pb.numpy().view(np.uint32), boxes.numpy().view(np.uint32)), (name, v)
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def full_size_digests():
    out = {}
    for name, geom in geometry.GEOMETRIES.items():
        grid = ref_make_grid(geom.world_size, cube_LW=list(geom.cube_size[:2]), dataset=name)
        calibs = synthetic.ring_calibs(geom, in_field=True)
        m = ref_module(geom, 1, 0)
        sizes = geom.feature_sizes()
        for v in range(calibs.shape[0]):
            feat = torch.zeros(1, 1, *sizes[0])
            _, boxes = run_ref_fp32(m, feat, calibs[v], grid)
            bn = boxes.numpy()
            out[f'{name}/boxes{v}'] = digest(bn.view(np.uint32))
            for s, (fh, fw) in enumerate(sizes):
                b = boxes[None]
                area = ((b[..., 2:] - b[..., :2]).prod(dim=-1) * fh * fw + ref_op.EPSILON).unsqueeze(1)
                vis = torch.logical_and(area > ref_op.EPSILON, area < (fh * fw * ref_op.MAXIMUM_AREA_RATIO))
                out[f'{name}/visible{v}_{s}'] = digest(vis[0, 0].numpy().astype(np.uint8))
                out[f'{name}/visible_count{v}_{s}'] = int(vis.sum())
                # integer tap indices floor(((c+1)*S-1)/2) in fp32 (ATen GridSampler.h:27-35)
                ix = torch.floor(((boxes[..., [0, 2]] + 1) * fw - 1) / 2)
                iy = torch.floor(((boxes[..., [1, 3]] + 1) * fh - 1) / 2)
                taps = torch.stack([ix[..., 0], iy[..., 0], ix[..., 1], iy[..., 1]], -1).to(torch.int32)
                out[f'{name}/taps{v}_{s}'] = digest(taps.numpy())
        out[f'{name}/grid'] = digest(grid.numpy().view(np.uint32))
    return out


if __name__ == '__main__':
    small = small_cases()
    np.savez_compressed(os.path.join(HERE, 'small_cases.npz'), **small)
    dg = full_size_digests()
    import json
    with open(os.path.join(HERE, 'full_size_digests.json'), 'w') as f:
        json.dump({'torch': torch.__version__, 'digests': dg}, f, indent=1, sort_keys=True)
    print('wrote', len(small), 'arrays,', len(dg), 'digests')
    print('size', os.path.getsize(os.path.join(HERE, 'small_cases.npz')))
