"""Generate tests/golden/network_case.npz from the UNMODIFIED reference network (run in the build container only).

    python tests/golden/make_golden_net.py

Builds reference `vfa.model.vfanet.VFANet` (resnet18, 3D mode; matplotlib stubbed -- it only serves the
`visualize=True` branches, reference vfanet.py:84-125), gives it `vfa_b200.network.procedural_state` weights (values
depend on key names only, so the GPU box re-creates them without a checkpoint), and records its eval-mode outputs on a
small rig in fp32 (the reference as shipped) and in fp64 on the fp32 boxes (the "hybrid oracle" of SURVEY.md section
8(c) at network level: the same modules after `.double()`, with the `torch.cat` of reference vfa_op.py:81 answered by
the boxes the fp32 run handed to `F.grid_sample`, so visibility decisions are the fp32 ones): the BEV feature map
entering `fuse` (every 16th channel) and the four head outputs (`rotation`: every 45th bin).

The rig and the inputs live in tests/golden/net_case_inputs.py; tests/test_gpu_network.py rebuilds them from there.
"""
import copy
import os
import sys
import types
import warnings
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for _n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec'):
    sys.modules.setdefault(_n, types.ModuleType(_n))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import vfa.model.vfanet as ref_net                    # noqa: E402  (the reference)
import vfa.model.vfa_op as ref_op                     # noqa: E402
from vfa.utils import make_grid as ref_make_grid      # noqa: E402

from vfa_b200.network import procedural_state        # noqa: E402

from net_case_inputs import IMAGE, STEP, VIEWS, case_inputs   # noqa: E402


class _CatFromQueue:
    """Stands in for the `torch` global of the reference aggregation module: answers its one torch.cat
    (vfa_op.py:81) with the next recorded box tensor, forwards everything else."""
    def __init__(self, queue):
        self._queue = queue

    def cat(self, tensors, dim=0):
        assert dim == -1 and len(tensors) == 4
        return self._queue.pop(0).reshape(tuple(tensors[0].shape[:-1]) + (4,)).to(tensors[0].dtype)

    def __getattr__(self, name):
        return getattr(torch, name)


def run(net, images, calibs, grid, boxes_out=None, boxes_in=None):
    seen = {}
    orig_gs, orig_torch, calls = ref_op.F.grid_sample, ref_op.torch, []
    if boxes_out is not None:                      # record (L,T) and (R,B) of every aggregation call (vfa_op.py:112-113)
        def spy(inp, g, *a, **k):
            calls.append(g.detach().clone())
            return orig_gs(inp, g, *a, **k)
        ref_op.F.grid_sample = spy
    if boxes_in is not None:
        ref_op.torch = _CatFromQueue(list(boxes_in))
    hook = net.fuse.register_forward_pre_hook(lambda m, inp: seen.__setitem__('ortho', inp[0].detach()))
    with torch.no_grad():
        pred = net(images, calibs, grid)
    hook.remove()
    ref_op.F.grid_sample, ref_op.torch = orig_gs, orig_torch
    if boxes_out is not None:
        for i in range(0, len(calls), 4):
            boxes_out.append(torch.cat([calls[i], calls[i + 1]], dim=-1))
    out = {'ortho': seen['ortho'][:, ::16], 'heatmap': pred['heatmap'], 'loc_offset': pred['loc_offset'],
           'dim_offset': pred['dim_offset'], 'rotation': pred['rotation'][..., ::45]}
    return {k: v.contiguous().numpy() for k, v in out.items()}


def main():
    g, images, calibs, grid = case_inputs()
    ref_grid = ref_make_grid(g.world_size, cube_LW=list(g.cube_size[:2]), dataset=g.name)
    assert torch.equal(ref_grid[::STEP, ::STEP][None], grid), 'geometry.grid_for deviates from the reference make_grid'
    args = SimpleNamespace(data=g.name, image_size=IMAGE)
    torch.manual_seed(0)
    net = ref_net.VFANet(args, 'resnet18', g.grid_height, g.cube_size, 360, '3D', False).eval()
    net.load_state_dict(procedural_state(net.state_dict()))
    boxes = []
    f32 = run(net, images, calibs, grid, boxes_out=boxes)
    assert len(boxes) == 3 * VIEWS
    f64 = run(copy.deepcopy(net).double(), images.double(), calibs.double(), grid.double(), boxes_in=boxes)
    store = {}
    for k in f32:
        d = np.abs(f32[k].astype(np.float64) - f64[k])
        print(f'{k:11s} {f32[k].shape}  max|ref64| {np.abs(f64[k]).max():.4g}  max|ref32 - ref64| {d.max():.3g}')
        store['f32_' + k] = f32[k]
        store['f64_' + k] = f64[k]
    np.savez_compressed(os.path.join(HERE, 'network_case.npz'), **store)
    print('wrote network_case.npz', os.path.getsize(os.path.join(HERE, 'network_case.npz')), 'bytes')


if __name__ == '__main__':
    main()
