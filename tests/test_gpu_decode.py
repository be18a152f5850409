"""GPU parity of the decode tail (vfa_decode_topk: sigmoid -> 5 x 5 max-pool NMS -> top-k -> decode at the selected cells)
against the golden outputs of the unmodified reference `ObjectEncoder.decode3d` / `decode2d` and the oracle port.

Integer results (selected cells, orientation bins, number of detections) are compared exactly; floating-point ones within
2e-6 relative (the device's expf against the host's vectorised exp inside torch.sigmoid / torch.exp).  Cells of EQUAL
confidence (plateaus of equal logits all survive the NMS) have no defined order in torch.topk: both sides are put into a
canonical order (confidence descending, then cell) before comparing.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from decode_case_inputs import CLS_THRESH, DIM_MEAN, GRID_SIZE, TOPK, WORLD_SIZE, L, W, case_pred   # noqa: E402
from oracle import decode_port                    # noqa: E402
import vfa_b200                                   # noqa: E402


def _canon(conf, *cols):
    """rows sorted by (conf descending, then the remaining columns ascending)."""
    m = np.stack([-np.asarray(conf, np.float64)] + [np.asarray(c, np.float64).reshape(len(conf), -1)[:, i]
                                                    for c in cols for i in range(np.asarray(c).reshape(len(conf), -1).shape[1])], 1)
    order = np.lexsort(m.T[::-1])
    return m[order]


def _cuda(pred):
    return {k: v.cuda() for k, v in pred.items()}


@pytest.mark.parametrize('seed', [0, 1])
def test_decode_matches_reference_golden(seed):
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'decode_case.npz'))
    pred = case_pred(seed)
    d3 = vfa_b200.decode3d(_cuda(pred), CLS_THRESH, TOPK, GRID_SIZE, WORLD_SIZE, DIM_MEAN)
    want = {k: gold[f's{seed}/3d/{k}'] for k in ('conf', 'location', 'dimension', 'rotation')}
    assert d3['conf'].shape == want['conf'].shape                        # same number of detections above the threshold
    a = _canon(d3['conf'].cpu(), d3['location'].cpu(), d3['dimension'].cpu(), d3['rotation'].cpu())
    b = _canon(want['conf'], want['location'], want['dimension'], want['rotation'])
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-6)
    assert np.array_equal(a[:, -1], b[:, -1].astype(np.float32).astype(np.float64))      # orientation bins: exact
    for name, wild in (('2d', False), ('2dw', True)):
        d2 = vfa_b200.decode2d(_cuda(pred), CLS_THRESH, TOPK, GRID_SIZE, WORLD_SIZE, wildtrack=wild)
        a = _canon(d2['conf'].cpu(), d2['location'].cpu())
        b = _canon(gold[f's{seed}/{name}/conf'], gold[f's{seed}/{name}/location'])
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-6)


def test_decode_cells_layouts_and_batches():
    """Selected cells equal the oracle's top-k indices exactly (per frame of a batch of 3), for heads handed as the
    reference's [B,L,W,C] views, as [B,C,L,W] and as channels-last storage; a frame with fewer survivors than k is padded
    with cell -1 / conf 0."""
    B = 3
    pred = case_pred(5, batch=B)
    for b in range(B):                                         # distinct peak heights: no ties at the top
        pred['heatmap'][b].add_(0.01 * torch.randn(1, L, W, generator=torch.Generator().manual_seed(b)))
    want = [decode_port.topk_lists({k: v[b:b + 1] for k, v in pred.items()}, TOPK, GRID_SIZE, WORLD_SIZE, DIM_MEAN)
            for b in range(B)]
    layouts = {
        'views': _cuda(pred),
        'nchw': {k: (v.cuda() if k == 'heatmap' else v.permute(0, 3, 1, 2).contiguous().cuda()) for k, v in pred.items()},
        'cl': {k: (v.cuda() if k == 'heatmap' else v.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last).cuda())
               for k, v in pred.items()},
    }
    for name, p in layouts.items():
        vals, cells = vfa_b200.decode_topk(p, TOPK, GRID_SIZE, WORLD_SIZE, DIM_MEAN, heads_last=name == 'views')
        for b in range(B):
            o = want[b]
            assert np.array_equal(cells[b].cpu().numpy(), o['index'][0].numpy().astype(np.int32)), (name, b)
            assert np.array_equal(vals[b, :, 6].cpu().numpy(), o['orient_idx'][0].numpy().astype(np.float32)), (name, b)
            for col, key in ((0, 'conf'), (1, 'cy'), (2, 'cx'), (3, 'h'), (4, 'w'), (5, 'l')):
                np.testing.assert_allclose(vals[b, :, col].cpu().numpy(), o[key][0].numpy(), rtol=2e-6, atol=1e-6)
    # fewer NMS survivors than k
    small = {'heatmap': torch.randn(1, 1, 8, 8), 'loc_offset': torch.randn(1, 8, 8, 2)}
    vals, cells = vfa_b200.decode_topk(_cuda(small), 50, (8.0, 8.0), (200.0, 200.0))
    n = int((cells[0] >= 0).sum())
    assert 0 < n < 50 and bool((cells[0, n:] == -1).all()) and float(vals[0, n:].abs().max()) == 0.0
    o = decode_port.topk_lists(small, 50, (8.0, 8.0), (200.0, 200.0))
    assert np.array_equal(cells[0, :n].cpu().numpy(), o['index'][0, :n].numpy().astype(np.int32))
    assert float(o['conf'][0, n:].max()) == 0.0            # what the reference's top-k pads with: suppressed cells
    with pytest.raises(vfa_b200.VFAError, match='topk'):
        vfa_b200.decode_topk(_cuda(small), 5000, (8.0, 8.0), (200.0, 200.0))
