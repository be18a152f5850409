"""Pin the oracle (oracle/) to vectors produced by the unmodified reference (tests/golden/make_golden.py).

Bit-exact items: boxes, area, visible, tap indices.  Floating point: fp64 hybrid features to 1e-11.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_port
from oracle import vfa_oracle as onp
from vfa_b200 import geometry, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NAMES = ['MultiviewC', 'MultiviewX', 'Wildtrack']
SMALL_SIZES = [(45, 80), (30, 52), (23, 40)]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('name', NAMES)
def test_boxes_area_visible_bit_exact_small(golden, name):
    geom = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    for v in range(calibs.shape[0]):
        want = golden[f'{name}/boxes{v}']
        got = onp.project_boxes(calibs[v], grid, geom.grid_height, geom.cube_size, name, geom.image_size)
        assert np.array_equal(_bits(got), _bits(want)), f'numpy oracle boxes differ: {name} view {v}'
        gport = ref_port.boxes_fp32(torch.from_numpy(calibs[v]), torch.from_numpy(grid), geom.grid_height,
                                    geom.cube_size, name, geom.image_size).numpy()
        assert np.array_equal(_bits(gport), _bits(want)), f'torch port boxes differ: {name} view {v}'
        for s, (fh, fw) in enumerate(SMALL_SIZES):
            area, vis = onp.area_visible(got, fh, fw)
            assert np.array_equal(_bits(area), _bits(golden[f'{name}/area{v}_{s}']))
            assert np.array_equal(vis, golden[f'{name}/visible{v}_{s}'])


@pytest.mark.parametrize('name', NAMES)
def test_full_size_digests(digests, name):
    """Full grids (156x156x5, 160x250x8, 120x360x8), ring + in-field camera: digests of the reference's bits."""
    geom = geometry.GEOMETRIES[name]
    grid = onp.make_grid(geom.world_size, geom.cube_size[:2], name)
    assert grid.shape[:2] == geom.grid_shape
    assert _digest(_bits(grid)) == digests[f'{name}/grid']
    assert np.array_equal(_bits(geometry.grid_for(geom).numpy()), _bits(grid))
    calibs = synthetic.ring_calibs(geom, in_field=True).numpy()
    for v in range(calibs.shape[0]):
        boxes = onp.project_boxes(calibs[v], grid, geom.grid_height, geom.cube_size, name, geom.image_size)
        assert _digest(_bits(boxes)) == digests[f'{name}/boxes{v}'], (name, v)
        for s, (fh, fw) in enumerate(geom.feature_sizes()):
            _, vis = onp.area_visible(boxes, fh, fw)
            assert _digest(vis.astype(np.uint8)) == digests[f'{name}/visible{v}_{s}']
            assert int(vis.sum()) == digests[f'{name}/visible_count{v}_{s}']
            taps = np.stack([onp.tap_index(boxes[..., 0], fw), onp.tap_index(boxes[..., 1], fh),
                             onp.tap_index(boxes[..., 2], fw), onp.tap_index(boxes[..., 3], fh)], -1)
            assert _digest(taps.astype(np.int32)) == digests[f'{name}/taps{v}_{s}']


@pytest.mark.parametrize('name', NAMES)
def test_hybrid_features_match_reference_fp64(golden, name):
    geom = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    L, W = grid.shape[:2]
    checked = 0
    for v in range(calibs.shape[0]):
        boxes = golden[f'{name}/boxes{v}']
        for s in range(3):
            key = f'{name}/out64_{v}_{s}'
            if key not in golden:
                continue
            feat, w, b = golden[f'{name}/feat{s}'], golden[f'{name}/weight{s}'], golden[f'{name}/bias{s}']
            want = golden[key]
            # numpy restatement, reference form (integral image + 4 bilinear samples)
            vis = golden[f'{name}/visible{v}_{s}']
            got = onp.collapse_relu(onp.vox_features_integral(feat, boxes, vis), w, b).reshape(-1, L, W)
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
            # numpy restatement, direct coverage-weighted form (what the CUDA kernels evaluate)
            got_d = onp.collapse_relu(onp.vox_features_direct(feat, boxes, vis), w, b).reshape(-1, L, W)
            np.testing.assert_allclose(got_d, want, rtol=1e-9, atol=1e-9)
            # torch port in float64 on the fp32 boxes
            got_p = ref_port.vfa_forward(torch.from_numpy(feat).double(), None, torch.from_numpy(grid),
                                         torch.from_numpy(w).double(), torch.from_numpy(b).double(),
                                         geom.grid_height, geom.cube_size, name, geom.image_size,
                                         boxes=torch.from_numpy(boxes))[0].numpy()
            np.testing.assert_allclose(got_p, want, rtol=1e-9, atol=1e-9)
            checked += 1
    assert checked >= 9


@pytest.mark.parametrize('name', NAMES)
def test_port_fp32_matches_reference_fp32(golden, name):
    """The fp32 port runs the reference's operator sequence: same bits on the same torch build, and in any case
    far inside the reference's own fp32 noise."""
    geom = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    for v in range(calibs.shape[0]):
        for s in range(3):
            feat, w, b = golden[f'{name}/feat{s}'], golden[f'{name}/weight{s}'], golden[f'{name}/bias{s}']
            got = ref_port.vfa_forward(torch.from_numpy(feat), torch.from_numpy(calibs[v]), torch.from_numpy(grid),
                                       torch.from_numpy(w), torch.from_numpy(b), geom.grid_height,
                                       geom.cube_size, name, geom.image_size)[0].numpy()
            want = golden[f'{name}/out32_{v}_{s}']
            ok = np.isfinite(want)
            np.testing.assert_allclose(got[ok], want[ok], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name', NAMES)
def test_port_gradients_match_reference_autograd(golden, name):
    geom = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    checked = 0
    for v in range(calibs.shape[0]):
        for s in range(3):
            key = f'{name}/dfeat_{v}_{s}'
            if key not in golden:
                continue
            feat = torch.from_numpy(golden[f'{name}/feat{s}']).double().requires_grad_(True)
            w = torch.from_numpy(golden[f'{name}/weight{s}']).double().requires_grad_(True)
            b = torch.from_numpy(golden[f'{name}/bias{s}']).double().requires_grad_(True)
            out = ref_port.vfa_forward(feat, None, torch.from_numpy(grid), w, b, geom.grid_height, geom.cube_size,
                                       name, geom.image_size, boxes=torch.from_numpy(golden[f'{name}/boxes{v}']))
            out.backward(torch.from_numpy(golden[f'{name}/gout_{v}_{s}'])[None])
            np.testing.assert_allclose(feat.grad.numpy(), golden[key], rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(w.grad.numpy(), golden[f'{name}/dweight_{v}_{s}'], rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(b.grad.numpy(), golden[f'{name}/dbias_{v}_{s}'], rtol=1e-10, atol=1e-10)
            checked += 1
    assert checked >= 3


def test_golden_has_ghosts_and_infinite_projections(golden):
    """The fixtures must exercise the reference's missing depth test and the x/0 path."""
    name = 'MultiviewC'
    geom = geometry.GEOMETRIES[name]
    grid, calibs = golden[f'{name}/grid'], golden[f'{name}/calibs']
    # in-field camera (index 2): some voxels are behind the camera yet pass `visible`
    P = calibs[2].astype(np.float64)
    centre = grid.reshape(-1, 3).astype(np.float64) + np.array([0, 0, 16.0])
    depth = centre @ P[2, :3] + P[2, 3]
    vis = golden[f'{name}/visible2_0'][0]
    assert (depth < 0).any() and vis[depth < 0].any()
    # principal-plane camera (index 3): h2 == 0 on layer-0 top corners -> boxes pinned at the clamp limits
    b = golden[f'{name}/boxes3'][0]
    assert np.isfinite(b).all() and ((b[:, 2] == np.float32(0.95)) | (b[:, 0] == -1)).any()


def test_decode_port_matches_the_unmodified_reference_decode():
    """oracle/decode_port.py vs the outputs of the reference's own `ObjectEncoder.decode3d` / `decode2d`
    (tests/golden/make_golden_decode.py -> decode_case.npz): identical detections, bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from decode_case_inputs import CLS_THRESH, DIM_MEAN, GRID_SIZE, TOPK, WORLD_SIZE, case_pred
    from oracle import decode_port
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'decode_case.npz'))
    for seed in (0, 1):
        pred = case_pred(seed)
        got = {'3d': decode_port.decode3d(pred, CLS_THRESH, TOPK, GRID_SIZE, WORLD_SIZE, DIM_MEAN),
               '2d': decode_port.decode2d(pred, CLS_THRESH, TOPK, GRID_SIZE, WORLD_SIZE),
               '2dw': decode_port.decode2d(pred, CLS_THRESH, TOPK, GRID_SIZE, WORLD_SIZE, wildtrack=True)}
        for name, d in got.items():
            for k, v in d.items():
                want = gold[f's{seed}/{name}/{k}']
                assert v.shape == want.shape and 0 < want.shape[0] < TOPK, (seed, name, k, want.shape)
                assert np.array_equal(v.numpy(), want), (seed, name, k)
