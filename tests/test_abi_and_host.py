"""CPU-side checks: the C-ABI library loads and exports every symbol include/vfa_b200.h declares (no compute calls
without a GPU), ctypes structs match the header layout, and the host mirror fails loudly off-GPU."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import vfa_b200
from vfa_b200 import _lib, geometry, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, 'include', 'vfa_b200.h')) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r'^(?:int|void|size_t|const char\*)\s+(vfa_\w+)\s*\(', _header(), flags=re.M))
    assert declared, 'no prototypes found in include/vfa_b200.h'
    L = _lib.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, f'libvfa_b200.so does not export {missing}'
    assert declared == set(_lib.EXPORTS), 'vfa_b200/_lib.py EXPORTS is out of sync with the header'
    assert L.vfa_version() == int(re.search(r'#define VFA_ABI_VERSION (\d+)', _header()).group(1))


def test_struct_layouts_match_header():
    # vfa_geometry_t: 4 int32 + 1 float + 3 + 3 + 16 floats + 4 floats ; vfa_shape_t: 4 + 3 + 3 int32
    assert ctypes.sizeof(_lib.Geometry) == 4 * 4 + 4 * (1 + 3 + 3 + 16 + 4)
    assert ctypes.sizeof(_lib.Shape) == 4 * 10
    h = _header()
    assert int(re.search(r'#define VFA_MAX_LAYERS (\d+)', h).group(1)) == _lib.VFA_MAX_LAYERS
    assert int(re.search(r'#define VFA_MAX_SCALES (\d+)', h).group(1)) == _lib.VFA_MAX_SCALES
    for name, val in (('FORCE_SIMT', _lib.FLAG_FORCE_SIMT), ('FORCE_UMMA', _lib.FLAG_FORCE_UMMA),
                      ('BF16_MMA', _lib.FLAG_BF16_MMA), ('WEIGHTS_PREPARED', _lib.FLAG_WEIGHTS_PREPARED),
                      ('BF16_FEATURES', _lib.FLAG_BF16_FEATURES), ('GRID_SIDE', _lib.FLAG_GRID_SIDE),
                      ('TABLE_PREPARED', _lib.FLAG_TABLE_PREPARED), ('OUT_NHWC', _lib.FLAG_OUT_NHWC),
                      ('OUT_ACCUMULATE', _lib.FLAG_OUT_ACCUMULATE), ('OUT_MULTICAST', _lib.FLAG_OUT_MULTICAST),
                      ('WS_FORWARD', _lib.FLAG_WS_FORWARD), ('WS_BACKWARD', _lib.FLAG_WS_BACKWARD)):
        assert int(re.search(rf'#define VFA_FLAG_{name}\s+(\d+)u', h).group(1)) == val


def test_argument_validation_needs_no_gpu():
    """Invalid arguments are rejected before any CUDA call, with a message."""
    L = _lib.lib()
    g = vfa_b200.make_geometry(5, (25, 25, 32), [0, 32, 64, 96, 128], (156, 156), 'MultiviewC', (720, 1280))
    assert L.vfa_table_build(ctypes.byref(g), 0, None, None, None, None) == -1
    assert b'n_views' in L.vfa_last_error()
    g.n_layers = 99
    assert L.vfa_table_build(ctypes.byref(g), 1, None, None, None, None) == -1
    assert b'n_layers' in L.vfa_last_error()
    assert L.vfa_table_scale(None, 10, 90, 160, None, None, None, None) == -1
    if not torch.cuda.is_available():
        # with valid arguments and no device the library reports NO_DEVICE instead of computing anything
        g.n_layers = 5
        buf = (ctypes.c_float * 64)()
        rc = L.vfa_table_build(ctypes.byref(g), 1, ctypes.addressof(buf), ctypes.addressof(buf), ctypes.addressof(buf), None)
        assert rc == -5 and b'no CPU fallback' in L.vfa_last_error() or rc == -5


def test_module_mirrors_reference_interface():
    g = geometry.MULTIVIEWC
    args = SimpleNamespace(data=g.name, image_size=g.image_size)
    m = vfa_b200.VFA(256, g.grid_height, np.array(g.cube_size), 1 / 8., args)
    sd = m.state_dict()
    assert list(sd) == ['z_corners', 'corners_offset', 'collapse.weight', 'collapse.bias']
    assert sd['z_corners'].dtype == torch.int64 and sd['z_corners'][:, 0, 0, 2].tolist() == [0, 32, 64, 96, 128]
    assert tuple(sd['corners_offset'].shape) == (1, 1, 1, 1, 8, 3)
    assert sd['corners_offset'][0, 0, 0, 0, 0].tolist() == [-12.5, -12.5, 0.0]
    assert sd['corners_offset'][0, 0, 0, 0, 6].tolist() == [12.5, 12.5, 32.0]
    assert [n for n, _ in m.named_parameters()] == ['collapse.weight', 'collapse.bias']
    # no CPU path: loud failure instead of a silent fallback
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(torch.zeros(1, 256, 90, 160), torch.zeros(3, 4), geometry.grid_for(g)[None])
    with pytest.raises(ValueError, match='unknown dataset'):
        vfa_b200.VFA(8, 160, (25, 25, 32), 1, SimpleNamespace(data='KITTI', image_size=(1, 1))).geometry((4, 4))


def test_geometries_match_configs_of_record():
    assert geometry.MULTIVIEWC.grid_shape == (156, 156) and geometry.MULTIVIEWC.n_layers == 5
    assert geometry.MULTIVIEWX.grid_shape == (160, 250) and geometry.MULTIVIEWX.n_layers == 8
    assert geometry.WILDTRACK.grid_shape == (120, 360) and geometry.WILDTRACK.n_layers == 8
    for g in geometry.GEOMETRIES.values():
        assert g.feature_sizes() == [(90, 160), (45, 80), (23, 40)]
        assert tuple(geometry.grid_for(g).shape) == g.grid_shape + (3,)
    # Wildtrack grid has its axes swapped (reference utils.py:24-25, :33-34)
    wt = geometry.grid_for(geometry.WILDTRACK)
    assert wt[1, 0].tolist() == [4.0, 0.0, 0.0] and wt[0, 1].tolist() == [0.0, 4.0, 0.0]
    mc = geometry.grid_for(geometry.MULTIVIEWC)
    assert mc[0, 1].tolist() == [25.0, 0.0, 0.0] and mc[1, 0].tolist() == [0.0, 25.0, 0.0]


def test_synthetic_inputs_are_deterministic():
    g = geometry.MULTIVIEWC
    a, b = synthetic.ring_calibs(g), synthetic.ring_calibs(g)
    assert torch.equal(a, b) and a.shape == (7, 3, 4) and a.dtype == torch.float32
    f1 = synthetic.features(g, channels=4, seed=3)
    f2 = synthetic.features(g, channels=4, seed=3)
    assert all(torch.equal(x, y) for x, y in zip(f1, f2)) and float(f1[0].min()) >= 0.0
    p = synthetic.collapse_params(g, channels=4)
    assert p[0][0].shape == (4, 20) and p[0][1].shape == (4,)


def test_workspace_sizing_is_host_only(libenv):
    """vfa_aggregate_workspace_bytes is pure host arithmetic: the default C = 256 forward carries the image-plane
    intermediate Y of one frame chunk (bounded by the 6 GiB budget), VFA_FLAG_GRID_SIDE does not; the backward adds its
    CSR / gm / Gs buffers; other channel counts use the generic kernels' small workspace."""
    g = geometry.MULTIVIEWC
    L = _lib.lib()
    zs = list(range(0, g.grid_height, g.cube_size[2]))
    geom = vfa_b200.make_geometry(len(zs), g.cube_size, zs, g.grid_shape, g.name, g.image_size)

    def shape(batch, channels=256):
        sh = _lib.Shape()
        sh.batch, sh.n_views, sh.channels, sh.n_scales = batch, g.n_views, channels, 3
        for s, (h, w) in enumerate(g.feature_sizes()):
            sh.feat_h[s], sh.feat_w[s] = h, w
        return sh

    def ws(batch, flags=0, channels=256):
        sh = shape(batch, channels)
        return L.vfa_aggregate_workspace_bytes(ctypes.byref(geom), ctypes.byref(sh), flags)

    libenv.delenv('VFA_FSIDE_Y_BUDGET_MB', raising=False)
    y_frame = g.n_views * len(zs) * sum(h * w for h, w in g.feature_sizes()) * 256 * 4          # 678 MB
    one = ws(1)
    assert one > y_frame                                       # Y of one frame + CSR / gm / Gs of the backward
    assert ws(4) > one + 3 * y_frame                           # grows with the chunk ...
    assert 0 <= ws(64) - ws(32) < 1 << 20                      # ... up to the 6 GiB budget (chunks are looped; only the
                                                               # per-tile need bytes of the backward still grow with B)
    assert ws(4, channels=16) < 200 << 20                      # generic kernels: weights + small scratch
    # sizing honours flags and direction: the grid-side forward needs no intermediate, a forward-only caller is not
    # told to allocate the backward's CSR / Gs, and vice versa
    fwd, bwd = ws(1, _lib.FLAG_WS_FORWARD), ws(1, _lib.FLAG_WS_BACKWARD)
    assert max(fwd, bwd) == one and y_frame < fwd < one
    assert ws(1, _lib.FLAG_WS_FORWARD | _lib.FLAG_GRID_SIDE) < 128 << 20 < ws(4, _lib.FLAG_WS_FORWARD)
    assert ws(4, _lib.FLAG_WS_FORWARD | _lib.FLAG_FORCE_SIMT) < 200 << 20
    libenv.setenv('VFA_FSIDE_Y_BUDGET_MB', '700')         # one frame per chunk
    assert 0 <= ws(4) - ws(1) < 1 << 20


def test_multiscale_module_state_dict_matches_vfanet_keys():
    """MultiScaleVFA carries the reference network's attribute names, so the vfa8/16/32 entries of a VFANet checkpoint
    (reference vfanet.py:30-32) load into it unchanged."""
    g = geometry.MULTIVIEWX
    m = vfa_b200.MultiScaleVFA(256, g.grid_height, g.cube_size, SimpleNamespace(data=g.name, image_size=g.image_size))
    keys = list(m.state_dict())
    want = [f'vfa{s}.{k}' for s in (8, 16, 32) for k in ('z_corners', 'corners_offset', 'collapse.weight', 'collapse.bias')]
    assert keys == want
    assert tuple(m.vfa16.collapse.weight.shape) == (256, 256 * g.n_layers)
    with pytest.raises(ValueError, match='three FPN scales'):
        m([torch.zeros(1, 1, 256, 4, 4)], torch.zeros(1, 3, 4), geometry.grid_for(g))


def test_network_module_has_the_reference_checkpoint_layout():
    """vfa_b200.network.VFANet: key names / shapes of the reference checkpoint (vfanet.py:27-56, resnet.py:100-110); the
    key-for-key comparison with the reference class itself is in test_reference_compat.py (build container only)."""
    from types import SimpleNamespace
    from vfa_b200.network import VFANet, procedural_state
    args = SimpleNamespace(data='MultiviewC', image_size=(720, 1280))
    net = VFANet(args, 'resnet18', 160, (25, 25, 32), 360, '3D', False)
    sd = net.state_dict()
    assert len(sd) == 112
    want = {'base.conv1.weight': (64, 3, 7, 7), 'base.layer2.0.downsample.0.weight': (128, 64, 1, 1),
            'base.layer4.1.bn2.bias': (512,), 'lat32.weight': (256, 512, 1, 1), 'bn16.weight': (256,),
            'vfa8.collapse.weight': (256, 1280), 'vfa16.z_corners': (5, 1, 1, 3), 'vfa32.corners_offset': (1, 1, 1, 1, 8, 3),
            'fuse.4.running_var': (256,), 'map_classifier.0.weight': (1, 256, 3, 3), 'tytx_pred.3.weight': (2, 256, 3, 3),
            'orient_pred.0.weight': (360, 256, 3, 3), 'thtwtl_pred.1.weight': (256,), 'mean': (3,), 'std': (3,)}
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp, k
    assert 'orient_pred.0.weight' not in VFANet(args, 'resnet34', 160, (25, 25, 32), 360, '2D', False).state_dict()
    # procedural weights: a function of the key names only
    a, b = procedural_state(sd), procedural_state(dict(reversed(list(sd.items()))))
    assert all(torch.equal(a[k], b[k]) for k in sd)
    assert torch.equal(a['vfa16.z_corners'], sd['vfa16.z_corners']) and torch.equal(a['mean'], sd['mean'])
    net.load_state_dict(a)
    # no CPU path: the aggregation raises on host tensors
    with pytest.raises(Exception):
        net(torch.rand(1, 3, 704, 704), torch.rand(1, 3, 4), torch.zeros(1, 4, 4, 3))


def test_network_pretrained_trunk_takes_the_matching_checkpoint_entries(monkeypatch):
    """pretrained=True (reference train.py:90, resnet.py:150-180): entries of the ImageNet checkpoint whose names and
    shapes the trunk knows are loaded (BatchNorm scale / shift -> GroupNorm), everything else is ignored."""
    from types import SimpleNamespace
    from vfa_b200 import network
    fake = {'conv1.weight': torch.full((64, 3, 7, 7), 0.25), 'bn1.weight': torch.full((64,), 2.0),
            'bn1.running_mean': torch.zeros(64), 'layer1.0.conv1.weight': torch.full((64, 64, 3, 3), -0.5),
            'layer2.0.downsample.1.bias': torch.full((128,), 0.125), 'fc.weight': torch.zeros(1000, 512),
            'layer4.1.conv2.weight': torch.zeros(3, 3)}                       # wrong shape: skipped
    seen = []
    monkeypatch.setattr(torch.hub, 'load_state_dict_from_url', lambda url, **kw: (seen.append(url), fake)[1])
    args = SimpleNamespace(data='MultiviewC', image_size=(720, 1280))
    net = network.VFANet(args, 'resnet18', 160, (25, 25, 32), 360, '2D', pretrained=True)
    assert seen == [network._IMAGENET['resnet18']]
    sd = net.base.state_dict()
    assert 'fc.weight' not in sd and 'bn1.running_mean' not in sd
    assert float(sd['conv1.weight'].mean()) == 0.25 and float(sd['bn1.weight'][3]) == 2.0
    assert float(sd['layer1.0.conv1.weight'][0, 0, 0, 0]) == -0.5 and float(sd['layer2.0.downsample.1.bias'][5]) == 0.125
    assert tuple(sd['layer4.1.conv2.weight'].shape) == (512, 512, 3, 3) and float(sd['layer4.1.conv2.weight'].abs().sum()) > 0
