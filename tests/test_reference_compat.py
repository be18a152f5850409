"""Drop-in check at the network level (build container only: needs the reference tree, absent on the GPU box):
the reference `VFANet` constructed on top of `vfa_b200.compat.install()` has exactly the reference's state dict."""
import os
import sys
import types
from types import SimpleNamespace

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'vfa')), reason='reference tree not present')


def _import_vfanet(with_compat):
    for n in [k for k in sys.modules if k == 'vfa' or k.startswith('vfa.')]:
        del sys.modules[n]
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec'):
        sys.modules.setdefault(n, types.ModuleType(n))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if with_compat:
        import vfa_b200.compat as compat
        compat.install()
    import vfa.model.vfanet as vfanet
    return vfanet


def test_reference_vfanet_builds_on_the_drop_in_module():
    import vfa_b200
    args = SimpleNamespace(data='MultiviewC', image_size=(720, 1280))
    ref = _import_vfanet(False).VFANet(args, 'resnet18', 160, (25, 25, 32), 360, '3D', False)
    ref_sd = {k: (tuple(v.shape), v.dtype) for k, v in ref.state_dict().items()}
    net_mod = _import_vfanet(True)
    ours = net_mod.VFANet(args, 'resnet18', 160, (25, 25, 32), 360, '3D', False)
    assert isinstance(ours.vfa8, vfa_b200.VFA) and isinstance(ours.vfa32, vfa_b200.VFA)
    our_sd = {k: (tuple(v.shape), v.dtype) for k, v in ours.state_dict().items()}
    assert our_sd == ref_sd
    # a reference checkpoint loads into the drop-in network and back, key for key
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    for k in ('vfa8.collapse.weight', 'vfa16.z_corners', 'vfa32.corners_offset'):
        assert torch.equal(ours.state_dict()[k], ref.state_dict()[k])
    # and the optimiser sees the collapse parameters (reference train.py:256)
    names = {n for n, _ in ours.named_parameters()}
    assert {'vfa8.collapse.weight', 'vfa8.collapse.bias', 'vfa16.collapse.weight', 'vfa32.collapse.bias'} <= names
    for n in [k for k in sys.modules if k == 'vfa' or k.startswith('vfa.')]:
        del sys.modules[n]


def test_own_network_module_has_the_reference_state_dict():
    """vfa_b200.network.VFANet (the batched, fused network of BASELINE config 5) against the reference class: same keys in
    the same order, same shapes / dtypes, for both backbones and both modes; the trunks are bit-identical on CPU."""
    from vfa_b200.network import VFANet, procedural_state
    args = SimpleNamespace(data='Wildtrack', image_size=(1080, 1920))
    ref_mod = _import_vfanet(False)
    for base, mode in (('resnet18', '3D'), ('resnet34', '2D')):
        ref = ref_mod.VFANet(args, base, 32, (4, 4, 4), 360, mode, False)
        ours = VFANet(args, base, 32, (4, 4, 4), 360, mode, False)
        a = [(k, tuple(v.shape), v.dtype) for k, v in ref.state_dict().items()]
        b = [(k, tuple(v.shape), v.dtype) for k, v in ours.state_dict().items()]
        assert a == b
    sd = procedural_state(ref.state_dict())
    ref.load_state_dict(sd)
    ours.load_state_dict(sd)
    x = torch.rand(2, 3, 64, 96)
    with torch.no_grad():
        for r, o in zip(ref.base(x), ours.base(x)):
            assert torch.equal(r, o)
    for n in [k for k in sys.modules if k == 'vfa' or k.startswith('vfa.')]:
        del sys.modules[n]


def test_compat_install_can_replace_the_whole_network_module():
    """compat.install(network=True): the reference scripts' `from vfa.model.vfanet import VFANet` resolves to the B200
    network, the rest of the reference package (vfa.utils, vfa.config ...) stays the reference's own."""
    import vfa_b200.compat as compat
    from vfa_b200.network import VFANet as Ours
    for n in [k for k in sys.modules if k == 'vfa' or k.startswith('vfa.')]:
        del sys.modules[n]
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec'):
        sys.modules.setdefault(n, types.ModuleType(n))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    compat.install(network=True)
    from vfa.model.vfanet import VFANet          # what reference train.py:13 / evaluate.py:9 execute
    from vfa.utils import make_grid               # still the reference's
    assert VFANet is Ours and 'reference' in make_grid.__code__.co_filename
    args = SimpleNamespace(data='MultiviewX', image_size=(1080, 1920))
    net = VFANet(args=args, grid_height=64, cube_size=(4, 4, 8), angle_range=360, mode='2D', pretrained=False)   # train.py:249
    assert net.vfa8.num_grid_layer == 8 and not hasattr(net, 'orient_pred')
    for n in [k for k in sys.modules if k == 'vfa' or k.startswith('vfa.')]:
        del sys.modules[n]
