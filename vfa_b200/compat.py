"""Make the reference code base use the B200 aggregation without editing it.

The reference imports the module with `from vfa.model.vfa_op import VFA` (reference vfa/model/vfanet.py:10).
`install()` registers a stand-in `vfa.model.vfa_op` module in `sys.modules` *before* `vfa.model.vfanet` is imported, so
`VFANet(...)` (reference vfanet.py:30-32) builds `vfa_b200.VFA` modules: same constructor, same state-dict keys, and
`train.py` / `evaluate.py` run unchanged on a B200.  The reference's own module also imports matplotlib at the top for
its `visualize=True` branch (vfa_op.py:7-8); the stand-in does not need it.
"""
from __future__ import annotations

import sys
import types

from .vfa_op import VFA

EPSILON = 1e-6               # reference vfa_op.py:14 (kept for code that imports the constants)
MAXIMUM_AREA_RATIO = 0.3     # reference vfa_op.py:15


def install(force: bool = False, network: bool = False) -> types.ModuleType:
    """Register the stand-in.  Call before importing `vfa.model.vfanet`; with force=True an already imported reference
    `vfa.model.vfa_op` (and `vfa.model.vfanet`'s binding of VFA) is replaced as well.

    network=True additionally registers a stand-in `vfa.model.vfanet` whose `VFANet` is `vfa_b200.network.VFANet` (same
    constructor, forward and checkpoint keys; laterals batched over cameras, one fused aggregation call), so the
    reference's `from vfa.model.vfanet import VFANet` (train.py:13, evaluate.py:9) picks up the whole B200 network."""
    name = 'vfa.model.vfa_op'
    if name in sys.modules and not force and getattr(sys.modules[name], '__vfa_b200__', False) is False:
        raise RuntimeError(f'{name} is already imported; call install(force=True) to replace it')
    mod = types.ModuleType(name)
    mod.__doc__ = 'vfa_b200 stand-in for the reference aggregation module'
    mod.__vfa_b200__ = True
    mod.VFA = VFA
    mod.EPSILON = EPSILON
    mod.MAXIMUM_AREA_RATIO = MAXIMUM_AREA_RATIO
    sys.modules[name] = mod
    net = sys.modules.get('vfa.model.vfanet')
    if net is not None and force:
        net.VFA = VFA
    parent = sys.modules.get('vfa.model')
    if parent is not None:
        parent.vfa_op = mod
    if network:
        from .network import VFANet
        net_name = 'vfa.model.vfanet'
        if net_name in sys.modules and not force and getattr(sys.modules[net_name], '__vfa_b200__', False) is False:
            raise RuntimeError(f'{net_name} is already imported; call install(force=True, network=True) to replace it')
        net_mod = types.ModuleType(net_name)
        net_mod.__doc__ = 'vfa_b200 stand-in for the reference network module'
        net_mod.__vfa_b200__ = True
        net_mod.VFANet = VFANet
        net_mod.VFA = VFA
        sys.modules[net_name] = net_mod
        if parent is not None:
            parent.vfanet = net_mod
    return mod
