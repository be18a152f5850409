"""vfa_b200 -- B200-native voxelized 3D feature aggregation (the hot path of Jiahao-Ma/VFA).

    from vfa_b200 import VFA, aggregate, build_table

`VFA` is a drop-in for the reference module `vfa.model.vfa_op.VFA`; `aggregate` is the fused batched multi-view
entry, `MultiScaleVFA` the same as an nn.Module (feats [B,V,C,H,W] per scale, calibs, grid -> [B,C,L,W]).  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/vfa_b200.h
(libvfa_b200.so, loaded with ctypes); there is no CPU, Triton or PyTorch-op fallback.
"""
from . import geometry, synthetic                                    # noqa: F401
from ._lib import (VFAError, reload_env, FLAG_BF16_MMA, FLAG_FORCE_SIMT, FLAG_FORCE_UMMA,   # noqa: F401
                   FLAG_WEIGHTS_PREPARED, FLAG_BF16_FEATURES, FLAG_GRID_SIDE, FLAG_TABLE_PREPARED, FLAG_OUT_NHWC,
                   FLAG_OUT_ACCUMULATE, FLAG_OUT_MULTICAST, FLAG_OUT_PEERS)
from .vfa_op import (VFA, ProjectionTable, aggregate, aggregate_forward_raw, build_table, last_kernel_path,  # noqa: F401
                     make_geometry, make_shape, prepare_weights, to_channels_last, workspace_for)

from .streaming import StreamingAggregator                          # noqa: F401
from .graphed import GraphedAggregator                              # noqa: F401
from .vfanet import MultiScaleVFA                                   # noqa: F401
from .decode import decode_topk, decode3d, decode2d                 # noqa: F401

__version__ = '0.1.0'
