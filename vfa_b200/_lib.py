"""ctypes binding of libvfa_b200.so (the C ABI declared in include/vfa_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not an sm_100 GPU every
entry point raises.  Build the library with `make` (or `python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os

VFA_MAX_LAYERS = 16
VFA_MAX_SCALES = 3

FLAG_FORCE_SIMT = 1
FLAG_FORCE_UMMA = 2
FLAG_BF16_MMA = 4
FLAG_WEIGHTS_PREPARED = 8
FLAG_BF16_FEATURES = 16
FLAG_GRID_SIDE = 32
FLAG_TABLE_PREPARED = 64
FLAG_OUT_NHWC = 128
FLAG_OUT_ACCUMULATE = 256
FLAG_OUT_MULTICAST = 512
FLAG_WS_FORWARD = 1024
FLAG_OUT_PEERS = 4096
FLAG_WS_BACKWARD = 2048

# VFA_B200_LIB: another build of the same library (A/B timing of compile-time kernel variants, scripts/build_variant.sh)
LIB_PATH = os.environ.get('VFA_B200_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib',
                                                           'libvfa_b200.so')

EXPORTS = ['vfa_version', 'vfa_last_error', 'vfa_last_path', 'vfa_reload_env', 'vfa_table_build', 'vfa_table_scale',
           'vfa_nchw_to_nhwc', 'vfa_nhwc_to_nchw', 'vfa_aggregate_workspace_bytes', 'vfa_prepare_weights', 'vfa_aggregate_fwd',
           'vfa_aggregate_bwd', 'vfa_decode_workspace_bytes', 'vfa_decode_topk', 'vfa_multicast_copy']


class Geometry(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('grid_l', C.c_int32), ('grid_w', C.c_int32), ('convert_kind', C.c_int32),
                ('convert_scale', C.c_float), ('convert_offset', C.c_float * 3), ('cube', C.c_float * 3),
                ('layer_z', C.c_float * VFA_MAX_LAYERS), ('image_w', C.c_float), ('image_h', C.c_float),
                ('clamp_lo', C.c_float), ('clamp_hi', C.c_float)]


class Shape(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n_views', C.c_int32), ('channels', C.c_int32), ('n_scales', C.c_int32),
                ('feat_h', C.c_int32 * VFA_MAX_SCALES), ('feat_w', C.c_int32 * VFA_MAX_SCALES)]


class Decode(C.Structure):
    _fields_ = [('batch', C.c_int32), ('grid_l', C.c_int32), ('grid_w', C.c_int32), ('topk', C.c_int32),
                ('n_angles', C.c_int32), ('heatmap', C.c_void_p), ('loc_offset', C.c_void_p), ('loc_stride', C.c_int64 * 3),
                ('dim_offset', C.c_void_p), ('dim_stride', C.c_int64 * 3), ('rotation', C.c_void_p),
                ('rot_stride', C.c_int64 * 3), ('grid_size', C.c_float * 2), ('world_size', C.c_float * 2),
                ('dim_mean', C.c_float * 3)]


class VFAError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libvfa_b200 error {code}: {message}')
        self.code = code


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} not found: build it with `make` at the repository root. '
                          'vfa_b200 has no CPU / PyTorch fallback for the aggregation path.')
    L = C.CDLL(LIB_PATH)
    vp, fp, i32, i64, u32, sz = C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
    L.vfa_version.restype = C.c_int
    L.vfa_reload_env.restype = None
    L.vfa_last_error.restype = C.c_char_p
    L.vfa_last_path.restype = C.c_char_p
    L.vfa_table_build.argtypes = [C.POINTER(Geometry), i32, fp, fp, fp, vp]
    L.vfa_table_scale.argtypes = [fp, i64, i32, i32, fp, vp, vp, vp]
    L.vfa_nchw_to_nhwc.argtypes = [fp, fp, i64, i32, i64, vp]
    L.vfa_nhwc_to_nchw.argtypes = [fp, fp, i64, i32, i64, vp]
    L.vfa_aggregate_workspace_bytes.argtypes = [C.POINTER(Geometry), C.POINTER(Shape), u32]
    L.vfa_aggregate_workspace_bytes.restype = sz
    PP = C.POINTER(C.c_void_p)
    L.vfa_prepare_weights.argtypes = [C.POINTER(Geometry), C.POINTER(Shape), PP, vp, sz, u32, vp]
    L.vfa_aggregate_fwd.argtypes = [C.POINTER(Geometry), C.POINTER(Shape), fp, PP, PP, PP, fp, vp, vp, sz, u32, vp]
    L.vfa_aggregate_bwd.argtypes = [C.POINTER(Geometry), C.POINTER(Shape), fp, PP, PP, vp, fp, PP, PP, PP, vp, sz,
                                    u32, vp]
    L.vfa_multicast_copy.argtypes = [vp, vp, sz, vp]
    L.vfa_multicast_copy.restype = C.c_int
    L.vfa_decode_workspace_bytes.argtypes = [i32]
    L.vfa_decode_workspace_bytes.restype = sz
    L.vfa_decode_topk.argtypes = [C.POINTER(Decode), fp, vp, vp, sz, vp]
    L.vfa_decode_topk.restype = C.c_int
    for name in ('vfa_prepare_weights', 'vfa_table_build', 'vfa_table_scale', 'vfa_nchw_to_nhwc', 'vfa_nhwc_to_nchw', 'vfa_aggregate_fwd',
                 'vfa_aggregate_bwd'):
        getattr(L, name).restype = C.c_int
    if L.vfa_version() != 1:
        raise ImportError(f'{LIB_PATH}: ABI version {L.vfa_version()} != 1')
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise VFAError(rc, lib().vfa_last_error().decode())


def reload_env():
    """Make the library re-read its debug / A-B switches (VFA_* environment variables; it reads them once per process)."""
    lib().vfa_reload_env()


def last_path() -> str:
    return lib().vfa_last_path().decode()


def ptr_array(ptrs):
    """void*[VFA_MAX_SCALES] from a list of ints / None."""
    arr = (C.c_void_p * VFA_MAX_SCALES)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
