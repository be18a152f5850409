// Shared host/device helpers of libvfa_b200: error plumbing and the per-box tap derivation that BOTH the
// parity-checked table kernels and the aggregation kernels call (so what is checked is what is used).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vfa_b200.h"

namespace vfa {

// ---------------------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void set_path(const char* name);

#define VFA_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::vfa::set_error(__VA_ARGS__);        \
      return (code);                        \
    }                                       \
  } while (0)

#define VFA_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::vfa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return VFA_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define VFA_LAUNCH_CHECK(name)                                                                  \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) {                                                                   \
      ::vfa::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));               \
      return VFA_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

int check_device();   // VFA_OK iff the current device is sm_100

// ---------------------------------------------------------------------------------------------------------
// process-wide state (host): everything the library keeps besides the thread-local error string
// ---------------------------------------------------------------------------------------------------------
// Debug / A-B switches, read from the environment ONCE (first call) -- never on the launch path.  vfa_reload_env()
// re-reads them (tests and timing scripts that flip a switch inside one process).
struct RuntimeConfig {
  int pool_tile;          // VFA_POOL_TILE      1 (default): staged-tile pooling for batch >= 2; 2: any batch; 0: off
  int pool_tile_cap;      // VFA_POOL_TILE_CAP  pool sizes of the tiles' chunk lists in percent (100)
  int tile_variant;       // VFA_TILE_VARIANT   pool_tile_kernel debug bits
  int tile_order;         // VFA_TILE_ORDER     tile_order_kernel: 0 = by size (default), 1 = force the bitonic sort, 2 = identity
  int pool_list;          // VFA_POOL_LIST      0: walking kernel instead of the quads' texel lists
  int pool_quad;          // VFA_POOL_QUAD      0: one-cell-per-warp comparison kernel
  int pool_list_cap;      // VFA_POOL_LIST_CAP  list entries per (view, scale, layer) of a quad's slot (0 = default)
  int fside_compact;      // VFA_FSIDE_COMPACT  0: whole 256-row GEMM tiles instead of compacted rows
  int fside_no_skip;      // VFA_FSIDE_NO_SKIP  1: multiply every (tile, layer)
  long long y_budget_mb;  // VFA_FSIDE_Y_BUDGET_MB  Y bytes held per frame chunk (0 = 6 GiB)
  int umma_variant;       // VFA_UMMA_VARIANT   debug bits of the forward launchers
  int fwd_gridside;       // VFA_FWD_GRIDSIDE   1: grid-side fused kernel by default
  int bwd_scatter, bwd_untiled, bwd_generic, bwd_cublas_dw;   // VFA_BWD_*  (set = 1)
  int bwd_csr_per_box;    // VFA_BWD_CSR_PER_BOX  CSR capacity per box (0 = default)
};
const RuntimeConfig& runtime_config();

// Per-device caches (a process may drive several GPUs: nn.DataParallel, `with torch.cuda.device(d)`): keyed by the
// current device, filled under a mutex.  `slot` names the cached quantity.
enum DeviceCacheSlot { DC_YGEMM_CLUSTERS = 0, DC_YGEMM_COMPACT_CLUSTERS, DC_DWEIGHT_CLUSTERS, DC_SM_COUNT, DC_SLOTS };
int device_cache_get(int slot);            // 0 = not cached yet
void device_cache_set(int slot, int value);

// ---------------------------------------------------------------------------------------------------------
// per-box taps (device)
// ---------------------------------------------------------------------------------------------------------
#define VFA_EPSILON_F 1e-6f          // reference vfa_op.py:14
#define VFA_AREA_RATIO 0.3           // reference vfa_op.py:15

// Constants of one feature scale, prepared on the host.
struct ScaleConst {
  int fh, fw;
  float fhf, fwf;      // (float)fh, (float)fw
  float area_max;      // (float)(fh*fw*0.3) : the Python double threshold rounded to fp32 by the comparison
  double fhd, fwd;
};

__host__ inline ScaleConst make_scale_const(int fh, int fw) {
  ScaleConst s;
  s.fh = fh;
  s.fw = fw;
  s.fhf = (float)fh;
  s.fwf = (float)fw;
  s.area_max = (float)((double)(fh * fw) * VFA_AREA_RATIO);
  s.fhd = (double)fh;
  s.fwd = (double)fw;
  return s;
}

// fp32, separately rounded, left to right: ((R-L)*(B-T)) * fH * fW + EPS        (reference vfa_op.py:104-105)
__device__ __forceinline__ float box_area_f32(float4 box, const ScaleConst& sc) {
  float w = __fsub_rn(box.z, box.x);
  float h = __fsub_rn(box.w, box.y);
  float a = __fmul_rn(w, h);
  a = __fmul_rn(a, sc.fhf);
  a = __fmul_rn(a, sc.fwf);
  return __fadd_rn(a, VFA_EPSILON_F);
}

// (area > EPS) & (area < fH*fW*0.3); NaN compares false on both sides            (reference vfa_op.py:106)
__device__ __forceinline__ bool box_visible(float area, const ScaleConst& sc) {
  return (area > VFA_EPSILON_F) && (area < sc.area_max);
}

// fp32 unnormalised sampling coordinate of F.grid_sample, align_corners=False: ((c+1)*S-1)/2
// (torch ATen/native/GridSampler.h:27-35 behind reference vfa_op.py:112-115)
__device__ __forceinline__ float unnormalize_f32(float c, float size) {
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.0f), size), 1.0f), 2.0f);
}

// What the pooling needs for one (box, scale): the separable coverage weights of the box over the texel
// lattice, with 1/area and the visibility mask folded into the row weights.
//   sum_{i<ny} sum_{j<nx} wy(i) * wx(j) * f[y0+i][x0+j]  ==  visible * (LT + RB - RT - LB) / area
// where wx(j) = wx_first for j == 0, wx_last for j == nx-1 (nx > 1), 1 otherwise (same for wy with wy_mid).
// Indices and fractions come from a float64 evaluation of the fp32 box edges (the fp32 floor is what
// vfa_table_scale reports for the bit-exact check; the two differ only when fp32 rounding of the coordinate
// crosses an integer, where the float64 one is the faithful evaluation of the reference's formula).
struct BoxTaps {
  int x0, y0;        // first texel column / row
  int nx, ny;        // tap counts (0 when the box is not visible)
  float wx_first, wx_last;
  float wy_first, wy_last, wy_mid;
};

__device__ __forceinline__ void axis_taps(double lo, double hi, double size, int limit, int& first, int& count,
                                          double& w_first, double& w_last) {
  double xl = ((lo + 1.0) * size - 1.0) * 0.5;
  double xr = ((hi + 1.0) * size - 1.0) * 0.5;
  double fl = floor(xl), fr = floor(xr);
  int j0 = (int)fl + 1, j1 = (int)fr + 1;
  if (j0 == j1) {
    w_first = xr - xl;
    w_last = 0.0;
  } else {
    w_first = 1.0 - (xl - fl);
    w_last = xr - fr;
  }
  if (j0 < 0) j0 = 0;                 // cannot happen for clamp_lo >= -1; keeps every index in range regardless
  if (j1 > limit - 1) j1 = limit - 1;  // cannot happen for the validated clamp_hi (host check in vfa_api.cu)
  first = j0;
  count = j1 - j0 + 1;
}

__device__ __forceinline__ BoxTaps derive_taps(float4 box, const ScaleConst& sc) {
  BoxTaps t;
  float area = box_area_f32(box, sc);
  bool vis = box_visible(area, sc);
  if (!vis) {
    t.x0 = t.y0 = 0;
    t.nx = t.ny = 0;
    t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
    return t;
  }
  double L = (double)box.x, T = (double)box.y, R = (double)box.z, B = (double)box.w;
  double inv_area = 1.0 / ((R - L) * (B - T) * sc.fhd * sc.fwd + 1e-6);
  double wxf, wxl, wyf, wyl;
  axis_taps(L, R, sc.fwd, sc.fw, t.x0, t.nx, wxf, wxl);
  axis_taps(T, B, sc.fhd, sc.fh, t.y0, t.ny, wyf, wyl);
  t.wx_first = (float)wxf;
  t.wx_last = (float)wxl;
  t.wy_first = (float)(wyf * inv_area);
  t.wy_last = (float)(wyl * inv_area);
  t.wy_mid = (float)inv_area;
  return t;
}

__device__ __forceinline__ float tap_wx(const BoxTaps& t, int j) {
  return j == 0 ? t.wx_first : (j == t.nx - 1 ? t.wx_last : 1.0f);
}
__device__ __forceinline__ float tap_wy(const BoxTaps& t, int i) {
  return i == 0 ? t.wy_first : (i == t.ny - 1 ? t.wy_last : t.wy_mid);
}

// Problem description handed to the aggregation kernels by value.
struct AggParams {
  int B, V, C, nl, S, L, W;          // batch, views, channels, layers, scales, BEV size
  int LW, K;                         // L*W, C*nl
  ScaleConst sc[VFA_MAX_SCALES];
  const float* feats[VFA_MAX_SCALES];    // [B, V, fh, fw, C]
  const float* wprep[VFA_MAX_SCALES];    // prepared weights (layout depends on the kernel family)
  const float* bias[VFA_MAX_SCALES];     // [C]
  const float* boxes;                    // [V, nl, LW, 4]
  float* out;                            // [B, C, L, W]
  int y_bf16;                            // feature-side forward: the intermediate Y is stored in bf16 (VFA_FLAG_BF16_MMA)
  uint32_t* mask;                        // nullptr or [B, V, S, ceil(C/32), LW]: ReLU pass bits (word o/32, bit o%32)
};

// Per-(view, scale, layer, cell) gather recipe in its compact global form (32 bytes), built by a small pre-pass so the
// hot kernel carries no float64 arithmetic (the fp64 pipe throttled the producers: 15 % of their stall samples).
struct __align__(16) TapRec {
  int xy;          // x0 | y0 << 16
  int nxy;         // nx | ny << 16   (0 = not visible)
  float wx_first, wx_last, wy_first, wy_last, wy_mid;
  int pad;
};

}  // namespace vfa
