// extern "C" entry points of libvfa_b200.so (include/vfa_b200.h): argument validation, dispatch, no state.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "vfa_common.cuh"

namespace vfa {

static thread_local char g_error[512] = "";
static thread_local char g_path[64] = "none";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void set_path(const char* name) {
  strncpy(g_path, name, sizeof(g_path) - 1);
  g_path[sizeof(g_path) - 1] = 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
static RuntimeConfig g_config;
static std::once_flag g_config_once;
static std::mutex g_state_mutex;
static void load_config() {
  RuntimeConfig c;
  c.pool_tile = env_int("VFA_POOL_TILE", 1);
  c.pool_tile_cap = env_int("VFA_POOL_TILE_CAP", 100);
  c.tile_variant = env_int("VFA_TILE_VARIANT", 0);
  c.tile_order = env_int("VFA_TILE_ORDER", 0);
  c.pool_list = env_int("VFA_POOL_LIST", 1);
  c.pool_quad = env_int("VFA_POOL_QUAD", 1);
  c.pool_list_cap = env_int("VFA_POOL_LIST_CAP", 0);
  c.fside_compact = env_int("VFA_FSIDE_COMPACT", 1);
  c.fside_no_skip = env_int("VFA_FSIDE_NO_SKIP", 0);
  const char* mb = getenv("VFA_FSIDE_Y_BUDGET_MB");
  c.y_budget_mb = mb != nullptr ? atoll(mb) : 0;
  c.umma_variant = env_int("VFA_UMMA_VARIANT", 0);
  c.fwd_gridside = env_int("VFA_FWD_GRIDSIDE", 0);
  c.bwd_scatter = getenv("VFA_BWD_SCATTER") != nullptr;
  c.bwd_untiled = getenv("VFA_BWD_UNTILED") != nullptr;
  c.bwd_generic = getenv("VFA_BWD_GENERIC") != nullptr;
  c.bwd_cublas_dw = getenv("VFA_BWD_CUBLAS_DW") != nullptr;
  c.bwd_csr_per_box = env_int("VFA_BWD_CSR_PER_BOX", 0);
  std::lock_guard<std::mutex> lock(g_state_mutex);
  g_config = c;
}
const RuntimeConfig& runtime_config() {
  std::call_once(g_config_once, load_config);
  return g_config;
}

constexpr int MAX_DEVICES = 64;
static int g_device_cache[MAX_DEVICES][DC_SLOTS];
int device_cache_get(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 0;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  return g_device_cache[dev][slot];
}
void device_cache_set(int slot, int value) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  g_device_cache[dev][slot] = value;
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libvfa_b200 has no CPU fallback)", cudaGetErrorString(e));
    return VFA_ERR_NO_DEVICE;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("device %d has compute capability %d.x; libvfa_b200 is built for sm_100a only", dev, major);
    return VFA_ERR_NO_DEVICE;
  }
  return VFA_OK;
}

// implemented in the kernel translation units
int launch_table_build(const vfa_geometry_t*, int, const float*, const float*, float*, cudaStream_t);
int launch_table_scale(const float*, long long, int, int, float*, uint8_t*, int32_t*, cudaStream_t);
int launch_transpose(const float*, float*, long long, int, long long, cudaStream_t);
size_t simt_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*);
int prep_weights_simt(const AggParams&, const float* const*, void*, cudaStream_t);
int launch_fwd_simt(AggParams, const float* const*, void*, uint32_t, cudaStream_t);
int prep_weights_umma(const AggParams&, const float* const*, void*, uint32_t, cudaStream_t);
size_t umma_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*, uint32_t);
bool umma_supported(const vfa_geometry_t*, const vfa_shape_t*, uint32_t);
int launch_fwd_umma(AggParams, const float* const*, void*, size_t, uint32_t, cudaStream_t);
size_t bwd_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*);
int launch_bwd(AggParams, const float* const*, const float*, float* const*, float* const*, float* const*, void*, size_t,
               uint32_t, cudaStream_t);

size_t decode_workspace_bytes(int B);
int launch_multicast_copy(const void* src, void* mc_dst, size_t n_bytes, cudaStream_t st);
int launch_decode(const vfa_decode_t* d, float* out_vals, int32_t* out_cell, void* ws, cudaStream_t st);

static int validate_geometry(const vfa_geometry_t* g) {
  VFA_REQUIRE(g != nullptr, VFA_ERR_INVALID_ARGUMENT, "geometry is NULL");
  VFA_REQUIRE(g->n_layers >= 1 && g->n_layers <= VFA_MAX_LAYERS, VFA_ERR_INVALID_ARGUMENT,
              "n_layers=%d outside [1,%d]", g->n_layers, VFA_MAX_LAYERS);
  VFA_REQUIRE(g->grid_l >= 1 && g->grid_w >= 1, VFA_ERR_INVALID_ARGUMENT, "empty BEV grid %dx%d", g->grid_l, g->grid_w);
  VFA_REQUIRE(g->convert_kind == VFA_CONVERT_DIV || g->convert_kind == VFA_CONVERT_AFFINE, VFA_ERR_INVALID_ARGUMENT,
              "unknown convert_kind %d", g->convert_kind);
  VFA_REQUIRE(g->image_w > 0.f && g->image_h > 0.f, VFA_ERR_INVALID_ARGUMENT, "image size must be positive");
  VFA_REQUIRE(g->clamp_lo < g->clamp_hi, VFA_ERR_INVALID_ARGUMENT, "crange (%g, %g) is empty", g->clamp_lo, g->clamp_hi);
  return VFA_OK;
}

// The pooling kernels evaluate the box mean by the direct coverage-weighted sum, which equals the reference's
// four-corner integral-image sampling as long as no bilinear tap falls outside the map (SURVEY.md section 8(a)
// row A4): needs clamp_lo >= -1 and floor(unnormalise(clamp_hi)) + 1 <= S - 1 on both axes of every scale.
static int validate_shape(const vfa_geometry_t* g, const vfa_shape_t* sh) {
  VFA_REQUIRE(sh != nullptr, VFA_ERR_INVALID_ARGUMENT, "shape is NULL");
  VFA_REQUIRE(sh->batch >= 1 && sh->n_views >= 1 && sh->channels >= 1, VFA_ERR_INVALID_ARGUMENT,
              "batch=%d views=%d channels=%d must be positive", sh->batch, sh->n_views, sh->channels);
  VFA_REQUIRE(sh->n_scales >= 1 && sh->n_scales <= VFA_MAX_SCALES, VFA_ERR_INVALID_ARGUMENT,
              "n_scales=%d outside [1,%d]", sh->n_scales, VFA_MAX_SCALES);
  VFA_REQUIRE(sh->batch <= 65535, VFA_ERR_UNSUPPORTED, "batch %d > 65535", sh->batch);
  VFA_REQUIRE(g->clamp_lo >= -1.0f, VFA_ERR_UNSUPPORTED,
              "crange[0]=%g < -1: left/top taps would fall outside the feature map", g->clamp_lo);
  for (int s = 0; s < sh->n_scales; ++s) {
    VFA_REQUIRE(sh->feat_h[s] >= 2 && sh->feat_w[s] >= 2, VFA_ERR_INVALID_ARGUMENT, "scale %d: feature map %dx%d too small",
                s, sh->feat_h[s], sh->feat_w[s]);
    const int dims[2] = {sh->feat_h[s], sh->feat_w[s]};
    for (int a = 0; a < 2; ++a) {
      const double hi = (((double)g->clamp_hi + 1.0) * dims[a] - 1.0) * 0.5;
      VFA_REQUIRE(floor(hi) + 1.0 <= dims[a] - 1, VFA_ERR_UNSUPPORTED,
                  "scale %d: crange[1]=%g on a %d-texel axis puts a bilinear tap outside the map; the reference then "
                  "reads zero padding (vfa_op.py:112-115) which the direct pooling form does not reproduce",
                  s, g->clamp_hi, dims[a]);
    }
  }
  return VFA_OK;
}

static int fill_params(const vfa_geometry_t* g, const vfa_shape_t* sh, const float* d_boxes, const float* const* d_feats,
                       const float* const* d_bias, AggParams& p) {
  p.B = sh->batch;
  p.V = sh->n_views;
  p.C = sh->channels;
  p.nl = g->n_layers;
  p.S = sh->n_scales;
  p.L = g->grid_l;
  p.W = g->grid_w;
  p.LW = g->grid_l * g->grid_w;
  p.K = sh->channels * g->n_layers;
  p.boxes = d_boxes;
  p.y_bf16 = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    p.feats[s] = nullptr;
    p.wprep[s] = nullptr;
    p.bias[s] = nullptr;
    p.sc[s] = make_scale_const(2, 2);
  }
  for (int s = 0; s < sh->n_scales; ++s) {
    VFA_REQUIRE(d_feats[s] != nullptr && (d_bias == nullptr || d_bias[s] != nullptr), VFA_ERR_INVALID_ARGUMENT,
                "scale %d: NULL tensor", s);
    VFA_REQUIRE((reinterpret_cast<uintptr_t>(d_feats[s]) & 15) == 0, VFA_ERR_INVALID_ARGUMENT,
                "scale %d: feature pointer not 16-byte aligned", s);
    p.feats[s] = d_feats[s];
    p.bias[s] = d_bias ? d_bias[s] : nullptr;
    p.sc[s] = make_scale_const(sh->feat_h[s], sh->feat_w[s]);
  }
  return VFA_OK;
}

}  // namespace vfa

using namespace vfa;

extern "C" {

int vfa_version(void) { return VFA_ABI_VERSION; }
void vfa_reload_env(void) {
  (void)runtime_config();
  load_config();
}
const char* vfa_last_error(void) { return g_error; }
const char* vfa_last_path(void) { return g_path; }

int vfa_table_build(const vfa_geometry_t* geom, int32_t n_views, const float* d_calibs, const float* d_grid,
                    float* d_boxes, void* stream) {
  if (int rc = validate_geometry(geom)) return rc;
  VFA_REQUIRE(n_views >= 1, VFA_ERR_INVALID_ARGUMENT, "n_views=%d", n_views);
  VFA_REQUIRE(d_calibs && d_grid && d_boxes, VFA_ERR_INVALID_ARGUMENT, "NULL device pointer");
  VFA_REQUIRE((reinterpret_cast<uintptr_t>(d_boxes) & 15) == 0, VFA_ERR_INVALID_ARGUMENT, "d_boxes not 16-byte aligned");
  if (int rc = check_device()) return rc;
  return launch_table_build(geom, n_views, d_calibs, d_grid, d_boxes, (cudaStream_t)stream);
}

int vfa_table_scale(const float* d_boxes, int64_t n_boxes, int32_t feat_h, int32_t feat_w, float* d_area,
                    uint8_t* d_visible, int32_t* d_taps, void* stream) {
  VFA_REQUIRE(d_boxes != nullptr && n_boxes >= 0, VFA_ERR_INVALID_ARGUMENT, "bad boxes");
  VFA_REQUIRE(feat_h >= 1 && feat_w >= 1, VFA_ERR_INVALID_ARGUMENT, "bad feature size %dx%d", feat_h, feat_w);
  if (n_boxes == 0) return VFA_OK;
  if (int rc = check_device()) return rc;
  return launch_table_scale(d_boxes, n_boxes, feat_h, feat_w, d_area, d_visible, d_taps, (cudaStream_t)stream);
}

int vfa_nchw_to_nhwc(const float* d_src, float* d_dst, int64_t n, int32_t channels, int64_t hw, void* stream) {
  VFA_REQUIRE(d_src && d_dst && n >= 1 && channels >= 1 && hw >= 1, VFA_ERR_INVALID_ARGUMENT, "bad transpose arguments");
  if (int rc = check_device()) return rc;
  return launch_transpose(d_src, d_dst, n, channels, hw, (cudaStream_t)stream);
}

int vfa_nhwc_to_nchw(const float* d_src, float* d_dst, int64_t n, int32_t channels, int64_t hw, void* stream) {
  VFA_REQUIRE(d_src && d_dst && n >= 1 && channels >= 1 && hw >= 1, VFA_ERR_INVALID_ARGUMENT, "bad transpose arguments");
  VFA_REQUIRE(hw <= 0x7fffffff, VFA_ERR_UNSUPPORTED, "hw too large");
  if (int rc = check_device()) return rc;
  return launch_transpose(d_src, d_dst, n, (int)hw, channels, (cudaStream_t)stream);
}

size_t vfa_aggregate_workspace_bytes(const vfa_geometry_t* geom, const vfa_shape_t* shape, uint32_t flags) {
  if (!geom || !shape) return 0;
  // forward: the one kernel family `flags` selects (the same dispatch as vfa_aggregate_fwd)
  const bool use_umma = !(flags & VFA_FLAG_FORCE_SIMT) && umma_supported(geom, shape, flags);
  const size_t fwd = use_umma ? umma_workspace_bytes(geom, shape, flags) : simt_workspace_bytes(geom, shape);
  const size_t bwd = bwd_workspace_bytes(geom, shape);
  size_t m;
  if (flags & VFA_FLAG_WS_FORWARD) m = fwd;
  else if (flags & VFA_FLAG_WS_BACKWARD) m = bwd;
  else m = fwd > bwd ? fwd : bwd;                    // one buffer for both directions
  return (m + 255) & ~(size_t)255;
}

int vfa_aggregate_fwd(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* d_boxes,
                      const float* const* d_feats, const float* const* d_weight, const float* const* d_bias,
                      float* d_out, uint32_t* d_relu_mask, void* d_workspace, size_t workspace_bytes, uint32_t flags,
                      void* stream) {
  if (int rc = validate_geometry(geom)) return rc;
  if (int rc = validate_shape(geom, shape)) return rc;
  VFA_REQUIRE(d_boxes && d_feats && d_weight && d_bias && d_out, VFA_ERR_INVALID_ARGUMENT, "NULL device pointer");
  VFA_REQUIRE((reinterpret_cast<uintptr_t>(d_boxes) & 15) == 0, VFA_ERR_INVALID_ARGUMENT, "d_boxes not 16-byte aligned");
  VFA_REQUIRE(!((flags & VFA_FLAG_FORCE_SIMT) && (flags & VFA_FLAG_FORCE_UMMA)), VFA_ERR_INVALID_ARGUMENT,
              "FORCE_SIMT and FORCE_UMMA are exclusive");
  for (int s = 0; s < shape->n_scales; ++s)
    VFA_REQUIRE(d_weight[s] != nullptr, VFA_ERR_INVALID_ARGUMENT, "scale %d: NULL weight", s);
  if (int rc = check_device()) return rc;
  AggParams p;
  if (int rc = fill_params(geom, shape, d_boxes, d_feats, d_bias, p)) return rc;
  p.out = d_out;
  p.mask = d_relu_mask;
  VFA_REQUIRE(d_workspace != nullptr && (reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, VFA_ERR_WORKSPACE,
              "workspace must be a 256-byte aligned device pointer");
  const bool use_umma = !(flags & VFA_FLAG_FORCE_SIMT) && umma_supported(geom, shape, flags);
  VFA_REQUIRE(!(flags & VFA_FLAG_BF16_FEATURES) || (use_umma && d_relu_mask == nullptr), VFA_ERR_UNSUPPORTED,
              "bf16 feature maps are supported by the tcgen05 forward (C = 256) only, without the backward mask");
  VFA_REQUIRE(!(flags & VFA_FLAG_OUT_NHWC) || (use_umma && !(flags & VFA_FLAG_GRID_SIDE)), VFA_ERR_UNSUPPORTED,
              "VFA_FLAG_OUT_NHWC is implemented by the feature-side forward (C = 256, without VFA_FLAG_GRID_SIDE) only");
  VFA_REQUIRE(!(flags & (VFA_FLAG_OUT_ACCUMULATE | VFA_FLAG_OUT_MULTICAST | VFA_FLAG_OUT_PEERS)) || (flags & VFA_FLAG_OUT_NHWC),
              VFA_ERR_UNSUPPORTED, "VFA_FLAG_OUT_ACCUMULATE / _MULTICAST / _PEERS need VFA_FLAG_OUT_NHWC (16-byte channel vectors)");
  VFA_REQUIRE(!(flags & VFA_FLAG_BF16_MMA) || (use_umma && !(flags & VFA_FLAG_GRID_SIDE) && d_relu_mask == nullptr),
              VFA_ERR_UNSUPPORTED, "VFA_FLAG_BF16_MMA is a forward-only variant of the feature-side path (C = 256, without "
              "VFA_FLAG_GRID_SIDE, no ReLU mask)");
  if ((flags & VFA_FLAG_FORCE_UMMA) && !use_umma) {
    set_error("tcgen05 path requested but unsupported for channels=%d layers=%d", shape->channels, geom->n_layers);
    return VFA_ERR_UNSUPPORTED;
  }
  if (use_umma) {
    VFA_REQUIRE(workspace_bytes >= umma_workspace_bytes(geom, shape, flags), VFA_ERR_WORKSPACE,
                "workspace %zu < required %zu", workspace_bytes, umma_workspace_bytes(geom, shape, flags));
    return launch_fwd_umma(p, d_weight, d_workspace, workspace_bytes, flags, (cudaStream_t)stream);
  }
  VFA_REQUIRE(workspace_bytes >= simt_workspace_bytes(geom, shape), VFA_ERR_WORKSPACE, "workspace %zu < required %zu",
              workspace_bytes, simt_workspace_bytes(geom, shape));
  return launch_fwd_simt(p, d_weight, d_workspace, flags, (cudaStream_t)stream);
}

int vfa_prepare_weights(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* const* d_weight,
                        void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (int rc = validate_geometry(geom)) return rc;
  if (int rc = validate_shape(geom, shape)) return rc;
  VFA_REQUIRE(d_weight != nullptr, VFA_ERR_INVALID_ARGUMENT, "NULL weight array");
  for (int s = 0; s < shape->n_scales; ++s)
    VFA_REQUIRE(d_weight[s] != nullptr, VFA_ERR_INVALID_ARGUMENT, "scale %d: NULL weight", s);
  VFA_REQUIRE(d_workspace != nullptr && (reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, VFA_ERR_WORKSPACE,
              "workspace must be a 256-byte aligned device pointer");
  if (int rc = check_device()) return rc;
  AggParams p;
  p.C = shape->channels;
  p.nl = geom->n_layers;
  p.S = shape->n_scales;
  const bool use_umma = !(flags & VFA_FLAG_FORCE_SIMT) && umma_supported(geom, shape, flags);
  const size_t need = use_umma ? umma_workspace_bytes(geom, shape, flags) : simt_workspace_bytes(geom, shape);
  VFA_REQUIRE(workspace_bytes >= need, VFA_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  return use_umma ? prep_weights_umma(p, d_weight, d_workspace, flags, (cudaStream_t)stream)
                  : prep_weights_simt(p, d_weight, d_workspace, (cudaStream_t)stream);
}

int vfa_multicast_copy(const void* d_src, void* d_mc_dst, size_t n_bytes, void* stream) {
  VFA_REQUIRE(d_src != nullptr && d_mc_dst != nullptr, VFA_ERR_INVALID_ARGUMENT, "NULL pointer");
  VFA_REQUIRE((n_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(d_src) | reinterpret_cast<uintptr_t>(d_mc_dst)) & 15) == 0,
              VFA_ERR_INVALID_ARGUMENT, "multicast copy needs 16-byte aligned pointers and size");
  if (n_bytes == 0) return VFA_OK;
  if (int rc = check_device()) return rc;
  return launch_multicast_copy(d_src, d_mc_dst, n_bytes, (cudaStream_t)stream);
}

size_t vfa_decode_workspace_bytes(int32_t batch) { return batch > 0 ? decode_workspace_bytes(batch) : 0; }

int vfa_decode_topk(const vfa_decode_t* dec, float* d_out_vals, int32_t* d_out_cell, void* d_workspace,
                    size_t workspace_bytes, void* stream) {
  VFA_REQUIRE(dec != nullptr && d_out_vals != nullptr && d_out_cell != nullptr, VFA_ERR_INVALID_ARGUMENT, "NULL argument");
  VFA_REQUIRE(dec->batch >= 1 && dec->grid_l >= 1 && dec->grid_w >= 1, VFA_ERR_INVALID_ARGUMENT, "empty heatmap %d x %d x %d",
              dec->batch, dec->grid_l, dec->grid_w);
  VFA_REQUIRE(dec->topk >= 1 && dec->topk <= 1024, VFA_ERR_INVALID_ARGUMENT, "topk=%d outside [1, 1024]", dec->topk);
  VFA_REQUIRE(dec->heatmap != nullptr && dec->loc_offset != nullptr, VFA_ERR_INVALID_ARGUMENT, "heatmap / loc_offset is NULL");
  VFA_REQUIRE(dec->rotation == nullptr || dec->n_angles >= 1, VFA_ERR_INVALID_ARGUMENT, "n_angles=%d", dec->n_angles);
  VFA_REQUIRE((long long)dec->grid_l * dec->grid_w < 0x7fffffffll, VFA_ERR_UNSUPPORTED, "grid too large");
  VFA_REQUIRE(d_workspace != nullptr && workspace_bytes >= decode_workspace_bytes(dec->batch), VFA_ERR_WORKSPACE,
              "decode workspace %zu < required %zu", workspace_bytes, decode_workspace_bytes(dec->batch));
  if (int rc = check_device()) return rc;
  return launch_decode(dec, d_out_vals, d_out_cell, d_workspace, (cudaStream_t)stream);
}

int vfa_aggregate_bwd(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* d_boxes,
                      const float* const* d_feats, const float* const* d_weight, const uint32_t* d_relu_mask,
                      const float* d_grad_out, float* const* d_grad_feats, float* const* d_grad_weight,
                      float* const* d_grad_bias, void* d_workspace, size_t workspace_bytes, uint32_t flags,
                      void* stream) {
  if (int rc = validate_geometry(geom)) return rc;
  if (int rc = validate_shape(geom, shape)) return rc;
  VFA_REQUIRE(d_boxes && d_feats && d_weight && d_relu_mask && d_grad_out, VFA_ERR_INVALID_ARGUMENT,
              "NULL device pointer");
  VFA_REQUIRE(d_grad_feats && d_grad_weight && d_grad_bias, VFA_ERR_INVALID_ARGUMENT, "NULL gradient pointer array");
  for (int s = 0; s < shape->n_scales; ++s)
    VFA_REQUIRE(d_weight[s] != nullptr, VFA_ERR_INVALID_ARGUMENT, "scale %d: NULL weight", s);
  if (int rc = check_device()) return rc;
  AggParams p;
  if (int rc = fill_params(geom, shape, d_boxes, d_feats, nullptr, p)) return rc;
  p.out = nullptr;
  p.mask = const_cast<uint32_t*>(d_relu_mask);
  VFA_REQUIRE(d_workspace != nullptr && (reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, VFA_ERR_WORKSPACE,
              "workspace must be a 256-byte aligned device pointer");
  VFA_REQUIRE(workspace_bytes >= bwd_workspace_bytes(geom, shape), VFA_ERR_WORKSPACE, "workspace %zu < required %zu",
              workspace_bytes, bwd_workspace_bytes(geom, shape));
  return launch_bwd(p, d_weight, d_grad_out, d_grad_feats, d_grad_weight, d_grad_bias, d_workspace, workspace_bytes, flags,
                    (cudaStream_t)stream);
}

}  // extern "C"
