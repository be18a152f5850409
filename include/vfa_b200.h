/* vfa_b200.h -- C ABI of the B200-native voxelized-feature-aggregation library (libvfa_b200.so).
 *
 * The reference (Jiahao-Ma/VFA) has no FFI for this path: the boundary is a Python nn.Module,
 * `VFA.forward(feature, calib, grid)` (reference vfa/model/vfa_op.py:61-125) called in the per-view / per-scale
 * loop of `VFANet.forward` (reference vfa/model/vfanet.py:64-82).  These entry points are what a binding for
 * that path binds (INTEGRATION.md shows the ctypes stub); each one cites the reference lines it replaces.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; the library never allocates, frees or
 *     synchronises the device.  Process-wide state, all of it: the thread-local last-error / last-path strings; the
 *     debug switches read once from the environment (vfa_reload_env); per-device caches filled under a mutex (resident
 *     cluster counts of the persistent kernels, SM count, and -- for the generic C < 256 backward only -- one lazily
 *     created cuBLAS handle per device).  Any number of devices may be driven from one process;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued on it;
 *   - every function returns VFA_OK (0) or a negative vfa_status_t; vfa_last_error() describes the failure;
 *   - tensors are dense, row-major in the index order written in the comment.
 */
#ifndef VFA_B200_H_
#define VFA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFA_ABI_VERSION 1
#define VFA_MAX_LAYERS 16
#define VFA_MAX_SCALES 3

typedef enum {
  VFA_OK = 0,
  VFA_ERR_INVALID_ARGUMENT = -1,   /* null pointer, non-positive size, unknown enum value             */
  VFA_ERR_UNSUPPORTED = -2,        /* valid request outside what the kernels implement (see message)   */
  VFA_ERR_WORKSPACE = -3,          /* workspace too small / misaligned                                 */
  VFA_ERR_CUDA = -4,               /* a CUDA runtime call or kernel launch failed                      */
  VFA_ERR_NO_DEVICE = -5           /* no sm_100 device: this library has no CPU or other-arch fallback */
} vfa_status_t;

/* grid -> world conversion of the reference's per-dataset `convert` (reference vfa/model/vfa_op.py:23-44).
 * The operation kind matters for bit-exactness: MultiviewX divides by 40, it does not multiply by 0.025. */
typedef enum {
  VFA_CONVERT_DIV = 0,     /* world = p / scale                       MultiviewC (1.0), MultiviewX (40.0)       */
  VFA_CONVERT_AFFINE = 1   /* world.xy = p.xy*scale - offset.xy ; world.z = p.z*scale      Wildtrack (2.5)   */
} vfa_convert_t;

/* Static description of the voxel grid and the camera normalisation (reference VFA.__init__,
 * vfa_op.py:47-59, plus the `args` fields it reads, vfa_op.py:38-43, :75). */
typedef struct {
  int32_t n_layers;                  /* nl = len(arange(0, grid_height, cube_h))                     */
  int32_t grid_l, grid_w;            /* BEV cells: grid tensor is [L, W, 3]                          */
  int32_t convert_kind;              /* vfa_convert_t                                                */
  float convert_scale;
  float convert_offset[3];
  float cube[3];                     /* voxel (l, w, h) in grid units -> the 8 corner offsets        */
  float layer_z[VFA_MAX_LAYERS];     /* z_n = n * cube_h as fp32 (buffer `z_corners`)                */
  float image_w, image_h;            /* args.image_size[::-1]                                        */
  float clamp_lo, clamp_hi;          /* `crange`, default (-1, 0.95)                                 */
} vfa_geometry_t;

/* One frame batch of multi-view, multi-scale features, channels-last. */
typedef struct {
  int32_t batch;                     /* B frames                                                     */
  int32_t n_views;                   /* V cameras                                                    */
  int32_t channels;                  /* C (input channels == output channels, reference vfanet.py:30) */
  int32_t n_scales;                  /* 1..VFA_MAX_SCALES                                            */
  int32_t feat_h[VFA_MAX_SCALES];
  int32_t feat_w[VFA_MAX_SCALES];
} vfa_shape_t;

/* flags of vfa_aggregate_fwd / _bwd */
#define VFA_FLAG_FORCE_SIMT   1u     /* never take the tcgen05 path (generic fp32 FFMA kernel)         */
#define VFA_FLAG_FORCE_UMMA   2u     /* fail with VFA_ERR_UNSUPPORTED instead of falling to SIMT       */
#define VFA_FLAG_BF16_MMA     4u     /* reserved: single-pass bf16 collapse; returns VFA_ERR_UNSUPPORTED */
#define VFA_FLAG_WEIGHTS_PREPARED 8u /* workspace already holds vfa_prepare_weights output for these flags */
#define VFA_FLAG_BF16_FEATURES 16u  /* d_feats point to bf16 [B,V,fH,fW,C] maps (forward, C = 256 only): half the
                                        gather bytes; pooling / collapse arithmetic is unchanged (fp32, 3xTF32)    */
#define VFA_FLAG_GRID_SIDE    32u    /* C = 256: use the fully fused grid-side kernel (pool, then contract; no
                                        intermediate at all) instead of the default feature-side pair of kernels
                                        (contract on the image plane, then pool from per-quad texel lists; Y of a
                                        frame chunk and the lists live in the workspace)                            */

#define VFA_FLAG_TABLE_PREPARED 64u  /* C = 256 forward: the workspace still holds what the previous vfa_aggregate_fwd
                                        call derived from the SAME boxes, shapes and flags (tap records, coverage bitmap,
                                        row lists, the quads' texel lists) -- static cameras: skip rebuilding them      */

#define VFA_FLAG_OUT_NHWC 128u        /* C = 256 feature-side forward: d_out is [B, L, W, C] (channels-last, what a cuDNN
                                        head in torch.channels_last reads without a permute; reference consumer:
                                        vfanet.py:131-139) instead of [B, C, L, W]; the backward reads d_grad_out in the
                                        layout it is handed with the same flag                                          */

#define VFA_FLAG_OUT_ACCUMULATE 256u  /* with VFA_FLAG_OUT_NHWC: the result is ADDED into d_out (red.global.add.v4.f32, system
                                        scope) instead of stored; the caller zero-initialises d_out.  d_out may be a PEER
                                        GPU's memory mapped over NVLink: camera-sharded ranks reduce their partial maps
                                        inside the pooling kernel, tile by tile, instead of calling a collective afterwards
                                        (the sum over cameras of reference vfanet.py:82 is taken across GPUs)             */
#define VFA_FLAG_OUT_MULTICAST 512u   /* as OUT_ACCUMULATE, but d_out is the MULTICAST address of a symmetric allocation
                                        (cuMulticast / torch symmetric memory): multimem.red.add -- the NVSwitch adds the
                                        tile into every GPU's replica (in-switch all-reduce fused into the kernel)         */

#define VFA_FLAG_OUT_PEERS 4096u      /* as OUT_ACCUMULATE, but d_out points to a vfa_peer_outputs_t in DEVICE memory: the BEV rows
                                        are split into bands, one per rank, and every finished tile is red.add-ed into the
                                        replica of the rank that owns its band (peer memory over NVLink): a reduce-scatter
                                        fused into the pooling kernel.  vfa_multicast_copy then broadcasts each rank's band  */
#define VFA_FLAG_WS_FORWARD 1024u     /* vfa_aggregate_workspace_bytes only: size for vfa_aggregate_fwd alone (the kernel
                                        family the other flags select: e.g. 90 MB with VFA_FLAG_GRID_SIDE)               */
#define VFA_FLAG_WS_BACKWARD 2048u    /* vfa_aggregate_workspace_bytes only: size for vfa_aggregate_bwd alone            */

int vfa_version(void);
const char* vfa_last_error(void);

/* The library reads its debug / A-B switches (VFA_POOL_TILE, VFA_FSIDE_Y_BUDGET_MB, ...: csrc/vfa_common.cuh,
 * RuntimeConfig) from the environment once, at the first call.  This re-reads them (tests and timing scripts). */
void vfa_reload_env(void);

/* Human-readable name of the kernel family the last vfa_aggregate_fwd on this thread dispatched to. */
const char* vfa_last_path(void);

/* Projection table: clamped normalised bounding box of every voxel's 8 projected corners, bit-identical to
 * the reference's `box_corners` (reference vfa_op.py:64-88 + vfa/utils.py:50-59).
 *   d_calibs [V,3,4] fp32, d_grid [L,W,3] fp32  ->  d_boxes [V, nl, L*W, 4] fp32 (left, top, right, bottom). */
int vfa_table_build(const vfa_geometry_t* geom, int32_t n_views, const float* d_calibs, const float* d_grid,
                    float* d_boxes, void* stream);

/* Per-scale derived table for parity checking, produced by the same device function the aggregation kernels
 * call: area (reference vfa_op.py:104-105), visible (:106) and the fp32 sampling tap indices
 * floor(((c+1)*S-1)/2) of F.grid_sample (:112-115).  Any output pointer may be NULL.
 *   d_boxes [n_boxes,4] -> d_area [n_boxes] fp32, d_visible [n_boxes] u8, d_taps [n_boxes,4] i32 (xl,yt,xr,yb) */
int vfa_table_scale(const float* d_boxes, int64_t n_boxes, int32_t feat_h, int32_t feat_w, float* d_area,
                    uint8_t* d_visible, int32_t* d_taps, void* stream);

/* Layout helpers: [n, C, HW] <-> [n, HW, C] fp32 transposes (the reference hands NCHW feature maps,
 * reference vfanet.py:72-78; the gather kernels read channels-last). */
int vfa_nchw_to_nhwc(const float* d_src, float* d_dst, int64_t n, int32_t channels, int64_t hw, void* stream);
int vfa_nhwc_to_nchw(const float* d_src, float* d_dst, int64_t n, int32_t channels, int64_t hw, void* stream);

/* Bytes of scratch for this problem and these flags.  Without a VFA_FLAG_WS_* flag: enough for vfa_aggregate_fwd with
 * `flags` AND vfa_aggregate_bwd (one buffer serving both directions); VFA_FLAG_WS_FORWARD / _BACKWARD size one direction.
 * Contents: prepared weights, tap records, coverage bitmap and
 * row lists, the quads' texel lists, Y of one frame chunk (forward); CSR, masked gradients, Gs (backward).  Pure host
 * arithmetic on (geom, shape, flags). */
size_t vfa_aggregate_workspace_bytes(const vfa_geometry_t* geom, const vfa_shape_t* shape, uint32_t flags);

/* Re-lays the collapse weights for the kernel family `flags` selects (column permutation c*nl+n -> n*C+c,
 * operand splitting / swizzling for the tensor-core path) into the workspace.  vfa_aggregate_fwd does this
 * itself unless VFA_FLAG_WEIGHTS_PREPARED is set; inference callers with frozen weights call it once. */
int vfa_prepare_weights(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* const* d_weight,
                        void* d_workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Fused aggregation forward = the whole loop of reference vfanet.py:64-82 minus the lateral convs:
 *   out[b] = sum_v sum_s relu( collapse_s( pooled voxels of view v at scale s ) )       (vfa_op.py:104-124)
 *   d_boxes   [V, nl, L*W, 4] fp32 from vfa_table_build
 *   d_feats[s] [B, V, fH_s, fW_s, C] fp32 channels-last
 *   d_weight[s] [C, C*nl] fp32 with the reference's column order c*nl + n, d_bias[s] [C]   (vfa_op.py:59, :120)
 *   d_out     [B, C, L, W] fp32 (fully overwritten)
 *   d_relu_mask  NULL, or [B, V, S, ceil(C/32), L*W] u32 (fully overwritten): bit (o % 32) of word o/32 is set iff
 *             output channel o of that (frame, view, scale, cell) passed the ReLU -- what the backward needs
 * With batch = n_views = n_scales = 1 this is exactly one reference `VFA.forward`. */
int vfa_aggregate_fwd(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* d_boxes,
                      const float* const* d_feats, const float* const* d_weight, const float* const* d_bias,
                      float* d_out, uint32_t* d_relu_mask, void* d_workspace, size_t workspace_bytes, uint32_t flags,
                      void* stream);

/* Backward of vfa_aggregate_fwd (what autograd derives for reference vfa_op.py:110-124): given d_grad_out
 * [B,C,L,W] ([B,L,W,C] with VFA_FLAG_OUT_NHWC, C = 256: read in place, no transpose) accumulates into d_grad_feats[s] [B,V,fH,fW,C] (must be zero-initialised by the caller),
 * d_grad_weight[s] [C, C*nl] and d_grad_bias[s] [C] (both fully overwritten).  Any gradient pointer array
 * entry may be NULL to skip it.  Needs the ReLU mask the forward wrote; recomputes the pooled voxels instead of
 * saving them (the reference keeps 0.6-1.8 GB of intermediates per call for autograd). */
int vfa_aggregate_bwd(const vfa_geometry_t* geom, const vfa_shape_t* shape, const float* d_boxes,
                      const float* const* d_feats, const float* const* d_weight, const uint32_t* d_relu_mask,
                      const float* d_grad_out, float* const* d_grad_feats, float* const* d_grad_weight,
                      float* const* d_grad_bias, void* d_workspace, size_t workspace_bytes, uint32_t flags,
                      void* stream);

/* Destination table of VFA_FLAG_OUT_PEERS (lives in device memory; all fields 64-bit): BEV row cy belongs to rank
 * min(cy / band_rows, n_ranks - 1); out[r] is rank r's [B, L, W, C] replica as mapped into THIS device's address space. */
typedef struct {
  uint64_t n_ranks;
  uint64_t band_rows;
  uint64_t out[16];
} vfa_peer_outputs_t;

/* d_mc_dst[i] = d_src[i] for n_bytes (a multiple of 16) where d_mc_dst is the MULTICAST address of a symmetric allocation:
 * multimem.st -- one read of local memory, the NVSwitch writes every GPU's replica (the all-gather half of the fused
 * all-reduce: each rank broadcasts the band of the BEV map it owns). */
int vfa_multicast_copy(const void* d_src, void* d_mc_dst, size_t n_bytes, void* stream);

/* ---- decode tail --------------------------------------------------------------------
 * The heads' maps -> at most `topk` detections per frame: sigmoid, 5 x 5 max-pool NMS, top-k by confidence, centre / size /
 * orientation decoded at the selected cells only.  Replaces reference vfa/data/encoder.py:230-273 (decode3d; pass
 * dim_offset = rotation = NULL for decode2d, :275-305) up to the final `conf > cls_thresh` mask, which the caller applies.
 * Head tensors are addressed as p[b * stride[0] + c * stride[1] + cell * stride[2]] (elements), cell = y * W + x, so
 * [B,C,L,W], [B,L,W,C] (the reference's permuted views) and channels-last storage all pass without a copy. */
typedef struct {
  int32_t batch, grid_l, grid_w;     /* heatmap is [B, 1, L, W] contiguous fp32 logits                               */
  int32_t topk;                      /* 1..1024 (reference default 50, train.py:126)                                 */
  int32_t n_angles;                  /* channels of `rotation` (reference: 360)                                      */
  const float* heatmap;
  const float* loc_offset;           /* 2 channels (ty, tx) logits                                                   */
  int64_t loc_stride[3];
  const float* dim_offset;           /* 3 channels (th, tw, tl), or NULL                                             */
  int64_t dim_stride[3];
  const float* rotation;             /* n_angles channels of logits, or NULL                                         */
  int64_t rot_stride[3];
  float grid_size[2], world_size[2]; /* encoder.grid_size, encoder.world_size (reference encoder.py:244-245)         */
  float dim_mean[3];                 /* classAverage mean (h, w, l) (encoder.py:247-250)                             */
} vfa_decode_t;

size_t vfa_decode_workspace_bytes(int32_t batch);

/* d_out_vals [B, topk, 7] fp32 = (conf, cy, cx, h, w, l, orientation bin), sorted by conf descending (equal conf: lower
 * cell first); d_out_cell [B, topk] int32 = cell index y * W + x, or -1 (and conf = 0) where a frame has fewer than topk
 * NMS survivors. */
int vfa_decode_topk(const vfa_decode_t* dec, float* d_out_vals, int32_t* d_out_cell, void* d_workspace,
                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VFA_B200_H_ */
