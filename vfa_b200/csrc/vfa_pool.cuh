// Helpers shared by the pooling kernels of the feature-side forward (vfa_fwd_fside.cu: walking / list kernels,
// vfa_pool_tile.cu: shared-memory staged tile kernel): channel <-> lane map, packed FMAs, ReLU-mask packing, arguments.
#pragma once

#include "vfa_common.cuh"

namespace vfa {

// Bitmap of the texels some visible box covers, per (scale, view, layer) plane: the row-compacted GEMM multiplies only
// those rows.  Filled by cover_mark_kernel (vfa_fwd_fside.cu) or, on its way, by tile_build_kernel (vfa_pool_tile.cu).
struct CoverMap {
  int word_base[VFA_MAX_SCALES];     // first 32-bit word of scale s; planes (v * nl + n) follow each other, `words` apart
  int words[VFA_MAX_SCALES];         // ceil(hw / 32)
  int hw[VFA_MAX_SCALES], fw[VFA_MAX_SCALES];
  int total_words;
};

inline CoverMap make_cover_map(const AggParams& p) {
  CoverMap cm;
  int base = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    const int ss = s < p.S ? s : 0;
    cm.hw[s] = p.sc[ss].fh * p.sc[ss].fw;
    cm.fw[s] = p.sc[ss].fw;
    cm.words[s] = (cm.hw[s] + 31) / 32;
    cm.word_base[s] = base;
    if (s < p.S) base += p.V * p.nl * cm.words[s];
  }
  cm.total_words = base;
  return cm;
}

namespace fside {

constexpr int CH = 256;

// kernel-parameter arrays indexed by a run-time scale: a select chain instead of a local-memory copy of the struct
template <typename T>
__device__ __forceinline__ T pick(const T (&arr)[VFA_MAX_SCALES], int s) {
  static_assert(VFA_MAX_SCALES == 3, "");
  return s == 0 ? arr[0] : (s == 1 ? arr[1] : arr[2]);
}

struct PoolArgs {
  AggParams p;
  const void* y[VFA_MAX_SCALES];   // [plane][layer][texel][256] fp32 (bf16 with VFA_FLAG_BF16_MMA)
  const TapRec* recs;        // [V][S][nl][LW]
  int b0;                    // first frame of this chunk (output / mask index); Y planes are chunk-relative
  int tiles_x;
  // texel lists of the quads (pool_list_kernel; nullptr = not built)
  const uint32_t* seg_off;   // [quads][V*S + 1] first entry of segment (view, scale) of a quad; last = end of the quad's list,
                             // or LIST_OVERFLOW when the quad's texels did not fit its slot
  const uint32_t* ent_off;   // [quads * slot] texel index inside the (view, scale) stack of nl planes
  const float4* ent_w;       // [quads * slot] weights of the four cells of the quad
  int quads_x;
  // staged-tile pooling (pool_tile_kernel): one byte per 8 x 8-cell tile, 1 = the tile's chunk lists overflowed their pools and
  // the walking kernel pools its quads (nullptr = lists of the quads decide, above)
  const uint8_t* tile_ovf;
  int ptiles_x;
  int out_nhwc;              // 0: out [B, C, L, W]; 1: out [B, L, W, C]
  int out_mode;              // 0: store; 1: red.add into `out` (own, or a peer GPU's memory over NVLink); 2: multimem.red.add
                             // on a multicast address (every GPU's replica receives the add); 3: red.add into the replica of the
                             // rank that OWNS the cell's band of BEV rows (p.out = vfa_peer_outputs_t in device memory).
                             // 1 - 3: [B, L, W, C] only
};

// mode 3: base pointer of the replica that owns BEV row `cy` (bands of `band_rows` rows, the last rank takes the rest)
__device__ __forceinline__ float* owner_base(const float* desc_ptr, int cy) {
  const unsigned long long* d = reinterpret_cast<const unsigned long long*>(desc_ptr);
  const int n = (int)__ldg(d), band = (int)__ldg(d + 1);
  const int owner = min(cy / band, n - 1);
  return reinterpret_cast<float*>(__ldg(d + 2 + owner));
}
constexpr uint32_t LIST_OVERFLOW = 0xffffffffu;

__device__ __forceinline__ void fma8(float (&acc)[8], float w, const float4& a, const float4& b) {
  const float2 w2 = make_float2(w, w);
  const float2 r0 = __ffma2_rn(w2, make_float2(a.x, a.y), make_float2(acc[0], acc[1]));
  const float2 r1 = __ffma2_rn(w2, make_float2(a.z, a.w), make_float2(acc[2], acc[3]));
  const float2 r2 = __ffma2_rn(w2, make_float2(b.x, b.y), make_float2(acc[4], acc[5]));
  const float2 r3 = __ffma2_rn(w2, make_float2(b.z, b.w), make_float2(acc[6], acc[7]));
  acc[0] = r0.x; acc[1] = r0.y; acc[2] = r1.x; acc[3] = r1.y;
  acc[4] = r2.x; acc[5] = r2.y; acc[6] = r3.x; acc[7] = r3.y;
}

// Channel <-> lane map of the pooling kernels: lane l owns channels [4l, 4l+4) (acc[0..3]) and [128 + 4l, 128 + 4l + 4)
// (acc[4..7]), so each of the two LDG.128 of a warp covers 512 contiguous bytes of a texel row = 4 L1 wavefronts (with 8
// consecutive channels per lane the 32-byte lane stride touches every 128-byte line twice: 8 wavefronts per load).
__device__ __forceinline__ int chan_of(int lane, int i) { return (i < 4 ? 0 : CH / 2 - 4) + lane * 4 + i; }

// ReLU pass bits of one cell (bit i of `bits` = channel chan_of(lane, i)) -> mask words (word o/32, bit o%32) at
// words[k * word_stride]: lanes 8k .. 8k+7 hold the eight nibbles of word k (low channels) and of word 4 + k (high)
__device__ __forceinline__ void store_mask_words(uint32_t* words, size_t word_stride, int lane, uint32_t bits, bool valid) {
  uint32_t lo = (bits & 0xfu) << (4 * (lane & 7)), hi = (bits >> 4) << (4 * (lane & 7));
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    lo |= __shfl_xor_sync(0xffffffffu, lo, d);
    hi |= __shfl_xor_sync(0xffffffffu, hi, d);
  }
  if ((lane & 7) == 0 && valid) {
    words[(size_t)(lane >> 3) * word_stride] = lo;
    words[(size_t)(4 + (lane >> 3)) * word_stride] = hi;
  }
}

// 16-byte reduction into global memory: plain (mode 1) or through the NVLink multicast object (mode 2)
__device__ __forceinline__ void red_add_v4(float* dst, const float4& v, int mode) {      // mode 1 / 3: plain red
  if (mode == 2)
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  else
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// output element (frame b, channel c, cell) in either layout
__device__ __forceinline__ size_t out_index(int nhwc, int b, int c, int cell, int LW) {
  return nhwc ? ((size_t)b * LW + cell) * CH + c : ((size_t)b * CH + c) * LW + cell;
}

}  // namespace fside
}  // namespace vfa
