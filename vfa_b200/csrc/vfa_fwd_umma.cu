// Fused aggregation forward on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
// accumulator in TMEM, 3xTF32 operand splitting for fp32-level accuracy.   C = 256 channels.
//
// CTAs run as pairs (cluster of 2, tcgen05 cta_group::2, M = 256 cells across the pair).  One CTA (896 threads,
// 1 per SM) owns an 8 x 16 block of BEV cells of one frame (= its 128 TMEM lanes) and a group of views.  For every
// (view, scale) it runs a K loop over (height layer n, 32-channel chunk), 3 shared-memory stages of 64 KB:
//
//   warp 0       weight loader   per stage two cp.async.bulk (UBLKCP) of this CTA's half (128 rows) of the pre-split,
//                                pre-swizzled hi and lo slabs of collapse.weight; the MMA reads the other half from
//                                the peer CTA's shared memory
//   warp 1       leader: MMA issuer, one thread, per stage 4 k-steps x {A_lo*B_hi, A_hi*B_lo, A_hi*B_hi}
//                                tcgen05.mma.cta_group::2.kind::tf32, M=256 N=256 K=8, D in TMEM cols [0,256);
//                                tcgen05.commit (multicast) frees the stage / publishes the layer in both CTAs
//                follower: relays "my stage is full" to the leader with a remote mbarrier arrive
//   warps 4-19   pool producers  build per-(cell, layer) gather recipes in shared memory (shared device function with
//                                the parity-checked table kernel), gather 4 channels per thread with predicated
//                                128-bit loads from the channels-last map (8 lanes = one 128 B row of a texel;
//                                4 adjacent cells per warp instruction), split into tf32 hi/lo and write the
//                                K-major SWIZZLE_128B operand tile
//   warps 20-27  epilogue        after every height layer: tcgen05.ld the accumulator and fold it into an fp32
//                                running sum in TMEM cols [256,512) (the tensor-core accumulator truncates, so
//                                only 32 k-steps are chained in it); after the last layer + bias, ReLU, add into
//                                [B,C,L,W] (read-modify-write; atomics when views are split over CTAs)
//
// mbarriers: full[stage] (16 producer warps + loader tx), empty[stage] (tcgen05.commit), peer_full[stage] (relay),
// acc_full (commit), acc_empty (epilogue warps of both CTAs).  The [L*W, C*nl] matrix of the reference
// (vfa_op.py:118-120) only ever exists as 16 KB operand tiles; the collapse (vfa_op.py:123) is the tensor-core
// contraction; ReLU and the sums over scales and views (vfa_op.py:124, vfanet.py:79-82) are the epilogue.
// Debug switches (environment, read at launch): VFA_UMMA_VARIANT knock-out bits, see UmmaArgs::variant.
#include <stdlib.h>

#include "vfa_common.cuh"
#include "vfa_umma_ptx.cuh"

namespace vfa {

namespace umma {

constexpr int CH = 256;                 // channels == MMA N
constexpr int TILE_M = 128;             // cells per CTA == TMEM lanes
// A tile is a compact TILE_H x TILE_W block of the BEV grid, not 128 consecutive cells: neighbouring cells in both
// directions project to overlapping texel windows, so the texels one stage touches shrink ~3x (L1 / L2 reuse).
constexpr int TILE_W = 16;
constexpr int TILE_H = TILE_M / TILE_W;
constexpr int KCH = 32;                 // K elements per stage (one 128-byte swizzle row of tf32)
#ifndef VFA_CTA_PAIR
#define VFA_CTA_PAIR 1
#endif
// PAIR: two CTAs of a cluster (one TPC) run tcgen05.mma.cta_group::2 -- M = 256 cells across the pair, each CTA stages
// its own pooled rows and only HALF of the weight slab (the tensor cores fetch the other half from the peer), which
// halves the weight traffic from L2 and the operand reads from shared memory, and makes room for a third stage.
constexpr bool PAIR = VFA_CTA_PAIR != 0;
#ifndef VFA_STAGES
#define VFA_STAGES (VFA_CTA_PAIR ? 3 : 2)
#endif
constexpr int STAGES = VFA_STAGES;
constexpr int A_BYTES = TILE_M * KCH * 4;          // 16 KB per hi / lo tile
constexpr int B_BYTES = CH * KCH * 4;              // 32 KB per hi / lo slab (whole N = 256)
constexpr int B_LOCAL_BYTES = PAIR ? B_BYTES / 2 : B_BYTES;   // rows of the slab this CTA stages
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_LOCAL_BYTES;  // 64 KB (pair) / 96 KB
#ifndef VFA_PRODUCER_WARPS
#define VFA_PRODUCER_WARPS 16
#endif
constexpr int NUM_PRODUCER_WARPS = VFA_PRODUCER_WARPS;   // 8 (deep per-warp prefetch) or 16 (more warps, shallower)
constexpr int ROUNDS = TILE_M / (4 * NUM_PRODUCER_WARPS);  // items (4 cells each) per producer warp and K chunk
constexpr int ROWS_PER_WARP = 4 * ROUNDS;
constexpr int FIRST_PRODUCER_WARP = 4;     // warpgroup 0 = {loader, MMA, 2 idle warps} gives its registers away
constexpr int FIRST_EPILOGUE_WARP = FIRST_PRODUCER_WARP + NUM_PRODUCER_WARPS;
constexpr int NUM_EPILOGUE_WARPS = 8;
constexpr int THREADS = (FIRST_EPILOGUE_WARP + NUM_EPILOGUE_WARPS) * 32;
// register redistribution (setmaxnreg): {warpgroup 0, producers, epilogue}; 0 = leave at the launch value
constexpr int REGS_WG0 = NUM_PRODUCER_WARPS == 8 ? 32 : 24;
constexpr int REGS_PRODUCER = NUM_PRODUCER_WARPS == 8 ? 128 : 0;
constexpr int REGS_EPILOGUE = NUM_PRODUCER_WARPS == 8 ? 0 : 96;
constexpr int DEPTH_2X2 = NUM_PRODUCER_WARPS == 8 ? 4 : 2;
constexpr int DEPTH_3X3 = NUM_PRODUCER_WARPS == 8 ? 2 : 1;
constexpr bool WINDOW_DOUBLE_BUFFER = NUM_PRODUCER_WARPS == 8;
constexpr int TMEM_COLS = 512;

// Gather recipe of one (cell, layer) row at one scale, prepared once per (view, scale, layer) in shared memory:
// element offset of the first texel, tap counts and the separable edge weights (1/area and visibility folded
// into the row weights; interior taps weigh 1 resp. wy_mid), plus the nine products of a <= 3x3 box and the first 8 weights per axis.  144 bytes.
struct __align__(16) RowDesc {
  int base;        // ((y0 * fw) + x0) * CH
  int nx, ny;      // tap counts (0 = not visible)
  float wx_first, wx_last, wy_first, wy_last, wy_mid;
  float w9[12];    // wy(ty) * wx(tx) at [ty*3+tx], ty,tx in {0,1,2} (0 = tap not used); last 3 are padding
  float wx8[8];    // wx(0..7), wy(0..7) (0 beyond the box): window weights without the select chain for boxes up to
  float wy8[8];    // 8x8 taps; larger boxes fall back to desc_wx / desc_wy
};
__device__ __forceinline__ float desc_wx(const RowDesc& d, int i) {
  return i >= d.nx ? 0.f : (i == 0 ? d.wx_first : (i == d.nx - 1 ? d.wx_last : 1.0f));
}
__device__ __forceinline__ float desc_wy(const RowDesc& d, int i) {
  return i >= d.ny ? 0.f : (i == 0 ? d.wy_first : (i == d.ny - 1 ? d.wy_last : d.wy_mid));
}

struct __align__(16) SmemTail {
  RowDesc desc[NUM_PRODUCER_WARPS][ROWS_PER_WARP];
  float bias[VFA_MAX_SCALES][CH];
  unsigned long long full[STAGES];       // 8 producer warps + weight loader (tx) of THIS CTA
  unsigned long long empty[STAGES];      // tcgen05.commit (multicast to both CTAs of a pair)
  unsigned long long peer_full[STAGES];  // pair leader only: the peer's relay thread reports the peer's full[st]
  unsigned long long acc_full;
  unsigned long long acc_empty;          // epilogue warps (of both CTAs of a pair, on the leader's barrier)
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + sizeof(SmemTail);

constexpr uint32_t IDESC = make_idesc_tf32(CH, PAIR ? 2 * TILE_M : TILE_M);
__device__ __forceinline__ void tc_commit(void* bar) { tc_commit_t<PAIR>(bar); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  tc_mma_tf32_t<PAIR>(tmem_d, desc_a, desc_b, idesc, accumulate);
}

}  // namespace umma

using namespace umma;

// collapse.weight [C, C*nl] (column c*nl+n)  ->  per scale, per K chunk kc = n*(C/32) + c/32 a 64 KB block
// [hi: 256 rows x 128 B swizzled][lo: same], hi = tf32(w), lo = tf32(w - hi).
struct PrepWeightArgs {
  const float* w[VFA_MAX_SCALES];
  uint8_t* wp[VFA_MAX_SCALES];
};
// grid.y = scale: one launch re-lays the collapse weights of every scale
__global__ void __launch_bounds__(256) prep_weight_umma_kernel(const PrepWeightArgs a, int nl) {
  const float* __restrict__ w = blockIdx.y == 0 ? a.w[0] : (blockIdx.y == 1 ? a.w[1] : a.w[2]);
  uint8_t* __restrict__ wp = blockIdx.y == 0 ? a.wp[0] : (blockIdx.y == 1 ? a.wp[1] : a.wp[2]);
  const int K = CH * nl;
  const long long total = (long long)CH * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % KCH);
    const int o = (int)((idx / KCH) % CH);
    const int kc = (int)(idx / ((long long)KCH * CH));
    const int n = kc / (CH / KCH), cc = kc % (CH / KCH);
    const int c = cc * KCH + kk;
    const float v = w[(long long)o * K + (long long)c * nl + n];
    const uint32_t hi = to_tf32(v);
    const uint32_t lo = to_tf32(v - __uint_as_float(hi));
    uint8_t* blk = wp + (long long)kc * (2 * B_BYTES);
    const uint32_t off = swz(o, kk >> 2) + (kk & 3) * 4;
    *reinterpret_cast<uint32_t*>(blk + off) = hi;
    *reinterpret_cast<uint32_t*>(blk + B_BYTES + off) = lo;
  }
}

__global__ void __launch_bounds__(256) taps_table_kernel(AggParams p, TapRec* __restrict__ recs) {
  const long long total = (long long)p.V * p.S * p.nl * p.LW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(idx % p.LW);
    const int n = (int)((idx / p.LW) % p.nl);
    const int s = (int)((idx / ((long long)p.LW * p.nl)) % p.S);
    const int v = (int)(idx / ((long long)p.LW * p.nl * p.S));
    const BoxTaps t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)v * p.nl + n) * p.LW + cell], p.sc[s]);
    TapRec r;
    r.xy = t.x0 | (t.y0 << 16);
    r.nxy = t.nx | (t.ny << 16);
    r.wx_first = t.wx_first;
    r.wx_last = t.wx_last;
    r.wy_first = t.wy_first;
    r.wy_last = t.wy_last;
    r.wy_mid = t.wy_mid;
    r.pad = 0;
    recs[idx] = r;
  }
}

int launch_taps_table(const AggParams& p, TapRec* recs, cudaStream_t st) {
  taps_table_kernel<<<148 * 8, 256, 0, st>>>(p, recs);
  VFA_LAUNCH_CHECK("taps_table_kernel");
  return VFA_OK;
}

struct UmmaArgs {
  AggParams p;
  const uint8_t* wprep[VFA_MAX_SCALES];
  int n_groups;        // view groups (grid.x = tiles_padded * n_groups); > 1 -> atomic accumulation into a zeroed output
  int tiles_padded;    // cell tiles, rounded up to a multiple of the cluster size
  int tiles_x;         // tiles are TILE_H x TILE_W blocks of BEV cells: tile -> (tile / tiles_x, tile % tiles_x)
  const TapRec* recs;  // [V][S][nl][LW] gather recipes (taps_table_kernel)
  int views_per_group;
  int variant;         // debug bits: 1 = hi*hi only, 2 = no gather loads, 4 = no MMA, 8 = no weight loads, 16 = no output pass, 32 = no drain,
                       // 64 = feature-side: GEMM only, 128 = feature-side: pooling only, 256 = reuse the tap records
};

// ---- pooling producers ----------------------------------------------------------------------------------------
// One work item = 4 cells (one per quarter-warp) x 32 channels of one K chunk; 8 lanes x float4 read one 128-byte
// row of a texel, so every load instruction moves 4 full cache lines.  A producer warp owns the tile rows
// (128/ROUNDS)*round + 4*pw + {0,1,2,3}: the 4 cells of one instruction are adjacent (their boxes overlap, so
// their loads coalesce / hit L1) while the rounds are spread over the tile (box sizes vary smoothly along a BEV
// row, so every warp gets the same mix of cheap and expensive rows).  Loads are predicated on the tap weight (taps with weight 0 are never
// fetched) and issued ahead of their use so several items / windows are in flight per warp.

// Feature element type of the gather.  fp32: a thread's 4 channels are one 128-bit load (float4).  bf16 storage
// (VFA_FLAG_BF16_FEATURES): one 64-bit load (4 x bf16), widened to fp32 in registers -- half the gather bytes and half
// the registers per tap in flight; pooling, split and contraction stay fp32 / 3xTF32.
//   ldg_if<TX>:  @(w != 0) load the tap TX texels to the right of ptr; the slot keeps its old (finite) contents when
//                the tap is off.  With VFA_L1_PREFETCH the same predicate also prefetches the texel's next 128-byte
//                line into L1 (what the same rows load for the next K chunk); measured slower, off by default.
//   fma4:        acc += w * slot on 4 channels with two packed fp32x2 FMAs (sm_100), predicated on w != 0 so a stale
//                Inf / NaN left in an unused slot can never leak in.
#ifndef VFA_L1_PREFETCH
#define VFA_L1_PREFETCH 0
#endif
template <bool BF16>
struct Feat;

template <>
struct Feat<false> {
  using Slot = float4;
  static constexpr int ES = 4;      // bytes per element
  __device__ static __forceinline__ Slot zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  template <int TX>
  __device__ static __forceinline__ void ldg_if(Slot& v, const uint8_t* ptr, float w) {
#if VFA_L1_PREFETCH
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.neu.f32 p, %5, 0f00000000;\n\t"
        "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+%6];\n\t"
        "@p prefetch.global.L1 [%4+%7];\n\t"
        "}\n"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(ptr), "f"(w), "n"(TX * CH * ES), "n"(TX * CH * ES + KCH * ES));
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.neu.f32 p, %5, 0f00000000;\n\t"
        "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+%6];\n\t"
        "}\n"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(ptr), "f"(w), "n"(TX * CH * ES));
#endif
  }
  __device__ static __forceinline__ void fma4(float4& acc, float w, const Slot& v) {
    if (w != 0.f) {
      const float2 w2 = make_float2(w, w);
      const float2 lo = __ffma2_rn(w2, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
      const float2 hi = __ffma2_rn(w2, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
      acc = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
  }
};

template <>
struct Feat<true> {
  using Slot = uint2;               // 4 x bf16
  static constexpr int ES = 2;
  __device__ static __forceinline__ Slot zero() { return make_uint2(0u, 0u); }
  template <int TX>
  __device__ static __forceinline__ void ldg_if(Slot& v, const uint8_t* ptr, float w) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.neu.f32 p, %3, 0f00000000;\n\t"
        "@p ld.global.nc.v2.b32 {%0, %1}, [%2+%4];\n\t"
        "}\n"
        : "+r"(v.x), "+r"(v.y)
        : "l"(ptr), "f"(w), "n"(TX * CH * ES));
  }
  __device__ static __forceinline__ void fma4(float4& acc, float w, const Slot& v) {
    if (w != 0.f) {
      const float2 w2 = make_float2(w, w);      // bf16 -> fp32 is a 16-bit shift (exact)
      const float2 a = make_float2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
      const float2 c = make_float2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
      const float2 lo = __ffma2_rn(w2, a, make_float2(acc.x, acc.y));
      const float2 hi = __ffma2_rn(w2, c, make_float2(acc.z, acc.w));
      acc = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
  }
};

// Store one item's pooled values as the tf32 hi/lo pair into the swizzled operand tiles.
__device__ __forceinline__ void store_split(uint8_t* a_hi, uint32_t off, const float4& acc) {
  uint4 hi, lo;
  hi.x = to_tf32(acc.x);
  hi.y = to_tf32(acc.y);
  hi.z = to_tf32(acc.z);
  hi.w = to_tf32(acc.w);
  // lo = a - hi is exact in fp32; the tensor core reads its top 19 bits (truncation error 2^-11 of lo = 2^-22 of a)
  const float2 l01 = __fadd2_rn(make_float2(acc.x, acc.y), make_float2(-__uint_as_float(hi.x), -__uint_as_float(hi.y)));
  const float2 l23 = __fadd2_rn(make_float2(acc.z, acc.w), make_float2(-__uint_as_float(hi.z), -__uint_as_float(hi.w)));
  lo.x = __float_as_uint(l01.x);
  lo.y = __float_as_uint(l01.y);
  lo.z = __float_as_uint(l23.x);
  lo.w = __float_as_uint(l23.y);
  *reinterpret_cast<uint4*>(a_hi + off) = hi;
  *reinterpret_cast<uint4*>(a_hi + A_BYTES + off) = lo;
}

// Position of a role in the ring of STAGES smem slots: slot index and the parity its barriers are at.
struct Pipe {
  uint32_t st, ph;
  __device__ __forceinline__ void advance() {
    if (++st == STAGES) {
      st = 0;
      ph ^= 1u;
    }
  }
};

struct ProducerCtx {
  uint8_t* smem;
  SmemTail* tail;
  const RowDesc* wdesc;     // this warp's recipes (shared memory), rows of round r at [4*r .. 4*r+3]
  const uint8_t* feat;      // [fh, fw, CH] map of this (frame, view, scale), + this thread's channel offset (bytes)
  size_t row_stride;        // fw * CH * element size (bytes)
  uint32_t a_off[ROUNDS];   // swizzled byte offset of (this thread's row of round r, its 16-byte chunk) in an A tile
  int lane, q;
  bool no_gather;           // debug knock-out
};

// Called when an item is complete: waits for the stage slot (first round only), stores, signals (last round).
__device__ __forceinline__ void finish_item(const ProducerCtx& c, int round, uint32_t a_off, float4& acc, Pipe& pipe) {
  if (round == 0) mbar_wait(&c.tail->empty[pipe.st], pipe.ph ^ 1u);
  store_split(c.smem + (size_t)pipe.st * STAGE_BYTES, a_off, acc);
  acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (round == ROUNDS - 1) {
    fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.tail->full[pipe.st]);
    pipe.advance();
  }
}

// Layers where every row of the warp has at most TxT taps (T = 2: the common case at strides 16 and 32; T = 3:
// most of stride 8): 32 items, DEPTH in flight, tap weights precomputed in the recipe.
template <int T, int DEPTH, bool BF16>
__device__ __forceinline__ void produce_layer_small(const ProducerCtx& c, Pipe& pipe) {
  using F = Feat<BF16>;
  using Slot = typename F::Slot;
  constexpr int ITEMS = (CH / KCH) * ROUNDS;
  static_assert(ROUNDS % DEPTH == 0, "DEPTH must divide ROUNDS");
  static_assert(ITEMS % ROUNDS == 0, "");
  const RowDesc* dq = c.wdesc + c.q;          // this quarter-warp's row of round 0; round r is dq[4*r]
  const uint8_t* feat = c.feat;
  const size_t rs = c.row_stride;
  Slot buf[DEPTH][T][T];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d)
#pragma unroll
    for (int ty = 0; ty < T; ++ty)
#pragma unroll
      for (int tx = 0; tx < T; ++tx) buf[d][ty][tx] = F::zero();
  auto issue = [&](Slot(&v)[T][T], int round, int cc) {
    if (c.no_gather) return;
    const RowDesc& d = dq[4 * round];
    const uint8_t* r = feat + (size_t)(d.base + cc * KCH) * F::ES;
#pragma unroll
    for (int ty = 0; ty < T; ++ty) {
      F::template ldg_if<0>(v[ty][0], r, d.w9[ty * 3 + 0]);
      F::template ldg_if<1>(v[ty][1], r, d.w9[ty * 3 + 1]);
      if (T > 2) F::template ldg_if<2>(v[ty][T - 1], r, d.w9[ty * 3 + 2]);
      r += rs;
    }
  };
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) issue(buf[d], d % ROUNDS, d / ROUNDS);
  // The DEPTH resident items are processed phase-batched (all reductions, then all loads of the items DEPTH ahead,
  // then all stores) so their dependency chains interleave instead of running back to back.
#pragma unroll 1
  for (int cc = 0; cc < CH / KCH; ++cc) {
#pragma unroll
    for (int g = 0; g < ROUNDS; g += DEPTH) {
      float4 accr[DEPTH];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        accr[d] = make_float4(0.f, 0.f, 0.f, 0.f);
        const RowDesc& rd = dq[4 * (g + d)];
#pragma unroll
        for (int ty = 0; ty < T; ++ty)
#pragma unroll
          for (int tx = 0; tx < T; ++tx) F::fma4(accr[d], rd.w9[ty * 3 + tx], buf[d][ty][tx]);
      }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int nr = (g + d + DEPTH) % ROUNDS, ncc = cc + (g + d + DEPTH) / ROUNDS;
        if (ncc < CH / KCH) issue(buf[d], nr, ncc);
      }
      if (g == 0) mbar_wait(&c.tail->empty[pipe.st], pipe.ph ^ 1u);
      uint8_t* a_hi = c.smem + (size_t)pipe.st * STAGE_BYTES;
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) store_split(a_hi, c.a_off[g + d], accr[d]);
      if (g + DEPTH == ROUNDS) {
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.tail->full[pipe.st]);
        pipe.advance();
      }
    }
  }
}

// General layers: every item walks its box in 3x3 windows (boxes grow to 8x8 texels and beyond for voxels close
// to a camera before the reference's area cap hides them); the next window's 9 predicated loads are in flight
// while the current one is reduced.  nwx / nwy hold the per-round window counts (8 bits per round, warp-uniform).
struct WinPos {
  int cc, rd, sy, sx;
};

template <bool BF16>
__device__ __forceinline__ void produce_layer_win(const ProducerCtx& c, uint32_t nwx, uint32_t nwy, Pipe* pipe_io) {
  using F = Feat<BF16>;
  using Slot = typename F::Slot;
  Pipe pipe = *pipe_io;
  Slot bufA[3][3], bufB[3][3];      // bufB is dead (eliminated) when WINDOW_DOUBLE_BUFFER is false
  float wA[3][3], wB[3][3];
#pragma unroll
  for (int ty = 0; ty < 3; ++ty)
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      bufA[ty][tx] = F::zero();
      if (WINDOW_DOUBLE_BUFFER) bufB[ty][tx] = F::zero();
    }

  auto issue = [&](Slot(&v)[3][3], float(&w)[3][3], const WinPos& p) {
    const RowDesc& d = c.wdesc[p.rd * 4 + c.q];
    const bool single = (((nwx >> (8 * p.rd)) & 255u) == 1u) && (((nwy >> (8 * p.rd)) & 255u) == 1u);   // warp-uniform
    if (single) {            // all 4 boxes of this round fit one 3x3 window: precomputed products
#pragma unroll
      for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) w[ty][tx] = d.w9[ty * 3 + tx];
    } else {
      float wx[3], wy[3];
      if (3 * p.sx + 2 < 8 && 3 * p.sy + 2 < 8) {          // warp-uniform
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wx[k] = d.wx8[3 * p.sx + k];
          wy[k] = d.wy8[3 * p.sy + k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wx[k] = desc_wx(d, 3 * p.sx + k);
          wy[k] = desc_wy(d, 3 * p.sy + k);
        }
      }
#pragma unroll
      for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) w[ty][tx] = wy[ty] * wx[tx];
    }
    if (c.no_gather) return;
    const uint8_t* r = c.feat + (size_t)(3 * p.sy) * c.row_stride +
                       ((size_t)d.base + (size_t)(3 * p.sx) * CH + p.cc * KCH) * F::ES;
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) {
      F::template ldg_if<0>(v[ty][0], r, w[ty][0]);
      F::template ldg_if<1>(v[ty][1], r, w[ty][1]);
      F::template ldg_if<2>(v[ty][2], r, w[ty][2]);
      r += c.row_stride;
    }
  };
  auto advance = [&](WinPos p, bool& valid) {
    valid = true;
    if (++p.sx < (int)((nwx >> (8 * p.rd)) & 255u)) return p;
    p.sx = 0;
    if (++p.sy < (int)((nwy >> (8 * p.rd)) & 255u)) return p;
    p.sy = 0;
    if (++p.rd < ROUNDS) return p;
    p.rd = 0;
    if (++p.cc < CH / KCH) return p;
    valid = false;
    return p;
  };
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto consume = [&](const Slot(&v)[3][3], const float(&w)[3][3], const WinPos& p) {
#pragma unroll
    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
      for (int tx = 0; tx < 3; ++tx) F::fma4(acc, w[ty][tx], v[ty][tx]);
    if (p.sx + 1 == (int)((nwx >> (8 * p.rd)) & 255u) && p.sy + 1 == (int)((nwy >> (8 * p.rd)) & 255u))
      finish_item(c, p.rd, c.a_off[0] + (uint32_t)p.rd * ((TILE_M / ROUNDS) * 128u), acc, pipe);
  };

  WinPos cur{0, 0, 0, 0}, nxt;
  bool more;
  issue(bufA, wA, cur);
  if (WINDOW_DOUBLE_BUFFER) {
    while (true) {
      nxt = advance(cur, more);
      if (more) issue(bufB, wB, nxt);
      consume(bufA, wA, cur);
      if (!more) break;
      cur = nxt;
      nxt = advance(cur, more);
      if (more) issue(bufA, wA, nxt);
      consume(bufB, wB, cur);
      if (!more) break;
      cur = nxt;
    }
  } else {
    while (true) {
      consume(bufA, wA, cur);
      nxt = advance(cur, more);
      if (!more) break;
      cur = nxt;
      issue(bufA, wA, cur);
    }
  }
  *pipe_io = pipe;
}

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1) aggregate_fwd_umma_kernel(const UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // align by pointer arithmetic on the __shared__ array (keeps the shared address space visible to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const AggParams& p = a.p;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // grid.x = tiles_padded * n_groups; consecutive blockIdx.x (= the two CTAs of a pair) are adjacent tiles of a group
  const int tile = blockIdx.x % a.tiles_padded, group = blockIdx.x / a.tiles_padded;
  const int b = blockIdx.y;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;      // 0 = pair leader (issues the MMAs)
  // cell index of tile row r (0..127), or -1 outside the grid (edge tiles, the padding tile of an odd pair)
  const int tile_y0 = (tile / a.tiles_x) * TILE_H, tile_x0 = (tile % a.tiles_x) * TILE_W;
  auto cell_of_row = [&](int r) -> int {
    const int y = tile_y0 + r / TILE_W, x = tile_x0 + r % TILE_W;
    return (y < p.L && x < p.W) ? y * p.W + x : -1;
  };
  const int v_begin = group * a.views_per_group;
  const int v_end = min(p.V, v_begin + a.views_per_group);
  const int n_vs = (v_end - v_begin) * p.S;          // (view, scale) iterations of this CTA
  constexpr int CHUNKS_PER_LAYER = CH / KCH;

  // ---- one-time setup ----
  for (int i = tid; i < p.S * CH; i += THREADS) tail->bias[i / CH][i % CH] = __ldg(p.bias[i / CH] + (i % CH));
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&tail->full[s], NUM_PRODUCER_WARPS + 1);
      mbar_init(&tail->empty[s], 1);
      mbar_init(&tail->peer_full[s], 1);
    }
    mbar_init(&tail->acc_full, 1);
    mbar_init(&tail->acc_empty, PAIR ? 2 * NUM_EPILOGUE_WARPS : NUM_EPILOGUE_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {   // both CTAs of the pair, same warp, same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // barriers + TMEM of both CTAs exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp < FIRST_PRODUCER_WARP) {
    // warpgroup 0 (loader, MMA issuer, two idle warps) hands registers to the producer warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_WG0));
  if (warp == 0) {
    // ================= weight loader =================
    if (lane == 0) {
      int it = 0;
      for (int v = v_begin; v < v_end; ++v)
        for (int s = 0; s < p.S; ++s)
          for (int kc = 0; kc < p.nl * CHUNKS_PER_LAYER; ++kc, ++it) {
            const int st = it % STAGES;
            mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
            uint8_t* dst = smem + (size_t)st * STAGE_BYTES + 2 * A_BYTES;
            if (a.variant & 8) {   // debug knock-out: no weight traffic
              mbar_arrive(&tail->full[st]);
            } else {
              mbar_arrive_expect_tx(&tail->full[st], 2 * B_LOCAL_BYTES);
              const uint8_t* src = a.wprep[s] + (size_t)kc * (2 * B_BYTES);
              if (PAIR) {   // this CTA's half of the output-channel rows: rows [128*rank, 128*rank+128) of hi and of lo
                bulk_g2s(dst, src + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
                bulk_g2s(dst + B_LOCAL_BYTES, src + B_BYTES + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
              } else {
                bulk_g2s(dst, src, 2 * B_BYTES, &tail->full[st]);
              }
            }
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The TMEM accumulator is only trusted for ONE height layer (32 k-steps): the tensor core truncates every
    // accumulation (measured: error grows linearly in K, biased toward zero), so layer partials are summed by the
    // epilogue warps with round-to-nearest FADDs (drain) instead of inside the tensor core.
    if (lane == 0 && cta_rank != 0) {
      // pair follower: its tensor core is driven by the leader's instructions; this thread only relays "my stage is
      // full" (own producers + own half of the weight slab) to the leader
      const int total = n_vs * p.nl * CHUNKS_PER_LAYER;
      for (int it = 0; it < total; ++it) {
        const int st = it % STAGES;
        mbar_wait(&tail->full[st], (it / STAGES) & 1);
        mbar_arrive_remote(&tail->peer_full[st], 0);
      }
    } else if (lane == 0) {
      int it = 0, drain = 0;
      for (int vs = 0; vs < n_vs; ++vs) {
        for (int n = 0; n < p.nl; ++n, ++drain) {
          if (PAIR) mbar_wait_cluster(&tail->acc_empty, (drain & 1) ^ 1); else mbar_wait(&tail->acc_empty, (drain & 1) ^ 1);
          tc_fence_after();
          for (int cc = 0; cc < CHUNKS_PER_LAYER; ++cc, ++it) {
            const int st = it % STAGES;
            mbar_wait(&tail->full[st], (it / STAGES) & 1);
            if (PAIR) mbar_wait_cluster(&tail->peer_full[st], (it / STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
            const uint64_t b_hi = make_desc(sa + 2 * A_BYTES), b_lo = make_desc(sa + 2 * A_BYTES + B_LOCAL_BYTES);
#pragma unroll
            for (int ks = 0; ks < KCH / 8; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);      // 8 tf32 = 32 bytes along K inside the swizzle row
              if (a.variant & 4) {
                // debug knock-out: no MMA
              } else if (a.variant & 1) {
                tc_mma_tf32(tmem, a_hi + adv, b_hi + adv, IDESC, (cc | ks) ? 1u : 0u);
              } else {
                tc_mma_tf32(tmem, a_lo + adv, b_hi + adv, IDESC, (cc | ks) ? 1u : 0u);
                tc_mma_tf32(tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
                tc_mma_tf32(tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
              }
            }
            tc_commit(&tail->empty[st]);          // frees the stage when these MMAs have read it
          }
          tc_commit(&tail->acc_full);             // this layer's partial sum is complete
        }
      }
    }
  }
  } else if (warp < FIRST_EPILOGUE_WARP) {
    // ================= pool producers =================
    if (REGS_PRODUCER > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER > 0 ? REGS_PRODUCER : 128));
    const int pw = warp - FIRST_PRODUCER_WARP;       // tile rows (128/ROUNDS)*round + 4*pw + {0..3}
    RowDesc* wdesc = tail->desc[pw];
    ProducerCtx c;
    c.smem = smem;
    c.tail = tail;
    c.wdesc = wdesc;
    c.lane = lane;
    c.q = lane >> 3;
    c.no_gather = (a.variant & 2) != 0;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
      c.a_off[r] = swz((uint32_t)((TILE_M / ROUNDS) * r + 4 * pw + (lane >> 3)), (uint32_t)(lane & 7));
    Pipe pipe{0u, 0u};
    for (int v = v_begin; v < v_end; ++v) {
      for (int s = 0; s < p.S; ++s) {
        const ScaleConst sc = p.sc[s];
        c.feat = reinterpret_cast<const uint8_t*>(p.feats[s]) +
                 (((size_t)(b * p.V + v) * sc.fh * sc.fw) * CH + (lane & 7) * 4) * Feat<BF16>::ES;
        c.row_stride = (size_t)sc.fw * CH * Feat<BF16>::ES;
        for (int n = 0; n < p.nl; ++n) {
          __syncwarp();                              // everyone is done with the previous layer's recipes
          int nx = 0, ny = 0;
          if (lane < ROWS_PER_WARP) {
            const int cell = cell_of_row((TILE_M / ROUNDS) * (lane >> 2) + 4 * pw + (lane & 3));
            RowDesc d;
            d.base = 0;
            d.nx = d.ny = 0;
            d.wx_first = d.wx_last = d.wy_first = d.wy_last = d.wy_mid = 0.f;
            if (cell >= 0) {
              const uint4* rp = reinterpret_cast<const uint4*>(a.recs + (((size_t)v * p.S + s) * p.nl + n) * p.LW + cell);
              const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              d.base = (((int)r0.x >> 16) * sc.fw + ((int)r0.x & 0xffff)) * CH;
              d.nx = (int)r0.y & 0xffff;
              d.ny = (int)r0.y >> 16;
              d.wx_first = __uint_as_float(r0.z);
              d.wx_last = __uint_as_float(r0.w);
              d.wy_first = __uint_as_float(r1.x);
              d.wy_last = __uint_as_float(r1.y);
              d.wy_mid = __uint_as_float(r1.z);
            }
#pragma unroll
            for (int ty = 0; ty < 3; ++ty)
#pragma unroll
              for (int tx = 0; tx < 3; ++tx) d.w9[ty * 3 + tx] = desc_wy(d, ty) * desc_wx(d, tx);
            d.w9[9] = d.w9[10] = d.w9[11] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              d.wx8[k] = desc_wx(d, k);
              d.wy8[k] = desc_wy(d, k);
            }
            wdesc[lane] = d;
            nx = d.nx;
            ny = d.ny;
          }
          // per-round (4 consecutive lanes) maxima, then the window counts of the 4 rounds packed 8 bits each
          nx = max(nx, __shfl_xor_sync(0xffffffffu, nx, 1));
          nx = max(nx, __shfl_xor_sync(0xffffffffu, nx, 2));
          ny = max(ny, __shfl_xor_sync(0xffffffffu, ny, 1));
          ny = max(ny, __shfl_xor_sync(0xffffffffu, ny, 2));
          uint32_t nwx = 0, nwy = 0;
          int extent = 0;
#pragma unroll
          for (int r = 0; r < ROUNDS; ++r) {
            const int mx = __shfl_sync(0xffffffffu, nx, 4 * r), my = __shfl_sync(0xffffffffu, ny, 4 * r);
            extent = max(extent, max(mx, my));
            nwx |= (uint32_t)min(255, max(1, (mx + 2) / 3)) << (8 * r);
            nwy |= (uint32_t)min(255, max(1, (my + 2) / 3)) << (8 * r);
          }
          __syncwarp();
          if (extent <= 2)
            produce_layer_small<2, (BF16 ? ROUNDS : DEPTH_2X2), BF16>(c, pipe);
          else if (extent <= 3)
            produce_layer_small<3, (BF16 ? ROUNDS : DEPTH_3X3), BF16>(c, pipe);
          else
            produce_layer_win<BF16>(c, nwx, nwy, &pipe);
        }
      }
    }
  } else {
    // ================= epilogue: layer drains + per-(view, scale) finalisation =================
    if (REGS_EPILOGUE > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPILOGUE > 0 ? REGS_EPILOGUE : 96));
    const int e = warp - FIRST_EPILOGUE_WARP;      // 0..7
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int col_begin = (e >> 2) * (CH / 2);     // two warps per quarter split the 256 columns
    const int row = quarter * 32 + lane;
    const int cell = cell_of_row(row);
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* outp = p.out + (size_t)b * CH * p.LW + cell;
    int drain = 0;
    for (int vs = 0; vs < n_vs; ++vs) {
      const int s = vs % p.S;
      for (int n = 0; n < p.nl; ++n, ++drain) {
        mbar_wait_sleep(&tail->acc_full, drain & 1);
        tc_fence_after();
        const bool first_layer = (n == 0), last_layer = (n == p.nl - 1);
        // pass 1: fold this layer's partial into the fp32 running sum kept in TMEM columns [256, 512)
#pragma unroll 1
        for (int c0 = col_begin; c0 < col_begin + CH / 2 && !(a.variant & 32); c0 += 32) {
          float acc[32], pre[32];
          tc_ld32(lane_addr + c0, acc);
          if (!first_layer) tc_ld32(lane_addr + CH + c0, pre);
          tc_wait_ld();
          if (!first_layer) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] += pre[i];                   // round-to-nearest layer sum
          }
          tc_st32(lane_addr + CH + c0, acc);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                   // accumulator columns may be overwritten by the next layer
          if (PAIR && cta_rank != 0) mbar_arrive_remote(&tail->acc_empty, 0);
          else mbar_arrive(&tail->acc_empty);
        }
        if (last_layer && !(a.variant & 16)) {
          // pass 2 (overlaps the next layer's MMAs): + bias, ReLU (vfa_op.py:123-124), then the sum over scales and
          // views (vfanet.py:79, :82) straight into the [B, C, L*W] output
#pragma unroll 1
          for (int c0 = col_begin; c0 < col_begin + CH / 2; c0 += 32) {
            float y[32];
            tc_ld32(lane_addr + CH + c0, y);
            tc_wait_ld();
            if (cell >= 0) {
              if (p.mask != nullptr) {      // ReLU pass bits of channels c0 .. c0+31 for the backward
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) bits |= (y[i] + tail->bias[s][c0 + i] > 0.f ? 1u : 0u) << i;
                const int v_cur = v_begin + vs / p.S;
                p.mask[((((size_t)b * p.V + v_cur) * p.S + s) * (CH / 32) + c0 / 32) * p.LW + cell] = bits;
              }
              float* o = outp + (size_t)c0 * p.LW;
              if (a.n_groups == 1) {
                // this CTA is the only writer of its output tile: first (view, scale) stores, later ones add with a
                // plain read-modify-write (L2-resident).  Atomics are kept for the view-split case only: 131 M
                // RED.ADD per frame ran into the L2 atomic-unit throughput (~95 G/s) and cost 1.4 ms of a 2.5 ms frame.
                if (vs == 0) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[(size_t)i * p.LW] = fmaxf(y[i] + tail->bias[s][c0 + i], 0.f);
                } else {
                  float prev[32];
#pragma unroll
                  for (int i = 0; i < 32; ++i) prev[i] = o[(size_t)i * p.LW];
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[(size_t)i * p.LW] = prev[i] + fmaxf(y[i] + tail->bias[s][c0 + i], 0.f);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(o + (size_t)i * p.LW, fmaxf(y[i] + tail->bias[s][c0 + i], 0.f));
              }
            }
          }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // the peer's shared memory / TMEM stay valid until both are done
  if (warp == 1) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
bool umma_supported(const vfa_geometry_t* g, const vfa_shape_t* sh, uint32_t flags) {
  (void)g;
  (void)flags;
  return sh->channels == CH;
}

// feature-side forward (vfa_fwd_fside.cu)
size_t fside_y_bytes_per_frame(const AggParams& p);
int fside_chunk_frames(const AggParams& p);
size_t fside_workspace_bytes(const AggParams& p);
int launch_fwd_fside(const AggParams& p, const uint8_t* const* wprep, const TapRec* recs, void* fs_ws, size_t fs_bytes,
                     uint32_t flags, int variant, cudaStream_t st);

static bool grid_side_requested(uint32_t flags) {
  return (flags & VFA_FLAG_GRID_SIDE) != 0 || runtime_config().fwd_gridside != 0;
}

// workspace: [prepared weights][tap records][Y of one frame chunk (feature-side forward only)]
static size_t umma_fixed_bytes(int S, int nl, int V, size_t LW) {
  const size_t b = (size_t)S * nl * (CH / KCH) * (2 * B_BYTES) + (size_t)V * S * nl * LW * sizeof(TapRec);
  return (b + 255) & ~(size_t)255;
}

size_t umma_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh, uint32_t flags) {
  if (!umma_supported(g, sh, flags)) return 0;
  size_t bytes = umma_fixed_bytes(sh->n_scales, g->n_layers, sh->n_views, (size_t)g->grid_l * g->grid_w);
  if (!grid_side_requested(flags)) {
    AggParams p;
    p.B = sh->batch;
    p.V = sh->n_views;
    p.nl = g->n_layers;
    p.S = sh->n_scales;
    p.L = g->grid_l;                    // the quads' texel lists are sized by the grid
    p.W = g->grid_w;
    p.LW = g->grid_l * g->grid_w;
    p.y_bf16 = (flags & VFA_FLAG_BF16_MMA) ? 1 : 0;
    for (int s = 0; s < sh->n_scales; ++s) {
      p.sc[s].fh = sh->feat_h[s];
      p.sc[s].fw = sh->feat_w[s];
    }
    bytes += fside_workspace_bytes(p);
  }
  return bytes;
}

int prep_weights_bf16(const AggParams& p, const float* const* d_weight, void* ws, size_t per_scale, cudaStream_t st);

int prep_weights_umma(const AggParams& p, const float* const* d_weight, void* ws, uint32_t flags, cudaStream_t st) {
  const size_t per_scale = (size_t)p.nl * (CH / KCH) * (2 * B_BYTES);
  if ((flags & VFA_FLAG_BF16_MMA) && !grid_side_requested(flags))      // bf16 slabs in the same (larger) per-scale slots
    return prep_weights_bf16(p, d_weight, ws, per_scale, st);
  PrepWeightArgs a = {};
  for (int s = 0; s < p.S; ++s) {
    a.w[s] = d_weight[s];
    a.wp[s] = reinterpret_cast<uint8_t*>(ws) + s * per_scale;
  }
  prep_weight_umma_kernel<<<dim3(148 * 2, p.S), 256, 0, st>>>(a, p.nl);
  VFA_LAUNCH_CHECK("prep_weight_umma_kernel");
  return VFA_OK;
}

int launch_fwd_umma(AggParams p, const float* const* d_weight, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  if (!(flags & VFA_FLAG_WEIGHTS_PREPARED)) {
    if (int rc = prep_weights_umma(p, d_weight, ws, flags, st)) return rc;
  }
  const bool bf16 = (flags & VFA_FLAG_BF16_FEATURES) != 0;
  if (grid_side_requested(flags)) {
    VFA_CUDA(cudaFuncSetAttribute(aggregate_fwd_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SMEM_BYTES));
    VFA_CUDA(cudaFuncSetAttribute(aggregate_fwd_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SMEM_BYTES));
  }
  UmmaArgs a;
  a.p = p;
  a.variant = runtime_config().umma_variant;
  if (flags & VFA_FLAG_TABLE_PREPARED) a.variant |= 256;      // static cameras: records / coverage / lists are still valid
  const size_t per_scale = (size_t)p.nl * (CH / KCH) * (2 * B_BYTES);
  for (int s = 0; s < VFA_MAX_SCALES; ++s) a.wprep[s] = reinterpret_cast<const uint8_t*>(ws) + (s < p.S ? s : 0) * per_scale;
  // gather recipes of every (view, scale, layer, cell): behind the prepared weights in the workspace
  {
    TapRec* recs = reinterpret_cast<TapRec*>(reinterpret_cast<uint8_t*>(ws) + (size_t)p.S * per_scale);
    if (!(a.variant & 256)) {     // (debug bit 256: reuse the records already in the workspace)
      taps_table_kernel<<<148 * 8, 256, 0, st>>>(p, recs);
      VFA_LAUNCH_CHECK("taps_table_kernel");
    }
    a.recs = recs;
  }
  if (!grid_side_requested(flags)) {
    const size_t fixed = umma_fixed_bytes(p.S, p.nl, p.V, (size_t)p.LW);
    if (ws_bytes < fixed) {
      set_error("workspace %zu < required %zu", ws_bytes, fixed);
      return VFA_ERR_WORKSPACE;
    }
    if (int rc = launch_fwd_fside(p, a.wprep, a.recs, reinterpret_cast<uint8_t*>(ws) + fixed,
                                  ws_bytes - fixed, flags, a.variant, st))
      return rc;
    set_path((flags & VFA_FLAG_BF16_MMA) ? (bf16 ? "fside_bf16mma_bf16feat" : "fside_bf16mma")
                                         : (bf16 ? "fside_tf32x3_bf16feat" : "fside_tf32x3"));
    return VFA_OK;
  }
  a.tiles_x = (p.W + TILE_W - 1) / TILE_W;
  const int tiles = a.tiles_x * ((p.L + TILE_H - 1) / TILE_H);
  a.tiles_padded = PAIR ? (tiles + 1) / 2 * 2 : tiles;
  // enough CTAs for >= ~4 waves over 148 SMs, otherwise split the views and accumulate atomically
  a.n_groups = ((long long)tiles * p.B >= 4 * 148 || p.V == 1) ? 1 : p.V;
  a.views_per_group = (p.V + a.n_groups - 1) / a.n_groups;
  if (a.n_groups > 1) VFA_CUDA(cudaMemsetAsync(p.out, 0, (size_t)p.B * CH * p.LW * sizeof(float), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.tiles_padded * a.n_groups, p.B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (bf16)
    VFA_CUDA(cudaLaunchKernelEx(&cfg, aggregate_fwd_umma_kernel<true>, a));
  else
    VFA_CUDA(cudaLaunchKernelEx(&cfg, aggregate_fwd_umma_kernel<false>, a));
  VFA_LAUNCH_CHECK("aggregate_fwd_umma_kernel");
  set_path(bf16 ? "umma_tf32x3_bf16feat" : "umma_tf32x3");
  return VFA_OK;
}

}  // namespace vfa
