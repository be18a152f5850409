// Feature-side forward (C = 256): the collapse is applied BEFORE the pooling.
//
// Box pooling is linear in the feature map, so  collapse(pool(f))  ==  sum_n pool_n(f * W_n^T):  instead of pooling
// every (cell, layer) and contracting [L*W, C*nl] x [C*nl, C] on the grid side (2*L*W*C*nl*C flops per view and
// scale, 335 GFLOP per MultiviewC frame), the per-layer products  Y_n[texel, o] = sum_c f[texel, c] * W[o, c*nl + n]
// are formed once on the image plane (2*fH*fW*C*nl*C: 87 GFLOP per frame) and the pooling runs on Y_n, summing the
// height layers in registers; bias, ReLU and the sums over scales and views (vfa_op.py:123-124, vfanet.py:79-82) close
// the pooling kernel.  Same arithmetic class as the reference (fp32 features, fp32-accurate products), 3.9x fewer
// flops (the backward in vfa_bwd.cu uses the transposed identity).  Two kernels per chunk of frames:
//
//   ygemm_kernel   tcgen05 3xTF32 GEMM  [B*V*fH*fW, 256] x [256, nl*256] for the three scales in one launch.  CTA
//                  pairs (cta_group::2, M = 256 texel rows per cluster), 3 stages of 64 KB: 8 producer warps load the
//                  feature rows (coalesced 128-bit), split them into tf32 hi/lo and write the SWIZZLE_128B A tiles;
//                  one thread bulk-copies this CTA's half of the pre-split weight slab; one thread issues
//                  {A_lo*B_hi, A_hi*B_lo, A_hi*B_hi}; the accumulator of a layer (256 TMEM columns) is double-buffered,
//                  so the 8 epilogue warps store layer n to Y while layer n+1 is multiplied.  K = 256 per output, i.e.
//                  32 k-steps chained in the (truncating) tensor-core accumulator -- the same bound as the per-layer
//                  drain of the grid-side kernel.
//   pool_y_kernel  one warp per BEV cell, 8 channels per lane (a warp-wide load = the 1 KB row of one texel): walks
//                  (view, scale, layer) with the tap records of taps_table_kernel, accumulates the box taps of Y_n
//                  with fp32 FMAs, adds bias, ReLU, sums; optional ReLU mask for the backward.  No shared-memory
//                  staging: the whole L1 serves the overlapping boxes of neighbouring cells.
//
// Y (fp32, [plane, layer, texel, o]) lives in the caller's workspace for a chunk of frames (678 MB per MultiviewC frame);
// it is 5x the feature maps but a quarter of the [V, C, nl, L, W] tensor the reference materialises per scale.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "vfa_common.cuh"
#include "vfa_pool.cuh"
#include "vfa_umma_ptx.cuh"

namespace vfa {

namespace fside {

using namespace umma;

constexpr int TILE_M = 128;              // texel rows per CTA == TMEM lanes
constexpr int KCH = 32;                  // K elements per stage
constexpr int STAGES = 3;
constexpr int A_BYTES = TILE_M * KCH * 4;            // 16 KB per hi / lo tile
constexpr int B_BYTES = CH * KCH * 4;                // 32 KB per hi / lo slab (N = 256)
constexpr int B_LOCAL_BYTES = B_BYTES / 2;           // the rows this CTA of the pair stages
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_LOCAL_BYTES;   // 64 KB
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int FIRST_PRODUCER_WARP = 4;
constexpr int FIRST_EPILOGUE_WARP = FIRST_PRODUCER_WARP + NUM_PRODUCER_WARPS;   // 12: (warp & 3) == TMEM lane quarter
constexpr int NUM_EPILOGUE_WARPS = 8;
constexpr int THREADS = (FIRST_EPILOGUE_WARP + NUM_EPILOGUE_WARPS) * 32;       // 640
constexpr int ITEMS = TILE_M * (KCH / 4) / (NUM_PRODUCER_WARPS * 32);          // float4 items per producer thread and stage
constexpr int TMEM_COLS = 512;           // two accumulator slots of 256 columns
constexpr int CHUNKS = CH / KCH;         // stages per layer
constexpr uint32_t IDESC = make_idesc_tf32(CH, 2 * TILE_M);

struct __align__(16) SmemTail {
  unsigned long long full[STAGES];        // producer warps + weight loader (tx) of THIS CTA
  unsigned long long empty[STAGES];       // tcgen05.commit, multicast to both CTAs
  unsigned long long peer_full[STAGES];   // leader only: the follower's relay
  unsigned long long acc_full[2];         // tcgen05.commit per accumulator slot
  unsigned long long acc_empty[2];        // epilogue warps of both CTAs, on the leader's barrier
  uint32_t tmem_base;
};
constexpr int TAIL_BYTES = 256;           // SmemTail, padded
constexpr int PATCH_BYTES = 32 * 128;     // per epilogue warp: 32 rows x 32 fp32, swizzled
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + TAIL_BYTES + (size_t)NUM_EPILOGUE_WARPS * PATCH_BYTES;
static_assert(sizeof(SmemTail) <= TAIL_BYTES, "");
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");

// MODE 0 (forward): A = feature rows [rows, 256], re-read for every layer; output = one [rows-of-plane, 256] slab per
//                   layer, Y[plane][layer][texel][256].
// MODE 1 (backward, dFeature): A = Gs rows [rows, nl*256], the K loop runs through the layers; the layer partials are
//                   summed into ONE output row block out[rows, 256] (first layer stores, later layers add with fp32
//                   read-modify-write: the tensor-core accumulator is only trusted for 32 k-steps).
struct YGemmArgs {
  const uint8_t* feats[VFA_MAX_SCALES];   // A rows of this chunk (fp32; bf16 in MODE 0 with VFA_FLAG_BF16_FEATURES)
  float* y[VFA_MAX_SCALES];               // MODE 0: [plane][layer][texel][256]; MODE 1: [rows][256]
  const uint8_t* wprep[VFA_MAX_SCALES];   // prepared weights (prep_weight_umma_kernel layout)
  int rows[VFA_MAX_SCALES];               // planes * fh * fw
  int hw[VFA_MAX_SCALES];                 // fh * fw
  int tile_begin[VFA_MAX_SCALES + 1];     // cluster tiles (256 rows) of scale s: [tile_begin[s], tile_begin[s+1])
  int nl, S;
  int n_tiles;                            // tile_begin[S]
  int need_tile0[VFA_MAX_SCALES];         // first tile of scale s inside `need` (== tile_begin[s] unless a scale has no rows here)
  const uint8_t* need;                    // [tile][layer] (tiles of all scales concatenated): 0 = no box of that layer touches
                                          // the tile's texels, the product is never pooled -> the whole (tile, layer) is skipped
};

__device__ __forceinline__ void store_split(uint8_t* a_hi, uint32_t off, const float4& v) {
  uint4 hi, lo;
  hi.x = to_tf32(v.x);
  hi.y = to_tf32(v.y);
  hi.z = to_tf32(v.z);
  hi.w = to_tf32(v.w);
  // lo = v - hi is exact in fp32; the tensor core reads its top 19 bits
  const float2 l01 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-__uint_as_float(hi.x), -__uint_as_float(hi.y)));
  const float2 l23 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-__uint_as_float(hi.z), -__uint_as_float(hi.w)));
  lo.x = __float_as_uint(l01.x);
  lo.y = __float_as_uint(l01.y);
  lo.z = __float_as_uint(l23.x);
  lo.w = __float_as_uint(l23.y);
  *reinterpret_cast<uint4*>(a_hi + off) = hi;
  *reinterpret_cast<uint4*>(a_hi + A_BYTES + off) = lo;
}

template <bool BF16>
__device__ __forceinline__ float4 load_feat4(const uint8_t* p) {
  if (BF16) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));      // 4 x bf16; bf16 -> fp32 is a 16-bit shift (exact)
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
  }
  return __ldg(reinterpret_cast<const float4*>(p));
}

// Geometry of one cluster tile (256 texel rows of one scale), decoded from the flat tile index.
struct TileInfo {
  int s, row0, rows, hw;
  const uint8_t* feats;
  const uint8_t* wprep;
  const uint8_t* need;     // need bytes of this tile: [layer]
  float* y;
};
__device__ __forceinline__ TileInfo tile_info(const YGemmArgs& a, int ctile, uint32_t cta_rank) {
  TileInfo t;
  t.s = (a.S > 2 && ctile >= a.tile_begin[2]) ? 2 : ((a.S > 1 && ctile >= a.tile_begin[1]) ? 1 : 0);
  const int tile0 = t.s == 0 ? 0 : (t.s == 1 ? a.tile_begin[1] : a.tile_begin[2]);
  t.row0 = (ctile - tile0) * (2 * TILE_M) + (int)cta_rank * TILE_M;      // first texel row of this CTA
  t.need = a.need + (size_t)(pick(a.need_tile0, t.s) + ctile - tile0) * a.nl;
  t.rows = pick(a.rows, t.s);
  t.hw = pick(a.hw, t.s);
  t.feats = pick(a.feats, t.s);
  t.wprep = pick(a.wprep, t.s);
  t.y = pick(a.y, t.s);
  return t;
}

// Persistent: cluster c works on tiles c, c + n_clusters, ... ; every role keeps its stage / layer counters running
// across tiles, so the last layer's epilogue and the next tile's first loads overlap the tensor-core work.
template <bool BF16, int MODE>
__global__ void __launch_bounds__(THREADS, 1) ygemm_kernel(const YGemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tiles = a.n_tiles;
  const int total = a.nl * CHUNKS;                                       // stages per tile

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&tail->full[i], NUM_PRODUCER_WARPS + 1);
      mbar_init(&tail->empty[i], 1);
      mbar_init(&tail->peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tail->acc_full[i], 1);
      mbar_init(&tail->acc_empty[i], 2 * NUM_EPILOGUE_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ================= weight loader =================
    if (lane == 0) {
      int it = 0;
      for (int ctile = cluster; ctile < n_tiles; ctile += n_clusters) {
        const TileInfo t = tile_info(a, ctile, cta_rank);
        for (int k = 0; k < total; ++k) {
          if (!__ldg(t.need + k / CHUNKS)) continue;
          const int st = it % STAGES;
          mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
          uint8_t* dst = smem + (size_t)st * STAGE_BYTES + 2 * A_BYTES;
          mbar_arrive_expect_tx(&tail->full[st], 2 * B_LOCAL_BYTES);
          const uint8_t* src = t.wprep + (size_t)k * (2 * B_BYTES);      // kc = n * CHUNKS + cc == k
          bulk_g2s(dst, src + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
          bulk_g2s(dst + B_LOCAL_BYTES, src + B_BYTES + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank != 0) {
      // pair follower: relays "my stage is full" to the leader
      int it = 0;
      for (int ctile = cluster; ctile < n_tiles; ctile += n_clusters) {
        const TileInfo t = tile_info(a, ctile, cta_rank);
        for (int k = 0; k < total; ++k) {
          if (!__ldg(t.need + k / CHUNKS)) continue;
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_arrive_remote(&tail->peer_full[st], 0);
          ++it;
        }
      }
    } else if (lane == 0) {
      // ================= MMA issuer (pair leader) =================
      int it = 0, ln = 0;                               // running stage / layer counters
      for (int ctile = cluster; ctile < n_tiles; ctile += n_clusters) {
        const TileInfo t = tile_info(a, ctile, cta_rank);
        for (int n = 0; n < a.nl; ++n) {
          if (!__ldg(t.need + n)) continue;
          const int slot = ln & 1;
          mbar_wait_cluster(&tail->acc_empty[slot], ((ln >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + (uint32_t)slot * CH;
          for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
            const int st = it % STAGES;
            mbar_wait(&tail->full[st], (it / STAGES) & 1);
            mbar_wait_cluster(&tail->peer_full[st], (it / STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
            const uint64_t b_hi = make_desc(sa + 2 * A_BYTES), b_lo = make_desc(sa + 2 * A_BYTES + B_LOCAL_BYTES);
#pragma unroll
            for (int ks = 0; ks < KCH / 8; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);
              tc_mma_tf32_t<true>(d_tmem, a_lo + adv, b_hi + adv, IDESC, (cc | ks) ? 1u : 0u);
              tc_mma_tf32_t<true>(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
              tc_mma_tf32_t<true>(d_tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
            }
            tc_commit_t<true>(&tail->empty[st]);
          }
          tc_commit_t<true>(&tail->acc_full[slot]);
          ++ln;
        }
      }
    }
  } else if (warp >= FIRST_PRODUCER_WARP && warp < FIRST_EPILOGUE_WARP) {
    // ================= A producers: feature rows -> tf32 hi / lo operand tiles =================
    constexpr int ES = BF16 ? 2 : 4;
    const int pw = warp - FIRST_PRODUCER_WARP;
    const int j = lane & 7;                          // 16-byte chunk of the 128-byte K row
    const uint8_t* src[ITEMS];
    uint32_t off[ITEMS];
    bool ok[ITEMS];
    auto bind_tile = [&](int ctile) {                // source rows of this thread's items in tile `ctile`
      const TileInfo t = tile_info(a, ctile, cta_rank);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int r = 32 * i + 4 * pw + (lane >> 3); // 4 rows x 128 B per warp instruction
        ok[i] = t.row0 + r < t.rows;
        src[i] = t.feats + ((size_t)(t.row0 + r) * (MODE == 0 ? CH : a.nl * CH) + j * 4) * ES;
      }
    };
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) off[i] = swz((uint32_t)(32 * i + 4 * pw + (lane >> 3)), (uint32_t)j);
    float4 cur[ITEMS], nxt[ITEMS];
    // cursor over the NEEDED stages of this cluster's tiles: (tile, k), k = layer * CHUNKS + K chunk
    auto advance = [&](int& ct, int& k) -> bool {
      ++k;
      while (true) {
        if (k >= total) {
          ct += n_clusters;
          k = 0;
          if (ct >= n_tiles) return false;
        }
        if (k % CHUNKS != 0 || __ldg(tile_info(a, ct, cta_rank).need + k / CHUNKS)) return true;
        k += CHUNKS;                                   // nothing pools from this (tile, layer): skip its 8 stages
      }
    };
    auto load_stage = [&](float4(&v)[ITEMS], int k) {
      const int cc = MODE == 0 ? k % CHUNKS : k;       // K chunk inside an A row
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[i]) v[i] = load_feat4<BF16>(src[i] + (size_t)cc * KCH * ES);
      }
    };
    int ct = cluster, k = -1;
    bool have = ct < n_tiles && advance(ct, k);
    if (have) {
      bind_tile(ct);
      load_stage(cur, k);
    }
    int it = 0;
    while (have) {
      const int st = it % STAGES;
      // loads of the next needed stage (possibly in the next tile) fly while this one is stored
      int nct = ct, nk = k;
      const bool more = advance(nct, nk);
      if (more) {
        if (nct != ct) bind_tile(nct);
        load_stage(nxt, nk);
      }
      mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
      uint8_t* a_hi = smem + (size_t)st * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) store_split(a_hi, off[i], cur[i]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->full[st]);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) cur[i] = nxt[i];
      ct = nct;
      k = nk;
      have = more;
      ++it;
    }
  } else if (warp >= FIRST_EPILOGUE_WARP) {
    // ================= epilogue: accumulator of layer n -> Y[plane][n][texel][:] =================
    // A thread owns one accumulator row (TMEM lane); written directly, a warp store would touch 32 different 1 KB
    // rows of Y.  Each 32 x 32 chunk is therefore transposed through a swizzled 4 KB shared-memory patch of the warp,
    // so that a store instruction writes 4 rows x 128 contiguous bytes.
    const int e = warp - FIRST_EPILOGUE_WARP;
    const int quarter = warp & 3;
    const int col_begin = (e >> 2) * (CH / 2);
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* patch = smem + (size_t)STAGES * STAGE_BYTES + TAIL_BYTES + (size_t)e * PATCH_BYTES;
    const int sub = lane >> 3, chunk = lane & 7;      // store phase: row 4*rr + sub of the patch, 16-byte chunk
    int ln = 0;
    for (int ctile = cluster; ctile < n_tiles; ctile += n_clusters) {
      const TileInfo t = tile_info(a, ctile, cta_rank);
      uint32_t yrow[8];                               // 1 KB output row (layer 0) of patch row 4*rr + sub; ~0u = outside
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int r = t.row0 + quarter * 32 + 4 * rr + sub;
        if (MODE == 0)
          yrow[rr] = r < t.rows ? (uint32_t)((r / t.hw) * a.nl * t.hw + r % t.hw) : 0xffffffffu;
        else
          yrow[rr] = r < t.rows ? (uint32_t)r : 0xffffffffu;
      }
      bool first = true;                              // MODE 1: no layer of this tile has been written yet
      for (int n = 0; n < a.nl; ++n) {
        if (!__ldg(t.need + n)) continue;
        const int slot = ln & 1;
        mbar_wait_sleep(&tail->acc_full[slot], (ln >> 1) & 1);
        tc_fence_after();
        float* dst = t.y + (MODE == 0 ? (size_t)n * t.hw * CH : (size_t)0) + col_begin + chunk * 4;
#pragma unroll 1
        for (int c0 = 0; c0 < CH / 2; c0 += 32) {
          float v[32];
          tc_ld32(lane_addr + (uint32_t)(slot * CH + col_begin + c0), v);
          tc_wait_ld();
          __syncwarp();                               // the previous chunk has left the patch
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(patch + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int prow = 4 * rr + sub;
            float4 q = *reinterpret_cast<const float4*>(patch + prow * 128 + ((chunk ^ (prow & 7)) << 4));
            if (yrow[rr] != 0xffffffffu) {
              float4* o4 = reinterpret_cast<float4*>(dst + (size_t)yrow[rr] * CH + c0);
              if (MODE == 1 && !first) {              // layer partials summed in fp32 (this thread owns the element)
                const float4 prev = *o4;
                q.x += prev.x; q.y += prev.y; q.z += prev.z; q.w += prev.w;
              }
              *o4 = q;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cta_rank != 0) mbar_arrive_remote(&tail->acc_empty[slot], 0);
          else mbar_arrive(&tail->acc_empty[slot]);
        }
        first = false;
        ++ln;
      }
      if (MODE == 1 && first) {                       // no layer touches these rows: the sum over layers is zero
        float* dst = t.y + col_begin + chunk * 4;
#pragma unroll 1
        for (int c0 = 0; c0 < CH / 2; c0 += 32)
#pragma unroll
          for (int rr = 0; rr < 8; ++rr)
            if (yrow[rr] != 0xffffffffu)
              *reinterpret_cast<float4*>(dst + (size_t)yrow[rr] * CH + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ---- pooling of Y ------------------------------------------------------------------------------------------------
constexpr int POOL_TW = 4, POOL_TH = 4;                 // cells per CTA: 4 x 4, one warp each
constexpr int POOL_WARPS = POOL_TW * POOL_TH;
constexpr int TAP_BATCH = 3;                            // taps of a box row in flight per warp

template <bool MASK>
__global__ void __launch_bounds__(POOL_WARPS * 32, 2) pool_y_kernel(const PoolArgs a) {
  const AggParams& p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cy = (blockIdx.x / a.tiles_x) * POOL_TH + warp / POOL_TW;
  const int cx = (blockIdx.x % a.tiles_x) * POOL_TW + warp % POOL_TW;
  if (cy >= p.L || cx >= p.W) return;                 // no block-level synchronisation below
  const int cell = cy * p.W + cx;
  const int bl = blockIdx.y;                          // frame inside the chunk
  const int b = a.b0 + bl;

  float out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = 0.f;

  const uint4* rp = reinterpret_cast<const uint4*>(a.recs + cell);
  const size_t rec_stride = (size_t)p.LW * (sizeof(TapRec) / sizeof(uint4));
  uint4 n0 = __ldg(rp), n1 = __ldg(rp + 1);           // record of (v, s, n) = (0, 0, 0), prefetched one step ahead
  const int total = p.V * p.S * p.nl;
  int j = 0;
  for (int v = 0; v < p.V; ++v) {
    for (int s = 0; s < p.S; ++s) {
      const int fw = p.sc[s].fw, hw = p.sc[s].fh * p.sc[s].fw;
      const float* yplane = static_cast<const float*>(a.y[s]) + ((size_t)(bl * p.V + v) * p.nl) * hw * CH + lane * 4;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int n = 0; n < p.nl; ++n, ++j, yplane += (size_t)hw * CH) {
        const uint4 r0 = n0, r1 = n1;
        if (j + 1 < total) {
          rp += rec_stride;
          n0 = __ldg(rp);
          n1 = __ldg(rp + 1);
        }
        const int nx = (int)r0.y & 0xffff, ny = (int)r0.y >> 16;
        if (nx == 0) continue;                        // not visible (warp-uniform)
        const float wx_first = __uint_as_float(r0.z), wx_last = __uint_as_float(r0.w);
        const float wy_first = __uint_as_float(r1.x), wy_last = __uint_as_float(r1.y), wy_mid = __uint_as_float(r1.z);
        const float* row = yplane + (size_t)(((int)r0.x >> 16) * fw + ((int)r0.x & 0xffff)) * CH;
        for (int ty = 0; ty < ny; ++ty, row += (size_t)fw * CH) {
          const float wy = ty == 0 ? wy_first : (ty == ny - 1 ? wy_last : wy_mid);
          if (wy == 0.f) continue;
          for (int tx = 0; tx < nx; tx += TAP_BATCH) {
            float w[TAP_BATCH];
            float4 va[TAP_BATCH], vb[TAP_BATCH];
#pragma unroll
            for (int k = 0; k < TAP_BATCH; ++k) {
              const int t = tx + k;
              w[k] = t < nx ? wy * (t == 0 ? wx_first : (t == nx - 1 ? wx_last : 1.0f)) : 0.f;
              va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (w[k] != 0.f) {                      // taps of weight 0 are never fetched
                va[k] = __ldg(reinterpret_cast<const float4*>(row + (size_t)t * CH));
                vb[k] = __ldg(reinterpret_cast<const float4*>(row + (size_t)t * CH + CH / 2));
              }
            }
#pragma unroll
            for (int k = 0; k < TAP_BATCH; ++k)
              if (w[k] != 0.f) fma8(acc, w[k], va[k], vb[k]);
          }
        }
      }
      // + bias, ReLU (vfa_op.py:123-124), sum over scales and views (vfanet.py:79, :82)
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + lane * 4));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + CH / 2 + lane * 4));
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t bits = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = acc[i] + bb[i];
        bits |= (t > 0.f ? 1u : 0u) << i;
        out[i] += fmaxf(t, 0.f);
      }
      if (MASK)
        store_mask_words(p.mask + (((size_t)b * p.V + v) * p.S + s) * (CH / 32) * p.LW + cell, (size_t)p.LW, lane, bits, true);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) p.out[out_index(a.out_nhwc, b, chan_of(lane, i), cell, p.LW)] = out[i];
}


// ---- pooling of Y, four cells per warp ------------------------------------------------------------------------------
// The boxes of neighbouring BEV cells overlap heavily (the union bounding box of a 2 x 2 block of cells holds 31-36 % of
// the texels its four boxes hold one by one), and the one-cell-per-warp kernel above runs into the L1 data pipe
// (128 B / clk / SM, 80 % busy).  Here a warp owns a 2 x 2 quad of cells: per (view, scale, layer) it walks the UNION of
// the four boxes once, loads every texel once (2 x LDG.128 per lane = the 1 KB row) and applies it to each cell whose
// box contains it.  Lane 4*j + c computes the weight of cell c for column j of the current 8-column block; the weight
// of (cell, texel) then reaches all lanes by one shuffle.  Output partial sums live in shared memory (32 KB per CTA).
#ifndef VFA_QUAD_X
#define VFA_QUAD_X 4
#endif
#ifndef VFA_QUAD_Y
#define VFA_QUAD_Y 2
#endif
#ifndef VFA_QUAD_BATCH
#define VFA_QUAD_BATCH 2
#endif
#ifndef VFA_QUAD_MINBLOCKS
#define VFA_QUAD_MINBLOCKS 3
#endif
#ifndef VFA_QUAD_PREFETCH
#define VFA_QUAD_PREFETCH 0
#endif
constexpr int QX = VFA_QUAD_X, QY = VFA_QUAD_Y;         // quads per CTA: 4 x 8 cells, 3 CTAs (24 warps) per SM
constexpr int QWARPS = QX * QY;
constexpr int QTB = VFA_QUAD_BATCH;                     // texels of a union row in flight per warp (2 x LDG.128 each)
static_assert(8 % QTB == 0, "");

// OVF = true: the completion pass behind pool_list_kernel -- only the quads whose texel list did not fit its slot are
// walked (normally none: the warp returns at once); the list kernel leaves exactly those quads alone.
// row of Y as two float4 per lane (channels chan_of(lane, 0..7)); YB: Y is stored in bf16 (widening is a 16-bit shift)
template <bool YB>
__device__ __forceinline__ void load_y_row(const void* ybase, size_t row, int lane, float4& va, float4& vb) {
  if (YB) {
    const uint8_t* r = static_cast<const uint8_t*>(ybase) + row * (CH * 2);
    const uint2 x = __ldg(reinterpret_cast<const uint2*>(r + lane * 8)), y = __ldg(reinterpret_cast<const uint2*>(r + CH + lane * 8));
    va = make_float4(__uint_as_float(x.x << 16), __uint_as_float(x.x & 0xffff0000u), __uint_as_float(x.y << 16),
                     __uint_as_float(x.y & 0xffff0000u));
    vb = make_float4(__uint_as_float(y.x << 16), __uint_as_float(y.x & 0xffff0000u), __uint_as_float(y.y << 16),
                     __uint_as_float(y.y & 0xffff0000u));
  } else {
    const float* r = static_cast<const float*>(ybase) + row * CH + lane * 4;
    va = __ldg(reinterpret_cast<const float4*>(r));
    vb = __ldg(reinterpret_cast<const float4*>(r + CH / 2));
  }
}

template <bool MASK, bool OVF = false, bool YB = false>
__global__ void __launch_bounds__(QWARPS * 32, VFA_QUAD_MINBLOCKS) pool_quad_kernel(const PoolArgs a) {
  __shared__ float out_s[QWARPS][4][CH];
  const AggParams& p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cy0 = (blockIdx.x / a.tiles_x) * (2 * QY) + 2 * (warp / QX);
  const int cx0 = (blockIdx.x % a.tiles_x) * (2 * QX) + 2 * (warp % QX);
  if (cy0 >= p.L || cx0 >= p.W) return;                 // no block-level synchronisation below
  if (OVF) {          // completion pass: only the quads the list / tile kernel left alone
    if (a.tile_ovf != nullptr) {
      if (!__ldg(a.tile_ovf + (cy0 >> 3) * a.ptiles_x + (cx0 >> 3))) return;
    } else if (__ldg(a.seg_off + (size_t)((cy0 >> 1) * a.quads_x + (cx0 >> 1)) * (p.V * p.S + 1) + p.V * p.S) != LIST_OVERFLOW) {
      return;
    }
  }
  const int bl = blockIdx.y;
  const int b = a.b0 + bl;
  const int cl = lane & 3, jl = lane >> 2;              // this lane's cell (records, weights) and block column
  const bool my_valid = cy0 + (cl >> 1) < p.L && cx0 + (cl & 1) < p.W;
  const int my_cell = my_valid ? (cy0 + (cl >> 1)) * p.W + cx0 + (cl & 1) : cy0 * p.W + cx0;

  float* const outw = &out_s[warp][0][lane * 8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<float4*>(outw + c * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(outw + c * CH + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  const uint4* rp = reinterpret_cast<const uint4*>(a.recs + my_cell);
  const size_t rec_stride = (size_t)p.LW * (sizeof(TapRec) / sizeof(uint4));
  uint4 n0 = __ldg(rp), n1 = __ldg(rp + 1);             // record of (v, s, n) = (0, 0, 0), prefetched one step ahead
  const int total = p.V * p.S * p.nl;
  int j_rec = 0;
  for (int v = 0; v < p.V; ++v) {
    for (int s = 0; s < p.S; ++s) {
      const int fw = p.sc[s].fw, hw = p.sc[s].fh * p.sc[s].fw;
      size_t yplane = ((size_t)(bl * p.V + v) * p.nl) * hw;      // first row of the layer plane inside Y of scale s
      float acc[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
      for (int n = 0; n < p.nl; ++n, ++j_rec, yplane += (size_t)hw) {
        const uint4 r0 = n0, r1 = n1;
        if (j_rec + 1 < total) {
          rp += rec_stride;
          n0 = __ldg(rp);
          n1 = __ldg(rp + 1);
        }
        const int nx = my_valid ? ((int)r0.y & 0xffff) : 0, ny = (int)r0.y >> 16;
        const int x0 = (int)r0.x & 0xffff, y0 = (int)r0.x >> 16;
        const bool vis = nx != 0;
        // union of the visible boxes of the quad (warp-uniform)
        const int ux1 = __reduce_max_sync(0xffffffffu, vis ? x0 + nx - 1 : -1);
        if (ux1 < 0) continue;
        const int ux0 = __reduce_min_sync(0xffffffffu, vis ? x0 : 0x7fff);
        const int uy0 = __reduce_min_sync(0xffffffffu, vis ? y0 : 0x7fff);
        const int uy1 = __reduce_max_sync(0xffffffffu, vis ? y0 + ny - 1 : -1);
        const float wx_first = __uint_as_float(r0.z), wx_last = __uint_as_float(r0.w);
        const float wy_first = __uint_as_float(r1.x), wy_last = __uint_as_float(r1.y), wy_mid = __uint_as_float(r1.z);
        // column blocks outermost (nearly always one): the column weight of this lane is computed once per block
        for (int cb = 0; cb <= ux1 - ux0; cb += 8) {
          const int rx = ux0 + cb + jl - x0;
          const float wx = (vis && rx >= 0 && rx < nx) ? (rx == 0 ? wx_first : (rx == nx - 1 ? wx_last : 1.0f)) : 0.f;
          size_t rowp = yplane + (size_t)uy0 * fw + ux0;
          for (int ty = uy0; ty <= uy1; ++ty, rowp += (size_t)fw) {
            const int ry = ty - y0;
            const float wy = (ry >= 0 && ry < ny) ? (ry == 0 ? wy_first : (ry == ny - 1 ? wy_last : wy_mid)) : 0.f;
            float wl = wy * wx;                         // weight of (cell cl, column cb + jl) in this row
            asm volatile("" : "+f"(wl));                // keep it in a register (no rematerialisation per shuffle)
            const uint32_t bm = __ballot_sync(0xffffffffu, wl != 0.f);     // bit 4*j + c
#pragma unroll          // j static: the shuffles get immediate lane indices
            for (int j = 0; j < 8; j += QTB) {
              if ((bm >> (4 * j)) == 0u) break;
              const size_t tp = rowp + (size_t)(cb + j);
              float4 va[QTB], vb[QTB];
#pragma unroll
              for (int k = 0; k < QTB; ++k) {
                if ((bm >> (4 * (j + k))) & 0xfu) {     // texels no box of the quad covers are never fetched
                  load_y_row<YB>(a.y[s], tp + k, lane, va[k], vb[k]);
                }
              }
#pragma unroll
              for (int k = 0; k < QTB; ++k) {
                if ((bm >> (4 * (j + k))) & 0xfu) {
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    // weight 0 for a cell whose box does not hold this texel: the FMA is issued anyway (a predicate
                    // would cost an extra instruction per cell and texel)
                    fma8(acc[c], __shfl_sync(0xffffffffu, wl, 4 * (j + k) + c), va[k], vb[k]);
                  }
                }
              }
            }
          }
        }
      }
      // + bias, ReLU (vfa_op.py:123-124), sum over scales and views (vfanet.py:79, :82)
      const float4 bi0 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + lane * 4));
      const float4 bi1 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + CH / 2 + lane * 4));
      const float bb[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 o0 = *reinterpret_cast<const float4*>(outw + c * CH);
        float4 o1 = *reinterpret_cast<const float4*>(outw + c * CH + 4);
        float t[8];
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] = acc[c][i] + bb[i];
          bits |= (t[i] > 0.f ? 1u : 0u) << i;
        }
        o0.x += fmaxf(t[0], 0.f); o0.y += fmaxf(t[1], 0.f); o0.z += fmaxf(t[2], 0.f); o0.w += fmaxf(t[3], 0.f);
        o1.x += fmaxf(t[4], 0.f); o1.y += fmaxf(t[5], 0.f); o1.z += fmaxf(t[6], 0.f); o1.w += fmaxf(t[7], 0.f);
        *reinterpret_cast<float4*>(outw + c * CH) = o0;
        *reinterpret_cast<float4*>(outw + c * CH + 4) = o1;
        if (MASK) {
          const int cy = cy0 + (c >> 1), cx = cx0 + (c & 1);
          const bool ok = cy < p.L && cx < p.W;
          store_mask_words(p.mask + (((size_t)b * p.V + v) * p.S + s) * (CH / 32) * p.LW + (ok ? cy * p.W + cx : 0),
                           (size_t)p.LW, lane, bits, ok);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cy = cy0 + (c >> 1), cx = cx0 + (c & 1);
    if (cy >= p.L || cx >= p.W) continue;
    const float4 o0 = *reinterpret_cast<const float4*>(outw + c * CH);
    const float4 o1 = *reinterpret_cast<const float4*>(outw + c * CH + 4);
    if (a.out_nhwc) {          // 512 contiguous bytes per warp store
      float* o = (a.out_mode == 3 ? owner_base(p.out, cy) : p.out) + ((size_t)b * p.LW + cy * p.W + cx) * CH + lane * 4;
      if (a.out_mode == 0) {
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + CH / 2) = o1;
      } else {
        red_add_v4(o, o0, a.out_mode);
        red_add_v4(o + CH / 2, o1, a.out_mode);
      }
      continue;
    }
    float* o = p.out + (size_t)b * CH * p.LW + cy * p.W + cx;
    const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) o[(size_t)chan_of(lane, i) * p.LW] = ov[i];
  }
}

// ---- pooling of Y from precomputed texel lists ------------------------------------------------------------------------
// pool_quad_kernel spends ~60 instructions per texel it loads -- records, union bounds, row / column weights, ballots and
// shuffles around 2 loads and 16 packed FMAs -- and every pass over (at most) two texels of a union row waits for a full
// memory round trip.  The projection is static, so that bookkeeping is done ONCE per table by qlist_build_kernel: per
// (quad, view, scale) segment the covered texels of all layers in the order pool_quad_kernel visits them, each with its
// texel index and the four cell weights (qlist_build_kernel).  The hot kernel is then a flat loop -- uniform loads of (index, weights), two
// 128-bit loads of the texel row, 16 packed FMAs -- with LIST_BATCH texels in flight whatever the shape of the union,
// and its sums are bit-identical to the walking kernel's (same weights, same order).
#ifndef VFA_LIST_BATCH
#define VFA_LIST_BATCH 3
#endif
#ifndef VFA_LIST_MINBLOCKS
#define VFA_LIST_MINBLOCKS 3
#endif
constexpr int LB = VFA_LIST_BATCH;
constexpr uint32_t LIST_PER_ITER = 16;      // slot of a quad: entries per (view, scale, layer) on average

// One (quad, view, scale, layer) iteration of the walk, done by ONE thread (the warp-cooperative form of pool_quad_kernel
// costs ~300 warp instructions per iteration; here a warp instruction serves 32 iterations).  Weights and visiting order
// are those of pool_quad_kernel: wl = wy * wx per cell, column blocks of 8, rows, columns; a texel is listed when any of
// the four weights is non-zero.  FILL = false counts the entries, FILL = true writes them from absolute index `first`.
template <bool FILL>
__device__ __forceinline__ uint32_t qlist_iter(const AggParams& p, const TapRec* __restrict__ recs, int quads_x, int q, int it,
                                               uint32_t first, uint32_t* __restrict__ ent_off, float4* __restrict__ ent_w) {
  const int n = it % p.nl, s = (it / p.nl) % p.S;
  const int fw = p.sc[s].fw, hw = p.sc[s].fh * p.sc[s].fw;
  const int cy0 = 2 * (q / quads_x), cx0 = 2 * (q % quads_x);
  int nx[4], ny[4], x0[4], y0[4];
  float wxf[4], wxl[4], wyf[4], wyl[4], wym[4];
  int ux0 = 0x7fff, uy0 = 0x7fff, ux1 = -1, uy1 = -1;
  bool odd = false;                                      // a non-finite row weight: no shortcut on the columns
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const bool valid = cy0 + (c >> 1) < p.L && cx0 + (c & 1) < p.W;
    const int cell = valid ? (cy0 + (c >> 1)) * p.W + cx0 + (c & 1) : cy0 * p.W + cx0;
    const uint4* rp = reinterpret_cast<const uint4*>(recs + (size_t)it * p.LW + cell);      // recs[v][s][n][cell]
    const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
    nx[c] = valid ? ((int)r0.y & 0xffff) : 0;
    ny[c] = (int)r0.y >> 16;
    x0[c] = (int)r0.x & 0xffff;
    y0[c] = (int)r0.x >> 16;
    wxf[c] = __uint_as_float(r0.z);
    wxl[c] = __uint_as_float(r0.w);
    wyf[c] = __uint_as_float(r1.x);
    wyl[c] = __uint_as_float(r1.y);
    wym[c] = __uint_as_float(r1.z);
    odd = odd || !(isfinite(wyf[c]) && isfinite(wyl[c]) && isfinite(wym[c]));
    if (nx[c] != 0) {
      ux0 = min(ux0, x0[c]);
      uy0 = min(uy0, y0[c]);
      ux1 = max(ux1, x0[c] + nx[c] - 1);
      uy1 = max(uy1, y0[c] + ny[c] - 1);
    }
  }
  if (ux1 < 0) return 0u;
  uint32_t total = 0;
  for (int cb = 0; cb <= ux1 - ux0; cb += 8) {
    // past ux1 every column weight is 0, so a finite row weight gives wl = 0: those columns hold no entry
    const int jmax = odd ? 7 : min(7, ux1 - ux0 - cb);
    for (int ty = uy0; ty <= uy1; ++ty) {
      float wy[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ry = ty - y0[c];
        wy[c] = (ry >= 0 && ry < ny[c]) ? (ry == 0 ? wyf[c] : (ry == ny[c] - 1 ? wyl[c] : wym[c])) : 0.f;
      }
      for (int j = 0; j <= jmax; ++j) {
        const int tx = ux0 + cb + j;
        float wl[4];
        bool any = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int rx = tx - x0[c];
          const float wx = (nx[c] != 0 && rx >= 0 && rx < nx[c]) ? (rx == 0 ? wxf[c] : (rx == nx[c] - 1 ? wxl[c] : 1.0f)) : 0.f;
          wl[c] = __fmul_rn(wy[c], wx);
          any = any || wl[c] != 0.f;
        }
        if (any) {
          if (FILL) {
            const size_t k = (size_t)first + total;
            ent_off[k] = (uint32_t)(n * hw + ty * fw + tx);
            ent_w[k] = make_float4(wl[0], wl[1], wl[2], wl[3]);
          }
          ++total;
        }
      }
    }
  }
  return total;
}

// LIST_QPB (8: 761 CTAs of 3 per SM on the MultiviewC grid; 16 quads at 2 CTAs per SM: 0.19 instead of 0.145 ms)
// consecutive quads per CTA, one thread per (iteration, quad) with the quad index fastest -- the records of one
// iteration are contiguous along a BEV row, so a warp's record loads coalesce (walking one quad per warp reads a different
// 32-byte record 0.8 MB apart in every iteration): count, per-quad exclusive scan in shared memory, segment table, fill.
// The quad's entries go into its own slot of `slot` entries (no global scan); a quad whose texels do not fit is marked
// LIST_OVERFLOW, gets no entries and is pooled by pool_quad_kernel<.., OVF> -- exact for any rig, no host synchronisation,
// static workspace.
#ifndef VFA_LIST_QPB
#define VFA_LIST_QPB 8
#endif
#ifndef VFA_LIST_BUILD_MINBLOCKS
#define VFA_LIST_BUILD_MINBLOCKS 3
#endif
constexpr int LIST_QPB = VFA_LIST_QPB;
constexpr int LIST_PAD = LIST_QPB + 1;                   // row stride of the counters: conflict-free both ways
__global__ void __launch_bounds__(256, VFA_LIST_BUILD_MINBLOCKS) qlist_build_kernel(const AggParams p, const TapRec* __restrict__ recs, int quads_x, int n_quads,
                                                          uint32_t slot, uint32_t* __restrict__ seg_off,
                                                          uint32_t* __restrict__ ent_off, float4* __restrict__ ent_w) {
  extern __shared__ uint32_t pre_s[];                    // [iters][LIST_PAD] counts -> exclusive prefixes, [LIST_QPB] totals
  const int VS = p.V * p.S, iters = VS * p.nl;
  uint32_t* const tot_s = pre_s + iters * LIST_PAD;
  const int q0 = blockIdx.x * LIST_QPB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int item = threadIdx.x; item < LIST_QPB * iters; item += blockDim.x) {
    const int ql = item % LIST_QPB, it = item / LIST_QPB, q = q0 + ql;
    pre_s[it * LIST_PAD + ql] = q < n_quads ? qlist_iter<false>(p, recs, quads_x, q, it, 0u, nullptr, nullptr) : 0u;
  }
  __syncthreads();
  for (int ql = warp; ql < LIST_QPB; ql += blockDim.x >> 5) {        // a warp scans the iterations of one quad
    uint32_t* c = pre_s + ql;
    uint32_t carry = 0;
    for (int i0 = 0; i0 < iters; i0 += 32) {
      const int i = i0 + lane;
      const uint32_t x = i < iters ? c[i * LIST_PAD] : 0u;
      uint32_t inc = x;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (i < iters) c[i * LIST_PAD] = carry + inc - x;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) tot_s[ql] = carry;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < LIST_QPB * (VS + 1); item += blockDim.x) {
    const int ql = item / (VS + 1), i = item % (VS + 1), q = q0 + ql;
    if (q >= n_quads) continue;
    const uint32_t base = (uint32_t)q * slot, tot = tot_s[ql];
    seg_off[(size_t)q * (VS + 1) + i] = i < VS ? base + pre_s[i * p.nl * LIST_PAD + ql] : (tot <= slot ? base + tot : LIST_OVERFLOW);
  }
  for (int item = threadIdx.x; item < LIST_QPB * iters; item += blockDim.x) {
    const int ql = item % LIST_QPB, it = item / LIST_QPB, q = q0 + ql;
    if (q >= n_quads || tot_s[ql] > slot) continue;
    qlist_iter<true>(p, recs, quads_x, q, it, (uint32_t)q * slot + pre_s[it * LIST_PAD + ql], ent_off, ent_w);
  }
}

template <bool MASK>
__global__ void __launch_bounds__(QWARPS * 32, VFA_LIST_MINBLOCKS) pool_list_kernel(const PoolArgs a) {
  __shared__ float out_s[QWARPS][4][CH];
  const AggParams& p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cy0 = (blockIdx.x / a.tiles_x) * (2 * QY) + 2 * (warp / QX);
  const int cx0 = (blockIdx.x % a.tiles_x) * (2 * QX) + 2 * (warp % QX);
  if (cy0 >= p.L || cx0 >= p.W) return;                 // no block-level synchronisation below
  const uint32_t* so = a.seg_off + (size_t)((cy0 >> 1) * a.quads_x + (cx0 >> 1)) * (p.V * p.S + 1);
  if (__ldg(so + p.V * p.S) == LIST_OVERFLOW) return;   // pooled by the completion pass (pool_quad_kernel<.., OVF>)
  const int bl = blockIdx.y;
  const int b = a.b0 + bl;

  float* const outw = &out_s[warp][0][lane * 8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<float4*>(outw + c * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(outw + c * CH + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  uint32_t e = __ldg(so);
  for (int v = 0; v < p.V; ++v) {
    for (int s = 0; s < p.S; ++s) {
      const uint32_t end = __ldg(++so);
      const int hw = p.sc[s].fh * p.sc[s].fw;
      const float* ybase = static_cast<const float*>(a.y[s]) + ((size_t)(bl * p.V + v) * p.nl) * hw * CH + lane * 4;
      float acc[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
      for (; e < end; e += LB) {
        uint32_t off[LB];
        float4 va[LB], vb[LB];
#pragma unroll
        for (int k = 0; k < LB; ++k) off[k] = e + k < end ? __ldg(a.ent_off + e + k) : 0xffffffffu;
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          if (off[k] != 0xffffffffu) {
            const float* tp = ybase + (size_t)off[k] * CH;
            va[k] = __ldg(reinterpret_cast<const float4*>(tp));
            vb[k] = __ldg(reinterpret_cast<const float4*>(tp + CH / 2));
          }
        }
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          if (off[k] != 0xffffffffu) {
            const float4 w = __ldg(a.ent_w + e + k);
            fma8(acc[0], w.x, va[k], vb[k]);
            fma8(acc[1], w.y, va[k], vb[k]);
            fma8(acc[2], w.z, va[k], vb[k]);
            fma8(acc[3], w.w, va[k], vb[k]);
          }
        }
      }
      e = end;
      // + bias, ReLU (vfa_op.py:123-124), sum over scales and views (vfanet.py:79, :82)
      const float4 bi0 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + lane * 4));
      const float4 bi1 = __ldg(reinterpret_cast<const float4*>(p.bias[s] + CH / 2 + lane * 4));
      const float bb[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 o0 = *reinterpret_cast<const float4*>(outw + c * CH);
        float4 o1 = *reinterpret_cast<const float4*>(outw + c * CH + 4);
        float t[8];
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] = acc[c][i] + bb[i];
          bits |= (t[i] > 0.f ? 1u : 0u) << i;
        }
        o0.x += fmaxf(t[0], 0.f); o0.y += fmaxf(t[1], 0.f); o0.z += fmaxf(t[2], 0.f); o0.w += fmaxf(t[3], 0.f);
        o1.x += fmaxf(t[4], 0.f); o1.y += fmaxf(t[5], 0.f); o1.z += fmaxf(t[6], 0.f); o1.w += fmaxf(t[7], 0.f);
        *reinterpret_cast<float4*>(outw + c * CH) = o0;
        *reinterpret_cast<float4*>(outw + c * CH + 4) = o1;
        if (MASK) {
          const int cy = cy0 + (c >> 1), cx = cx0 + (c & 1);
          const bool ok = cy < p.L && cx < p.W;
          store_mask_words(p.mask + (((size_t)b * p.V + v) * p.S + s) * (CH / 32) * p.LW + (ok ? cy * p.W + cx : 0),
                           (size_t)p.LW, lane, bits, ok);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cy = cy0 + (c >> 1), cx = cx0 + (c & 1);
    if (cy >= p.L || cx >= p.W) continue;
    const float4 o0 = *reinterpret_cast<const float4*>(outw + c * CH);
    const float4 o1 = *reinterpret_cast<const float4*>(outw + c * CH + 4);
    if (a.out_nhwc) {          // 512 contiguous bytes per warp store
      float* o = (a.out_mode == 3 ? owner_base(p.out, cy) : p.out) + ((size_t)b * p.LW + cy * p.W + cx) * CH + lane * 4;
      if (a.out_mode == 0) {
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + CH / 2) = o1;
      } else {
        red_add_v4(o, o0, a.out_mode);
        red_add_v4(o + CH / 2, o1, a.out_mode);
      }
      continue;
    }
    float* o = p.out + (size_t)b * CH * p.LW + cy * p.W + cx;
    const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) o[(size_t)chan_of(lane, i) * p.LW] = ov[i];
  }
}

}  // namespace fside

using namespace fside;


// ---- which (texel tile, layer) products are ever pooled ------------------------------------------------------------
// A camera looking at the field sees the voxel grid in part of its image only (sky, ground beyond the grid), and every
// height layer projects to a different band: on the three rigs 54 % / 42 % / 25 % of the (texel, layer) products are
// pooled by some visible box, 68 % / 64 % / 44 % of the (256-texel tile, layer) pairs.  The tap records give a bitmap of
// covered texels per (scale, view, layer) plane; from it one byte per (tile, layer) tells ygemm_kernel (and the backward's
// dFeature / dWeight products) which tiles to skip altogether.  Zero-weight taps are marked too (superset of what the
// pooling kernels fetch).
// (CoverMap / make_cover_map: vfa_pool.cuh -- the chunk-list builder of vfa_pool_tile.cu fills the same bitmap)

__global__ void __launch_bounds__(256) cover_mark_kernel(AggParams p, const TapRec* __restrict__ recs, CoverMap cm,
                                                         uint32_t* __restrict__ bits) {
  const long long total = (long long)p.V * p.S * p.nl * p.LW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const TapRec r = recs[idx];
    const int nx = r.nxy & 0xffff, ny = r.nxy >> 16;
    if (nx == 0) continue;
    const int n = (int)((idx / p.LW) % p.nl);
    const int s = (int)((idx / ((long long)p.LW * p.nl)) % p.S);
    const int v = (int)(idx / ((long long)p.LW * p.nl * p.S));
    const int fw = s == 0 ? cm.fw[0] : (s == 1 ? cm.fw[1] : cm.fw[2]);
    const int words = s == 0 ? cm.words[0] : (s == 1 ? cm.words[1] : cm.words[2]);
    const int wbase = s == 0 ? cm.word_base[0] : (s == 1 ? cm.word_base[1] : cm.word_base[2]);
    uint32_t* plane = bits + wbase + (size_t)(v * p.nl + n) * words;
    for (int ty = 0; ty < ny; ++ty) {
      const int t0 = ((r.xy >> 16) + ty) * fw + (r.xy & 0xffff), t1 = t0 + nx - 1;      // inclusive texel range of the row
      for (int w = t0 >> 5; w <= t1 >> 5; ++w) {
        const int lo = max(t0, w << 5) & 31, hi = min(t1, (w << 5) + 31) & 31;
        const uint32_t m = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
        if ((plane[w] & m) != m) atomicOr(plane + w, m);
      }
    }
  }
}

struct NeedArgs {
  CoverMap cm;
  int rows[VFA_MAX_SCALES];
  int tile_begin[VFA_MAX_SCALES + 1];
  int V, nl, S;
};

// need[tile * nl + n] = any texel row of the 256-row tile (rows = chunk-relative plane * hw + texel) is covered in layer n
__global__ void __launch_bounds__(256) tile_need_kernel(NeedArgs q, const uint32_t* __restrict__ bits, uint8_t* __restrict__ need) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_tiles = q.S == 1 ? q.tile_begin[1] : (q.S == 2 ? q.tile_begin[2] : q.tile_begin[3]);   // no dynamic indexing
  if (idx >= n_tiles * q.nl) return;
  const int tile = idx / q.nl, n = idx % q.nl;
  const int s = (q.S > 2 && tile >= q.tile_begin[2]) ? 2 : ((q.S > 1 && tile >= q.tile_begin[1]) ? 1 : 0);
  const int tb = s == 0 ? 0 : (s == 1 ? q.tile_begin[1] : q.tile_begin[2]);
  const int rows = s == 0 ? q.rows[0] : (s == 1 ? q.rows[1] : q.rows[2]);
  const int hw = s == 0 ? q.cm.hw[0] : (s == 1 ? q.cm.hw[1] : q.cm.hw[2]);
  const int words = s == 0 ? q.cm.words[0] : (s == 1 ? q.cm.words[1] : q.cm.words[2]);
  const int wbase = s == 0 ? q.cm.word_base[0] : (s == 1 ? q.cm.word_base[1] : q.cm.word_base[2]);
  const int r0 = (tile - tb) * (2 * TILE_M), r1 = min(rows, r0 + 2 * TILE_M);
  uint32_t any = 0;
  for (int r = r0; r < r1 && !any; ++r) {
    const int v = (r / hw) % q.V, t = r % hw;
    any = (__ldg(bits + wbase + (size_t)(v * q.nl + n) * words + (t >> 5)) >> (t & 31)) & 1u;
  }
  need[idx] = (uint8_t)any;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// bytes for the bitmap + the need bytes of a chunk of `frames` frames
// layout of the coverage workspace: [bitmap][need bytes of a chunk][cnt per plane list][row lists][unit table][unit count]
struct CoverLayout {
  size_t off_need, off_cnt, off_rowlist, off_tab, off_nunits, total;
  int planes, max_units;
};
static CoverLayout cover_layout(const AggParams& p, int frames) {
  const CoverMap cm = make_cover_map(p);
  CoverLayout L;
  size_t tiles = 0, units = 0;
  for (int s = 0; s < p.S; ++s) {
    tiles += ((size_t)frames * p.V * cm.hw[s] + 2 * TILE_M - 1) / (2 * TILE_M);
    units += (size_t)p.V * p.nl * ((cm.hw[s] + 2 * TILE_M - 1) / (2 * TILE_M));
  }
  L.planes = p.S * p.V * p.nl;
  L.max_units = (int)units;
  size_t o = align256((size_t)cm.total_words * sizeof(uint32_t));
  L.off_need = o;    o += align256(tiles * p.nl);
  L.off_cnt = o;     o += align256((size_t)L.planes * sizeof(int));
  L.off_rowlist = o; o += align256((size_t)cm.total_words * 32 * sizeof(int));
  L.off_tab = o;     o += align256(units * 16);                 // RowUnit entries (16 bytes each)
  L.off_nunits = o;  o += 256;
  L.total = o;
  return L;
}

size_t fside_cover_bytes(const AggParams& p, int frames) { return cover_layout(p, frames).total; }

// bitmap of covered texels from the tap records (once per call)
int launch_cover_mark(const AggParams& p, const TapRec* recs, void* cover_ws, cudaStream_t st) {
  const CoverMap cm = make_cover_map(p);
  VFA_CUDA(cudaMemsetAsync(cover_ws, 0, (size_t)cm.total_words * sizeof(uint32_t), st));
  cover_mark_kernel<<<148 * 8, 256, 0, st>>>(p, recs, cm, reinterpret_cast<uint32_t*>(cover_ws));
  VFA_LAUNCH_CHECK("cover_mark_kernel");
  return VFA_OK;
}

// need bytes of the 256-row tiles of a chunk of `frames` frames (tiles of the scales concatenated); returns the pointer
// run = false: only return the pointer (the bytes of an earlier call with the same shapes are still in the workspace).
// VFA_FSIDE_NO_SKIP=1 marks every tile as needed (measurement of the GEMMs without the skipping).
int launch_tile_need(const AggParams& p, void* cover_ws, int frames, const uint8_t** need_out, cudaStream_t st, bool run) {
  NeedArgs q;
  q.cm = make_cover_map(p);
  q.V = p.V;
  q.nl = p.nl;
  q.S = p.S;
  q.tile_begin[0] = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    q.rows[s] = frames * p.V * q.cm.hw[s];
    if (s < p.S) q.tile_begin[s + 1] = q.tile_begin[s] + (q.rows[s] + 2 * TILE_M - 1) / (2 * TILE_M);
    else if (s + 1 <= VFA_MAX_SCALES) q.tile_begin[s + 1] = q.tile_begin[s];
  }
  uint8_t* need = reinterpret_cast<uint8_t*>(cover_ws) + cover_layout(p, frames).off_need;
  const int total = q.tile_begin[p.S] * p.nl;
  *need_out = need;
  if (!run) return VFA_OK;
  if (runtime_config().fside_no_skip != 0) {
    VFA_CUDA(cudaMemsetAsync(need, 1, (size_t)total, st));
    return VFA_OK;
  }
  tile_need_kernel<<<(total + 255) / 256, 256, 0, st>>>(q, reinterpret_cast<const uint32_t*>(cover_ws), need);
  VFA_LAUNCH_CHECK("tile_need_kernel");
  return VFA_OK;
}

// ---- row-compacted image-plane GEMM (forward) ----------------------------------------------------------------------------
// Whole 256-row tiles still carry uncovered texels (70 / 66 / 46 % of the tiles are needed, but only 54 / 42 / 25 % of the
// texel rows).  The forward therefore multiplies COMPACTED rows: per (scale, view, layer) the covered texels are listed
// (rowlist_kernel: popcount + block scan over the bitmap, sorted), a unit = 256 consecutive entries of one list for one
// frame, and the producers / the epilogue go through the list (indirect A rows, indirect Y rows).  Units are uniform
// (8 stages, one accumulator), so dealing them round-robin over the resident CTA pairs is balanced; Y keeps its layout.
struct __align__(16) RowUnit {
  int s, vn, j, cnt;          // scale, view * nl + layer, chunk of 256 list entries, entries in the list
};
static_assert(sizeof(RowUnit) == 16, "cover_layout sizes the unit table with 16-byte entries");

__global__ void __launch_bounds__(256) rowlist_kernel(CoverMap cm, int planes_per_scale, const uint32_t* __restrict__ bits,
                                                      int* __restrict__ rowlist, int* __restrict__ cnt) {
  __shared__ int warp_sums[8];
  __shared__ int running_s;
  const int pl = blockIdx.x, s = pl / planes_per_scale, vn = pl % planes_per_scale;
  const int words = s == 0 ? cm.words[0] : (s == 1 ? cm.words[1] : cm.words[2]);
  const int base = (s == 0 ? cm.word_base[0] : (s == 1 ? cm.word_base[1] : cm.word_base[2])) + vn * words;
  int* out = rowlist + (size_t)base * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running_s = 0;
  __syncthreads();
  for (int w0 = 0; w0 < words; w0 += 256) {
    const int w = w0 + threadIdx.x;
    uint32_t m = w < words ? __ldg(bits + base + w) : 0u;
    const int c = __popc(m);
    int v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    int before = running_s;
    for (int k = 0; k < warp; ++k) before += warp_sums[k];
    int pos = before + v - c;
    while (m) {
      const int b = __ffs(m) - 1;
      out[pos++] = w * 32 + b;
      m &= m - 1;
    }
    __syncthreads();
    if (threadIdx.x == 255) running_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[pl] = running_s;
}

// per-frame unit table: for every plane list its ceil(cnt / 256) chunks (single block)
__global__ void __launch_bounds__(256) unit_table_kernel(const int* __restrict__ cnt, int planes, int planes_per_scale,
                                                         RowUnit* __restrict__ tab, int* __restrict__ n_units_frame) {
  __shared__ int warp_sums[8];
  __shared__ int running_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running_s = 0;
  __syncthreads();
  for (int p0 = 0; p0 < planes; p0 += 256) {
    const int pl = p0 + threadIdx.x;
    const int n = pl < planes ? cnt[pl] : 0;
    const int c = (n + 2 * TILE_M - 1) / (2 * TILE_M);
    int v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    int before = running_s;
    for (int k = 0; k < warp; ++k) before += warp_sums[k];
    int pos = before + v - c;
    for (int j = 0; j < c; ++j) {
      RowUnit u;
      u.s = pl / planes_per_scale;
      u.vn = pl % planes_per_scale;
      u.j = j;
      u.cnt = n;
      tab[pos + j] = u;
    }
    __syncthreads();
    if (threadIdx.x == 255) running_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_units_frame = running_s;
}

struct YCompactArgs {
  const uint8_t* feats[VFA_MAX_SCALES];   // [plane][texel][256] of this chunk (fp32 or bf16)
  float* y[VFA_MAX_SCALES];               // [plane][layer][texel][256]
  const uint8_t* wprep[VFA_MAX_SCALES];
  int hw[VFA_MAX_SCALES];
  int rl_base[VFA_MAX_SCALES];            // first rowlist entry of scale s; lists of plane vn follow `rl_stride` apart
  int rl_stride[VFA_MAX_SCALES];
  const int* rowlist;
  const RowUnit* tab;
  const int* n_units_frame;
  int nb, V, nl;
};

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1) ygemm_compact_kernel(const YCompactArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nuf = __ldg(a.n_units_frame);
  const int n_units = nuf * a.nb;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&tail->full[i], NUM_PRODUCER_WARPS + 1);
      mbar_init(&tail->empty[i], 1);
      mbar_init(&tail->peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tail->acc_full[i], 1);
      mbar_init(&tail->acc_empty[i], 2 * NUM_EPILOGUE_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ================= weight loader: the 8 K-chunk slabs of the unit's (scale, layer) =================
    if (lane == 0) {
      int it = 0;
      for (int u = cluster; u < n_units; u += n_clusters) {
        const RowUnit w = a.tab[u % nuf];
        const uint8_t* wp = pick(a.wprep, w.s) + (size_t)(w.vn % a.nl) * CHUNKS * (2 * B_BYTES);
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
          uint8_t* dst = smem + (size_t)st * STAGE_BYTES + 2 * A_BYTES;
          mbar_arrive_expect_tx(&tail->full[st], 2 * B_LOCAL_BYTES);
          const uint8_t* src = wp + (size_t)cc * (2 * B_BYTES);
          bulk_g2s(dst, src + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
          bulk_g2s(dst + B_LOCAL_BYTES, src + B_BYTES + cta_rank * B_LOCAL_BYTES, B_LOCAL_BYTES, &tail->full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank != 0) {
      int it = 0;
      for (int u = cluster; u < n_units; u += n_clusters)
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_arrive_remote(&tail->peer_full[st], 0);
        }
    } else if (lane == 0) {
      // ================= MMA issuer (pair leader): one accumulator slot per unit =================
      int it = 0, ln = 0;
      for (int u = cluster; u < n_units; u += n_clusters, ++ln) {
        const int slot = ln & 1;
        mbar_wait_cluster(&tail->acc_empty[slot], ((ln >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)slot * CH;
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_wait_cluster(&tail->peer_full[st], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
          const uint64_t b_hi = make_desc(sa + 2 * A_BYTES), b_lo = make_desc(sa + 2 * A_BYTES + B_LOCAL_BYTES);
#pragma unroll
          for (int ks = 0; ks < KCH / 8; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
            tc_mma_tf32_t<true>(d_tmem, a_lo + adv, b_hi + adv, IDESC, (cc | ks) ? 1u : 0u);
            tc_mma_tf32_t<true>(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
            tc_mma_tf32_t<true>(d_tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
          }
          tc_commit_t<true>(&tail->empty[st]);
        }
        tc_commit_t<true>(&tail->acc_full[slot]);
      }
    }
  } else if (warp >= FIRST_PRODUCER_WARP && warp < FIRST_EPILOGUE_WARP) {
    // ================= A producers: listed texel rows -> tf32 hi / lo operand tiles =================
    constexpr int ES = BF16 ? 2 : 4;
    const int pw = warp - FIRST_PRODUCER_WARP;
    const int j = lane & 7;
    const uint8_t* src[ITEMS];
    uint32_t off[ITEMS];
    bool ok[ITEMS];
    int nidx[ITEMS];                                   // list entries (texels) of this thread's rows in the NEXT unit
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) off[i] = swz((uint32_t)(32 * i + 4 * pw + (lane >> 3)), (uint32_t)j);
    auto fetch_rows = [&](int u, int(&idx)[ITEMS]) {   // texel of each of this thread's rows, -1 beyond the list
      const RowUnit w = a.tab[u % nuf];
      const int* rl = a.rowlist + pick(a.rl_base, w.s) + (size_t)w.vn * pick(a.rl_stride, w.s);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = w.j * (2 * TILE_M) + (int)cta_rank * TILE_M + 32 * i + 4 * pw + (lane >> 3);
        idx[i] = e < w.cnt ? __ldg(rl + e) : -1;
      }
    };
    auto bind = [&](int u, const int(&idx)[ITEMS]) {
      const RowUnit w = a.tab[u % nuf];
      const int plane = (u / nuf) * a.V + w.vn / a.nl;
      const uint8_t* f = pick(a.feats, w.s) + ((size_t)plane * pick(a.hw, w.s)) * CH * ES + j * 4 * ES;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        ok[i] = idx[i] >= 0;
        src[i] = f + (size_t)(ok[i] ? idx[i] : 0) * CH * ES;
      }
    };
    auto load_stage = [&](float4(&v)[ITEMS], int cc) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[i]) v[i] = load_feat4<BF16>(src[i] + (size_t)cc * KCH * ES);
      }
    };
    float4 cur[ITEMS], nxt[ITEMS];
    int u = cluster;
    if (u < n_units) {
      int idx0[ITEMS];
      fetch_rows(u, idx0);
      bind(u, idx0);
      load_stage(cur, 0);
      if (u + n_clusters < n_units) fetch_rows(u + n_clusters, nidx);
    }
    int it = 0;
    for (; u < n_units; u += n_clusters) {
      for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
        const int st = it % STAGES;
        bool more = true;
        if (cc + 1 < CHUNKS) {
          load_stage(nxt, cc + 1);
        } else if (u + n_clusters < n_units) {           // first stage of the next unit; then look one unit further
          bind(u + n_clusters, nidx);
          load_stage(nxt, 0);
          if (u + 2 * n_clusters < n_units) fetch_rows(u + 2 * n_clusters, nidx);
        } else {
          more = false;
        }
        mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
        uint8_t* a_hi = smem + (size_t)st * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) store_split(a_hi, off[i], cur[i]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->full[st]);
        if (more) {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) cur[i] = nxt[i];
        }
      }
    }
  } else if (warp >= FIRST_EPILOGUE_WARP) {
    // ================= epilogue: accumulator -> the listed rows of Y[plane][layer] =================
    const int e = warp - FIRST_EPILOGUE_WARP;
    const int quarter = warp & 3;
    const int col_begin = (e >> 2) * (CH / 2);
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* patch = smem + (size_t)STAGES * STAGE_BYTES + TAIL_BYTES + (size_t)e * PATCH_BYTES;
    const int sub = lane >> 3, chunk = lane & 7;
    int ln = 0;
    for (int u = cluster; u < n_units; u += n_clusters, ++ln) {
      const RowUnit w = a.tab[u % nuf];
      const int hw = pick(a.hw, w.s);
      const int* rl = a.rowlist + pick(a.rl_base, w.s) + (size_t)w.vn * pick(a.rl_stride, w.s);
      const int plane = (u / nuf) * a.V + w.vn / a.nl;
      // 1 KB row of Y of patch row 4*rr + sub; ~0u = beyond the list
      uint32_t yrow[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int en = w.j * (2 * TILE_M) + (int)cta_rank * TILE_M + quarter * 32 + 4 * rr + sub;
        yrow[rr] = en < w.cnt ? (uint32_t)((plane * a.nl + w.vn % a.nl) * hw + __ldg(rl + en)) : 0xffffffffu;
      }
      const int slot = ln & 1;
      mbar_wait_sleep(&tail->acc_full[slot], (ln >> 1) & 1);
      tc_fence_after();
      float* dst = pick(a.y, w.s) + col_begin + chunk * 4;
#pragma unroll 1
      for (int c0 = 0; c0 < CH / 2; c0 += 32) {
        float v[32];
        tc_ld32(lane_addr + (uint32_t)(slot * CH + col_begin + c0), v);
        tc_wait_ld();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(patch + lane * 128 + ((i ^ (lane & 7)) << 4)) =
              make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int prow = 4 * rr + sub;
          const float4 q = *reinterpret_cast<const float4*>(patch + prow * 128 + ((chunk ^ (prow & 7)) << 4));
          if (yrow[rr] != 0xffffffffu) *reinterpret_cast<float4*>(dst + (size_t)yrow[rr] * CH + c0) = q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote(&tail->acc_empty[slot], 0);
        else mbar_arrive(&tail->acc_empty[slot]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ---- bf16 tensor-core variant (VFA_FLAG_BF16_MMA) -------------------------------------------------------------------------
// The same row-compacted image-plane GEMM with bf16 operands: ONE tcgen05.mma kind::f16 pass (fp32 accumulate in TMEM)
// instead of three TF32 passes, and Y stored in bf16 -- a sixth of the tensor-core work, half the Y bytes written here and
// read by the pooling.  Not the parity path: operands are rounded to 8 bits of mantissa (tolerance stated and tested in
// tests/test_gpu_frame_parity.py::test_bf16_mma_variant).  A stage = 64 K elements: A tile 128 rows x 128 B, this CTA's half
// of the weight slab 128 rows x 128 B (SWIZZLE_128B both); 4 stages per unit, 4 MMAs (K = 16) per stage.
namespace bfmma {
constexpr int KE = 64;                                   // K elements per stage (128 bytes of bf16)
constexpr int A_BYTES = TILE_M * 128;                    // 16 KB
constexpr int B_BYTES = CH * 128;                        // 32 KB slab (N = 256); each CTA of the pair stages half
constexpr int B_LOCAL_BYTES = B_BYTES / 2;
constexpr int STAGE_BYTES = A_BYTES + B_LOCAL_BYTES;     // 32 KB
constexpr int STAGES = 5;
constexpr int CHUNKS = CH / KE;                          // 4 stages per unit (one layer of 256 listed rows)
constexpr int ITEMS = TILE_M * 8 / (NUM_PRODUCER_WARPS * 32);      // 16-byte smem items per producer thread and stage
constexpr int PATCH_BYTES = 32 * 64;                     // per epilogue warp: 32 rows x 32 bf16
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 512 + (size_t)NUM_EPILOGUE_WARPS * PATCH_BYTES;
constexpr uint32_t IDESC = make_idesc_bf16(CH, 2 * TILE_M);
struct __align__(16) Tail {
  unsigned long long full[STAGES], empty[STAGES], peer_full[STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Tail) <= 512, "");

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // low half = a
  return r;
}

// collapse.weight [C, C*nl] (column c*nl + n) -> per K chunk kc = n * CHUNKS + c / 64 a 32 KB block [256 rows x 128 B], bf16
__global__ void __launch_bounds__(256) prep_weight_bf16_kernel(const float* __restrict__ w, uint8_t* __restrict__ wp, int nl) {
  const int K = CH * nl;
  const long long total = (long long)CH * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % KE);
    const int o = (int)((idx / KE) % CH);
    const int kc = (int)(idx / ((long long)KE * CH));
    const int n = kc / CHUNKS, c = (kc % CHUNKS) * KE + kk;
    const __nv_bfloat16 v = __float2bfloat16_rn(w[(long long)o * K + (long long)c * nl + n]);
    *reinterpret_cast<__nv_bfloat16*>(wp + (long long)kc * B_BYTES + swz(o, kk >> 3) + (kk & 7) * 2) = v;
  }
}

template <bool BF16>      // BF16: the feature maps are stored in bf16; else fp32, rounded to bf16 by the producers
__global__ void __launch_bounds__(THREADS, 1) ygemm_compact_bf16_kernel(const YCompactArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Tail* tail = reinterpret_cast<Tail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nuf = __ldg(a.n_units_frame);
  const int n_units = nuf * a.nb;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&tail->full[i], NUM_PRODUCER_WARPS + 1);
      mbar_init(&tail->empty[i], 1);
      mbar_init(&tail->peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tail->acc_full[i], 1);
      mbar_init(&tail->acc_empty[i], 2 * NUM_EPILOGUE_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ================= weight loader =================
    if (lane == 0) {
      int it = 0;
      for (int u = cluster; u < n_units; u += n_clusters) {
        const RowUnit w = a.tab[u % nuf];
        const uint8_t* wp = pick(a.wprep, w.s) + (size_t)(w.vn % a.nl) * CHUNKS * B_BYTES;
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&tail->full[st], B_LOCAL_BYTES);
          bulk_g2s(smem + (size_t)st * STAGE_BYTES + A_BYTES, wp + (size_t)cc * B_BYTES + cta_rank * B_LOCAL_BYTES,
                   B_LOCAL_BYTES, &tail->full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank != 0) {
      int it = 0;
      for (int u = cluster; u < n_units; u += n_clusters)
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_arrive_remote(&tail->peer_full[st], 0);
        }
    } else if (lane == 0) {
      // ================= MMA issuer (pair leader) =================
      int it = 0, ln = 0;
      for (int u = cluster; u < n_units; u += n_clusters, ++ln) {
        const int slot = ln & 1;
        mbar_wait_cluster(&tail->acc_empty[slot], ((ln >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)slot * CH;
        for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_wait_cluster(&tail->peer_full[st], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
          const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
          for (int ks = 0; ks < KE / 16; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
            tc_mma_f16_pair(d_tmem, da + adv, db + adv, IDESC, (cc | ks) ? 1u : 0u);
          }
          tc_commit_t<true>(&tail->empty[st]);
        }
        tc_commit_t<true>(&tail->acc_full[slot]);
      }
    }
  } else if (warp >= FIRST_PRODUCER_WARP && warp < FIRST_EPILOGUE_WARP) {
    // ================= A producers: listed texel rows -> bf16 operand tile =================
    const int pw = warp - FIRST_PRODUCER_WARP;
    const int j = lane & 7;                          // 16-byte chunk (8 K elements) of the 128-byte tile row
    const uint8_t* src[ITEMS];
    uint32_t off[ITEMS];
    bool ok[ITEMS];
    int nidx[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) off[i] = swz((uint32_t)(32 * i + 4 * pw + (lane >> 3)), (uint32_t)j);
    auto fetch_rows = [&](int u, int(&idx)[ITEMS]) {
      const RowUnit w = a.tab[u % nuf];
      const int* rl = a.rowlist + pick(a.rl_base, w.s) + (size_t)w.vn * pick(a.rl_stride, w.s);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int e = w.j * (2 * TILE_M) + (int)cta_rank * TILE_M + 32 * i + 4 * pw + (lane >> 3);
        idx[i] = e < w.cnt ? __ldg(rl + e) : -1;
      }
    };
    constexpr int ES = BF16 ? 2 : 4;
    auto bind = [&](int u, const int(&idx)[ITEMS]) {
      const RowUnit w = a.tab[u % nuf];
      const int plane = (u / nuf) * a.V + w.vn / a.nl;
      const uint8_t* f = pick(a.feats, w.s) + ((size_t)plane * pick(a.hw, w.s)) * CH * ES + j * 8 * ES;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        ok[i] = idx[i] >= 0;
        src[i] = f + (size_t)(ok[i] ? idx[i] : 0) * CH * ES;
      }
    };
    auto load_stage = [&](uint4(&v)[ITEMS], int cc) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        v[i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[i]) {
          const uint8_t* sp = src[i] + (size_t)cc * KE * ES;
          if (BF16) {
            v[i] = __ldg(reinterpret_cast<const uint4*>(sp));
          } else {
            const float4 lo = __ldg(reinterpret_cast<const float4*>(sp)), hi = __ldg(reinterpret_cast<const float4*>(sp) + 1);
            v[i] = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
          }
        }
      }
    };
    uint4 cur[ITEMS], nxt[ITEMS];
    int u = cluster;
    if (u < n_units) {
      int idx0[ITEMS];
      fetch_rows(u, idx0);
      bind(u, idx0);
      load_stage(cur, 0);
      if (u + n_clusters < n_units) fetch_rows(u + n_clusters, nidx);
    }
    int it = 0;
    for (; u < n_units; u += n_clusters) {
      for (int cc = 0; cc < CHUNKS; ++cc, ++it) {
        const int st = it % STAGES;
        bool more = true;
        if (cc + 1 < CHUNKS) {
          load_stage(nxt, cc + 1);
        } else if (u + n_clusters < n_units) {
          bind(u + n_clusters, nidx);
          load_stage(nxt, 0);
          if (u + 2 * n_clusters < n_units) fetch_rows(u + 2 * n_clusters, nidx);
        } else {
          more = false;
        }
        mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
        uint8_t* at = smem + (size_t)st * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) *reinterpret_cast<uint4*>(at + off[i]) = cur[i];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->full[st]);
        if (more) {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) cur[i] = nxt[i];
        }
      }
    }
  } else if (warp >= FIRST_EPILOGUE_WARP) {
    // ================= epilogue: accumulator -> bf16 -> the listed rows of Y[plane][layer] =================
    const int e = warp - FIRST_EPILOGUE_WARP;
    const int quarter = warp & 3;
    const int col_begin = (e >> 2) * (CH / 2);
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* patch = smem + (size_t)STAGES * STAGE_BYTES + 512 + (size_t)e * PATCH_BYTES;
    const int sub = lane >> 2, chunk = lane & 3;      // store phase: patch row 8 * rr + sub, 16-byte chunk of its 64 bytes
    int ln = 0;
    for (int u = cluster; u < n_units; u += n_clusters, ++ln) {
      const RowUnit w = a.tab[u % nuf];
      const int hw = pick(a.hw, w.s);
      const int* rl = a.rowlist + pick(a.rl_base, w.s) + (size_t)w.vn * pick(a.rl_stride, w.s);
      const int plane = (u / nuf) * a.V + w.vn / a.nl;
      uint32_t yrow[4];                               // Y row of patch row 8 * rr + sub; ~0u = beyond the list
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int en = w.j * (2 * TILE_M) + (int)cta_rank * TILE_M + quarter * 32 + 8 * rr + sub;
        yrow[rr] = en < w.cnt ? (uint32_t)((plane * a.nl + w.vn % a.nl) * hw + __ldg(rl + en)) : 0xffffffffu;
      }
      const int slot = ln & 1;
      mbar_wait_sleep(&tail->acc_full[slot], (ln >> 1) & 1);
      tc_fence_after();
      uint8_t* dst = reinterpret_cast<uint8_t*>(pick(a.y, w.s)) + (size_t)(col_begin) * 2 + chunk * 16;
#pragma unroll 1
      for (int c0 = 0; c0 < CH / 2; c0 += 32) {
        float v[32];
        tc_ld32(lane_addr + (uint32_t)(slot * CH + col_begin + c0), v);
        tc_wait_ld();
        __syncwarp();                                 // the previous chunk has left the patch
#pragma unroll
        for (int i = 0; i < 4; ++i)                   // row `lane`: 4 x 16 bytes, chunk XOR-swizzled with the row pair
          *reinterpret_cast<uint4*>(patch + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                         pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int prow = 8 * rr + sub;
          const uint4 q = *reinterpret_cast<const uint4*>(patch + prow * 64 + ((chunk ^ ((prow >> 1) & 3)) << 4));
          if (yrow[rr] != 0xffffffffu) *reinterpret_cast<uint4*>(dst + ((size_t)yrow[rr] * CH + c0) * 2) = q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote(&tail->acc_empty[slot], 0);
        else mbar_arrive(&tail->acc_empty[slot]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}
}  // namespace bfmma

// One persistent launch over all cluster tiles of `g` (tile_begin must be filled for g.S scales).
static int launch_ygemm(YGemmArgs& g, int mode, bool bf16, cudaStream_t st) {
  // function attributes are per device: set on every launch (a host-side table update, no device work)
  VFA_CUDA(cudaFuncSetAttribute(ygemm_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  VFA_CUDA(cudaFuncSetAttribute(ygemm_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  VFA_CUDA(cudaFuncSetAttribute(ygemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int resident_clusters = device_cache_get(DC_YGEMM_CLUSTERS);      // CTA pairs the device can hold at once (1 CTA per SM)
  if (resident_clusters == 0) {
    cfg.gridDim = dim3(2 * 148);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ygemm_kernel<false, 0>, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = 64;
    }
    resident_clusters = n;
    device_cache_set(DC_YGEMM_CLUSTERS, n);
  }
  g.n_tiles = g.tile_begin[g.S];
  if (g.n_tiles <= 0) return VFA_OK;
  cfg.gridDim = dim3(2 * (g.n_tiles < resident_clusters ? g.n_tiles : resident_clusters));
  if (mode == 1)
    VFA_CUDA(cudaLaunchKernelEx(&cfg, ygemm_kernel<false, 1>, g));
  else if (bf16)
    VFA_CUDA(cudaLaunchKernelEx(&cfg, ygemm_kernel<true, 0>, g));
  else
    VFA_CUDA(cudaLaunchKernelEx(&cfg, ygemm_kernel<false, 0>, g));
  VFA_LAUNCH_CHECK("ygemm_kernel");
  return VFA_OK;
}

// out_s[rows_s, 256] = sum over the nl*256 columns of a_s[rows_s, nl*256] against the TRANSPOSED prepared weights
// (prep_weight_umma_T_kernel): the dFeature product of the backward.  3xTF32, layer partials added in fp32.
// `rows` = texel rows per scale (the tiling the need bytes were built for); scales with out[s] == nullptr are skipped.
int launch_ygemm_accum(const float* const* a_rows, float* const* out, const uint8_t* const* wprep_t, const int* rows,
                       int nl, int S, const uint8_t* need, cudaStream_t st) {
  YGemmArgs g;
  g.nl = nl;
  g.S = S;
  g.need = need;
  g.tile_begin[0] = 0;
  int tile0 = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    const int ss = s < S ? s : 0;
    g.feats[s] = reinterpret_cast<const uint8_t*>(a_rows[ss]);
    g.y[s] = out[ss];
    g.wprep[s] = wprep_t[ss];
    g.rows[s] = out[ss] != nullptr ? rows[ss] : 0;
    g.hw[s] = 1;
    g.need_tile0[s] = tile0;
    if (s < S) {
      tile0 += (rows[s] + 2 * TILE_M - 1) / (2 * TILE_M);
      g.tile_begin[s + 1] = g.tile_begin[s] + (g.rows[s] + 2 * TILE_M - 1) / (2 * TILE_M);
    }
  }
  return launch_ygemm(g, 1, false, st);
}

// covered-texel lists and the per-frame unit table from the bitmap (once per call, after launch_cover_mark)
static int launch_rowlists(const AggParams& p, void* cover_ws, int frames, cudaStream_t st) {
  const CoverMap cm = make_cover_map(p);
  const CoverLayout L = cover_layout(p, frames);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(cover_ws);
  rowlist_kernel<<<L.planes, 256, 0, st>>>(cm, p.V * p.nl, reinterpret_cast<const uint32_t*>(cover_ws),
                                           reinterpret_cast<int*>(w8 + L.off_rowlist), reinterpret_cast<int*>(w8 + L.off_cnt));
  VFA_LAUNCH_CHECK("rowlist_kernel");
  unit_table_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const int*>(w8 + L.off_cnt), L.planes, p.V * p.nl,
                                       reinterpret_cast<RowUnit*>(w8 + L.off_tab), reinterpret_cast<int*>(w8 + L.off_nunits));
  VFA_LAUNCH_CHECK("unit_table_kernel");
  return VFA_OK;
}

static int launch_ygemm_compact(const AggParams& p, const YGemmArgs& g, void* cover_ws, int frames_layout, int nb, bool bf16,
                                cudaStream_t st) {
  const bool mma_bf16 = p.y_bf16 != 0;
  VFA_CUDA(cudaFuncSetAttribute(ygemm_compact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  VFA_CUDA(cudaFuncSetAttribute(ygemm_compact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  if (mma_bf16) {
    VFA_CUDA(cudaFuncSetAttribute(bfmma::ygemm_compact_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)bfmma::SMEM_BYTES));
    VFA_CUDA(cudaFuncSetAttribute(bfmma::ygemm_compact_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)bfmma::SMEM_BYTES));
  }
  const CoverMap cm = make_cover_map(p);
  const CoverLayout L = cover_layout(p, frames_layout);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(cover_ws);
  YCompactArgs a;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    a.feats[s] = g.feats[s];
    a.y[s] = g.y[s];
    a.wprep[s] = g.wprep[s];
    a.hw[s] = cm.hw[s];
    a.rl_base[s] = cm.word_base[s] * 32;
    a.rl_stride[s] = cm.words[s] * 32;
  }
  a.rowlist = reinterpret_cast<const int*>(w8 + L.off_rowlist);
  a.tab = reinterpret_cast<const RowUnit*>(w8 + L.off_tab);
  a.n_units_frame = reinterpret_cast<const int*>(w8 + L.off_nunits);
  a.nb = nb;
  a.V = p.V;
  a.nl = p.nl;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int resident_clusters = device_cache_get(DC_YGEMM_COMPACT_CLUSTERS);
  if (resident_clusters == 0) {
    cfg.gridDim = dim3(2 * 148);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ygemm_compact_kernel<false>, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = 64;
    }
    resident_clusters = n;
    device_cache_set(DC_YGEMM_COMPACT_CLUSTERS, n);
  }
  long long max_units = (long long)L.max_units * nb;     // the real count lives on the device; this bounds the grid
  cfg.gridDim = dim3(2 * (int)(max_units < resident_clusters ? max_units : resident_clusters));
  if (mma_bf16) {
    cfg.dynamicSmemBytes = bfmma::SMEM_BYTES;
    if (bf16)
      VFA_CUDA(cudaLaunchKernelEx(&cfg, bfmma::ygemm_compact_bf16_kernel<true>, a));
    else
      VFA_CUDA(cudaLaunchKernelEx(&cfg, bfmma::ygemm_compact_bf16_kernel<false>, a));
    VFA_LAUNCH_CHECK("ygemm_compact_bf16_kernel");
    return VFA_OK;
  }
  if (bf16)
    VFA_CUDA(cudaLaunchKernelEx(&cfg, ygemm_compact_kernel<true>, a));
  else
    VFA_CUDA(cudaLaunchKernelEx(&cfg, ygemm_compact_kernel<false>, a));
  VFA_LAUNCH_CHECK("ygemm_compact_kernel");
  return VFA_OK;
}

// bytes of Y for one frame
size_t fside_y_bytes_per_frame(const AggParams& p) {
  size_t texels = 0;
  for (int s = 0; s < p.S; ++s) texels += (size_t)p.sc[s].fh * p.sc[s].fw;
  return texels * p.V * p.nl * CH * (p.y_bf16 ? 2 : sizeof(float));
}

// frames per chunk for a Y budget (the workspace holds Y of one chunk)
int fside_chunk_frames(const AggParams& p) {
  size_t budget = (size_t)6 << 30;
  if (runtime_config().y_budget_mb > 0) budget = (size_t)runtime_config().y_budget_mb << 20;      // tests: several chunks
  const size_t per = fside_y_bytes_per_frame(p);
  size_t cb = budget / per;
  if (cb < 1) cb = 1;
  if (cb > (size_t)p.B) cb = (size_t)p.B;
  return (int)cb;
}

// ---- texel lists of the quads (pool_list_kernel) --------------------------------------------------------------------------
// VFA_POOL_LIST=0 selects the walking kernel (pool_quad_kernel) for the whole grid; VFA_POOL_LIST_CAP = list capacity in
// entries per (view, scale, layer) of a quad's slot (default 16: 2.8 - 4.4 x the rigs' average; tests force overflows with 1).
struct ListLayout {
  size_t off_segoff, off_entoff, off_entw, total;
  int n_quads, quads_x;
  uint32_t slot;
};
static bool pool_list_enabled() {
  // (the one-cell-per-warp comparison kernel has no lists)
  return runtime_config().pool_list != 0 && runtime_config().pool_quad != 0;
}
static ListLayout list_layout(const AggParams& p) {
  ListLayout L = {};
  // the builder keeps LIST_PAD * V * S * nl counters in (default-sized) shared memory
  if (!pool_list_enabled() || ((size_t)LIST_PAD * p.V * p.S * p.nl + LIST_QPB) * sizeof(uint32_t) > 40960) return L;
  L.quads_x = (p.W + 1) / 2;
  L.n_quads = L.quads_x * ((p.L + 1) / 2);
  unsigned long long per = LIST_PER_ITER;
  if (runtime_config().pool_list_cap > 0) per = (unsigned long long)runtime_config().pool_list_cap;
  unsigned long long slot = per * (unsigned long long)(p.V * p.S * p.nl);
  const unsigned long long most = 0xfffffff0ull / (unsigned long long)(L.n_quads > 0 ? L.n_quads : 1);
  if (slot > most) slot = most;                          // entry indices are 32-bit
  L.slot = (uint32_t)slot;
  const size_t entries = (size_t)L.slot * L.n_quads;
  size_t o = 0;
  L.off_segoff = o; o += align256((size_t)L.n_quads * (p.V * p.S + 1) * sizeof(uint32_t));
  L.off_entoff = o; o += align256(entries * sizeof(uint32_t));
  L.off_entw = o;   o += align256(entries * sizeof(float4));
  L.total = o;
  return L;
}

// texel lists from the tap records (once per call: the projection is static, the lists serve every frame)
static int launch_quad_lists(const AggParams& p, const TapRec* recs, void* list_ws, cudaStream_t st) {
  const ListLayout L = list_layout(p);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(list_ws);
  const size_t smem = ((size_t)LIST_PAD * p.V * p.S * p.nl + LIST_QPB) * sizeof(uint32_t);
  qlist_build_kernel<<<(L.n_quads + LIST_QPB - 1) / LIST_QPB, 256, smem, st>>>(
      p, recs, L.quads_x, L.n_quads, L.slot, reinterpret_cast<uint32_t*>(w8 + L.off_segoff),
      reinterpret_cast<uint32_t*>(w8 + L.off_entoff), reinterpret_cast<float4*>(w8 + L.off_entw));
  VFA_LAUNCH_CHECK("qlist_build_kernel");
  return VFA_OK;
}

// staged-tile pooling (vfa_pool_tile.cu): the default; 0 bytes = not available for this problem (the quads' lists serve it)
size_t tile_pool_workspace_bytes(const AggParams& p);
int launch_tile_build(const AggParams& p, const TapRec* recs, void* ws, void* cover_ws, cudaStream_t st);
int launch_pool_tile(fside::PoolArgs q, void* ws, int nb, bool y_bf16, cudaStream_t st);
void tile_pool_overflow_view(const AggParams& p, void* ws, const uint8_t** tile_ovf, int* tiles_x);

// bytes of the pooling lists: chunk lists of the tiles, else texel lists of the quads
static size_t pool_lists_bytes(const AggParams& p) {
  const size_t t = tile_pool_workspace_bytes(p);
  return t != 0 ? t : list_layout(p).total;
}

// workspace of the feature-side forward behind the prepared weights and tap records:
// [cover bitmap + need bytes + row lists][chunk lists of the tiles | texel lists of the quads][Y]
size_t fside_workspace_bytes(const AggParams& p) {
  const int cb = fside_chunk_frames(p);
  return fside_cover_bytes(p, cb) + pool_lists_bytes(p) + (size_t)cb * fside_y_bytes_per_frame(p);
}

// prepared bf16 weight slabs of the bf16 tensor-core variant (one slot of `per_scale` bytes per scale, as the tf32 slabs)
int prep_weights_bf16(const AggParams& p, const float* const* d_weight, void* ws, size_t per_scale, cudaStream_t st) {
  for (int s = 0; s < p.S; ++s) {
    bfmma::prep_weight_bf16_kernel<<<148 * 4, 256, 0, st>>>(d_weight[s], reinterpret_cast<uint8_t*>(ws) + s * per_scale, p.nl);
    VFA_LAUNCH_CHECK("prep_weight_bf16_kernel");
  }
  return VFA_OK;
}

int launch_fwd_fside(const AggParams& p_in, const uint8_t* const* wprep, const TapRec* recs, void* fs_ws, size_t fs_bytes,
                     uint32_t flags, int variant, cudaStream_t st) {
  AggParams p = p_in;
  p.y_bf16 = (flags & VFA_FLAG_BF16_MMA) ? 1 : 0;
  const bool bf16 = (flags & VFA_FLAG_BF16_FEATURES) != 0;
  {
    VFA_CUDA(cudaFuncSetAttribute(pool_y_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxL1));
    VFA_CUDA(cudaFuncSetAttribute(pool_y_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxL1));
    // three CTAs x 32 KB of partial sums per SM, the rest of the unified array as L1
    VFA_CUDA(cudaFuncSetAttribute(pool_quad_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_quad_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_quad_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_quad_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_quad_kernel<false, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_list_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    VFA_CUDA(cudaFuncSetAttribute(pool_list_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
  }
  const size_t per_frame = fside_y_bytes_per_frame(p);
  const size_t tile_bytes = tile_pool_workspace_bytes(p);
  ListLayout LL = list_layout(p);
  if (tile_bytes != 0) LL = ListLayout{};                  // the tiles' chunk lists replace the quads' texel lists
  const size_t lists_bytes = tile_bytes != 0 ? tile_bytes : LL.total;
  if (p.y_bf16 && tile_bytes == 0) {
    set_error("VFA_FLAG_BF16_MMA needs the staged-tile pooling (views * scales <= 32, feature maps below 4 M texel-layers)");
    return VFA_ERR_UNSUPPORTED;
  }
  int cb = fside_chunk_frames(p);
  while (cb > 1 && fside_cover_bytes(p, cb) + lists_bytes + (size_t)cb * per_frame > fs_bytes) --cb;
  const size_t cover_bytes = fside_cover_bytes(p, cb);
  if (cover_bytes + lists_bytes + per_frame > fs_bytes) {
    set_error("feature-side forward: workspace holds %zu bytes for Y, one frame needs %zu", fs_bytes,
              cover_bytes + lists_bytes + per_frame);
    return VFA_ERR_WORKSPACE;
  }
  uint8_t* list_ws = reinterpret_cast<uint8_t*>(fs_ws) + cover_bytes;
  float* y_ws = reinterpret_cast<float*>(list_ws + lists_bytes);
  // row-compacted GEMM unless VFA_FSIDE_COMPACT=0 (whole 256-row tiles, skipped by the need bytes) or VFA_FSIDE_NO_SKIP=1
  const bool compact = runtime_config().fside_compact != 0 && runtime_config().fside_no_skip == 0;
  if (!(variant & 256)) {       // (debug bit 256: reuse tap records, coverage bitmap, lists and need bytes of the previous call)
    // the chunk-list builder sees every covered (layer, texel) row of every tile: it fills the coverage bitmap on its way
    // (one pass over the records less); without it cover_mark_kernel does
    const bool build_tiles = tile_bytes != 0 && !(variant & 64);
    if (build_tiles) {
      if (int rc = launch_tile_build(p, recs, list_ws, fs_ws, st)) return rc;
    } else {
      if (int rc = launch_cover_mark(p, recs, fs_ws, st)) return rc;
    }
    if (compact)
      if (int rc = launch_rowlists(p, fs_ws, cb, st)) return rc;
    if (!build_tiles && LL.total != 0 && !(variant & 64)) {
      if (int rc = launch_quad_lists(p, recs, list_ws, st)) return rc;
    }
  }
  const size_t es = bf16 ? 2 : 4;
  for (int b0 = 0; b0 < p.B; b0 += cb) {
    const int nb = p.B - b0 < cb ? p.B - b0 : cb;
    YGemmArgs g;
    PoolArgs q = {};
    q.out_nhwc = (flags & VFA_FLAG_OUT_NHWC) ? 1 : 0;
    q.out_mode = (flags & VFA_FLAG_OUT_PEERS) ? 3 : ((flags & VFA_FLAG_OUT_MULTICAST) ? 2 : ((flags & VFA_FLAG_OUT_ACCUMULATE) ? 1 : 0));
    q.p = p;
    q.recs = recs;
    q.b0 = b0;
    q.tiles_x = (p.W + POOL_TW - 1) / POOL_TW;
    g.nl = p.nl;
    g.S = p.S;
    size_t y_off = 0;      // bytes
    const size_t y_es = p.y_bf16 ? 2 : 4;
    g.tile_begin[0] = 0;
    for (int s = 0; s < VFA_MAX_SCALES; ++s) {
      const int ss = s < p.S ? s : 0;
      const int hw = p.sc[ss].fh * p.sc[ss].fw;
      g.hw[s] = hw;
      g.rows[s] = nb * p.V * hw;
      g.feats[s] = reinterpret_cast<const uint8_t*>(p.feats[ss]) + (size_t)b0 * p.V * hw * CH * es;
      g.wprep[s] = wprep[ss];
      g.y[s] = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(y_ws) + (s < p.S ? y_off : 0));
      q.y[s] = g.y[s];
      g.need_tile0[s] = g.tile_begin[s < p.S ? s : 0];
      if (s < p.S) {
        y_off += (size_t)cb * p.V * p.nl * hw * CH * y_es;
        g.tile_begin[s + 1] = g.tile_begin[s] + (g.rows[s] + 2 * TILE_M - 1) / (2 * TILE_M);
      }
    }
    if (!(variant & 128)) {
      if (compact) {
        if (int rc = launch_ygemm_compact(p, g, fs_ws, cb, nb, bf16, st)) return rc;
      } else if (p.y_bf16) {
        set_error("VFA_FLAG_BF16_MMA runs the row-compacted GEMM only (VFA_FSIDE_COMPACT=0 / VFA_FSIDE_NO_SKIP=1 are fp32 switches)");
        return VFA_ERR_UNSUPPORTED;
      } else {
        if (int rc = launch_tile_need(p, fs_ws, nb, &g.need, st, !(variant & 256))) return rc;
        if (int rc = launch_ygemm(g, 0, bf16, st)) return rc;
      }
    }
    if (!(variant & 64)) {
      const int quad = runtime_config().pool_quad != 0;     // 0 selects the one-cell-per-warp kernel (debug / comparison)
      if (!quad && q.out_mode != 0) {
        set_error("VFA_POOL_QUAD=0 (debug kernel) does not implement VFA_FLAG_OUT_ACCUMULATE / _MULTICAST / _PEERS");
        return VFA_ERR_UNSUPPORTED;
      }
      q.tile_ovf = nullptr;
      q.ptiles_x = 0;
      if (quad && tile_bytes != 0) {
        q.tiles_x = (p.W + 2 * QX - 1) / (2 * QX);
        q.seg_off = nullptr;
        if (int rc = launch_pool_tile(q, list_ws, nb, p.y_bf16 != 0, st)) return rc;
        tile_pool_overflow_view(p, list_ws, &q.tile_ovf, &q.ptiles_x);
        const dim3 grid(q.tiles_x * ((p.L + 2 * QY - 1) / (2 * QY)), nb);
        if (p.y_bf16)
          pool_quad_kernel<false, true, true><<<grid, QWARPS * 32, 0, st>>>(q);      // (forward-only variant: no mask)
        else if (p.mask != nullptr)
          pool_quad_kernel<true, true><<<grid, QWARPS * 32, 0, st>>>(q);
        else
          pool_quad_kernel<false, true><<<grid, QWARPS * 32, 0, st>>>(q);
        VFA_LAUNCH_CHECK("pool_quad_kernel (completion pass)");
      } else if (quad && LL.total != 0) {
        q.tiles_x = (p.W + 2 * QX - 1) / (2 * QX);
        q.seg_off = reinterpret_cast<const uint32_t*>(list_ws + LL.off_segoff);
        q.ent_off = reinterpret_cast<const uint32_t*>(list_ws + LL.off_entoff);
        q.ent_w = reinterpret_cast<const float4*>(list_ws + LL.off_entw);
        q.quads_x = LL.quads_x;
        const dim3 grid(q.tiles_x * ((p.L + 2 * QY - 1) / (2 * QY)), nb);
        if (p.mask != nullptr) {
          pool_list_kernel<true><<<grid, QWARPS * 32, 0, st>>>(q);
          pool_quad_kernel<true, true><<<grid, QWARPS * 32, 0, st>>>(q);
        } else {
          pool_list_kernel<false><<<grid, QWARPS * 32, 0, st>>>(q);
          pool_quad_kernel<false, true><<<grid, QWARPS * 32, 0, st>>>(q);
        }
        VFA_LAUNCH_CHECK("pool_list_kernel");
      } else if (quad) {
        q.tiles_x = (p.W + 2 * QX - 1) / (2 * QX);
        q.seg_off = nullptr;
        const dim3 grid(q.tiles_x * ((p.L + 2 * QY - 1) / (2 * QY)), nb);
        if (p.mask != nullptr)
          pool_quad_kernel<true><<<grid, QWARPS * 32, 0, st>>>(q);
        else
          pool_quad_kernel<false><<<grid, QWARPS * 32, 0, st>>>(q);
        VFA_LAUNCH_CHECK("pool_quad_kernel");
      } else {
        const dim3 grid(q.tiles_x * ((p.L + POOL_TH - 1) / POOL_TH), nb);
        if (p.mask != nullptr)
          pool_y_kernel<true><<<grid, POOL_WARPS * 32, 0, st>>>(q);
        else
          pool_y_kernel<false><<<grid, POOL_WARPS * 32, 0, st>>>(q);
        VFA_LAUNCH_CHECK("pool_y_kernel");
      }
    }
  }
  return VFA_OK;
}

}  // namespace vfa
