// Fused aggregation forward on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
// accumulator in TMEM, 3xTF32 operand splitting for fp32-level accuracy.   C = 256 channels.
//
// One CTA (448 threads, 1 per SM) owns a tile of 128 consecutive BEV cells of one frame (= the 128 TMEM lanes)
// and a group of views.  For every (view, scale) it runs a K loop over (height layer n, 32-channel chunk):
//
//   warp 0      weight loader   one cp.async.bulk (UBLKCP) per stage: the pre-split, pre-swizzled [256 x 32]
//                               hi and lo slabs of collapse.weight (64 KB, contiguous in the prepared layout)
//   warp 1      MMA issuer      one thread: per stage 4 k-steps x {A_hi*B_hi, A_lo*B_hi, A_hi*B_lo}
//                               tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8, D in TMEM cols [0,256)
//   warps 2-9   pool producers  derive the box taps (shared device function with the parity-checked table
//                               kernel), gather 4 channels per thread with 128-bit loads from the channels-last
//                               map (8 lanes = one 128 B row of a texel; 4 cells per warp instruction), split
//                               into tf32 hi/lo and write the K-major SWIZZLE_128B operand tile in shared memory
//   warps 10-13 epilogue        tcgen05.ld the accumulator, + bias, ReLU, add to the running BEV sum kept in
//                               TMEM cols [256,512); after the last (view, scale) write [B,C,L,W] once
//
// mbarriers: full[stage] (8 producer warps + loader tx), empty[stage] (tcgen05.commit), acc_full (commit),
// acc_empty (epilogue).  The [L*W, C*nl] matrix of the reference (vfa_op.py:118-120) only ever exists as
// 16 KB operand tiles; the collapse (vfa_op.py:123) is the tensor-core contraction; ReLU and the sums over
// scales and views (vfa_op.py:124, vfanet.py:79-82) are the epilogue.
#include <stdlib.h>

#include "vfa_common.cuh"

namespace vfa {

namespace umma {

constexpr int CH = 256;                 // channels == MMA N
constexpr int TILE_M = 128;             // cells per CTA == TMEM lanes
constexpr int KCH = 32;                 // K elements per stage (one 128-byte swizzle row of tf32)
constexpr int STAGES = 2;
constexpr int A_BYTES = TILE_M * KCH * 4;          // 16 KB per hi / lo tile
constexpr int B_BYTES = CH * KCH * 4;              // 32 KB per hi / lo slab
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int FIRST_PRODUCER_WARP = 4;     // warpgroup 0 = {loader, MMA, 2 idle warps} gives its registers away
constexpr int FIRST_EPILOGUE_WARP = 12;
constexpr int NUM_EPILOGUE_WARPS = 8;
constexpr int THREADS = 20 * 32;
constexpr int TMEM_COLS = 512;

struct __align__(8) SmemTail {
  BoxTaps taps[NUM_PRODUCER_WARPS][16];
  float bias[VFA_MAX_SCALES][CH];
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
  unsigned long long acc_full;
  unsigned long long acc_empty;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + sizeof(SmemTail);

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 |
// SBO(1024 B >> 4)<<32 | version 1 <<46 | layout SWIZZLE_128B(2) <<61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both, N=256, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

// byte offset of (row, 16-byte chunk j) inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __host__ __forceinline__ uint32_t swz(uint32_t row, uint32_t j) { return row * 128u + ((j ^ (row & 7u)) << 4); }

}  // namespace umma

using namespace umma;

// collapse.weight [C, C*nl] (column c*nl+n)  ->  per scale, per K chunk kc = n*(C/32) + c/32 a 64 KB block
// [hi: 256 rows x 128 B swizzled][lo: same], hi = tf32(w), lo = tf32(w - hi).
__global__ void __launch_bounds__(256) prep_weight_umma_kernel(const float* __restrict__ w, uint8_t* __restrict__ wp, int nl) {
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

struct UmmaArgs {
  AggParams p;
  const uint8_t* wprep[VFA_MAX_SCALES];
  int n_groups;        // view groups (grid.x = tiles * n_groups); > 1 -> atomic accumulation into a zeroed output
  int views_per_group;
  int variant;         // debug: 0 = 3xTF32, 1 = hi*hi only (single-pass tf32)
};

// One producer work item = 4 cells (one per quarter-warp) x 32 channels of one K chunk.  The loads of an item are
// issued into registers DEPTH items ahead of their use, so a warp keeps DEPTH * T * T 128-bit loads in flight.
template <int T>
struct GatherBuf {
  float4 v[T][T];
};

template <int T>
__device__ __forceinline__ void issue_item(GatherBuf<T>& buf, const BoxTaps& t, const float* __restrict__ feat, int fw,
                                           int coff) {
  const float* base = feat + ((size_t)t.y0 * fw + t.x0) * CH + coff;
#pragma unroll
  for (int ty = 0; ty < T; ++ty)
#pragma unroll
    for (int tx = 0; tx < T; ++tx) {
      const bool on = (ty < t.ny) && (tx < t.nx);
      buf.v[ty][tx] = on ? __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty * fw + tx) * CH))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int T>
__device__ __forceinline__ float4 reduce_item(const GatherBuf<T>& buf, const BoxTaps& t, const float* __restrict__ feat,
                                              int fw, int coff) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t.nx <= T && t.ny <= T) {
#pragma unroll
    for (int ty = 0; ty < T; ++ty) {
      const float wy = (ty < t.ny) ? tap_wy(t, ty) : 0.f;
#pragma unroll
      for (int tx = 0; tx < T; ++tx) {
        const float w = (tx < t.nx) ? wy * tap_wx(t, tx) : 0.f;
        acc.x = fmaf(w, buf.v[ty][tx].x, acc.x);
        acc.y = fmaf(w, buf.v[ty][tx].y, acc.y);
        acc.z = fmaf(w, buf.v[ty][tx].z, acc.z);
        acc.w = fmaf(w, buf.v[ty][tx].w, acc.w);
      }
    }
  } else {  // rare large box (near-camera voxel): plain loops, not prefetched
    const float* base = feat + ((size_t)t.y0 * fw + t.x0) * CH + coff;
    for (int ty = 0; ty < t.ny; ++ty) {
      const float wy = tap_wy(t, ty);
      float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int tx = 0; tx < t.nx; ++tx) {
        const float wx = tap_wx(t, tx);
        const float4 f4 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty * fw + tx) * CH));
        rs.x = fmaf(wx, f4.x, rs.x);
        rs.y = fmaf(wx, f4.y, rs.y);
        rs.z = fmaf(wx, f4.z, rs.z);
        rs.w = fmaf(wx, f4.w, rs.w);
      }
      acc.x = fmaf(wy, rs.x, acc.x);
      acc.y = fmaf(wy, rs.y, acc.y);
      acc.z = fmaf(wy, rs.z, acc.z);
      acc.w = fmaf(wy, rs.w, acc.w);
    }
  }
  return acc;
}

// All 8 K chunks of one height layer for the 16 rows of one producer warp: 32 items, DEPTH in flight.
template <int T, int DEPTH>
__device__ __forceinline__ void produce_layer(uint8_t* smem, SmemTail* tail, const BoxTaps* __restrict__ wtaps,
                                              const float* __restrict__ feat, int fw, int pw, int lane, int& it) {
  constexpr int ITEMS = (CH / KCH) * 4;
  const int q = lane >> 3, j = lane & 7;
  GatherBuf<T> buf[DEPTH];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) issue_item<T>(buf[d], wtaps[(d & 3) * 4 + q], feat, fw, (d >> 2) * KCH + j * 4);
#pragma unroll 1
  for (int base = 0; base < ITEMS; base += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const int item = base + d;
      const int cc = item >> 2, round = item & 3;
      const int st = it % STAGES;
      if (round == 0) mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
      const int lr = round * 4 + q;                     // row within the warp's 16
      const BoxTaps t = wtaps[lr];
      const float4 acc = reduce_item<T>(buf[d], t, feat, fw, cc * KCH + j * 4);
      const int nxt = item + DEPTH;
      if (nxt < ITEMS) issue_item<T>(buf[d], wtaps[(nxt & 3) * 4 + q], feat, fw, (nxt >> 2) * KCH + j * 4);
      uint4 hi, lo;
      hi.x = to_tf32(acc.x);
      hi.y = to_tf32(acc.y);
      hi.z = to_tf32(acc.z);
      hi.w = to_tf32(acc.w);
      lo.x = to_tf32(acc.x - __uint_as_float(hi.x));
      lo.y = to_tf32(acc.y - __uint_as_float(hi.y));
      lo.z = to_tf32(acc.z - __uint_as_float(hi.z));
      lo.w = to_tf32(acc.w - __uint_as_float(hi.w));
      uint8_t* a_hi = smem + (size_t)st * STAGE_BYTES;
      const uint32_t off = swz((uint32_t)(pw * 16 + lr), (uint32_t)j);
      *reinterpret_cast<uint4*>(a_hi + off) = hi;
      *reinterpret_cast<uint4*>(a_hi + A_BYTES + off) = lo;
      if (round == 3) {
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->full[st]);
        ++it;
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) aggregate_fwd_umma_kernel(const UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const AggParams& p = a.p;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x / a.n_groups, group = blockIdx.x % a.n_groups;
  const int b = blockIdx.y;
  const int cell0 = tile * TILE_M;
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
    }
    mbar_init(&tail->acc_full, 1);
    mbar_init(&tail->acc_empty, NUM_EPILOGUE_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp < FIRST_PRODUCER_WARP) {
    // warpgroup 0 (loader, MMA issuer, two idle warps) hands registers to the producer warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
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
            mbar_arrive_expect_tx(&tail->full[st], 2 * B_BYTES);
            bulk_g2s(dst, a.wprep[s] + (size_t)kc * (2 * B_BYTES), 2 * B_BYTES, &tail->full[st]);
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The TMEM accumulator is only trusted for ONE height layer (32 k-steps): the tensor core truncates every
    // accumulation (measured: error grows linearly in K, biased toward zero), so layer partials are summed by the
    // epilogue warps with round-to-nearest FADDs (drain) instead of inside the tensor core.
    if (lane == 0) {
      int it = 0, drain = 0;
      for (int vs = 0; vs < n_vs; ++vs) {
        for (int n = 0; n < p.nl; ++n, ++drain) {
          mbar_wait(&tail->acc_empty, (drain & 1) ^ 1);
          tc_fence_after();
          for (int cc = 0; cc < CHUNKS_PER_LAYER; ++cc, ++it) {
            const int st = it % STAGES;
            mbar_wait(&tail->full[st], (it / STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
            const uint64_t b_hi = make_desc(sa + 2 * A_BYTES), b_lo = make_desc(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int ks = 0; ks < KCH / 8; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);      // 8 tf32 = 32 bytes along K inside the swizzle row
              if (a.variant == 1) {
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    const int pw = warp - FIRST_PRODUCER_WARP;       // 0..7 -> rows pw*16 .. pw*16+15
    BoxTaps* wtaps = tail->taps[pw];
    int it = 0;
    for (int v = v_begin; v < v_end; ++v) {
      for (int s = 0; s < p.S; ++s) {
        const ScaleConst sc = p.sc[s];
        const float* __restrict__ feat = p.feats[s] + ((size_t)(b * p.V + v) * sc.fh * sc.fw) * CH;
        for (int n = 0; n < p.nl; ++n) {
          __syncwarp();                              // everyone is done with the previous layer's taps
          int extent = 0;
          if (lane < 16) {
            const int cell = cell0 + pw * 16 + lane;
            BoxTaps t;
            if (cell < p.LW) {
              t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)v * p.nl + n) * p.LW + cell], sc);
            } else {
              t.x0 = t.y0 = t.nx = t.ny = 0;
              t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
            }
            wtaps[lane] = t;
            extent = max(t.nx, t.ny);
          }
          extent = __reduce_max_sync(0xffffffffu, extent);   // also orders the taps writes before the reads below
          __syncwarp();
          if (extent <= 2)
            produce_layer<2, 4>(smem, tail, wtaps, feat, sc.fw, pw, lane, it);
          else
            produce_layer<3, 2>(smem, tail, wtaps, feat, sc.fw, pw, lane, it);
        }
      }
    }
  } else {
    // ================= epilogue: layer drains + per-(view, scale) finalisation =================
    const int e = warp - FIRST_EPILOGUE_WARP;      // 0..7
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int col_begin = (e >> 2) * (CH / 2);     // two warps per quarter split the 256 columns
    const int row = quarter * 32 + lane;
    const int cell = cell0 + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* outp = p.out + (size_t)b * CH * p.LW + cell;
    int drain = 0;
    for (int vs = 0; vs < n_vs; ++vs) {
      const int s = vs % p.S;
      for (int n = 0; n < p.nl; ++n, ++drain) {
        mbar_wait(&tail->acc_full, drain & 1);
        tc_fence_after();
        const bool first_layer = (n == 0), last_layer = (n == p.nl - 1);
#pragma unroll 1
        for (int c0 = col_begin; c0 < col_begin + CH / 2; c0 += 32) {
          float acc[32], pre[32];
          tc_ld32(lane_addr + c0, acc);
          if (!first_layer) tc_ld32(lane_addr + CH + c0, pre);
          tc_wait_ld();
          if (!first_layer) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] += pre[i];                   // round-to-nearest layer sum
          }
          if (!last_layer) {
            tc_st32(lane_addr + CH + c0, acc);
          } else if (cell < p.LW) {
            // + bias, ReLU (vfa_op.py:123-124), then the sum over scales and views (vfanet.py:79, :82)
            if (a.n_groups == 1 && vs == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) outp[(size_t)(c0 + i) * p.LW] = fmaxf(acc[i] + tail->bias[s][c0 + i], 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                atomicAdd(outp + (size_t)(c0 + i) * p.LW, fmaxf(acc[i] + tail->bias[s][c0 + i], 0.f));
            }
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->acc_empty);      // accumulator columns may be overwritten
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
bool umma_supported(const vfa_geometry_t* g, const vfa_shape_t* sh, uint32_t flags) {
  (void)g;
  (void)flags;
  return sh->channels == CH;
}

size_t umma_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh, uint32_t flags) {
  if (!umma_supported(g, sh, flags)) return 0;
  return (size_t)sh->n_scales * g->n_layers * (CH / KCH) * (2 * B_BYTES);
}

int prep_weights_umma(const AggParams& p, const float* const* d_weight, void* ws, cudaStream_t st) {
  const size_t per_scale = (size_t)p.nl * (CH / KCH) * (2 * B_BYTES);
  for (int s = 0; s < p.S; ++s) {
    prep_weight_umma_kernel<<<148 * 4, 256, 0, st>>>(d_weight[s], reinterpret_cast<uint8_t*>(ws) + s * per_scale, p.nl);
    VFA_LAUNCH_CHECK("prep_weight_umma_kernel");
  }
  return VFA_OK;
}

int launch_fwd_umma(AggParams p, const float* const* d_weight, void* ws, uint32_t flags, cudaStream_t st) {
  if (!(flags & VFA_FLAG_WEIGHTS_PREPARED)) {
    if (int rc = prep_weights_umma(p, d_weight, ws, st)) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VFA_CUDA(cudaFuncSetAttribute(aggregate_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  UmmaArgs a;
  a.p = p;
  const char* var = getenv("VFA_UMMA_VARIANT");
  a.variant = var ? atoi(var) : 0;
  const size_t per_scale = (size_t)p.nl * (CH / KCH) * (2 * B_BYTES);
  for (int s = 0; s < VFA_MAX_SCALES; ++s) a.wprep[s] = reinterpret_cast<const uint8_t*>(ws) + (s < p.S ? s : 0) * per_scale;
  const int tiles = (p.LW + TILE_M - 1) / TILE_M;
  // enough CTAs for >= ~4 waves over 148 SMs, otherwise split the views and accumulate atomically
  a.n_groups = ((long long)tiles * p.B >= 4 * 148 || p.V == 1) ? 1 : p.V;
  a.views_per_group = (p.V + a.n_groups - 1) / a.n_groups;
  if (a.n_groups > 1) VFA_CUDA(cudaMemsetAsync(p.out, 0, (size_t)p.B * CH * p.LW * sizeof(float), st));
  dim3 grid(tiles * a.n_groups, p.B);
  aggregate_fwd_umma_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(a);
  VFA_LAUNCH_CHECK("aggregate_fwd_umma_kernel");
  set_path("umma_tf32x3");
  return VFA_OK;
}

}  // namespace vfa
