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
constexpr int FIRST_PRODUCER_WARP = 2;
constexpr int FIRST_EPILOGUE_WARP = 10;
constexpr int THREADS = 14 * 32;
constexpr int TMEM_COLS = 512;

struct __align__(8) SmemTail {
  BoxTaps taps[2][TILE_M];
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
};

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
  const int chunks_per_vs = p.nl * (CH / KCH);

  // ---- one-time setup ----
  for (int i = tid; i < p.S * CH; i += THREADS) tail->bias[i / CH][i % CH] = __ldg(p.bias[i / CH] + (i % CH));
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&tail->full[s], NUM_PRODUCER_WARPS + 1);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->acc_full, 1);
    mbar_init(&tail->acc_empty, 4);
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

  if (warp == 0) {
    // ================= weight loader =================
    if (lane == 0) {
      int it = 0;
      for (int v = v_begin; v < v_end; ++v)
        for (int s = 0; s < p.S; ++s)
          for (int kc = 0; kc < chunks_per_vs; ++kc, ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&tail->empty[st], ph ^ 1);
            uint8_t* dst = smem + (size_t)st * STAGE_BYTES + 2 * A_BYTES;
            mbar_arrive_expect_tx(&tail->full[st], 2 * B_BYTES);
            bulk_g2s(dst, a.wprep[s] + (size_t)kc * (2 * B_BYTES), 2 * B_BYTES, &tail->full[st]);
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int it = 0;
      for (int vs = 0; vs < n_vs; ++vs) {
        mbar_wait(&tail->acc_empty, (vs & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < chunks_per_vs; ++kc, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&tail->full[st], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
          const uint64_t b_hi = make_desc(sa + 2 * A_BYTES), b_lo = make_desc(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int ks = 0; ks < KCH / 8; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);      // 8 tf32 = 32 bytes along K inside the swizzle row
            // small cross terms first, the dominant hi*hi product last
            tc_mma_tf32(tmem, a_lo + adv, b_hi + adv, IDESC, (kc | ks) ? 1u : 0u);
            tc_mma_tf32(tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
            tc_mma_tf32(tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
          }
          tc_commit(&tail->empty[st]);          // frees the stage when these MMAs have read it
        }
        tc_commit(&tail->acc_full);             // accumulator of this (view, scale) is complete
      }
    }
  } else if (warp < FIRST_EPILOGUE_WARP) {
    // ================= pool producers =================
    const int pw = warp - FIRST_PRODUCER_WARP;       // 0..7 -> rows pw*16 .. pw*16+15
    const int ptid = tid - FIRST_PRODUCER_WARP * 32;  // 0..255
    const int q = lane >> 3, j = lane & 7;            // quarter-warp = one cell, 8 lanes x float4 = 32 channels
    int it = 0;
    int nbuf = 0;
    for (int v = v_begin; v < v_end; ++v) {
      for (int s = 0; s < p.S; ++s) {
        const ScaleConst sc = p.sc[s];
        const float* __restrict__ feat = p.feats[s] + ((size_t)(b * p.V + v) * sc.fh * sc.fw) * CH;
        for (int n = 0; n < p.nl; ++n, nbuf ^= 1) {
          if (ptid < TILE_M) {
            const int cell = cell0 + ptid;
            BoxTaps t;
            if (cell < p.LW) {
              t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)v * p.nl + n) * p.LW + cell], sc);
            } else {
              t.x0 = t.y0 = t.nx = t.ny = 0;
              t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
            }
            tail->taps[nbuf][ptid] = t;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");        // producers only
          for (int cc = 0; cc < CH / KCH; ++cc, ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&tail->empty[st], ph ^ 1);
            uint8_t* a_hi = smem + (size_t)st * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_BYTES;
            const int coff = cc * KCH + j * 4;
#pragma unroll 1
            for (int round = 0; round < 4; ++round) {
              const int r = pw * 16 + round * 4 + q;
              const BoxTaps t = tail->taps[nbuf][r];
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              const float* base = feat + ((size_t)t.y0 * sc.fw + t.x0) * CH + coff;
              const bool small = (t.nx <= 3) && (t.ny <= 3);
              if (__all_sync(0xffffffffu, small)) {
                float4 val[3][3];
                float wgt[3][3];
#pragma unroll
                for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                  for (int tx = 0; tx < 3; ++tx) {
                    const bool on = (ty < t.ny) && (tx < t.nx);
                    wgt[ty][tx] = on ? tap_wy(t, ty) * tap_wx(t, tx) : 0.f;
                    val[ty][tx] = on ? __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty * sc.fw + tx) * CH))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                  }
#pragma unroll
                for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                  for (int tx = 0; tx < 3; ++tx) {
                    acc.x = fmaf(wgt[ty][tx], val[ty][tx].x, acc.x);
                    acc.y = fmaf(wgt[ty][tx], val[ty][tx].y, acc.y);
                    acc.z = fmaf(wgt[ty][tx], val[ty][tx].z, acc.z);
                    acc.w = fmaf(wgt[ty][tx], val[ty][tx].w, acc.w);
                  }
              } else {
                for (int ty = 0; ty < t.ny; ++ty) {
                  const float wy = tap_wy(t, ty);
                  float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
                  for (int tx = 0; tx < t.nx; ++tx) {
                    const float wx = tap_wx(t, tx);
                    const float4 f4 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty * sc.fw + tx) * CH));
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
              uint4 hi, lo;
              hi.x = to_tf32(acc.x);
              hi.y = to_tf32(acc.y);
              hi.z = to_tf32(acc.z);
              hi.w = to_tf32(acc.w);
              lo.x = to_tf32(acc.x - __uint_as_float(hi.x));
              lo.y = to_tf32(acc.y - __uint_as_float(hi.y));
              lo.z = to_tf32(acc.z - __uint_as_float(hi.z));
              lo.w = to_tf32(acc.w - __uint_as_float(hi.w));
              const uint32_t off = swz((uint32_t)r, (uint32_t)j);
              *reinterpret_cast<uint4*>(a_hi + off) = hi;
              *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
            fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&tail->full[st]);
          }
        }
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;
    const int cell = cell0 + row;
    const uint32_t lane_addr = tmem + ((uint32_t)(ew * 32) << 16);
    float* outp = p.out + (size_t)b * CH * p.LW + cell;
    int vs = 0;
    for (int v = v_begin; v < v_end; ++v) {
      for (int s = 0; s < p.S; ++s, ++vs) {
        mbar_wait(&tail->acc_full, vs & 1);
        tc_fence_after();
        const bool first = (vs == 0), last = (vs == n_vs - 1);
#pragma unroll 1
        for (int c0 = 0; c0 < CH; c0 += 32) {
          float acc[32], sum[32];
          tc_ld32(lane_addr + c0, acc);
          if (!first) tc_ld32(lane_addr + CH + c0, sum);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = fmaxf(acc[i] + tail->bias[s][c0 + i], 0.f);      // vfa_op.py:123-124
            sum[i] = first ? y : sum[i] + y;                                 // vfanet.py:79, :82
          }
          if (last) {
            if (cell < p.LW) {
              if (a.n_groups == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) outp[(size_t)(c0 + i) * p.LW] = sum[i];
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(outp + (size_t)(c0 + i) * p.LW, sum[i]);
              }
            }
          } else {
            tc_st32(lane_addr + CH + c0, sum);
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
