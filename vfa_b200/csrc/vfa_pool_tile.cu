// Pooling of Y from shared memory: one CTA per 8 x 8 tile of BEV cells, texel rows staged with cp.async.bulk.
//
// The list kernel (pool_list_kernel) loads every texel row a 2 x 2 quad touches straight from L2 / L1: on the MultiviewC
// rig that is 3.7 M row visits (1 KB each) per frame for 1.0 M distinct rows per 8 x 8 tile and 0.36 M distinct rows per
// frame -- 11 GB per B = 4 launch through the L2 -> SM fabric (its cap is ~6300 B/clk = 12 TB/s: 0.9 ms) and the L1 data
// pipe (multi-line LDGs replay at ~2 cycles per line).  Here the 16 quads of a tile share the rows:
//
//   tile_build_kernel (once per table)  per (tile, view, scale): bitmap of the (layer, texel) rows any box of the tile
//       covers (ORed into the coverage bitmap of the row-compacted GEMM on the way) -> rank of every row -> CHUNKS of TR
//       consecutive ranks.  Per chunk a descriptor (its rows as runs of consecutive rows, blob address) and a blob: per
//       quad the entries (slot of the row inside the chunk, the four cell weights wy * wx), ordered by slot -- a
//       deterministic order, whatever the thread schedule of the builder.
//   tile_order_kernel  tiles by descending chunk count (8 levels), the order in which the scheduler deals them.
//   pool_tile_kernel  persistent, one CTA per SM: four producer warps walk the chunk descriptors of the CTA's tiles and
//       issue one bulk copy per run (+ one for the blob) into a ring of shared-memory stages (mbarrier complete_tx);
//       16 consumer warps (one per quad, lane l = channels [4l, 4l+4) and [128+4l, ...)) wait for a stage, apply the
//       rows their entries name with LDS.128 + packed FMAs, release the stage; after the last chunk of a (view, scale):
//       + bias, ReLU, ReLU mask, sum into the tile's partial sums (64 KB of shared memory, swizzled); after the last
//       chunk of the tile the 8 x 8 x 256 block is written with full 32-byte sectors ([B,C,L,W]) or 1 KB rows ([B,L,W,C]).
//
// L2 -> SM traffic drops 3.5x, the texel rows leave the LSU / L1 path (conflict-free LDS, no replays), and the loads a
// warp waits for are shared-memory loads.  Tiles whose chunk lists do not fit the pools are flagged and pooled by the
// walking kernel (pool_quad_kernel<.., OVF>) -- exact for any rig, no host synchronisation, static workspace.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "vfa_common.cuh"
#include "vfa_pool.cuh"
#include "vfa_umma_ptx.cuh"

namespace vfa {
namespace ptile {

using namespace umma;
using fside::CH;
using fside::PoolArgs;
using fside::pick;

#ifndef VFA_TILE_ROWS
#define VFA_TILE_ROWS 32
#endif
#ifndef VFA_TILE_STAGES
#define VFA_TILE_STAGES 4
#endif
constexpr int TC = 8;                       // tile = TC x TC cells
constexpr int TQ = 16;                      // quads per tile = consumer warps
constexpr int TR = VFA_TILE_ROWS;           // rows (ranks) per chunk
constexpr int TSTAGES = VFA_TILE_STAGES;
#ifndef VFA_TILE_MAXENT
#define VFA_TILE_MAXENT 384
#endif
constexpr int MAX_ENT = VFA_TILE_MAXENT;    // entries of a chunk (a quad lists a row at most once: TR * TQ is the most there
                                            // can be); a chunk with more -> the tile is left to the walking kernel
constexpr int BLOB_HDR = 80;                // segpair[16] (begin | end << 16), nent, info, 8 bytes pad
constexpr int META_CAP = (BLOB_HDR + 17 * MAX_ENT + 127) / 128 * 128;
constexpr int CAPV = 3584 / TR;              // chunks of one (tile, view, scale) the builder handles; more -> overflow (sized so
                                             // that four builder CTAs fit the shared memory of an SM on the shipped rigs)
#ifndef VFA_TILE_PRODUCERS
#define VFA_TILE_PRODUCERS 4
#endif
constexpr int NPROD = VFA_TILE_PRODUCERS;   // producer warps
#ifndef VFA_TILE_SPLIT_RUNS
#define VFA_TILE_SPLIT_RUNS 1
#endif
constexpr int CONSUMER_THREADS = TQ * 32;
constexpr int POOL_THREADS = CONSUMER_THREADS + 32 * NPROD;
constexpr int PREFETCH = 3;                 // chunk descriptors a producer warp holds in registers ahead of the ring
static_assert(NPROD >= 1 && (VFA_TILE_SPLIT_RUNS || NPROD <= VFA_TILE_STAGES), "");
static_assert(TR <= 32 && TR % 8 == 0, "one producer lane per row; slot masks are 32-bit");

constexpr uint32_t INFO_LAST_VS = 1u, INFO_LAST_TILE = 2u;     // info = flags | s << 8 | v << 16

template <typename YT>
struct RowBytes {
  static constexpr int value = CH * (int)sizeof(YT);
  static constexpr int stage = TR * value + META_CAP;                      // rows + blob of one chunk
  static constexpr size_t smem = (size_t)TC * TC * CH * 4 + (size_t)TSTAGES * stage + 2 * TSTAGES * 8 + 128 +
                                 (size_t)VFA_MAX_SCALES * CH * 4;
};
static_assert(RowBytes<float>::smem <= 232448, "exceeds the 227 KB of shared memory a CTA can have");

struct TileLists {
  const uint8_t* tile_ovf;     // [n_tiles]
  const uint2* tvs;            // [n_tiles][VS]: first chunk id, chunk count (>= 1) of a (tile, view, scale)
  const uint4* hdr;            // [chunk]: blob offset (16-byte units), blob bytes, nrows | nruns << 8 | nent << 16, info
  const uint32_t* rowoff;      // [chunk][TR]: the chunk's rows as RUNS of consecutive rows of Y (one bulk copy each):
                               // first row inside the (frame, view) stack of nl planes | first slot << 22 | (len - 1) << 27
  const uint8_t* blob;
  const int* order;            // [n_tiles]: tiles by descending chunk count (the dynamic scheduler deals the heavy ones first)
  int tiles_x, n_tiles, VS;
};

// ---- builder ----------------------------------------------------------------------------------------------------------
struct BuildArgs {
  AggParams p;
  const TapRec* recs;
  uint32_t* cursors;           // [0] chunk ids handed out, [1] blob pool used (16-byte units)
  uint8_t* tile_ovf;
  uint32_t* tile_work;         // [n_tiles]: chunks of the tile over its (view, scale) pairs (zeroed before the launch)
  uint2* tvs;
  uint4* hdr;
  uint32_t* rowoff;
  uint8_t* blob;
  uint32_t desc_cap, blob_cap16;
  int tiles_x, VS;
  int max_words;               // bitmap words of the largest scale (shared-memory layout)
  uint32_t* cover;             // coverage bitmap of the row-compacted GEMM (CoverMap layout; nullptr = not wanted)
  int cover_base[VFA_MAX_SCALES];
};

// The box of one (cell, layer), walked by one thread from the tile's records in shared memory: every texel with a non-zero
// weight wy * wx (the walking kernel's weights) is an entry of the cell's quad.  FILL = false: mark the row's slot in the
// (chunk, quad) mask.  FILL = true: write this cell's component of the entry's weights (the other three stay at the zero the
// blob was initialised with, or are written by their own cells) and the slot byte, at the entry's position = number of lower
// slots of the quad in the chunk.  Balanced (a box holds ~5 texels whatever the quad's union looks like) and, unlike a walk
// of the quad's union, without divergent loop bounds inside a warp.
template <bool FILL>
__device__ __forceinline__ void cell_walk(const uint4* __restrict__ recs_s, int cell, int n, int fw, int hw,
                                          const uint32_t* __restrict__ bits, const uint32_t* __restrict__ pref,
                                          uint32_t* __restrict__ qmask, const uint16_t* __restrict__ segbeg,
                                          const uint32_t* __restrict__ blob_off16, const uint16_t* __restrict__ nent_s,
                                          uint8_t* __restrict__ blob) {
  const uint4 r0 = recs_s[(n * TC * TC + cell) * 2], r1 = recs_s[(n * TC * TC + cell) * 2 + 1];
  const int nx = (int)r0.y & 0xffff, ny = (int)r0.y >> 16;
  if (nx == 0) return;
  const int x0 = (int)r0.x & 0xffff, y0 = (int)r0.x >> 16;
  const float wxf = __uint_as_float(r0.z), wxl = __uint_as_float(r0.w);
  const float wyf = __uint_as_float(r1.x), wyl = __uint_as_float(r1.y), wym = __uint_as_float(r1.z);
  const int cy = cell / TC, cx = cell % TC;
  const int q = (cy >> 1) * 4 + (cx >> 1), c = (cy & 1) * 2 + (cx & 1);       // quad of the tile, cell of the quad
  for (int ry = 0; ry < ny; ++ry) {
    const float wy = ry == 0 ? wyf : (ry == ny - 1 ? wyl : wym);
    // pass 1 marked the whole box row x0 .. x0 + nx - 1, so its ranks are consecutive: one bitmap lookup per box row
    const int idx0 = n * hw + (y0 + ry) * fw + x0;
    const uint32_t bit0 = (uint32_t)idx0 & 31u;
    const uint32_t rank0 = pref[idx0 >> 5] + __popc(bits[idx0 >> 5] & ((1u << bit0) - 1u));
    for (int rx = 0; rx < nx; ++rx) {
      const float wx = rx == 0 ? wxf : (rx == nx - 1 ? wxl : 1.0f);
      const float wl = __fmul_rn(wy, wx);
      if (!(wl != 0.f)) continue;                        // zero weights are not listed (NaN is: it propagates as in the reference)
      const uint32_t rank = rank0 + (uint32_t)rx;
      const uint32_t chunk = rank / TR, slot = rank % TR;
      if (!FILL) {
        atomicOr(&qmask[chunk * TQ + q], 1u << slot);
      } else {
        const uint32_t pos = segbeg[chunk * TQ + q] + __popc(qmask[chunk * TQ + q] & ((1u << slot) - 1u));
        uint8_t* b = blob + (size_t)blob_off16[chunk] * 16;
        reinterpret_cast<float*>(b + BLOB_HDR)[4 * pos + c] = wl;
        b[BLOB_HDR + 16 * (uint32_t)nent_s[chunk] + pos] = (uint8_t)slot;
      }
    }
  }
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
    const uint32_t v = warp_sums[k];
    if (k < warp) before += v;
    all += v;
  }
  __syncthreads();
  *total = all;
  return before + inc - x;
}

#ifdef VFA_BUILD_PROFILE
__device__ unsigned long long g_build_prof[16];
#define BUILD_PHASE(k)                                                        \
  do {                                                                        \
    __syncthreads();                                                          \
    if (tid == 0) {                                                           \
      const long long now_ = clock64();                                       \
      atomicAdd(&g_build_prof[k], (unsigned long long)(now_ - t_phase));      \
      t_phase = now_;                                                         \
    }                                                                         \
  } while (0)
#else
#define BUILD_PHASE(k)
#endif

// grid = n_tiles * VS CTAs of 256 threads; dynamic shared memory: records, bitmap, prefix, masks, segment starts.
__global__ void __launch_bounds__(256, 4) tile_build_kernel(const BuildArgs a) {
  extern __shared__ uint4 bsm[];
  __shared__ uint32_t warp_sums[8];
  __shared__ uint32_t base_s[2];
  __shared__ uint32_t blob_off16[CAPV];
  __shared__ uint16_t nent_s[CAPV];
  const AggParams& p = a.p;
  const int tile = blockIdx.x / a.VS, vs = blockIdx.x % a.VS;
  const int s = vs % p.S;
  const int fw = p.sc[s].fw, hw = p.sc[s].fh * p.sc[s].fw;
  const int nbits = p.nl * hw, nwords = (nbits + 31) >> 5;
  const int ty0 = (tile / a.tiles_x) * TC, tx0 = (tile % a.tiles_x) * TC;
  uint4* recs_s = bsm;                                                     // [nl * 64] x 32 B
  uint32_t* bits = reinterpret_cast<uint32_t*>(recs_s + (size_t)p.nl * TC * TC * 2);
  uint32_t* pref = bits + a.max_words;
  uint32_t* qmask = pref + a.max_words;                                    // [CAPV][TQ]
  uint16_t* segbeg = reinterpret_cast<uint16_t*>(qmask + CAPV * TQ);       // [CAPV][TQ]
  uint32_t* rank_idx = reinterpret_cast<uint32_t*>(segbeg + CAPV * TQ);    // [CAPV * TR] row of every rank
  const int tid = threadIdx.x;
#ifdef VFA_BUILD_PROFILE
  long long t_phase = clock64();
#endif

  for (int w = tid; w < nwords; w += 256) bits[w] = 0u;
  __syncthreads();
  BUILD_PHASE(0);
  // pass 1: the tile's records -> shared memory; rows of every visible box -> bitmap
  for (int item = tid; item < p.nl * TC * TC; item += 256) {
    const int n = item / (TC * TC), c = item % (TC * TC);
    const int cy = ty0 + c / TC, cx = tx0 + c % TC;
    uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
    if (cy < p.L && cx < p.W) {
      const uint4* rp = reinterpret_cast<const uint4*>(a.recs + ((size_t)vs * p.nl + n) * p.LW + cy * p.W + cx);
      r0 = __ldg(rp);
      r1 = __ldg(rp + 1);
    }
    recs_s[item * 2] = r0;
    recs_s[item * 2 + 1] = r1;
    const int nx = (int)r0.y & 0xffff, ny = (int)r0.y >> 16;
    if (nx == 0) continue;
    const int x0 = (int)r0.x & 0xffff, y0 = (int)r0.x >> 16;
    for (int ty = 0; ty < ny; ++ty) {
      const int t0 = n * hw + (y0 + ty) * fw + x0, t1 = t0 + nx - 1;      // inclusive bit range of the box row
      for (int w = t0 >> 5; w <= t1 >> 5; ++w) {
        const int lo = max(t0, w << 5) & 31, hi = min(t1, (w << 5) + 31) & 31;
        const uint32_t m = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
        if ((bits[w] & m) != m) atomicOr(&bits[w], m);
      }
    }
  }
  __syncthreads();
  BUILD_PHASE(1);
  // the tile's covered rows -> the coverage bitmap of the GEMM (plane n of the shared bitmap starts at bit n * hw, a plane
  // of the global one at a word boundary: funnel shift; most words are zero and cost two shared-memory reads)
  if (a.cover != nullptr) {
    const int cwords = (hw + 31) >> 5;
    uint32_t* cplane = a.cover + pick(a.cover_base, s) + (size_t)((vs / p.S) * p.nl) * cwords;
    for (int i = tid; i < p.nl * cwords; i += 256) {
      const int n = i / cwords, w = i - n * cwords;
      const int b0 = n * hw + w * 32;
      const int sw = b0 >> 5, sh = b0 & 31;
      uint32_t m = __funnelshift_r(bits[sw], sw + 1 < nwords ? bits[sw + 1] : 0u, sh);
      const int valid = hw - w * 32;
      if (valid < 32) m &= (1u << valid) - 1u;
      if (m != 0u) {
        uint32_t* dst = cplane + (size_t)n * cwords + w;
        if ((*dst & m) != m) atomicOr(dst, m);
      }
    }
  }
  // pass 2: rank of every covered row = exclusive prefix of the popcounts (a thread scans a contiguous span of words)
  const int span = ((nwords + 255) / 256) | 1;        // odd: thread t starts at word t * span -- no shared-memory bank conflicts
  uint32_t local = 0;
  for (int w = tid * span; w < min(nwords, (tid + 1) * span); ++w) local += __popc(bits[w]);
  uint32_t U = 0;
  uint32_t run = block_exclusive_scan(local, warp_sums, &U);
  for (int w = tid * span; w < min(nwords, (tid + 1) * span); ++w) {
    pref[w] = run;
    run += __popc(bits[w]);
  }
  const uint32_t nchunks = U == 0 ? 1u : (U + TR - 1) / TR;                // an empty (view, scale) still closes with bias + ReLU
  if (nchunks > CAPV) {                                                    // uniform: U is the same in every thread
    if (tid == 0) {
      a.tile_ovf[tile] = 1;
      a.tvs[(size_t)tile * a.VS + vs] = make_uint2(0u, 0u);
    }
    return;
  }
  if (tid == 0) base_s[0] = atomicAdd(&a.cursors[0], nchunks);            // chunk ids: the atomic's latency hides behind pass 3
  for (int i = tid; i < (int)nchunks * TQ; i += 256) qmask[i] = 0u;
  __syncthreads();
  BUILD_PHASE(2);
  // pass 3: which slots of which chunk every quad lists
  for (int item = tid; item < p.nl * TC * TC; item += 256)
    cell_walk<false>(recs_s, item % (TC * TC), item / (TC * TC), fw, hw, bits, pref, qmask, nullptr, nullptr, nullptr, nullptr);
  __syncthreads();
  BUILD_PHASE(3);
  // pass 4: entries per chunk, segment starts (16 lanes per chunk, one per quad), blob sizes -> blob offsets; one
  // allocation per CTA
  for (int base = 0; base < (int)nchunks * TQ; base += 256) {
    const int i = base + tid;
    const uint32_t cnt = i < (int)nchunks * TQ ? __popc(qmask[i]) : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < TQ; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d, TQ);
      if ((tid & (TQ - 1)) >= d) inc += t;
    }
    if (i < (int)nchunks * TQ) {
      segbeg[i] = (uint16_t)(inc - cnt);
      if ((i & (TQ - 1)) == TQ - 1) nent_s[i / TQ] = (uint16_t)inc;
    }
  }
  __syncthreads();
  uint32_t nent = 0, bytes16 = 0;
  if (tid < (int)nchunks) {
    nent = nent_s[tid];
    bytes16 = (BLOB_HDR + 16 * nent + ((nent + 15) & ~15u)) >> 4;
  }
  uint32_t total16 = 0;
  // (a chunk with more than MAX_ENT entries would not fit its stage: flagged through the top bit of the scanned sizes)
  const uint32_t my_off16 = block_exclusive_scan(bytes16 | (nent > (uint32_t)MAX_ENT ? 0x40000000u : 0u), warp_sums, &total16);
  if (tid == 0) base_s[1] = atomicAdd(&a.cursors[1], total16);
  __syncthreads();
  const uint32_t base_desc = base_s[0], base_blob16 = base_s[1];
  if (total16 >= 0x40000000u || base_desc + nchunks > a.desc_cap || base_blob16 + total16 > a.blob_cap16) {   // pools exhausted
    if (tid == 0) {
      a.tile_ovf[tile] = 1;
      a.tvs[(size_t)tile * a.VS + vs] = make_uint2(0u, 0u);
    }
    return;
  }
  BUILD_PHASE(4);
  // pass 5: descriptors, blob headers, row offsets
  if (tid < (int)nchunks) {
    const bool last = tid == (int)nchunks - 1;
    const uint32_t info = (last ? INFO_LAST_VS : 0u) | ((last && vs == a.VS - 1) ? INFO_LAST_TILE : 0u) |
                          ((uint32_t)s << 8) | ((uint32_t)(vs / p.S) << 16);
    blob_off16[tid] = base_blob16 + my_off16;
    uint32_t* hd = reinterpret_cast<uint32_t*>(a.hdr + base_desc + tid);    // .z (rows, runs, entries) comes with the runs
    hd[0] = base_blob16 + my_off16;
    hd[1] = bytes16 << 4;
    hd[3] = info;
    uint32_t* bh = reinterpret_cast<uint32_t*>(a.blob + (size_t)(base_blob16 + my_off16) * 16);
    bh[16] = nent;
    bh[17] = info;
    bh[18] = 0u;
    bh[19] = 0u;
  }
  if (tid == 0) {
    a.tvs[(size_t)tile * a.VS + vs] = make_uint2(base_desc, nchunks);
    atomicAdd(a.tile_work + tile, nchunks);
  }
  BUILD_PHASE(5);
  for (int w = tid; w < nwords; w += 256) {
    uint32_t m = bits[w], rank = pref[w];
    while (m) {
      const int b = __ffs(m) - 1;
      rank_idx[rank++] = (uint32_t)(w * 32 + b);
      m &= m - 1;
    }
  }
  __syncthreads();                                                         // blob_off16 is read by every walker
  BUILD_PHASE(6);
  // the quads' segments of every chunk -> blob header (one thread per (chunk, quad))
  for (int i = tid; i < (int)nchunks * TQ; i += 256) {
    const uint32_t beg = segbeg[i];
    reinterpret_cast<uint32_t*>(a.blob + (size_t)blob_off16[i / TQ] * 16)[i & (TQ - 1)] = beg | ((beg + __popc(qmask[i])) << 16);
  }
  // the chunk's rows as runs of consecutive rows (ranks follow the row index, so a run occupies consecutive slots): one
  // warp per chunk, one lane per slot; a lane that starts a run finds the run's end in the ballot of the starts
  {
    const int lane = tid & 31, warp = tid >> 5;
    for (int c = warp; c < (int)nchunks; c += 8) {
      const uint32_t nrows = U == 0 ? 0u : min((uint32_t)TR, U - (uint32_t)c * TR);
      const uint32_t mine = lane < (int)nrows ? rank_idx[c * TR + lane] : 0u;
      const uint32_t prev = __shfl_up_sync(0xffffffffu, mine, 1);
      const bool start = lane < (int)nrows && (lane == 0 || mine != prev + 1u);
      const uint32_t smask = __ballot_sync(0xffffffffu, start);
      if (start) {
        const uint32_t higher = lane == 31 ? 0u : smask & ~((2u << lane) - 1u);
        const uint32_t end = higher ? (uint32_t)__ffs(higher) - 1u : nrows;
        a.rowoff[(size_t)(base_desc + c) * TR + __popc(smask & ((1u << lane) - 1u))] =
            mine | ((uint32_t)lane << 22) | ((end - (uint32_t)lane - 1u) << 27);
      }
      if (lane == 0)
        reinterpret_cast<uint32_t*>(a.hdr + base_desc + c)[2] = nrows | ((uint32_t)__popc(smask) << 8) | ((uint32_t)nent_s[c] << 16);
    }
  }
  BUILD_PHASE(7);
  // pass 6: the entries -- weights zero-initialised, then every cell writes its own component
  for (int c = 0; c < (int)nchunks; ++c) {
    float4* wv = reinterpret_cast<float4*>(a.blob + (size_t)blob_off16[c] * 16 + BLOB_HDR);
    for (int e = tid; e < (int)nent_s[c]; e += 256) wv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  BUILD_PHASE(8);
  for (int item = tid; item < p.nl * TC * TC; item += 256)
    cell_walk<true>(recs_s, item % (TC * TC), item / (TC * TC), fw, hw, bits, pref, qmask, segbeg, blob_off16, nent_s, a.blob);
  BUILD_PHASE(9);
}

// ---- pooling ----------------------------------------------------------------------------------------------------------
struct TilePoolArgs {
  PoolArgs q;
  TileLists t;
  unsigned long long plane_bytes[VFA_MAX_SCALES];   // bytes of one (frame, view) stack of nl planes of Y
  int nb;                      // frames of this chunk
  uint32_t* next_item;         // work counter (zeroed before the launch): items = (frame, tile) pairs handed out in order
  int variant;                 // debug (VFA_TILE_VARIANT): 1 = no row copies (consumer time alone), 2 = consumers apply no
                               // entries (producer / memory time alone)
};

// Producer and consumer warps meet at the CTA-wide barrier that ends a tile from different places of the kernel.
// compute-sanitizer's synccheck reports "divergent thread(s) in block" whenever the threads of a block arrive at a barrier
// from different BAR instructions, protocol errors or not; built with -DVFA_NAMED_BAR_NOINLINE the helper is one function --
// one BAR instruction for every caller -- and synccheck is clean (profiles/r2_sanitizer.md), which is the evidence that the
// arrivals themselves match.  The shipped build inlines it (the call costs 2 % of the kernel in register allocation).
#ifdef VFA_NAMED_BAR_NOINLINE
__device__ __noinline__ void named_bar_sync(int id, int threads) {
#else
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
#endif
  __syncwarp();        // the lanes of a warp arrive together (a producer warp may still be split behind its `if (lane ...)` blocks)
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <typename YT>
__device__ __forceinline__ void load_row8(const uint8_t* row, int lane, float4& va, float4& vb);
template <>
__device__ __forceinline__ void load_row8<float>(const uint8_t* row, int lane, float4& va, float4& vb) {
  va = *reinterpret_cast<const float4*>(row + lane * 16);
  vb = *reinterpret_cast<const float4*>(row + CH * 2 + lane * 16);
}
template <>
__device__ __forceinline__ void load_row8<__nv_bfloat16>(const uint8_t* row, int lane, float4& va, float4& vb) {
  const uint2 a = *reinterpret_cast<const uint2*>(row + lane * 8);         // bf16 -> fp32 is a 16-bit shift (exact)
  const uint2 b = *reinterpret_cast<const uint2*>(row + CH + lane * 8);
  va = make_float4(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xffff0000u), __uint_as_float(a.y << 16),
                   __uint_as_float(a.y & 0xffff0000u));
  vb = make_float4(__uint_as_float(b.x << 16), __uint_as_float(b.x & 0xffff0000u), __uint_as_float(b.y << 16),
                   __uint_as_float(b.y & 0xffff0000u));
}

// out_s layout: cell (quad q, sub c) at [(q * 4 + c) * CH floats]; inside a cell the 16-byte chunk j sits at j ^ key with
// key = the cell's column inside the tile (0..7): the owning warp's accesses (fixed cell, 8 consecutive chunks per quarter
// warp) and the transposed read of the final store (fixed chunk, the 8 cells of a tile row per quarter warp) are both
// conflict-free.
__device__ __forceinline__ int cell_key(int q, int c) { return 2 * (q & 3) + (c & 1); }

template <bool MASK, typename YT>
__global__ void __launch_bounds__(POOL_THREADS, 1) pool_tile_kernel(const TilePoolArgs a) {
  constexpr int ROWB = RowBytes<YT>::value;
  constexpr int STAGE = RowBytes<YT>::stage;
  extern __shared__ uint8_t psm_raw[];
  uint8_t* psm = psm_raw + ((128u - (smem_u32(psm_raw) & 127u)) & 127u);
  float* out_s = reinterpret_cast<float*>(psm);
  uint8_t* stages = psm + (size_t)TC * TC * CH * 4;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(stages + (size_t)TSTAGES * STAGE);
  unsigned long long* empty = full + TSTAGES;
  float* bias_s = reinterpret_cast<float*>(empty + TSTAGES);               // [S][CH]
  const AggParams& p = a.q.p;
  const TileLists& t = a.t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_items = t.n_tiles * a.nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TSTAGES; ++i) {
      mbar_init(&full[i], VFA_TILE_SPLIT_RUNS ? NPROD : 1);
      mbar_init(&empty[i], TQ);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.S * CH; i += blockDim.x) bias_s[i] = __ldg(pick(p.bias, i / CH) + i % CH);
  // Dynamic tile scheduling: tiles differ 2x in cost, a static deal leaves the SMs idle 25 % of the time.  Thread 0 draws
  // the NEXT item from a global counter at the start of every tile (the atomic's latency hides behind the tile) and every
  // warp picks it up behind the CTA-wide barrier that ends the tile.
  __shared__ int item_s[2];
  if (threadIdx.x == 0) item_s[0] = (int)atomicAdd(a.next_item, 1u);
  __syncthreads();

  if (warp >= TQ) {
    // ================= producers: chunk descriptors -> bulk copies =================
    // A bulk copy is issued from uniform registers, one lane at a time, ~100 cycles each, and a chunk has ~9 runs: issued by
    // one warp they put ~900 cycles between "the consumers released the stage" and "the last copy is on its way", and that
    // refill latency is what the leading consumer warps wait for (the ring couples them to the slowest warp).  So every
    // producer warp visits every chunk and issues the runs pj, pj + NPROD, ... (VFA_TILE_SPLIT_RUNS; four warps: -0.06 ms
    // per launch against two warps taking alternate chunks, which is what VFA_TILE_SPLIT_RUNS=0 still builds).  The stage
    // of a chunk is its running index mod TSTAGES.
    const int pj = warp - TQ;
#if VFA_TILE_SPLIT_RUNS
    constexpr uint32_t PSTEP = 1;                  // every producer warp visits every chunk and issues runs pj, pj + NPROD, ...
    uint32_t seq = 0, st = 0, ph = 1u;
#else
    constexpr uint32_t PSTEP = NPROD;
    uint32_t seq = (uint32_t)pj;                   // running index of the next chunk this warp issues
    uint32_t st = (uint32_t)pj % TSTAGES, ph = 1u ^ (((uint32_t)pj / TSTAGES) & 1u);   // parity awaited on empty[st]
#endif
    uint32_t tile_seq0 = 0;                        // running index of the current tile's first chunk
#ifdef VFA_TILE_PROFILE
    long long t_desc = 0, t_empty = 0, t_issue = 0, n_chunks = 0;
    const long long t_begin = clock64();
#endif
    for (int round = 0;; ++round) {
      const int item = item_s[round & 1];
      if (item >= n_items) break;
      const int tile = __ldg(t.order + item / a.nb), bl = item % a.nb;      // heavy tiles first, their frames together
      if (__ldg(t.tile_ovf + tile)) {
        named_bar_sync(2, POOL_THREADS);
        continue;
      }
      const uint2 mine = lane < t.VS ? __ldg(t.tvs + (size_t)tile * t.VS + lane) : make_uint2(0u, 0u);
      uint32_t total = mine.y;                     // chunks of the tile
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
      const uint32_t tile_end = tile_seq0 + total;
      // cursor of the prefetcher: (view-scale, chunk inside it) of running index pf_seq; it runs PREFETCH of this warp's
      // chunks ahead of the issue loop
      uint32_t pf_seq = seq;
      int pf_vs = 0;
      uint32_t pf_k = seq - tile_seq0, pf_first = __shfl_sync(0xffffffffu, mine.x, 0), pf_cnt = __shfl_sync(0xffffffffu, mine.y, 0);
      auto settle = [&]() {                        // carry pf_k over the (view, scale) boundaries
        while (pf_vs < t.VS && pf_k >= pf_cnt) {
          pf_k -= pf_cnt;
          ++pf_vs;
          const int src = pf_vs < t.VS ? pf_vs : 0;
          pf_first = __shfl_sync(0xffffffffu, mine.x, src);
          pf_cnt = __shfl_sync(0xffffffffu, mine.y, src);
        }
      };
      settle();
      uint4 hq[PREFETCH];
      uint32_t rq[PREFETCH];
      auto fetch = [&](uint4& h, uint32_t& r) {
        if (pf_seq < tile_end) {
          const uint32_t id = pf_first + pf_k;
          h = __ldg(t.hdr + id);
          r = lane < TR ? __ldg(t.rowoff + (size_t)id * TR + lane) : 0u;
          pf_seq += PSTEP;
          pf_k += PSTEP;
          settle();
        } else {
          h = make_uint4(0u, 0u, 0u, 0u);
          r = 0u;
        }
      };
#pragma unroll
      for (int i = 0; i < PREFETCH; ++i) fetch(hq[i], rq[i]);
      // the descriptor ring is indexed statically (unrolled by PREFETCH): slot j is consumed and refilled in place, so no
      // register that a load in flight will write is ever moved (a rotating copy would wait for the load every chunk)
      bool more = seq < tile_end;
      while (more) {
#pragma unroll
        for (int j = 0; j < PREFETCH; ++j) {
#ifdef VFA_TILE_PROFILE
          const long long c0 = clock64();
          asm volatile("" ::"r"(hq[j].x), "r"(hq[j].w), "r"(rq[j]) : "memory");      // the descriptor loads have landed
          const long long c1 = clock64();
#endif
          const uint4 h = hq[j];
          const uint32_t r = rq[j];
          fetch(hq[j], rq[j]);
          mbar_wait(&empty[st], ph);
#ifdef VFA_TILE_PROFILE
          const long long c2 = clock64();
#endif
          uint8_t* sb = stages + (size_t)st * STAGE;
          const uint32_t nrows = h.z & 0xffu, nruns = (h.z >> 8) & 0xffu;
          const int s = (int)((h.w >> 8) & 0xffu), v = (int)(h.w >> 16);
#if VFA_TILE_SPLIT_RUNS
          // every producer warp arrives on full[st] with the bytes of ITS runs (warp 0: + the blob), so the stage cannot fill
          // -- and the ring cannot move a whole turn -- without this warp: a warp that had no run in a chunk could otherwise
          // be lapped by the consumers and would then wait on a parity of empty[st] that has come round again
          const bool mine_run = lane < (int)nruns && (lane % NPROD) == pj && !(a.variant & 1);
          const uint32_t my_rows = __reduce_add_sync(0xffffffffu, mine_run ? (r >> 27) + 1u : 0u);
          if (lane == 0) {
            mbar_arrive_expect_tx(&full[st], my_rows * ROWB + (pj == 0 ? h.y : 0u));
            if (pj == 0) bulk_g2s(sb + TR * ROWB, t.blob + (size_t)h.x * 16, h.y, &full[st]);
          }
          __syncwarp();
          if (mine_run) {
#else
          if (lane == 0) {
            mbar_arrive_expect_tx(&full[st], h.y + ((a.variant & 1) ? 0u : nrows * ROWB));
            bulk_g2s(sb + TR * ROWB, t.blob + (size_t)h.x * 16, h.y, &full[st]);
          }
          __syncwarp();
          if (lane < (int)nruns && !(a.variant & 1)) {   // one bulk copy per run of consecutive rows
#endif
            const uint8_t* ybase = reinterpret_cast<const uint8_t*>(pick(a.q.y, s)) +
                                   (size_t)(bl * p.V + v) * pick(a.plane_bytes, s);
            bulk_g2s(sb + ((r >> 22) & 31u) * ROWB, ybase + (size_t)(r & 0x3fffffu) * ROWB, ((r >> 27) + 1u) * ROWB, &full[st]);
          }
          st += PSTEP;
          if (st >= TSTAGES) {
            st -= TSTAGES;
            ph ^= 1u;
          }
          seq += PSTEP;
#ifdef VFA_TILE_PROFILE
          t_desc += c1 - c0;
          t_empty += c2 - c1;
          t_issue += clock64() - c2;
          ++n_chunks;
#endif
          if (seq >= tile_end) {
            more = false;
            break;
          }
        }
      }
      tile_seq0 = tile_end;
      named_bar_sync(2, POOL_THREADS);
    }
#ifdef VFA_TILE_PROFILE
    if (lane == 0 && blockIdx.x == 74)
      printf("cta %3d producer %d: %lld chunks, total %lld cyc; per chunk: descriptor wait %lld, empty wait %lld, issue %lld\n",
             blockIdx.x, pj, n_chunks, clock64() - t_begin, t_desc / max(n_chunks, 1ll), t_empty / max(n_chunks, 1ll),
             t_issue / max(n_chunks, 1ll));
#endif
  } else {
    // ================= consumers: one warp per quad =================
    const int q = warp;
    const int cy0 = 2 * (q >> 2), cx0 = 2 * (q & 3);              // inside the tile
    uint32_t st = 0, ph = 0;                       // stage and the parity the consumers wait for on full[st]
#ifdef VFA_TILE_PROFILE
    long long t_full = 0, t_work = 0, t_epi = 0, t_out = 0, n_chunks = 0, w_big = 0, n_big = 0, w_max = 0, n_first = 0, w_first = 0; bool first_chunk = true, after_vs = false; long long n_big_vs = 0, runs_big = 0, runs_all = 0;
    const long long t_begin = clock64();
#endif
    for (int round = 0;; ++round) {
      const int item = item_s[round & 1];
      if (item >= n_items) break;
      if (threadIdx.x == 0) item_s[(round + 1) & 1] = (int)atomicAdd(a.next_item, 1u);
      const int tile = __ldg(t.order + item / a.nb), bl = item % a.nb;      // heavy tiles first, their frames together
      if (__ldg(t.tile_ovf + tile)) {
        named_bar_sync(2, POOL_THREADS);
        continue;
      }
      const int b = a.q.b0 + bl;
      const int ty0 = (tile / t.tiles_x) * TC, tx0 = (tile % t.tiles_x) * TC;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* o = out_s + (q * 4 + c) * CH;
        *reinterpret_cast<float4*>(o + ((lane ^ cell_key(q, c)) << 2)) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(o + CH / 2 + ((lane ^ cell_key(q, c)) << 2)) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float acc[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
      while (true) {
#ifdef VFA_TILE_PROFILE
        const long long c0 = clock64();
#endif
        mbar_wait(&full[st], ph);
#ifdef VFA_TILE_PROFILE
        const long long c1 = clock64();
#endif
        const uint8_t* sb = stages + (size_t)st * STAGE;
        const uint8_t* meta = sb + TR * ROWB;
        const uint32_t seg = *reinterpret_cast<const uint32_t*>(meta + 4 * q);
        const uint32_t nent = *reinterpret_cast<const uint32_t*>(meta + 64);
        const uint32_t info = *reinterpret_cast<const uint32_t*>(meta + 68);
        const float4* wv = reinterpret_cast<const float4*>(meta + BLOB_HDR);
        const uint8_t* sl = meta + BLOB_HDR + 16 * nent;
        const uint32_t e1 = (a.variant & 2) ? 0u : seg >> 16;
        // entries two at a time: the six row / weight loads of a pair issue back to back ahead of its 32 packed FMAs (a
        // branch-free body -- with the odd tail inside the loop the compiler sinks the second entry's loads below the first
        // entry's FMAs and every entry pays a shared-memory round trip), and the slot bytes of the NEXT pair are fetched
        // before the FMAs so that the slot -> row address -> row chain of a pair starts one level down
        uint32_t e = seg & 0xffffu;
        uint32_t s0 = 0, s1 = 0;
        if (e + 2 <= e1) {
          s0 = sl[e];
          s1 = sl[e + 1];
        }
        for (; e + 2 <= e1; e += 2) {
          const float4 w0 = wv[e], w1 = wv[e + 1];
          float4 va0, vb0, va1, vb1;
          load_row8<YT>(sb + s0 * ROWB, lane, va0, vb0);
          load_row8<YT>(sb + s1 * ROWB, lane, va1, vb1);
          if (e + 4 <= e1) {
            s0 = sl[e + 2];
            s1 = sl[e + 3];
          }
          fside::fma8(acc[0], w0.x, va0, vb0);
          fside::fma8(acc[1], w0.y, va0, vb0);
          fside::fma8(acc[2], w0.z, va0, vb0);
          fside::fma8(acc[3], w0.w, va0, vb0);
          fside::fma8(acc[0], w1.x, va1, vb1);
          fside::fma8(acc[1], w1.y, va1, vb1);
          fside::fma8(acc[2], w1.z, va1, vb1);
          fside::fma8(acc[3], w1.w, va1, vb1);
        }
        if (e < e1) {                                      // odd tail
          const uint32_t st_ = sl[e];
          const float4 w0 = wv[e];
          float4 va0, vb0;
          load_row8<YT>(sb + st_ * ROWB, lane, va0, vb0);
          fside::fma8(acc[0], w0.x, va0, vb0);
          fside::fma8(acc[1], w0.y, va0, vb0);
          fside::fma8(acc[2], w0.z, va0, vb0);
          fside::fma8(acc[3], w0.w, va0, vb0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == TSTAGES) {
          st = 0;
          ph ^= 1u;
        }
#ifdef VFA_TILE_PROFILE
        const long long c2 = clock64();
        t_full += c1 - c0;
        if (c1 - c0 > 1000) { w_big += c1 - c0; ++n_big; if (after_vs) ++n_big_vs; }
        if (c1 - c0 > w_max) w_max = c1 - c0;
        if (first_chunk) { w_first += c1 - c0; ++n_first; first_chunk = false; }
        t_work += c2 - c1;
        ++n_chunks;
#endif
        if (info & INFO_LAST_VS) {
          // + bias, ReLU (vfa_op.py:123-124), sum over scales and views (vfanet.py:79, :82)
          const int s = (int)((info >> 8) & 0xffu), v = (int)(info >> 16);
          const float4 bi0 = *reinterpret_cast<const float4*>(bias_s + s * CH + lane * 4);
          const float4 bi1 = *reinterpret_cast<const float4*>(bias_s + s * CH + CH / 2 + lane * 4);
          const float bb[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
          uint32_t bits4 = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float* o = out_s + (q * 4 + c) * CH + ((lane ^ cell_key(q, c)) << 2);
            float4 o0 = *reinterpret_cast<const float4*>(o);
            float4 o1 = *reinterpret_cast<const float4*>(o + CH / 2);
            float tt[8];
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              tt[i] = acc[c][i] + bb[i];
              bits |= (tt[i] > 0.f ? 1u : 0u) << i;
              acc[c][i] = 0.f;
            }
            o0.x += fmaxf(tt[0], 0.f); o0.y += fmaxf(tt[1], 0.f); o0.z += fmaxf(tt[2], 0.f); o0.w += fmaxf(tt[3], 0.f);
            o1.x += fmaxf(tt[4], 0.f); o1.y += fmaxf(tt[5], 0.f); o1.z += fmaxf(tt[6], 0.f); o1.w += fmaxf(tt[7], 0.f);
            *reinterpret_cast<float4*>(o) = o0;
            *reinterpret_cast<float4*>(o + CH / 2) = o1;
            if (MASK) bits4 |= bits << (8 * c);
          }
          if (MASK) {
            // ReLU bits of the quad's four cells: a lane holds eight nibbles (cell c: channels [4l, 4l + 4) and [128 + 4l,
            // ...) -> nibble 2c, 2c + 1); mask word j of the lane's group of eight (cell j >> 1, half j & 1; global word
            // (j & 1) * 4 + lane / 8) is the nibble j of lanes 0 .. 7 of the group in lane order: an 8 x 8 nibble transpose
            // in three exchange steps, after which every lane stores one word (was: 24 shuffles, 8 stores by 4 lanes)
            uint32_t x = bits4;
            uint32_t y = __shfl_xor_sync(0xffffffffu, x, 4);
            x = (lane & 4) ? (x & 0xffff0000u) | (y >> 16) : (x & 0x0000ffffu) | (y << 16);
            y = __shfl_xor_sync(0xffffffffu, x, 2);
            x = (lane & 2) ? (x & 0xff00ff00u) | ((y >> 8) & 0x00ff00ffu) : (x & 0x00ff00ffu) | ((y << 8) & 0xff00ff00u);
            y = __shfl_xor_sync(0xffffffffu, x, 1);
            x = (lane & 1) ? (x & 0xf0f0f0f0u) | ((y >> 4) & 0x0f0f0f0fu) : (x & 0x0f0f0f0fu) | ((y << 4) & 0xf0f0f0f0u);
            const int j = lane & 7, c = j >> 1;
            const int cy = ty0 + cy0 + (c >> 1), cx = tx0 + cx0 + (c & 1);
            if (cy < p.L && cx < p.W)
              p.mask[((((size_t)b * p.V + v) * p.S + s) * (CH / 32) + (j & 1) * 4 + (lane >> 3)) * p.LW + cy * p.W + cx] = x;
          }
        }
#ifdef VFA_TILE_PROFILE
        t_epi += clock64() - c2;
        after_vs = (info & INFO_LAST_VS) != 0;
#endif
        if (info & INFO_LAST_TILE) break;
      }
#ifdef VFA_TILE_PROFILE
      const long long c3 = clock64();
      first_chunk = true;
#endif
      // ---- the tile's 8 x 8 x 256 block of partial sums -> global memory ----
      if (a.q.out_nhwc) {
        // [B, L, W, C]: a warp writes its own four cells, 2 x 512 contiguous bytes each (no cross-warp exchange)
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cy = ty0 + cy0 + (c >> 1), cx = tx0 + cx0 + (c & 1);
          if (cy >= p.L || cx >= p.W) continue;
          const float* o = out_s + (q * 4 + c) * CH + ((lane ^ cell_key(q, c)) << 2);
          float* g = (a.q.out_mode == 3 ? fside::owner_base(p.out, cy) : p.out) + ((size_t)b * p.LW + cy * p.W + cx) * CH + lane * 4;
          const float4 o0 = *reinterpret_cast<const float4*>(o), o1 = *reinterpret_cast<const float4*>(o + CH / 2);
          if (a.q.out_mode == 0) {
            *reinterpret_cast<float4*>(g) = o0;
            *reinterpret_cast<float4*>(g + CH / 2) = o1;
          } else {                 // fused collective: this GPU's cameras are ADDED into the (peer / multicast) map
            fside::red_add_v4(g, o0, a.q.out_mode);
            fside::red_add_v4(g + CH / 2, o1, a.q.out_mode);
          }
        }
      } else {
        // [B, C, L, W]: warp w stores channels [16w, 16w + 16); one store instruction = one channel x 4 tile rows x 8 cells
        // (four full 32-byte sectors); a lane reads its cell's four channels with one conflict-free LDS.128
        named_bar_sync(1, CONSUMER_THREADS);
        const int col = lane & 7;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = 4 * half + (lane >> 3);
          const int cy = ty0 + row, cx = tx0 + col;
          const bool ok = cy < p.L && cx < p.W;
          const int cell_s = ((row >> 1) * 4 + (col >> 1)) * 4 + (row & 1) * 2 + (col & 1);      // quad * 4 + sub
          float* g = p.out + ((size_t)b * CH) * p.LW + (size_t)cy * p.W + cx;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = 4 * warp + j;                                  // logical 16-byte chunk = channels [4 chunk, +4)
            const float4 o4 = *reinterpret_cast<const float4*>(out_s + cell_s * CH + (((chunk & ~7) | ((chunk ^ col) & 7)) << 2));
            if (ok) {
              // logical chunk j' < 32 holds channels [4j', 4j'+4); chunk 32 + j' holds channels [128 + 4j', ...)
              const int c0 = chunk < 32 ? 4 * chunk : CH / 2 + 4 * (chunk - 32);
              g[(size_t)(c0 + 0) * p.LW] = o4.x;
              g[(size_t)(c0 + 1) * p.LW] = o4.y;
              g[(size_t)(c0 + 2) * p.LW] = o4.z;
              g[(size_t)(c0 + 3) * p.LW] = o4.w;
            }
          }
        }
      }
#ifdef VFA_TILE_PROFILE
      t_out += clock64() - c3;
#endif
      named_bar_sync(2, POOL_THREADS);
    }
#ifdef VFA_TILE_PROFILE
    if (lane == 0 && blockIdx.x == 74 && (warp & 3) == 0)
      printf("cta %3d warp %2d: %lld chunks, total %lld cyc; per chunk: full wait %lld, work %lld, epilogue %lld; tile store total %lld; waits>1000: %lld totalling %lld, max %lld; first-chunk-of-tile waits %lld totalling %lld; big waits right after a (view, scale) end: %lld\n",
             blockIdx.x, warp, n_chunks, clock64() - t_begin, t_full / max(n_chunks, 1ll), t_work / max(n_chunks, 1ll),
             t_epi / max(n_chunks, 1ll), t_out, n_big, w_big, w_max, n_first, w_first, n_big_vs);
#endif
  }
}

// ---- schedule: tiles by descending work ---------------------------------------------------------------------------------
// The pooling kernel deals (tile, frame) items to its CTAs from a global counter; an item is ~1 / 11 of a CTA's share, so the
// order matters: with the heavy tiles (centre of the ground plane, seen by every camera) first and the light ones last the
// CTAs finish within a light tile of each other (longest-processing-time-first).  One CTA sorts (chunks of the tile, tile)
// descending in shared memory: by counting ranks up to RANK_CAP tiles, bitonic up to ORDER_CAP; beyond that identity order.
constexpr int ORDER_CAP = 8192, RANK_CAP = 1024;
#ifndef VFA_ORDER_BINS
#define VFA_ORDER_BINS 8
#endif
constexpr int ORDER_BINS = VFA_ORDER_BINS;

__global__ void __launch_bounds__(1024) tile_order_kernel(const uint32_t* __restrict__ tile_work,
                                                          const uint8_t* __restrict__ tile_ovf, int n_tiles,
                                                          int* __restrict__ order, int mode) {
  extern __shared__ unsigned long long okeys[];
  if (n_tiles > ORDER_CAP || mode == 2) {
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) order[i] = i;
    return;
  }
  if (n_tiles <= RANK_CAP && mode != 1) {
    // few tiles (every shipped rig): rank = number of larger keys, 32-bit keys read four at a time -- ~2 us instead of the
    // ~10 us of the 45 barrier-separated bitonic stages
    // The work is quantised to ORDER_BINS levels and tiles of one level keep their row-major order: neighbours (which
    // share a fifth of their rows) are still pooled at the same time by different SMs and meet in L2, and the tail of the
    // launch is made of the lightest level all the same.
    uint32_t* k32 = reinterpret_cast<uint32_t*>(okeys);
    __shared__ uint32_t wmax_s;
    if (threadIdx.x == 0) wmax_s = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x)
      if (!tile_ovf[i]) atomicMax(&wmax_s, tile_work[i]);
    __syncthreads();
    const uint32_t wmax = wmax_s + 1u;
    const int n4 = (n_tiles + 3) & ~3;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const uint32_t w = i < n_tiles && !tile_ovf[i] ? tile_work[i] : 0u;
      k32[i] = i < n_tiles ? (((uint32_t)((unsigned long long)w * ORDER_BINS / wmax)) << 16) | (0xffffu - (unsigned)i) : 0u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
      const uint32_t mine = k32[i];
      int r = 0;
      for (int j = 0; j < n4; j += 4) {
        const uint4 o = *reinterpret_cast<const uint4*>(k32 + j);
        r += (o.x > mine) + (o.y > mine) + (o.z > mine) + (o.w > mine);
      }
      order[r] = i;
    }
    return;
  }
  int np2 = 1;
  while (np2 < n_tiles) np2 <<= 1;
  // padding keys are 0 and sort behind every real key (the low word ~tile of a real key is never 0)
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    okeys[i] = i < n_tiles ? ((unsigned long long)(tile_ovf[i] ? 0u : tile_work[i]) << 32) | (unsigned)(~(unsigned)i) : 0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = okeys[i], y = okeys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? x < y : x > y) {
            okeys[i] = y;
            okeys[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) order[i] = (int)(~(unsigned)(okeys[i] & 0xffffffffull));
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct TileLayout {
  size_t off_cursors, off_ovf, off_work, off_tvs, off_order, off_hdr, off_rowoff, off_blob, total;
  int tiles_x, n_tiles, VS, max_words;
  uint32_t desc_cap, blob_cap16;
  size_t build_smem;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static bool tile_enabled() { return runtime_config().pool_tile != 0 && runtime_config().pool_quad != 0; }

static TileLayout tile_layout(const AggParams& p) {
  TileLayout L = {};
  L.VS = p.V * p.S;
  // The chunk lists (with the coverage bitmap) cost what the quads' texel lists + cover_mark_kernel cost (0.2 ms) and
  // the pooling of a frame is a third faster from them: they serve every batch size, single frames included (MultiviewC
  // 984 -> 1152 frames/s at batch 1).  VFA_POOL_TILE=0 selects the list kernel.
  if (!tile_enabled() || L.VS > 32) return L;
  size_t max_bits = 0;
  for (int s = 0; s < p.S; ++s) max_bits = max_bits > (size_t)p.nl * p.sc[s].fh * p.sc[s].fw ? max_bits : (size_t)p.nl * p.sc[s].fh * p.sc[s].fw;
  L.max_words = (int)((max_bits + 31) / 32);
  L.build_smem = (size_t)p.nl * TC * TC * 32 + (size_t)L.max_words * 8 + (size_t)CAPV * TQ * 4 + (size_t)CAPV * TQ * 2 +
                 (size_t)CAPV * TR * 4;
  if (L.build_smem > 200 * 1024 || max_bits >= (1u << 22)) {          // very large feature maps: the list kernel serves them
    L = {};
    return L;
  }
  L.tiles_x = (p.W + TC - 1) / TC;
  L.n_tiles = L.tiles_x * ((p.L + TC - 1) / TC);
  // pools: chunks and blob bytes per (tile, view, scale, layer); VFA_POOL_TILE_CAP scales both (tests force overflows)
  const int cap = runtime_config().pool_tile_cap;
  const unsigned long long iters = (unsigned long long)L.n_tiles * L.VS * p.nl;
  unsigned long long dc = iters * 3ull * cap / 100 + L.n_tiles * L.VS, bc16 = iters * (3072ull / 16) * cap / 100 + L.n_tiles * L.VS * 8ull;
  if (dc > 0x7fffffffull) dc = 0x7fffffffull;
  if (bc16 > 0x7fffffffull) bc16 = 0x7fffffffull;
  L.desc_cap = (uint32_t)dc;
  L.blob_cap16 = (uint32_t)bc16;
  size_t o = 0;
  L.off_cursors = o; o += 256;
  L.off_ovf = o;     o += align256((size_t)L.n_tiles);
  L.off_work = o;    o += align256((size_t)L.n_tiles * sizeof(uint32_t));
  L.off_tvs = o;     o += align256((size_t)L.n_tiles * L.VS * sizeof(uint2));
  L.off_order = o;   o += align256((size_t)L.n_tiles * sizeof(int));
  L.off_hdr = o;     o += align256((size_t)L.desc_cap * sizeof(uint4));
  L.off_rowoff = o;  o += align256((size_t)L.desc_cap * TR * sizeof(uint32_t));
  L.off_blob = o;    o += align256((size_t)L.blob_cap16 * 16);
  L.total = o;
  return L;
}

}  // namespace ptile

using namespace ptile;

size_t tile_pool_workspace_bytes(const AggParams& p) { return tile_layout(p).total; }

// chunk lists of every tile from the tap records (once per table; static for fixed cameras)
int launch_tile_build(const AggParams& p, const TapRec* recs, void* ws, void* cover_ws, cudaStream_t st) {
  const TileLayout L = tile_layout(p);
  if (L.total == 0) return VFA_OK;
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  VFA_CUDA(cudaMemsetAsync(w8, 0, L.off_tvs, st));                         // cursors + overflow flags + chunk counts
  BuildArgs a;
  a.p = p;
  a.recs = recs;
  a.cursors = reinterpret_cast<uint32_t*>(w8 + L.off_cursors);
  a.tile_ovf = w8 + L.off_ovf;
  a.tile_work = reinterpret_cast<uint32_t*>(w8 + L.off_work);
  a.tvs = reinterpret_cast<uint2*>(w8 + L.off_tvs);
  a.hdr = reinterpret_cast<uint4*>(w8 + L.off_hdr);
  a.rowoff = reinterpret_cast<uint32_t*>(w8 + L.off_rowoff);
  a.blob = w8 + L.off_blob;
  a.desc_cap = L.desc_cap;
  a.blob_cap16 = L.blob_cap16;
  a.tiles_x = L.tiles_x;
  a.VS = L.VS;
  a.max_words = L.max_words;
  a.cover = reinterpret_cast<uint32_t*>(cover_ws);
  if (cover_ws != nullptr) {
    const CoverMap cm = make_cover_map(p);
    for (int s = 0; s < VFA_MAX_SCALES; ++s) a.cover_base[s] = cm.word_base[s];
    VFA_CUDA(cudaMemsetAsync(cover_ws, 0, (size_t)cm.total_words * sizeof(uint32_t), st));
  }
  VFA_CUDA(cudaFuncSetAttribute(tile_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.build_smem));
  tile_build_kernel<<<L.n_tiles * L.VS, 256, L.build_smem, st>>>(a);
  VFA_LAUNCH_CHECK("tile_build_kernel");
#ifdef VFA_BUILD_PROFILE
  {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_build_prof, sizeof(h));
    const double n = (double)L.n_tiles * L.VS;
    static const char* names[10] = {"zero bitmap", "records + mark", "rank scan", "quad masks walk", "entry counts + alloc",
                                    "headers", "rank -> row index", "runs + descriptors", "zero weights", "fill walk"};
    for (int k = 0; k < 10; ++k) fprintf(stderr, "tile_build phase %d %-22s %8.0f cycles / CTA\n", k, names[k], (double)h[k] / n);
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_build_prof, z, sizeof(z));
  }
#endif
  int np2 = 1;
  while (np2 < L.n_tiles && np2 < ORDER_CAP) np2 <<= 1;
  const size_t osm = (size_t)np2 * sizeof(unsigned long long);
  if (osm > 48 * 1024)
    VFA_CUDA(cudaFuncSetAttribute(tile_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm));
  tile_order_kernel<<<1, np2 < 1024 ? (np2 < 32 ? 32 : np2) : 1024, osm, st>>>(a.tile_work, a.tile_ovf, L.n_tiles,
                                                                            reinterpret_cast<int*>(w8 + L.off_order),
                                                                            runtime_config().tile_order);
  VFA_LAUNCH_CHECK("tile_order_kernel");
  return VFA_OK;
}

// pooling of one frame chunk from the chunk lists; `q` carries the problem, Y, the output and the frame offset
int launch_pool_tile(fside::PoolArgs q, void* ws, int nb, bool y_bf16, cudaStream_t st) {
  const AggParams& p = q.p;
  const TileLayout L = tile_layout(p);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  TilePoolArgs a;
  a.t.tile_ovf = w8 + L.off_ovf;
  a.t.tvs = reinterpret_cast<const uint2*>(w8 + L.off_tvs);
  a.t.hdr = reinterpret_cast<const uint4*>(w8 + L.off_hdr);
  a.t.rowoff = reinterpret_cast<const uint32_t*>(w8 + L.off_rowoff);
  a.t.blob = w8 + L.off_blob;
  a.t.order = reinterpret_cast<const int*>(w8 + L.off_order);
  a.t.tiles_x = L.tiles_x;
  a.t.n_tiles = L.n_tiles;
  a.t.VS = L.VS;
  q.tile_ovf = a.t.tile_ovf;
  q.ptiles_x = L.tiles_x;
  a.q = q;
  for (int s = 0; s < VFA_MAX_SCALES; ++s)
    a.plane_bytes[s] = (unsigned long long)p.nl * p.sc[s].fh * p.sc[s].fw * CH * (y_bf16 ? 2 : 4);
  a.nb = nb;
  a.next_item = reinterpret_cast<uint32_t*>(w8 + L.off_cursors) + 2;
  VFA_CUDA(cudaMemsetAsync(a.next_item, 0, sizeof(uint32_t), st));
  a.variant = runtime_config().tile_variant;
  int sms = device_cache_get(DC_SM_COUNT);
  if (sms == 0) {
    int dev = 0;
    VFA_CUDA(cudaGetDevice(&dev));
    VFA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    device_cache_set(DC_SM_COUNT, sms);
  }
  const long long items = (long long)L.n_tiles * nb;
  const int grid = (int)(items < sms ? items : sms);
  const bool mask = p.mask != nullptr;
#define VFA_LAUNCH_POOL_TILE(M, T)                                                                                   \
  do {                                                                                                               \
    VFA_CUDA(cudaFuncSetAttribute(pool_tile_kernel<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                                  (int)RowBytes<T>::smem));                                                       \
    pool_tile_kernel<M, T><<<grid, POOL_THREADS, RowBytes<T>::smem, st>>>(a);                                     \
  } while (0)
  if (y_bf16) {
    if (mask) VFA_LAUNCH_POOL_TILE(true, __nv_bfloat16);
    else VFA_LAUNCH_POOL_TILE(false, __nv_bfloat16);
  } else {
    if (mask) VFA_LAUNCH_POOL_TILE(true, float);
    else VFA_LAUNCH_POOL_TILE(false, float);
  }
#undef VFA_LAUNCH_POOL_TILE
  VFA_LAUNCH_CHECK("pool_tile_kernel");
  return VFA_OK;
}

// ---- multicast copy (all-gather half of the fused all-reduce) ---------------------------------------------------------
__global__ void __launch_bounds__(256) multicast_copy_kernel(const float4* __restrict__ src, float* mc_dst, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + 4 * i), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

int launch_multicast_copy(const void* src, void* mc_dst, size_t n_bytes, cudaStream_t st) {
  const size_t n16 = n_bytes / 16;
  const int blocks = (int)((n16 + 255) / 256 < 148 * 8 ? (n16 + 255) / 256 : 148 * 8);
  multicast_copy_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float*>(mc_dst), n16);
  VFA_LAUNCH_CHECK("multicast_copy_kernel");
  return VFA_OK;
}

// the flags of the tiles the walking kernel has to complete (and the tile geometry it needs to find them)
void tile_pool_overflow_view(const AggParams& p, void* ws, const uint8_t** tile_ovf, int* tiles_x) {
  const TileLayout L = tile_layout(p);
  *tile_ovf = reinterpret_cast<const uint8_t*>(ws) + L.off_ovf;
  *tiles_x = L.tiles_x;
}

}  // namespace vfa
