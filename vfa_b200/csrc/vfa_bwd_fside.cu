// Backward of the aggregation for C = 256, on the image plane and without scatter atomics.
//
// Transposing the feature-side forward (vfa_fwd_fside.cu):  out = sum_{v,s} relu(b_s + sum_n pool_n(f W_n^T))  gives
//
//   gm[b,v,s][cell, o]       = dOut[b, o, cell] * relu_mask[b, v, s, o, cell]                            (mask_grad_kernel)
//   dBias_s[o]               = sum_{b, v, cell} gm
//   Gs[b,v,s][texel, n, o]   = sum_{cells whose layer-n box holds texel} wy * wx * gm[cell, o]           (dy_gather_kernel)
//   dFeat[b,v,s][texel, c]   = sum_{n,o} Gs[texel, n, o] * W_s[o, c*nl + n]          tcgen05 3xTF32 (ygemm_kernel, MODE 1)
//   dWeight_s[o, c*nl + n]  += sum_{b,v,texel} Gs[texel, n, o] * f[texel, c]          tcgen05 3xTF32, MN-major (dweight_kernel)
//
// The reference's autograd scatters through 4 grid_sampler backwards and two reverse cumsums (vfa_op.py:110-124); the
// first version of this path scattered Gs with atomics (vfa_bwd.cu, still used for C < 256) and spent its time in the L2
// atomic units.  Here the box -> texel relation is inverted once per call into a CSR list (texel row -> (cell, weight)
// entries: count, exclusive scan, fill; the projection is static, so the list is the same for every frame), and Gs is
// GATHERED: one warp per texel row of a layer plane, 8 channels per lane, every row written exactly once, no memset, no
// atomics.  The list has a fixed capacity (12 entries per box on average, 1.6x - 2.3x the rigs' need); rows beyond it
// (none on the three rigs) are completed by overflow_scatter_kernel with atomics, so the result is exact for any rig.
#include <stdlib.h>

#include "vfa_common.cuh"
#include "vfa_umma_ptx.cuh"

namespace vfa {

// vfa_fwd_umma.cu / vfa_fwd_fside.cu / vfa_bwd.cu / vfa_table.cu
int launch_taps_table(const AggParams& p, TapRec* recs, cudaStream_t st);
int launch_ygemm_accum(const float* const* a_rows, float* const* out, const uint8_t* const* wprep_t, const int* rows,
                       int nl, int S, const uint8_t* need, cudaStream_t st);
size_t fside_cover_bytes(const AggParams& p, int frames);
int launch_cover_mark(const AggParams& p, const TapRec* recs, void* cover_ws, cudaStream_t st);
int launch_tile_need(const AggParams& p, void* cover_ws, int frames, const uint8_t** need_out, cudaStream_t st, bool run);
int fs_sgemm_nt_acc(cudaStream_t st, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c, int ldc);
int launch_unprep_dweight(const float* dwr, float* dw, int C, int nl, cudaStream_t st);
int launch_transpose(const float*, float*, long long, int, long long, cudaStream_t);

namespace bfs {

using umma::swz;
using umma::to_tf32;

constexpr int CH = 256;
constexpr int KCH = 32;
constexpr int B_BYTES = CH * KCH * 4;
constexpr int CSR_PER_BOX = 12;            // capacity of the entry list, per (view, scale, layer, cell) box

struct __align__(8) CsrEntry {
  int cell;
  float w;
};

// Rows of the CSR = texels of every (scale, view, layer) plane: row = base[s] + (v * nl + n) * hw[s] + texel.
struct RowMap {
  int base[VFA_MAX_SCALES];
  int hw[VFA_MAX_SCALES];
  int fw[VFA_MAX_SCALES];
  int total;
};

// collapse.weight [C, C*nl] (column c*nl+n) -> per scale, per K chunk kc = n*(C/32) + o/32 a 64 KB block
// [hi: 256 rows (c) x 128 B (32 o), swizzled][lo: same]: the B operand of dFeat = Gs * W.
__global__ void __launch_bounds__(256) prep_weight_umma_t_kernel(const float* __restrict__ w, uint8_t* __restrict__ wp, int nl) {
  const int K = CH * nl;
  const long long total = (long long)CH * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % KCH);
    const int c = (int)((idx / KCH) % CH);
    const int kc = (int)(idx / ((long long)KCH * CH));
    const int n = kc / (CH / KCH), oc = kc % (CH / KCH);
    const int o = oc * KCH + kk;
    const float v = w[(long long)o * K + (long long)c * nl + n];
    const uint32_t hi = to_tf32(v);
    const uint32_t lo = to_tf32(v - __uint_as_float(hi));
    uint8_t* blk = wp + (long long)kc * (2 * B_BYTES);
    const uint32_t off = swz(c, kk >> 2) + (kk & 3) * 4;
    *reinterpret_cast<uint32_t*>(blk + off) = hi;
    *reinterpret_cast<uint32_t*>(blk + B_BYTES + off) = lo;
  }
}

// ---- CSR of the box -> texel relation ----------------------------------------------------------------------------
__device__ __forceinline__ float rec_wx(const TapRec& r, int nx, int i) {
  return i == 0 ? r.wx_first : (i == nx - 1 ? r.wx_last : 1.0f);
}
__device__ __forceinline__ float rec_wy(const TapRec& r, int ny, int i) {
  return i == 0 ? r.wy_first : (i == ny - 1 ? r.wy_last : r.wy_mid);
}

// FILL = false: counts[row] += 1 per tap;  FILL = true: entries[offsets[row] + cursor[row]++] = (cell, weight)
template <bool FILL>
__global__ void __launch_bounds__(256) csr_walk_kernel(AggParams p, const TapRec* __restrict__ recs, RowMap rm,
                                                       int* __restrict__ counts, const int* __restrict__ offsets,
                                                       CsrEntry* __restrict__ entries, int capacity) {
  const long long total = (long long)p.V * p.S * p.nl * p.LW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const TapRec r = recs[idx];
    const int nx = r.nxy & 0xffff, ny = r.nxy >> 16;
    if (nx == 0) continue;
    const int cell = (int)(idx % p.LW);
    const int n = (int)((idx / p.LW) % p.nl);
    const int s = (int)((idx / ((long long)p.LW * p.nl)) % p.S);
    const int v = (int)(idx / ((long long)p.LW * p.nl * p.S));
    const int fw = s == 0 ? rm.fw[0] : (s == 1 ? rm.fw[1] : rm.fw[2]);
    const int hw = s == 0 ? rm.hw[0] : (s == 1 ? rm.hw[1] : rm.hw[2]);
    const int base = s == 0 ? rm.base[0] : (s == 1 ? rm.base[1] : rm.base[2]);
    const int row0 = base + (v * p.nl + n) * hw + (r.xy >> 16) * fw + (r.xy & 0xffff);
    for (int ty = 0; ty < ny; ++ty) {
      const float wy = rec_wy(r, ny, ty);
      for (int tx = 0; tx < nx; ++tx) {
        const float w = wy * rec_wx(r, nx, tx);
        if (w == 0.f) continue;                          // the forward never fetches these taps either
        const int row = row0 + ty * fw + tx;
        if (!FILL) {
          atomicAdd(counts + row, 1);
        } else if (offsets[row + 1] <= capacity) {
          const int pos = atomicAdd(counts + row, 1);
          CsrEntry e;
          e.cell = cell;
          e.w = w;
          entries[offsets[row] + pos] = e;
        }
      }
    }
  }
}

// exclusive scan of n ints in three passes (1024 elements per block)
__global__ void __launch_bounds__(1024) scan_block_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                          int* __restrict__ block_sums, int n) {
  __shared__ int warp_sums[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = i < n ? in[i] : 0;
  int v = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int incl = v + (warp > 0 ? warp_sums[warp - 1] : 0);
  if (i < n) out[i] = incl - x;
  if (threadIdx.x == 1023) block_sums[blockIdx.x] = incl;
}
// Block offsets are summed in 64 bits and SATURATED at INT_MAX: a rig whose boxes hold more than 2^31 taps (close-up
// cameras, many views) must not wrap the 32-bit offsets negative -- a saturated offset exceeds every capacity, so the rows
// behind it are completed by the overflow kernel like any other row the list could not hold.
__device__ __forceinline__ int sat_int(long long x) { return x > 0x7fffffffll ? 0x7fffffff : (int)x; }
__global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ block_sums, int nblocks) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const long long x = i < nblocks ? block_sums[i] : 0;
    long long v = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long incl = v + (warp > 0 ? warp_sums[warp - 1] : 0);
    if (i < nblocks) block_sums[i] = sat_int(carry + incl - x);      // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
}
// out[i] += block offset; out[n] = grand total (both saturating)
__global__ void __launch_bounds__(1024) scan_add_kernel(int* __restrict__ out, const int* __restrict__ block_sums,
                                                        const int* __restrict__ in, int n) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) {
    const int o = sat_int((long long)out[i] + block_sums[blockIdx.x]);
    out[i] = o;
    if (i == n - 1) out[n] = sat_int((long long)o + in[i]);
  }
}

// ---- gm = dOut^T * mask, dBias ---------------------------------------------------------------------------------------
struct GradParams {
  AggParams p;
  const float* gt;           // [nb][LW][C]  dOut of the chunk, transposed
  float* gm;                 // [nb][V][S][LW][C]
  float* gbias[VFA_MAX_SCALES];
  int b0;
};

constexpr int MG_WARPS = 8, MG_CELLS = 16;       // cells per warp

__global__ void __launch_bounds__(MG_WARPS * 32) mask_grad_kernel(const GradParams q) {
  __shared__ float bsum[MG_WARPS][CH];
  const AggParams& p = q.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bvs = blockIdx.y;                                  // (bl * V + v) * S + s
  const int s = bvs % p.S, v = (bvs / p.S) % p.V, bl = bvs / (p.S * p.V);
  const int b = q.b0 + bl;
  // lane l owns channels [4l, 4l+4) and [128 + 4l, 128 + 4l + 4): contiguous 512-byte warp accesses (see pool_quad_kernel);
  // their ReLU bits are nibble (l & 7) of mask word l >> 3 and of word 4 + (l >> 3)
  const uint32_t* mlo = p.mask + ((((size_t)b * p.V + v) * p.S + s) * (CH / 32) + (lane >> 3)) * p.LW;
  const uint32_t* mhi = mlo + (size_t)4 * p.LW;
  const float* grow = q.gt + (size_t)bl * p.LW * CH + lane * 4;
  float* out = q.gm + (size_t)bvs * p.LW * CH + lane * 4;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int cell0 = (blockIdx.x * MG_WARPS + warp) * MG_CELLS;
  for (int k = 0; k < MG_CELLS; ++k) {
    const int cell = cell0 + k;
    if (cell >= p.LW) break;
    const uint32_t blo = (__ldg(mlo + cell) >> (4 * (lane & 7))) & 0xfu, bhi = (__ldg(mhi + cell) >> (4 * (lane & 7))) & 0xfu;
    float4 a = __ldg(reinterpret_cast<const float4*>(grow + (size_t)cell * CH));
    float4 c = __ldg(reinterpret_cast<const float4*>(grow + (size_t)cell * CH + CH / 2));
    a.x = (blo & 1u) ? a.x : 0.f;
    a.y = (blo & 2u) ? a.y : 0.f;
    a.z = (blo & 4u) ? a.z : 0.f;
    a.w = (blo & 8u) ? a.w : 0.f;
    c.x = (bhi & 1u) ? c.x : 0.f;
    c.y = (bhi & 2u) ? c.y : 0.f;
    c.z = (bhi & 4u) ? c.z : 0.f;
    c.w = (bhi & 8u) ? c.w : 0.f;
    *reinterpret_cast<float4*>(out + (size_t)cell * CH) = a;
    *reinterpret_cast<float4*>(out + (size_t)cell * CH + CH / 2) = c;
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += c.x; acc[5] += c.y; acc[6] += c.z; acc[7] += c.w;
  }
  float* gb = s == 0 ? q.gbias[0] : (s == 1 ? q.gbias[1] : q.gbias[2]);
  if (gb == nullptr) return;                                   // uniform per CTA
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[warp][(i < 4 ? 0 : CH / 2 - 4) + lane * 4 + i] = acc[i];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < MG_WARPS; ++w) t += bsum[w][threadIdx.x];
  if (t != 0.f) atomicAdd(gb + threadIdx.x, t);
}

// ---- Gs = gather of gm through the CSR ------------------------------------------------------------------------------
struct GatherParams {
  AggParams p;
  RowMap rm;
  const int* offsets;
  const CsrEntry* entries;
  const float* gm;                  // [nb][V][S][LW][C]
  float* gs[VFA_MAX_SCALES];        // per scale [nb*V planes][texel][nl][C]
  int tile_begin[VFA_MAX_SCALES + 1];   // CTA tiles (4 x 4 texels of one (view, layer) plane) per scale, prefix
  int tiles_x[VFA_MAX_SCALES], tiles[VFA_MAX_SCALES];
  int capacity;
  const uint8_t* need;              // [256-row tile][layer] (launch_tile_need): rows of tiles nobody reads are not written
  int need_tile0[VFA_MAX_SCALES];
};

constexpr int GT = 4;                            // texel tile side: one warp per texel
#ifndef VFA_GATHER_MINBLOCKS
#define VFA_GATHER_MINBLOCKS 3
#endif
#ifndef VFA_GATHER_BATCH
#define VFA_GATHER_BATCH 3
#endif
constexpr int GB = VFA_GATHER_BATCH;             // CSR entries in flight per warp

__device__ __forceinline__ void fma8w(float (&acc)[8], float w, const float4& a, const float4& b) {
  const float2 w2 = make_float2(w, w);
  const float2 r0 = __ffma2_rn(w2, make_float2(a.x, a.y), make_float2(acc[0], acc[1]));
  const float2 r1 = __ffma2_rn(w2, make_float2(a.z, a.w), make_float2(acc[2], acc[3]));
  const float2 r2 = __ffma2_rn(w2, make_float2(b.x, b.y), make_float2(acc[4], acc[5]));
  const float2 r3 = __ffma2_rn(w2, make_float2(b.z, b.w), make_float2(acc[6], acc[7]));
  acc[0] = r0.x; acc[1] = r0.y; acc[2] = r1.x; acc[3] = r1.y;
  acc[4] = r2.x; acc[5] = r2.y; acc[6] = r3.x; acc[7] = r3.y;
}

__global__ void __launch_bounds__(GT * GT * 32, VFA_GATHER_MINBLOCKS) dy_gather_kernel(const GatherParams q) {
  const AggParams& p = q.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bx = blockIdx.x, bl = blockIdx.y;
  const int s = (p.S > 2 && bx >= q.tile_begin[2]) ? 2 : ((p.S > 1 && bx >= q.tile_begin[1]) ? 1 : 0);
  const int tb = s == 0 ? 0 : (s == 1 ? q.tile_begin[1] : q.tile_begin[2]);
  const int tiles = s == 0 ? q.tiles[0] : (s == 1 ? q.tiles[1] : q.tiles[2]);
  const int tiles_x = s == 0 ? q.tiles_x[0] : (s == 1 ? q.tiles_x[1] : q.tiles_x[2]);
  const int fw = s == 0 ? q.rm.fw[0] : (s == 1 ? q.rm.fw[1] : q.rm.fw[2]);
  const int hw = s == 0 ? q.rm.hw[0] : (s == 1 ? q.rm.hw[1] : q.rm.hw[2]);
  const int base = s == 0 ? q.rm.base[0] : (s == 1 ? q.rm.base[1] : q.rm.base[2]);
  float* gs = s == 0 ? q.gs[0] : (s == 1 ? q.gs[1] : q.gs[2]);
  const int vn = (bx - tb) / tiles, tile = (bx - tb) % tiles;
  const int v = vn / p.nl, n = vn % p.nl;
  const int ty = (tile / tiles_x) * GT + warp / GT, tx = (tile % tiles_x) * GT + warp % GT;
  if (tx >= fw || ty * fw + tx >= hw) return;
  const int texel = ty * fw + tx;
  const int row = base + vn * hw + texel;
  const int e0 = __ldg(q.offsets + row), e1 = __ldg(q.offsets + row + 1);
  if (e0 == e1) {       // an empty row inside a tile that neither GEMM of the backward reads: nothing to write
    const int t0 = s == 0 ? q.need_tile0[0] : (s == 1 ? q.need_tile0[1] : q.need_tile0[2]);
    const int tile256 = ((bl * p.V + v) * hw + texel) >> 8;
    if (!__ldg(q.need + (size_t)(t0 + tile256) * p.nl + n)) return;
  }
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (e1 <= q.capacity) {
    // lane l owns channels [4l, 4l+4) and [128 + 4l, 128 + 4l + 4): each LDG.128 of the warp covers 512 contiguous bytes
    // (4 L1 wavefronts; with 8 consecutive channels per lane the 32-byte lane stride made it 8)
    const float* gm = q.gm + (((size_t)bl * p.V + v) * p.S + s) * p.LW * CH + lane * 4;
    const int2* ent = reinterpret_cast<const int2*>(q.entries);
    int e = e0;
    for (; e + GB <= e1; e += GB) {                    // GB entries (2 x LDG.128 each) in flight
      int2 en[GB];
      float4 ga[GB], gb[GB];
#pragma unroll
      for (int k = 0; k < GB; ++k) en[k] = __ldg(ent + e + k);
#pragma unroll
      for (int k = 0; k < GB; ++k) {
        ga[k] = __ldg(reinterpret_cast<const float4*>(gm + (size_t)en[k].x * CH));
        gb[k] = __ldg(reinterpret_cast<const float4*>(gm + (size_t)en[k].x * CH + CH / 2));
      }
#pragma unroll
      for (int k = 0; k < GB; ++k) fma8w(acc, __int_as_float(en[k].y), ga[k], gb[k]);
    }
    for (; e < e1; ++e) {
      const int2 a = __ldg(ent + e);
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(gm + (size_t)a.x * CH));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(gm + (size_t)a.x * CH + CH / 2));
      fma8w(acc, __int_as_float(a.y), a0, a1);
    }
  }
  float* dst = gs + ((((size_t)bl * p.V + v) * hw + texel) * p.nl + n) * CH + lane * 4;
  *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(dst + CH / 2) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// Rows the CSR could not hold (offsets[row + 1] > capacity): completed with atomics.  Exits at once when there are none.
__global__ void __launch_bounds__(256) overflow_scatter_kernel(const GatherParams q, const TapRec* __restrict__ recs, int nb) {
  const AggParams& p = q.p;
  if (__ldg(q.offsets + q.rm.total) <= q.capacity) return;
  const int lane = threadIdx.x & 31;
  const long long boxes = (long long)p.V * p.S * p.nl * p.LW;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < boxes; idx += warps) {
    const TapRec r = recs[idx];
    const int nx = r.nxy & 0xffff, ny = r.nxy >> 16;
    if (nx == 0) continue;
    const int cell = (int)(idx % p.LW);
    const int n = (int)((idx / p.LW) % p.nl);
    const int s = (int)((idx / ((long long)p.LW * p.nl)) % p.S);
    const int v = (int)(idx / ((long long)p.LW * p.nl * p.S));
    const int fw = q.rm.fw[s], hw = q.rm.hw[s];
    const int tex0 = (r.xy >> 16) * fw + (r.xy & 0xffff);
    const int row0 = q.rm.base[s] + (v * p.nl + n) * hw + tex0;
    for (int ty = 0; ty < ny; ++ty) {
      const float wy = rec_wy(r, ny, ty);
      for (int tx = 0; tx < nx; ++tx) {
        const float w = wy * rec_wx(r, nx, tx);
        const int row = row0 + ty * fw + tx;
        if (w == 0.f || __ldg(q.offsets + row + 1) <= q.capacity) continue;
        for (int bl = 0; bl < nb; ++bl) {
          const float* g = q.gm + ((((size_t)bl * p.V + v) * p.S + s) * p.LW + cell) * CH + lane * 8;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(g)), g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
          float* dst = q.gs[s] + ((((size_t)bl * p.V + v) * hw + tex0 + ty * fw + tx) * p.nl + n) * CH + lane * 8;
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(w * g0.x, w * g0.y, w * g0.z, w * g0.w));
          atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(w * g1.x, w * g1.y, w * g1.z, w * g1.w));
        }
      }
    }
  }
}

}  // namespace bfs

// ---- dWeight on the tensor cores ---------------------------------------------------------------------------------------
// dWr_s[(n,o), c] += sum_t Gs_s[t, (n,o)] * F_s[t, c]:  M = nl*256, N = 256, K = texel rows of the chunk (10^5 .. 4*10^5).
// Both operands are stored with the contraction index OUTERMOST ([t][m], [t][c]), i.e. they are MN-major for the MMA:
// a 128-byte run of a texel row (32 consecutive m or c) becomes one row of a SWIZZLE_128B atom, so the producers load
// coalesced 512-byte runs, split into tf32 hi / lo and store 128-bit -- no transposition anywhere (layout type
// SWIZZLE_128B_BASE32B, the one MN-major layout the tensor core accepts for 32-bit data).  CTA pairs
// (cta_group::2): a cluster owns a 256-row M tile and a range of 256-texel K blocks; 3 stages of 32 texels; after every
// K block (32 k-steps: the truncating accumulator is not trusted for more) the 8 epilogue warps fold the accumulator into an
// fp32 running sum in TMEM columns [256, 512); when the cluster's range leaves the (scale, M tile) segment the running sum is
// added to dWr with vector atomics (split-K).  The nl M tiles of one share of the K blocks run side by side.
namespace dwg {

using namespace umma;
using bfs::CH;

constexpr int TILE_M = 128, KST = 32, STAGES = 3;
constexpr int OP_BYTES = TILE_M * KST * 4;                    // 16 KB: one hi or lo operand tile (A or B)
constexpr int STAGE_BYTES = 4 * OP_BYTES;                     // A hi, A lo, B hi, B lo
constexpr int NUM_PRODUCER_WARPS = 16, FIRST_PRODUCER_WARP = 4;
constexpr int FIRST_EPILOGUE_WARP = FIRST_PRODUCER_WARP + NUM_PRODUCER_WARPS;   // 20: (warp & 3) == TMEM lane quarter
constexpr int NUM_EPILOGUE_WARPS = 8;
constexpr int THREADS = (FIRST_EPILOGUE_WARP + NUM_EPILOGUE_WARPS) * 32;       // 896
constexpr int STAGES_PER_BLOCK = 8;                           // 256 texels per K block
constexpr uint32_t LBO = 512, SBO = 2048;                     // 512-byte atoms: 4 along M/N, then 8 along K (4 texels each)
constexpr uint32_t IDESC = make_idesc_tf32(CH, 2 * TILE_M) | (1u << 15) | (1u << 16);     // A and B MN-major

struct __align__(16) SmemTail {
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
  unsigned long long peer_full[STAGES];
  unsigned long long acc_full;
  unsigned long long acc_empty;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(SmemTail);

struct Args {
  const float* gs[VFA_MAX_SCALES];     // [T_s][nl*256]
  const float* f[VFA_MAX_SCALES];      // [T_s][256]
  float* dwr[VFA_MAX_SCALES];          // [nl*256][256], zeroed / accumulated by the caller
  int T[VFA_MAX_SCALES];               // texel rows of the chunk
  int kb_begin[VFA_MAX_SCALES + 1];    // K blocks (256 texels) of scale s: [kb_begin[s], kb_begin[s+1]) of the concatenated list
  int need_tile0[VFA_MAX_SCALES];      // first 256-row tile of scale s in the need bytes (launch_tile_need)
  const uint8_t* need;                 // [tile][layer]: 0 = Gs of that (tile, layer) is all zero -> the K block is skipped
  int nl, S, n_kb;
};

// Work split: cluster c owns M tile (c % nl) and the c / nl -th share of the concatenated K-block list, so the nl
// clusters of a share stream the same feature rows at the same time (they are re-read from L2, not from HBM).
struct Unit {
  int s, kb;
};
__device__ __forceinline__ Unit decode_unit(const Args& a, int u) {
  Unit r;
  r.s = (a.S > 2 && u >= a.kb_begin[2]) ? 2 : ((a.S > 1 && u >= a.kb_begin[1]) ? 1 : 0);
  r.kb = u - (r.s == 0 ? 0 : (r.s == 1 ? a.kb_begin[1] : a.kb_begin[2]));
  return r;
}
__device__ __forceinline__ bool unit_needed(const Args& a, int u, int mt) {
  const Unit r = decode_unit(a, u);
  const int t0 = r.s == 0 ? a.need_tile0[0] : (r.s == 1 ? a.need_tile0[1] : a.need_tile0[2]);
  return __ldg(a.need + (size_t)(t0 + r.kb) * a.nl + mt) != 0;
}

__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint32_t off, const float4& v) {
  uint4 hi, lo;
  hi.x = to_tf32(v.x);
  hi.y = to_tf32(v.y);
  hi.z = to_tf32(v.z);
  hi.w = to_tf32(v.w);
  const float2 l01 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-__uint_as_float(hi.x), -__uint_as_float(hi.y)));
  const float2 l23 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-__uint_as_float(hi.z), -__uint_as_float(hi.w)));
  lo.x = __float_as_uint(l01.x);
  lo.y = __float_as_uint(l01.y);
  lo.z = __float_as_uint(l23.x);
  lo.w = __float_as_uint(l23.y);
  *reinterpret_cast<uint4*>(hi_tile + off) = hi;
  *reinterpret_cast<uint4*>(hi_tile + OP_BYTES + off) = lo;
}

__global__ void __launch_bounds__(THREADS, 1) dweight_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // this cluster's M tile and contiguous range of K blocks
  const int mt = cluster % a.nl, share = cluster / a.nl, shares = n_clusters / a.nl;
  const int u_begin = share < shares ? (int)((long long)a.n_kb * share / shares) : 0;
  const int u_end = share < shares ? (int)((long long)a.n_kb * (share + 1) / shares) : 0;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&tail->full[i], NUM_PRODUCER_WARPS);
      mbar_init(&tail->empty[i], 1);
      mbar_init(&tail->peer_full[i], 1);
    }
    mbar_init(&tail->acc_full, 1);
    mbar_init(&tail->acc_empty, 2 * NUM_EPILOGUE_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 1) {
    if (lane == 0 && cta_rank != 0) {
      int it = 0;
      for (int u = u_begin; u < u_end; ++u) {
        if (!unit_needed(a, u, mt)) continue;
        for (int k = 0; k < STAGES_PER_BLOCK; ++k, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_arrive_remote(&tail->peer_full[st], 0);
        }
      }
    } else if (lane == 0) {
      // ================= MMA issuer (pair leader) =================
      int it = 0, blk = 0;
      for (int u = u_begin; u < u_end; ++u) {
        if (!unit_needed(a, u, mt)) continue;
        mbar_wait_cluster(&tail->acc_empty, (blk & 1) ^ 1);
        tc_fence_after();
        for (int k = 0; k < STAGES_PER_BLOCK; ++k, ++it) {
          const int st = it % STAGES;
          mbar_wait(&tail->full[st], (it / STAGES) & 1);
          mbar_wait_cluster(&tail->peer_full[st], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
#pragma unroll
          for (int kg = 0; kg < KST / 8; ++kg) {           // one MMA K step = 8 texels = two atoms along K
            const uint64_t a_hi = make_desc_mn32(sa + kg * (2 * SBO), LBO, SBO);
            const uint64_t a_lo = make_desc_mn32(sa + OP_BYTES + kg * (2 * SBO), LBO, SBO);
            const uint64_t b_hi = make_desc_mn32(sa + 2 * OP_BYTES + kg * (2 * SBO), LBO, SBO);
            const uint64_t b_lo = make_desc_mn32(sa + 3 * OP_BYTES + kg * (2 * SBO), LBO, SBO);
            tc_mma_tf32_t<true>(tmem, a_lo, b_hi, IDESC, (k | kg) ? 1u : 0u);
            tc_mma_tf32_t<true>(tmem, a_hi, b_lo, IDESC, 1u);
            tc_mma_tf32_t<true>(tmem, a_hi, b_hi, IDESC, 1u);
          }
          tc_commit_t<true>(&tail->empty[st]);
        }
        tc_commit_t<true>(&tail->acc_full);
        ++blk;
      }
    }
  } else if (warp >= FIRST_PRODUCER_WARP && warp < FIRST_EPILOGUE_WARP) {
    // ================= producers: 512-byte runs of Gs rows (A) and F rows (B) -> MN-major hi / lo tiles =================
    const int pw = warp - FIRST_PRODUCER_WARP;
    const int j = lane & 7, mc = lane >> 3;              // 16-byte chunk, 32-element chunk of the 128 M (N) rows
    uint32_t off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = pw * 2 + i;                          // texel of the stage
      off[i] = (uint32_t)((mc + 4 * (k >> 2)) * 512) + swz_mn32((uint32_t)(k & 3), (uint32_t)j);
    }
    float4 ca[2], cbv[2], na[2], nb[2];
    auto load_stage = [&](int u, int k, float4(&va)[2], float4(&vb)[2]) {
      const Unit w = decode_unit(a, u);
      const int T = w.s == 0 ? a.T[0] : (w.s == 1 ? a.T[1] : a.T[2]);
      const float* gs = w.s == 0 ? a.gs[0] : (w.s == 1 ? a.gs[1] : a.gs[2]);
      const float* f = w.s == 0 ? a.f[0] : (w.s == 1 ? a.f[1] : a.f[2]);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int t = w.kb * (STAGES_PER_BLOCK * KST) + k * KST + pw * 2 + i;
        va[i] = vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) {
          va[i] = __ldg(reinterpret_cast<const float4*>(gs + (size_t)t * (a.nl * CH) + mt * (2 * TILE_M) +
                                                        (int)cta_rank * TILE_M + mc * 32 + j * 4));
          vb[i] = __ldg(reinterpret_cast<const float4*>(f + (size_t)t * CH + (int)cta_rank * TILE_M + mc * 32 + j * 4));
        }
      }
    };
    // cursor over the stages of the NEEDED K blocks of this cluster
    auto advance = [&](int& u, int& k) -> bool {
      ++k;
      while (true) {
        if (k >= STAGES_PER_BLOCK) {
          ++u;
          k = 0;
        }
        if (u >= u_end) return false;
        if (k != 0 || unit_needed(a, u, mt)) return true;
        k = STAGES_PER_BLOCK;                             // Gs is all zero here: skip the block
      }
    };
    int u = u_begin, k = -1;
    bool have = u_begin < u_end && advance(u, k);
    if (have) load_stage(u, k, ca, cbv);
    int it = 0;
    while (have) {
      const int st = it % STAGES;
      int nu = u, nk = k;
      const bool more = advance(nu, nk);
      if (more) load_stage(nu, nk, na, nb);
      mbar_wait(&tail->empty[st], ((it / STAGES) & 1) ^ 1);
      uint8_t* base = smem + (size_t)st * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        store_split(base, off[i], ca[i]);
        store_split(base + 2 * OP_BYTES, off[i], cbv[i]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->full[st]);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ca[i] = na[i];
        cbv[i] = nb[i];
      }
      u = nu;
      k = nk;
      have = more;
      ++it;
    }
  } else if (warp >= FIRST_EPILOGUE_WARP) {
    // ================= epilogue: per-K-block drain into the fp32 running sum, split-K flush =================
    const int e = warp - FIRST_EPILOGUE_WARP;
    const int quarter = warp & 3;
    const int col_begin = (e >> 2) * (CH / 2);
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    int blk = 0, cur_s = 0;
    bool first = true;                                    // the running sum holds nothing yet
    auto flush = [&](int s_of) {                          // running sum -> dWr of scale s_of (split-K: vector atomics)
      float* dwr = s_of == 0 ? a.dwr[0] : (s_of == 1 ? a.dwr[1] : a.dwr[2]);
      float* dst = dwr + (size_t)(mt * (2 * TILE_M) + (int)cta_rank * TILE_M + quarter * 32 + lane) * CH;
#pragma unroll 1
      for (int c0 = col_begin; c0 < col_begin + CH / 2; c0 += 32) {
        float v[32];
        tc_ld32(lane_addr + CH + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c0 + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
      }
    };
    for (int u = u_begin; u < u_end; ++u) {
      if (!unit_needed(a, u, mt)) continue;
      const Unit w = decode_unit(a, u);
      if (!first && w.s != cur_s) {                       // the K range moves on to another scale
        flush(cur_s);
        first = true;
      }
      cur_s = w.s;
      mbar_wait_sleep(&tail->acc_full, blk & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = col_begin; c0 < col_begin + CH / 2; c0 += 32) {
        float acc[32], pre[32];
        tc_ld32(lane_addr + c0, acc);
        if (!first) tc_ld32(lane_addr + CH + c0, pre);
        tc_wait_ld();
        if (!first) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += pre[i];
        }
        tc_st32(lane_addr + CH + c0, acc);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) mbar_arrive_remote(&tail->acc_empty, 0);
        else mbar_arrive(&tail->acc_empty);
      }
      first = false;
      ++blk;
    }
    if (!first) flush(cur_s);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace dwg

// dwr_s[nl*256][256] += Gs_s^T F_s for the S scales of one frame chunk (one persistent launch)
static int launch_dweight(const float* const* gs, const float* const* f, float* const* dwr, const int* T, int nl, int S,
                          const uint8_t* need, cudaStream_t st) {
  using namespace dwg;
  VFA_CUDA(cudaFuncSetAttribute(dweight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  Args a;
  a.nl = nl;
  a.S = S;
  a.kb_begin[0] = 0;
  a.need = need;
  int tile0 = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    const int ss = s < S ? s : 0;
    a.need_tile0[s] = tile0;                               // the need bytes cover every scale's tiles (launch_tile_need)
    if (s < S) tile0 += (T[s] + STAGES_PER_BLOCK * KST - 1) / (STAGES_PER_BLOCK * KST);
    a.gs[s] = gs[ss];
    a.f[s] = f[ss];
    a.dwr[s] = dwr[ss];
    a.T[s] = T[ss];
    const int kblocks = (dwr[ss] != nullptr && T[ss] > 0) ? (T[ss] + STAGES_PER_BLOCK * KST - 1) / (STAGES_PER_BLOCK * KST) : 0;
    if (s < S) a.kb_begin[s + 1] = a.kb_begin[s] + kblocks;
  }
  a.n_kb = a.kb_begin[S];
  if (a.n_kb == 0) return VFA_OK;
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
  int resident_clusters = device_cache_get(DC_DWEIGHT_CLUSTERS);
  if (resident_clusters == 0) {
    cfg.gridDim = dim3(2 * 148);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, dweight_kernel, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = 64;
    }
    resident_clusters = n;
    device_cache_set(DC_DWEIGHT_CLUSTERS, n);
  }
  int shares = resident_clusters / nl;
  if (shares < 1) shares = 1;
  if (shares > a.n_kb) shares = a.n_kb;
  cfg.gridDim = dim3(2 * shares * nl);
  VFA_CUDA(cudaLaunchKernelEx(&cfg, dweight_kernel, a));
  VFA_LAUNCH_CHECK("dweight_kernel");
  return VFA_OK;
}

namespace bfs {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Plan {
  size_t px;                 // texels of one view over all scales
  int rows;                  // CSR rows
  int capacity;
  size_t off_wprep, off_dwr, off_recs, off_counts, off_offsets, off_bsums, off_entries, off_cover, off_chunk;
  size_t per_frame;          // bytes of (gT + gm + Gs) for one frame
};

static Plan make_plan(int B, int V, int S, int nl, size_t LW, const int* fh, const int* fw) {
  Plan pl;
  pl.px = 0;
  for (int s = 0; s < S; ++s) pl.px += (size_t)fh[s] * fw[s];
  pl.rows = (int)(pl.px * V * nl);
  int per_box = CSR_PER_BOX;
  if (runtime_config().bwd_csr_per_box > 0) per_box = runtime_config().bwd_csr_per_box;     // tests: force the overflow path
  size_t cap = (size_t)per_box * V * S * nl * LW;
  if (cap > ((size_t)1 << 30)) cap = (size_t)1 << 30;      // entry offsets are 32-bit; rows beyond go to the overflow kernel
  pl.capacity = (int)cap;
  size_t o = 0;
  pl.off_wprep = o;   o += align256((size_t)S * nl * (CH / KCH) * (2 * B_BYTES));
  pl.off_dwr = o;     o += align256((size_t)S * CH * CH * nl * sizeof(float));
  pl.off_recs = o;    o += align256((size_t)V * S * nl * LW * sizeof(TapRec));
  pl.off_counts = o;  o += align256(((size_t)pl.rows + 1) * sizeof(int));
  pl.off_offsets = o; o += align256(((size_t)pl.rows + 1) * sizeof(int));
  pl.off_bsums = o;   o += align256(((size_t)pl.rows / 1024 + 2) * sizeof(int));
  pl.off_entries = o; o += align256((size_t)pl.capacity * sizeof(CsrEntry));
  {
    AggParams q;                                           // coverage bitmap + per-tile need bytes (vfa_fwd_fside.cu)
    q.B = B;
    q.V = V;
    q.S = S;
    q.nl = nl;
    for (int s = 0; s < S; ++s) {
      q.sc[s].fh = fh[s];
      q.sc[s].fw = fw[s];
    }
    pl.off_cover = o;
    o += align256(fside_cover_bytes(q, B));
  }
  pl.off_chunk = o;
  pl.per_frame = (LW * CH + (size_t)V * S * LW * CH + (size_t)V * pl.px * nl * CH) * sizeof(float);
  return pl;
}

static int chunk_frames(const Plan& pl, int B) {
  size_t budget = (size_t)6 << 30;
  if (runtime_config().y_budget_mb > 0) budget = (size_t)runtime_config().y_budget_mb << 20;
  size_t cb = budget / pl.per_frame;
  if (cb < 1) cb = 1;
  if (cb > (size_t)B) cb = (size_t)B;
  return (int)cb;
}

}  // namespace bfs

size_t bwd_fside_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh) {
  const bfs::Plan pl = bfs::make_plan(sh->batch, sh->n_views, sh->n_scales, g->n_layers, (size_t)g->grid_l * g->grid_w, sh->feat_h,
                                      sh->feat_w);
  return pl.off_chunk + (size_t)bfs::chunk_frames(pl, sh->batch) * pl.per_frame + 256;
}

int launch_bwd_fside(AggParams p, const float* const* d_weight, const float* d_grad_out, float* const* d_grad_feats,
                     float* const* d_grad_weight, float* const* d_grad_bias, void* ws, size_t ws_bytes, bool gout_nhwc,
                     cudaStream_t st) {
  using namespace bfs;
  int fh[VFA_MAX_SCALES] = {0, 0, 0}, fwv[VFA_MAX_SCALES] = {0, 0, 0};
  for (int s = 0; s < p.S; ++s) {
    fh[s] = p.sc[s].fh;
    fwv[s] = p.sc[s].fw;
  }
  const Plan pl = make_plan(p.B, p.V, p.S, p.nl, (size_t)p.LW, fh, fwv);
  int cb = chunk_frames(pl, p.B);
  if (ws_bytes < pl.off_chunk + pl.per_frame) {
    set_error("backward workspace %zu < required %zu", ws_bytes, pl.off_chunk + pl.per_frame);
    return VFA_ERR_WORKSPACE;
  }
  if (pl.off_chunk + (size_t)cb * pl.per_frame > ws_bytes) cb = (int)((ws_bytes - pl.off_chunk) / pl.per_frame);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  uint8_t* wprep = w8 + pl.off_wprep;
  float* dwr = reinterpret_cast<float*>(w8 + pl.off_dwr);
  TapRec* recs = reinterpret_cast<TapRec*>(w8 + pl.off_recs);
  int* counts = reinterpret_cast<int*>(w8 + pl.off_counts);
  int* offsets = reinterpret_cast<int*>(w8 + pl.off_offsets);
  int* bsums = reinterpret_cast<int*>(w8 + pl.off_bsums);
  CsrEntry* entries = reinterpret_cast<CsrEntry*>(w8 + pl.off_entries);
  float* gt = reinterpret_cast<float*>(w8 + pl.off_chunk);
  float* gm = gt + (size_t)cb * p.LW * CH;
  float* gs = gm + (size_t)cb * p.V * p.S * p.LW * CH;

  const size_t per_scale_w = (size_t)p.nl * (CH / KCH) * (2 * B_BYTES);
  const size_t per_scale = (size_t)CH * CH * p.nl;
  bool any_w = false, any_f = false;
  for (int s = 0; s < p.S; ++s) {
    VFA_REQUIRE((d_grad_weight[s] == nullptr) == (d_grad_bias[s] == nullptr), VFA_ERR_INVALID_ARGUMENT,
                "scale %d: pass both or neither of d_grad_weight / d_grad_bias", s);
    any_w |= d_grad_weight[s] != nullptr;
    any_f |= d_grad_feats[s] != nullptr;
    if (d_grad_bias[s] != nullptr) VFA_CUDA(cudaMemsetAsync(d_grad_bias[s], 0, CH * sizeof(float), st));
  }
  if (any_f)
    for (int s = 0; s < p.S; ++s) {
      prep_weight_umma_t_kernel<<<148 * 4, 256, 0, st>>>(d_weight[s], wprep + s * per_scale_w, p.nl);
      VFA_LAUNCH_CHECK("prep_weight_umma_t_kernel");
    }
  if (any_w) VFA_CUDA(cudaMemsetAsync(dwr, 0, p.S * per_scale * sizeof(float), st));

  // ---- the box -> texel relation, inverted once per call ----
  if (int rc = launch_taps_table(p, recs, st)) return rc;
  void* cover_ws = w8 + pl.off_cover;
  if (int rc = launch_cover_mark(p, recs, cover_ws, st)) return rc;
  RowMap rm;
  int base = 0;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    rm.base[s] = base;
    rm.hw[s] = s < p.S ? fh[s] * fwv[s] : 1;
    rm.fw[s] = s < p.S ? fwv[s] : 1;
    if (s < p.S) base += p.V * p.nl * rm.hw[s];
  }
  rm.total = base;
  const int nblk = (rm.total + 1023) / 1024;
  VFA_CUDA(cudaMemsetAsync(counts, 0, ((size_t)rm.total + 1) * sizeof(int), st));
  csr_walk_kernel<false><<<148 * 8, 256, 0, st>>>(p, recs, rm, counts, nullptr, nullptr, 0);
  VFA_LAUNCH_CHECK("csr_walk_kernel<count>");
  scan_block_kernel<<<nblk, 1024, 0, st>>>(counts, offsets, bsums, rm.total);
  VFA_LAUNCH_CHECK("scan_block_kernel");
  scan_sums_kernel<<<1, 1024, 0, st>>>(bsums, nblk);
  VFA_LAUNCH_CHECK("scan_sums_kernel");
  scan_add_kernel<<<nblk, 1024, 0, st>>>(offsets, bsums, counts, rm.total);
  VFA_LAUNCH_CHECK("scan_add_kernel");
  VFA_CUDA(cudaMemsetAsync(counts, 0, ((size_t)rm.total + 1) * sizeof(int), st));      // reused as the fill cursor
  csr_walk_kernel<true><<<148 * 8, 256, 0, st>>>(p, recs, rm, counts, offsets, entries, pl.capacity);
  VFA_LAUNCH_CHECK("csr_walk_kernel<fill>");

  VFA_CUDA(cudaFuncSetAttribute(dy_gather_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                (int)cudaSharedmemCarveoutMaxL1));

  for (int b0 = 0; b0 < p.B; b0 += cb) {
    const int nb = p.B - b0 < cb ? p.B - b0 : cb;
    // dOut[b0 .. b0+nb) : [C, LW] -> [LW, C]; a channels-last cotangent (VFA_FLAG_OUT_NHWC) is read in place
    if (!gout_nhwc)
      if (int rc = launch_transpose(d_grad_out + (size_t)b0 * CH * p.LW, gt, nb, CH, p.LW, st)) return rc;
    GradParams gq;
    gq.p = p;
    gq.gt = gout_nhwc ? d_grad_out + (size_t)b0 * p.LW * CH : gt;
    gq.gm = gm;
    gq.b0 = b0;
    for (int s = 0; s < VFA_MAX_SCALES; ++s) gq.gbias[s] = s < p.S ? d_grad_bias[s] : nullptr;
    {
      const dim3 grid((p.LW + MG_WARPS * MG_CELLS - 1) / (MG_WARPS * MG_CELLS), nb * p.V * p.S);
      mask_grad_kernel<<<grid, MG_WARPS * 32, 0, st>>>(gq);
      VFA_LAUNCH_CHECK("mask_grad_kernel");
    }
    const uint8_t* need = nullptr;                   // (256-row tile, layer) pairs whose Gs is not all zero
    if (int rc = launch_tile_need(p, cover_ws, nb, &need, st, true)) return rc;
    GatherParams q;
    q.p = p;
    q.rm = rm;
    q.need = need;
    {
      int t0 = 0;
      for (int s = 0; s < VFA_MAX_SCALES; ++s) {
        q.need_tile0[s] = t0;
        if (s < p.S) t0 += (nb * p.V * rm.hw[s] + 255) / 256;
      }
    }
    q.offsets = offsets;
    q.entries = entries;
    q.gm = gm;
    q.capacity = pl.capacity;
    q.tile_begin[0] = 0;
    size_t gs_off = 0;
    const float* a_rows[VFA_MAX_SCALES];
    float* outs[VFA_MAX_SCALES];
    const uint8_t* wts[VFA_MAX_SCALES];
    int rows[VFA_MAX_SCALES];
    for (int s = 0; s < VFA_MAX_SCALES; ++s) {
      const int ss = s < p.S ? s : 0;
      q.gs[s] = gs + (s < p.S ? gs_off : 0);
      q.tiles_x[s] = (fwv[ss] + GT - 1) / GT;
      q.tiles[s] = q.tiles_x[s] * ((fh[ss] + GT - 1) / GT);
      if (s < p.S) {
        q.tile_begin[s + 1] = q.tile_begin[s] + q.tiles[s] * p.V * p.nl;
        gs_off += (size_t)cb * p.V * rm.hw[s] * p.nl * CH;
      }
      a_rows[s] = q.gs[s];
      outs[s] = d_grad_feats[ss] != nullptr ? d_grad_feats[ss] + (size_t)b0 * p.V * rm.hw[ss] * CH : nullptr;
      wts[s] = wprep + ss * per_scale_w;
      rows[s] = nb * p.V * rm.hw[ss];
    }
    {
      const dim3 grid(q.tile_begin[p.S], nb);
      dy_gather_kernel<<<grid, GT * GT * 32, 0, st>>>(q);
      VFA_LAUNCH_CHECK("dy_gather_kernel");
      overflow_scatter_kernel<<<148 * 8, 256, 0, st>>>(q, recs, nb);
      VFA_LAUNCH_CHECK("overflow_scatter_kernel");
    }
    if (any_f) {
      if (int rc = launch_ygemm_accum(a_rows, outs, wts, rows, p.nl, p.S, need, st)) return rc;
    }
    if (any_w && !runtime_config().bwd_cublas_dw) {
      const float* gsp[VFA_MAX_SCALES];
      const float* fp[VFA_MAX_SCALES];
      float* dwp[VFA_MAX_SCALES];
      int Ts[VFA_MAX_SCALES];
      for (int s = 0; s < VFA_MAX_SCALES; ++s) {
        const int ss = s < p.S ? s : 0;
        gsp[s] = q.gs[ss];
        fp[s] = p.feats[ss] + (size_t)b0 * p.V * rm.hw[ss] * CH;
        dwp[s] = d_grad_weight[ss] != nullptr ? dwr + ss * per_scale : nullptr;
        Ts[s] = nb * p.V * rm.hw[ss];
      }
      if (int rc = launch_dweight(gsp, fp, dwp, Ts, p.nl, p.S, need, st)) return rc;
    } else if (any_w) {
      for (int s = 0; s < p.S; ++s) {
        if (d_grad_weight[s] == nullptr) continue;
        // row-major dWr[nl*C x C] += Gs^T[nl*C x T] * F[T x C]
        const int T = nb * p.V * rm.hw[s];
        if (int rc = fs_sgemm_nt_acc(st, CH, p.nl * CH, T, p.feats[s] + (size_t)b0 * p.V * rm.hw[s] * CH, CH, q.gs[s],
                                     p.nl * CH, dwr + s * per_scale, CH))
          return rc;
      }
    }
  }
  for (int s = 0; s < p.S; ++s) {
    if (d_grad_weight[s] == nullptr) continue;
    if (int rc = launch_unprep_dweight(dwr + s * per_scale, d_grad_weight[s], CH, p.nl, st)) return rc;
  }
  return VFA_OK;
}

}  // namespace vfa
