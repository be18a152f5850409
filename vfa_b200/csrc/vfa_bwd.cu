// Backward of the fused aggregation (sm_100a, fp32 FFMA, any channel count).
//
// What autograd derives for the reference (vfa_op.py:110-124: relu' -> addmm backward -> 4 x grid_sampler backward
// with atomic scatter -> 2 reverse cumsums, keeping 0.6-1.8 GB of intermediates per call) is restated on the direct
// pooling form, recomputing the pooled voxels instead of storing them:
//
//   g[cell, o]      = dOut[b, o, cell] * relu_mask[b, v, s, o, cell]                         (mask saved by the forward)
//   dBias_s[o]      = sum_{b, v, cell} g
//   dWeight_s[o,k]  = sum_{b, v, cell} g[cell, o] * vox_{b,v,s}[cell, k]           k = c*nl + n   (bwd_weight_kernel)
//   dVox[cell, k]   = sum_o g[cell, o] * W_s[o, k]
//   dFeat_{b,v,s}[y, x, c] += sum_{cell, n} wy * wx * dVox[cell, c*nl + n]   over the box taps     (bwd_feature_kernel)
//
// bwd_feature_kernel: one CTA = 64 cells of one frame, loops (view, scale, layer, 32-channel chunk); the per-chunk
// [64 x C] x [C x 32] product runs from shared memory, the result is scattered with coalesced 128-byte atomics
// (lane = channel) -- no reverse cumsum, no integral image.
// bwd_weight_kernel: one CTA owns the [C x 32] block of dWeight for (scale, layer, channel chunk) and a slice of
// the cell tiles; it re-pools its 32 channels for every tile and accumulates a rank-64 update in registers.
#include <stdlib.h>

#include <mutex>

#include "vfa_common.cuh"

namespace vfa {

namespace bwd {
constexpr int TM = 64;          // cells per tile
constexpr int KC = 32;          // channels per chunk
constexpr int OB = 64;          // output-channel block staged in shared memory (feature kernel)
constexpr int THREADS = 256;
constexpr int MAXC = 256;       // weight kernel keeps one [MAXC x KC] block per CTA
constexpr int GSTRIDE = MAXC + 1;   // padded row of the g tile: conflict-free for both the fill and the FFMA reads
}  // namespace bwd
using namespace bwd;

struct BwdParams {
  AggParams p;
  const float* weight[VFA_MAX_SCALES];    // original layout [C, C*nl]
  const float* wq[VFA_MAX_SCALES];        // prepared [nl][C(o)][C(c)]  (c contiguous)
  const float* gout;                      // [B, C, LW]
  float* gfeat[VFA_MAX_SCALES];           // [B, V, fh, fw, C] or nullptr
  float* gweight[VFA_MAX_SCALES];         // [C, C*nl] or nullptr
  float* gbias[VFA_MAX_SCALES];           // [C] or nullptr
  int cell_parts;                         // weight kernel: tiles are split into this many slices
};

// W[o, c*nl + n] -> Wq[n][o][c]
__global__ void __launch_bounds__(256) prep_weight_bwd_kernel(const float* __restrict__ w, float* __restrict__ wq, int C,
                                                              int nl) {
  const long long total = (long long)C * C * nl;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int o = (int)((idx / C) % C);
    const int n = (int)(idx / ((long long)C * C));
    wq[idx] = w[(long long)o * C * nl + (long long)c * nl + n];
  }
}

__device__ __forceinline__ float masked_grad(const BwdParams& q, int b, int v, int s, int o, int cell) {
  const AggParams& p = q.p;
  const uint32_t word =
      p.mask[((((size_t)b * p.V + v) * p.S + s) * ((p.C + 31) / 32) + (o >> 5)) * p.LW + cell];
  return ((word >> (o & 31)) & 1u) ? __ldg(q.gout + ((size_t)b * p.C + o) * p.LW + cell) : 0.f;
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 2) bwd_feature_kernel(const BwdParams q) {
  const AggParams& p = q.p;
  __shared__ BoxTaps taps[TM];
  __shared__ __align__(16) float Gs[OB][TM + 4];     // g[o][cell]
  __shared__ __align__(16) float Wsm[OB][KC];        // Wq[n][o][c0 + c]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cell0 = blockIdx.x * TM, b = blockIdx.y;

  for (int v = 0; v < p.V; ++v) {
    for (int s = 0; s < p.S; ++s) {
      if (q.gfeat[s] == nullptr) continue;
      const ScaleConst sc = p.sc[s];
      float* __restrict__ gfeat = q.gfeat[s] + ((size_t)(b * p.V + v) * sc.fh * sc.fw) * p.C;
      for (int n = 0; n < p.nl; ++n) {
        __syncthreads();
        if (tid < TM) {
          const int cell = cell0 + tid;
          BoxTaps t;
          if (cell < p.LW) {
            t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)v * p.nl + n) * p.LW + cell], sc);
          } else {
            t.x0 = t.y0 = t.nx = t.ny = 0;
            t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
          }
          taps[tid] = t;
        }
        for (int c0 = 0; c0 < p.C; c0 += KC) {
          float dv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = 0.f;
          for (int o0 = 0; o0 < p.C; o0 += OB) {
            __syncthreads();
            // stage g[o0..o0+63][64 cells] and Wq[n][o0..][c0..c0+31]
            for (int e = tid; e < OB * TM; e += THREADS) {
              const int oo = e / TM, r = e % TM;
              const int o = o0 + oo, cell = cell0 + r;
              Gs[oo][r] = (o < p.C && cell < p.LW) ? masked_grad(q, b, v, s, o, cell) : 0.f;
            }
            for (int e = tid; e < OB * KC; e += THREADS) {
              const int oo = e / KC, cc = e % KC;
              const int o = o0 + oo, c = c0 + cc;
              Wsm[oo][cc] = (o < p.C && c < p.C) ? __ldg(q.wq[s] + ((size_t)n * p.C + o) * p.C + c) : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int oo = 0; oo < OB; ++oo) {
              const float4 a0 = *reinterpret_cast<const float4*>(&Gs[oo][warp * 8]);
              const float4 a1 = *reinterpret_cast<const float4*>(&Gs[oo][warp * 8 + 4]);
              const float w = Wsm[oo][lane];
              dv[0] = fmaf(a0.x, w, dv[0]);
              dv[1] = fmaf(a0.y, w, dv[1]);
              dv[2] = fmaf(a0.z, w, dv[2]);
              dv[3] = fmaf(a0.w, w, dv[3]);
              dv[4] = fmaf(a1.x, w, dv[4]);
              dv[5] = fmaf(a1.y, w, dv[5]);
              dv[6] = fmaf(a1.z, w, dv[6]);
              dv[7] = fmaf(a1.w, w, dv[7]);
            }
          }
          // scatter: lane = channel, 8 cells per warp
          const int c = c0 + lane;
          if (c < p.C) {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const BoxTaps t = taps[warp * 8 + i];
              if (t.nx == 0 || dv[i] == 0.f) continue;
              for (int ty = 0; ty < t.ny; ++ty) {
                const float wy = tap_wy(t, ty) * dv[i];
                float* row = gfeat + ((size_t)(t.y0 + ty) * sc.fw + t.x0) * p.C + c;
                for (int tx = 0; tx < t.nx; ++tx) atomicAdd(row + (size_t)tx * p.C, wy * tap_wx(t, tx));
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// grid.x = S * nl * ceil(C/KC), grid.y = cell_parts.  dWeight / dBias must be zeroed before the launch.
__global__ void __launch_bounds__(THREADS, 1) bwd_weight_kernel(const BwdParams q) {
  const AggParams& p = q.p;
  extern __shared__ __align__(16) float sm[];
  float(*Gs)[GSTRIDE] = reinterpret_cast<float(*)[GSTRIDE]>(sm);                   // [TM][MAXC+1] g[cell][o]
  float(*As)[KC] = reinterpret_cast<float(*)[KC]>(sm + TM * GSTRIDE);              // [TM][KC]     vox[cell][c]
  BoxTaps* taps = reinterpret_cast<BoxTaps*>(sm + TM * GSTRIDE + TM * KC);         // [TM]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = (p.C + KC - 1) / KC;
  const int s = blockIdx.x / (p.nl * chunks);
  const int n = (blockIdx.x / chunks) % p.nl;
  const int c0 = (blockIdx.x % chunks) * KC;
  if (q.gweight[s] == nullptr) return;
  const ScaleConst sc = p.sc[s];
  const int tiles = (p.LW + TM - 1) / TM;
  const int total = p.B * p.V * tiles;
  const int per = (total + q.cell_parts - 1) / q.cell_parts;
  const int begin = blockIdx.y * per, end = min(total, begin + per);

  float acc[8][4];      // o = lane + 32*j, c = c0 + warp*4 + i
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  const bool do_bias = (n == 0 && c0 == 0 && q.gbias[s] != nullptr);

  for (int work = begin; work < end; ++work) {
    const int tile = work % tiles;
    const int v = (work / tiles) % p.V;
    const int b = work / (tiles * p.V);
    const int cell0 = tile * TM;
    const float* __restrict__ feat = p.feats[s] + ((size_t)(b * p.V + v) * sc.fh * sc.fw) * p.C;
    __syncthreads();
    if (tid < TM) {
      const int cell = cell0 + tid;
      BoxTaps t;
      if (cell < p.LW) {
        t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)v * p.nl + n) * p.LW + cell], sc);
      } else {
        t.x0 = t.y0 = t.nx = t.ny = 0;
        t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
      }
      taps[tid] = t;
    }
    for (int e = tid; e < TM * MAXC; e += THREADS) {
      const int r = e % TM, o = e / TM;          // consecutive threads -> consecutive cells (coalesced dOut / mask reads)
      const int cell = cell0 + r;
      Gs[r][o] = (o < p.C && cell < p.LW) ? masked_grad(q, b, v, s, o, cell) : 0.f;
    }
    __syncthreads();
    // re-pool the chunk: warp w handles cells w*8..w*8+7, lane = channel
    {
      const int c = c0 + lane;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i;
        const BoxTaps t = taps[r];
        float sum = 0.f;
        if (c < p.C) {
          for (int ty = 0; ty < t.ny; ++ty) {
            const float wy = tap_wy(t, ty);
            const float* row = feat + ((size_t)(t.y0 + ty) * sc.fw + t.x0) * p.C + c;
            float rs = 0.f;
            for (int tx = 0; tx < t.nx; ++tx) rs = fmaf(tap_wx(t, tx), __ldg(row + (size_t)tx * p.C), rs);
            sum = fmaf(wy, rs, sum);
          }
        }
        As[r][lane] = sum;
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < TM; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&As[r][warp * 4]);
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = Gs[r][lane + 32 * j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j][0] = fmaf(g[j], a.x, acc[j][0]);
        acc[j][1] = fmaf(g[j], a.y, acc[j][1]);
        acc[j][2] = fmaf(g[j], a.z, acc[j][2]);
        acc[j][3] = fmaf(g[j], a.w, acc[j][3]);
        if (do_bias && warp == 0) bsum[j] += g[j];
      }
    }
  }
  // dWeight[o, c*nl + n] += acc
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = lane + 32 * j;
    if (o >= p.C) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + warp * 4 + i;
      if (c < p.C) atomicAdd(q.gweight[s] + (size_t)o * p.K + (size_t)c * p.nl + n, acc[j][i]);
    }
    if (do_bias && warp == 0) atomicAdd(q.gbias[s] + o, bsum[j]);
  }
}

// ============================================================================================================
// Feature-side backward (C % 8 == 0): the pooling is linear, so instead of back-propagating through the collapse on
// the grid side (2 x 335 GFLOP per MultiviewC frame) the masked output gradient is first pooled BACK onto the image
// plane with the same box weights,
//     Gs_{b,v,s}[texel, n, o] = sum_{cells whose layer-n box covers texel} wy * wx * g[cell, o]         (scatter_g_kernel)
// and both gradients become plain dense GEMMs over texels (2 x 87 GFLOP, 3.9x fewer, no second scatter):
//     dFeat_{b,v,s}[texel, c]  = sum_{n,o} Gs[texel, n, o] * W_s[o, c*nl + n]                  [T x nl*C] x [nl*C x C]
//     dWeight_s[o, c*nl + n]  += sum_texel Gs[texel, n, o] * feat_{b,v,s}[texel, c]            [nl*C x T] x [T x C]
// The two GEMMs are plain library GEMMs (cuBLAS SGEMM, fp32 math; resolved with dlopen so the library has no link-time
// dependency); the scatter, the bias reduction and the layout kernels are this file's.
// ============================================================================================================
#include <dlfcn.h>

namespace fs {

typedef void* cublasHandle_t;
typedef int (*cublasCreate_t)(cublasHandle_t*);
typedef int (*cublasSetStream_t)(cublasHandle_t, cudaStream_t);
typedef int (*cublasSetMathMode_t)(cublasHandle_t, int);
typedef int (*cublasSgemm_t)(cublasHandle_t, int, int, int, int, int, const float*, const float*, int, const float*, int,
                             const float*, float*, int);
struct Cublas {
  cublasHandle_t handle = nullptr;
  cublasSetStream_t set_stream = nullptr;
  cublasSgemm_t sgemm = nullptr;
  bool tried = false;
};
static Cublas g_cublas[64];     // one lazily created handle per device, created under a mutex
static std::mutex g_cublas_mutex;

static int get_cublas(Cublas** out) {
  int dev = 0;
  VFA_CUDA(cudaGetDevice(&dev));
  VFA_REQUIRE(dev >= 0 && dev < 64, VFA_ERR_UNSUPPORTED, "device index %d beyond the 64 this library caches handles for", dev);
  std::lock_guard<std::mutex> lock(g_cublas_mutex);
  Cublas& c = g_cublas[dev];
  if (!c.tried) {
    c.tried = true;
    void* lib = dlopen("libcublas.so.12", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libcublas.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      auto create = (cublasCreate_t)dlsym(lib, "cublasCreate_v2");
      auto math = (cublasSetMathMode_t)dlsym(lib, "cublasSetMathMode");
      c.set_stream = (cublasSetStream_t)dlsym(lib, "cublasSetStream_v2");
      c.sgemm = (cublasSgemm_t)dlsym(lib, "cublasSgemm_v2");
      if (create && c.set_stream && c.sgemm && create(&c.handle) == 0) {
        if (math) math(c.handle, 0 /* CUBLAS_DEFAULT_MATH: fp32, no TF32 */);
      } else {
        c.handle = nullptr;
      }
    }
  }
  VFA_REQUIRE(c.handle != nullptr, VFA_ERR_CUDA, "libcublas.so.12 could not be loaded / initialised (needed by the backward)");
  *out = &c;
  return VFA_OK;
}

struct ScatterParams {
  AggParams p;
  const float* gt;          // dOut transposed: [LW, C] of frame b
  float* gs;                // Gs: scales concatenated, [texel][nl][C]
  size_t gs_off[VFA_MAX_SCALES];   // element offset of scale s inside gs
  float* gbias[VFA_MAX_SCALES];
  int b, v;
};

// One warp per (scale, layer, cell): lane l owns channels 8l .. 8l+7 (C <= 256, C % 8 == 0).
__global__ void __launch_bounds__(256) scatter_g_kernel(const ScatterParams q) {
  const AggParams& p = q.p;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)p.S * p.nl * p.LW;
  if (warp_global >= total) return;
  const int cell = (int)(warp_global % p.LW);
  const int n = (int)((warp_global / p.LW) % p.nl);
  const int s = (int)(warp_global / ((long long)p.LW * p.nl));
  const ScaleConst sc = p.sc[s];
  const BoxTaps t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)q.v * p.nl + n) * p.LW + cell], sc);
  const int c0 = lane * 8;
  const bool active = c0 < p.C;
  float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
  if (active) {
    const uint32_t word = p.mask[((((size_t)q.b * p.V + q.v) * p.S + s) * ((p.C + 31) / 32) + (c0 >> 5)) * p.LW + cell];
    const uint32_t bits = (word >> (c0 & 31)) & 0xffu;
    const float4 a = __ldg(reinterpret_cast<const float4*>(q.gt + (size_t)cell * p.C + c0));
    const float4 bq = __ldg(reinterpret_cast<const float4*>(q.gt + (size_t)cell * p.C + c0 + 4));
    g0 = make_float4((bits & 1) ? a.x : 0.f, (bits & 2) ? a.y : 0.f, (bits & 4) ? a.z : 0.f, (bits & 8) ? a.w : 0.f);
    g1 = make_float4((bits & 16) ? bq.x : 0.f, (bits & 32) ? bq.y : 0.f, (bits & 64) ? bq.z : 0.f,
                     (bits & 128) ? bq.w : 0.f);
    // dBias_s[o] = sum over (frame, view, cell) of g: layer 0 only so every cell counts once
    if (n == 0 && q.gbias[s] != nullptr) {
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (gv[i] != 0.f) atomicAdd(q.gbias[s] + c0 + i, gv[i]);
    }
  }
  if (!active || t.nx == 0) return;
  float* base = q.gs + q.gs_off[s] + (((size_t)t.y0 * sc.fw + t.x0) * p.nl + n) * p.C + c0;
  const size_t tx_stride = (size_t)p.nl * p.C, ty_stride = (size_t)sc.fw * p.nl * p.C;
  for (int ty = 0; ty < t.ny; ++ty) {
    const float wy = tap_wy(t, ty);
    for (int tx = 0; tx < t.nx; ++tx) {
      const float w = wy * tap_wx(t, tx);
      float* dst = base + ty * ty_stride + tx * tx_stride;
      atomicAdd(reinterpret_cast<float4*>(dst), make_float4(w * g0.x, w * g0.y, w * g0.z, w * g0.w));
      atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(w * g1.x, w * g1.y, w * g1.z, w * g1.w));
    }
  }
}

// Tile-local version of the scatter: a CTA takes a compact 8 x 16 block of BEV cells of one (scale, layer) and 64
// output channels.  Neighbouring cells pool from overlapping texel windows (4x / 7x / 15x overlap at strides 8 / 16 /
// 32 on the MultiviewC rig), so the contributions are first accumulated in shared memory over the block's bounding
// texel region and only the region is flushed with global atomics -- the L2 atomic units (~95 G fp32 adds/s) were the
// whole cost of the untiled kernel.  Blocks whose region does not fit (voxels next to a camera) scatter directly.
constexpr int SC_TW = 16, SC_TH = 8, SC_CELLS = SC_TW * SC_TH;     // cells per CTA
constexpr int SC_CH = 64;                                          // channels per CTA (2 per lane)
constexpr int SC_RMAX = 448;                                       // texels of the shared-memory region (112 KB)

__global__ void __launch_bounds__(256) scatter_g_tiled_kernel(const ScatterParams q, int tiles_x) {
  const AggParams& p = q.p;
  extern __shared__ __align__(16) float sacc[];                    // [SC_RMAX][SC_CH]
  __shared__ BoxTaps taps[SC_CELLS];
  __shared__ int cells[SC_CELLS];
  __shared__ int rb[4];                                            // region x0, y0, x1, y1 (inclusive-exclusive)
  __shared__ float bsum[8][SC_CH];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.y / p.nl, n = blockIdx.y % p.nl;
  const int c0 = blockIdx.z * SC_CH + lane * 2;
  const ScaleConst sc = p.sc[s];
  const int ty0 = (blockIdx.x / tiles_x) * SC_TH, tx0 = (blockIdx.x % tiles_x) * SC_TW;
  if (tid == 0) {
    rb[0] = rb[1] = 1 << 30;
    rb[2] = rb[3] = -1;
  }
  __syncthreads();
  if (tid < SC_CELLS) {
    const int y = ty0 + tid / SC_TW, x = tx0 + tid % SC_TW;
    BoxTaps t;
    t.x0 = t.y0 = t.nx = t.ny = 0;
    t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
    int cell = -1;
    if (y < p.L && x < p.W) {
      cell = y * p.W + x;
      t = derive_taps(reinterpret_cast<const float4*>(p.boxes)[((size_t)q.v * p.nl + n) * p.LW + cell], sc);
    }
    taps[tid] = t;
    cells[tid] = cell;
    if (t.nx > 0) {
      atomicMin(&rb[0], t.x0);
      atomicMin(&rb[1], t.y0);
      atomicMax(&rb[2], t.x0 + t.nx);
      atomicMax(&rb[3], t.y0 + t.ny);
    }
  }
  __syncthreads();
  const int rx0 = rb[0], ry0 = rb[1], rw = rb[2] - rb[0], rh = rb[3] - rb[1];
  const bool any = rb[2] > 0;
  const bool local = any && rw * rh <= SC_RMAX;
  if (local) {
    for (int i = tid; i < rw * rh * SC_CH / 4; i += 256) reinterpret_cast<float4*>(sacc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const bool chan_ok = c0 < p.C;
  float2 bias_acc = make_float2(0.f, 0.f);
  for (int i = 0; i < SC_CELLS / 8; ++i) {
    const int r = warp * (SC_CELLS / 8) + i;
    const int cell = cells[r];
    if (cell < 0 || !chan_ok) continue;
    const uint32_t word = p.mask[((((size_t)q.b * p.V + q.v) * p.S + s) * ((p.C + 31) / 32) + (c0 >> 5)) * p.LW + cell];
    const uint32_t bits = (word >> (c0 & 31)) & 3u;
    const float2 gv = __ldg(reinterpret_cast<const float2*>(q.gt + (size_t)cell * p.C + c0));
    const float g0 = (bits & 1) ? gv.x : 0.f, g1 = (bits & 2) ? gv.y : 0.f;
    bias_acc.x += g0;
    bias_acc.y += g1;
    const BoxTaps t = taps[r];
    if (t.nx == 0 || (g0 == 0.f && g1 == 0.f)) continue;
    if (local) {
      float* base = sacc + ((size_t)(t.y0 - ry0) * rw + (t.x0 - rx0)) * SC_CH + lane * 2;
      for (int ty = 0; ty < t.ny; ++ty) {
        const float wy = tap_wy(t, ty);
        for (int tx = 0; tx < t.nx; ++tx) {
          const float w = wy * tap_wx(t, tx);
          float* d = base + ((size_t)ty * rw + tx) * SC_CH;
          atomicAdd(d, w * g0);
          atomicAdd(d + 1, w * g1);
        }
      }
    } else {
      float* base = q.gs + q.gs_off[s] + (((size_t)t.y0 * sc.fw + t.x0) * p.nl + n) * p.C + c0;
      const size_t tx_stride = (size_t)p.nl * p.C, ty_stride = (size_t)sc.fw * p.nl * p.C;
      for (int ty = 0; ty < t.ny; ++ty) {
        const float wy = tap_wy(t, ty);
        for (int tx = 0; tx < t.nx; ++tx) {
          const float w = wy * tap_wx(t, tx);
          atomicAdd(reinterpret_cast<float2*>(base + ty * ty_stride + tx * tx_stride), make_float2(w * g0, w * g1));
        }
      }
    }
  }
  // dBias_s[o] = sum over (frame, view, cell) of g: layer 0 only so every cell counts once; one atomic per CTA channel
  if (n == 0 && q.gbias[s] != nullptr) {
    bsum[warp][lane * 2] = bias_acc.x;
    bsum[warp][lane * 2 + 1] = bias_acc.y;
  }
  __syncthreads();
  if (n == 0 && q.gbias[s] != nullptr && tid < SC_CH) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += bsum[w][tid];
    const int o = blockIdx.z * SC_CH + tid;
    if (o < p.C && t != 0.f) atomicAdd(q.gbias[s] + o, t);
  }
  if (local) {
    // flush the region: 16 float4 per texel, coalesced vector atomics
    const int quads = SC_CH / 4;
    for (int i = tid; i < rw * rh * quads; i += 256) {
      const int tex = i / quads, qd = i % quads;
      const float4 v4 = reinterpret_cast<const float4*>(sacc)[i];
      if (v4.x == 0.f && v4.y == 0.f && v4.z == 0.f && v4.w == 0.f) continue;
      const int o = blockIdx.z * SC_CH + qd * 4;
      if (o >= p.C) continue;
      const int ty = ry0 + tex / rw, tx = rx0 + tex % rw;
      float* d = q.gs + q.gs_off[s] + (((size_t)ty * sc.fw + tx) * p.nl + n) * p.C + o;
      atomicAdd(reinterpret_cast<float4*>(d), v4);
    }
  }
}

// dWr[n][o][c] -> dW[o, c*nl + n]
__global__ void __launch_bounds__(256) unprep_dweight_kernel(const float* __restrict__ dwr, float* __restrict__ dw, int C,
                                                             int nl) {
  const long long total = (long long)C * C * nl;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % nl);
    const int c = (int)((idx / nl) % C);
    const int o = (int)(idx / ((long long)nl * C));
    dw[idx] = dwr[((long long)n * C + o) * C + c];
  }
}

}  // namespace fs

int launch_transpose(const float*, float*, long long, int, long long, cudaStream_t);

// row-major C[n x m] += B^T[n x k] * A[k x m]   (A: [k, lda >= m], B: [k, ldb >= n]) through the lazily loaded cuBLAS:
// the dWeight product of the feature-side backwards (fp32 SGEMM, no TF32)
int fs_sgemm_nt_acc(cudaStream_t st, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c, int ldc) {
  fs::Cublas* cb = nullptr;
  if (int rc = fs::get_cublas(&cb)) return rc;
  VFA_REQUIRE(cb->set_stream(cb->handle, st) == 0, VFA_ERR_CUDA, "cublasSetStream failed");
  const float one = 1.f;
  VFA_REQUIRE(cb->sgemm(cb->handle, 0, 1, m, n, k, &one, a, lda, b, ldb, &one, c, ldc) == 0, VFA_ERR_CUDA,
              "cublasSgemm (dWeight) failed");
  return VFA_OK;
}

int launch_unprep_dweight(const float* dwr, float* dw, int C, int nl, cudaStream_t st) {
  const size_t per_scale = (size_t)C * C * nl;
  const int blocks = (int)((per_scale + 255) / 256 < 148 * 8 ? (per_scale + 255) / 256 : 148 * 8);
  fs::unprep_dweight_kernel<<<blocks, 256, 0, st>>>(dwr, dw, C, nl);
  VFA_LAUNCH_CHECK("unprep_dweight_kernel");
  return VFA_OK;
}

// C = 256: gather-form backward with the tcgen05 dFeature product (vfa_bwd_fside.cu)
size_t bwd_fside_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh);
int launch_bwd_fside(AggParams p, const float* const* d_weight, const float* d_grad_out, float* const* d_grad_feats,
                     float* const* d_grad_weight, float* const* d_grad_bias, void* ws, size_t ws_bytes, bool gout_nhwc, cudaStream_t st);
static bool gather_backward(int channels) { return channels == 256 && !runtime_config().bwd_scatter; }

static size_t fs_gs_elems(const AggParams& p) {
  size_t px = 0;
  for (int s = 0; s < p.S; ++s) px += (size_t)p.sc[s].fh * p.sc[s].fw;
  return px * p.nl * p.C;
}

static int launch_bwd_feature_side(AggParams p, const float* const* d_weight, const float* d_grad_out,
                                   float* const* d_grad_feats, float* const* d_grad_weight, float* const* d_grad_bias,
                                   void* ws, cudaStream_t st) {
  fs::Cublas* cb = nullptr;
  if (int rc = fs::get_cublas(&cb)) return rc;
  VFA_REQUIRE(cb->set_stream(cb->handle, st) == 0, VFA_ERR_CUDA, "cublasSetStream failed");
  const size_t per_scale = (size_t)p.C * p.C * p.nl;
  const int Kp = p.nl * p.C;
  // workspace: [wq: S x nl x C x C][dwr: S x nl x C x C][gt: LW x C][gs]
  float* wq = reinterpret_cast<float*>(ws);
  float* dwr = wq + p.S * per_scale;
  float* gt = dwr + p.S * per_scale;
  float* gs = gt + (size_t)p.LW * p.C;
  const size_t gs_elems = fs_gs_elems(p);
  fs::ScatterParams q;
  q.p = p;
  q.gt = gt;
  q.gs = gs;
  size_t off = 0;
  bool any_w = false;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    q.gs_off[s] = off;
    q.gbias[s] = nullptr;
    if (s < p.S) {
      off += (size_t)p.sc[s].fh * p.sc[s].fw * p.nl * p.C;
      q.gbias[s] = d_grad_bias[s];
      VFA_REQUIRE((d_grad_weight[s] == nullptr) == (d_grad_bias[s] == nullptr), VFA_ERR_INVALID_ARGUMENT,
                  "scale %d: pass both or neither of d_grad_weight / d_grad_bias", s);
      const int blocks = (int)((per_scale + 255) / 256 < 148 * 8 ? (per_scale + 255) / 256 : 148 * 8);
      prep_weight_bwd_kernel<<<blocks, 256, 0, st>>>(d_weight[s], wq + s * per_scale, p.C, p.nl);
      VFA_LAUNCH_CHECK("prep_weight_bwd_kernel");
      if (d_grad_weight[s] != nullptr) {
        any_w = true;
        VFA_CUDA(cudaMemsetAsync(d_grad_bias[s], 0, (size_t)p.C * sizeof(float), st));
      }
    }
  }
  if (any_w) VFA_CUDA(cudaMemsetAsync(dwr, 0, p.S * per_scale * sizeof(float), st));
  const float one = 1.f, zero = 0.f;
  const long long warps = (long long)p.S * p.nl * p.LW;
  const int sblocks = (int)((warps * 32 + 255) / 256);
  for (int b = 0; b < p.B; ++b) {
    // dOut[b] : [C, LW] -> [LW, C]
    if (int rc = launch_transpose(d_grad_out + (size_t)b * p.C * p.LW, gt, 1, p.C, p.LW, st)) return rc;
    for (int v = 0; v < p.V; ++v) {
      VFA_CUDA(cudaMemsetAsync(gs, 0, gs_elems * sizeof(float), st));
      q.b = b;
      q.v = v;
      if (runtime_config().bwd_untiled) {
        fs::scatter_g_kernel<<<sblocks, 256, 0, st>>>(q);
        VFA_LAUNCH_CHECK("scatter_g_kernel");
      } else {
        const int tiles_x = (p.W + fs::SC_TW - 1) / fs::SC_TW, tiles_y = (p.L + fs::SC_TH - 1) / fs::SC_TH;
        const size_t smem = (size_t)fs::SC_RMAX * fs::SC_CH * sizeof(float);
        VFA_CUDA(cudaFuncSetAttribute(fs::scatter_g_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(tiles_x * tiles_y, p.S * p.nl, (p.C + fs::SC_CH - 1) / fs::SC_CH);
        fs::scatter_g_tiled_kernel<<<grid, 256, smem, st>>>(q, tiles_x);
        VFA_LAUNCH_CHECK("scatter_g_tiled_kernel");
      }
      for (int s = 0; s < p.S; ++s) {
        const int T = p.sc[s].fh * p.sc[s].fw;
        const float* gs_s = gs + q.gs_off[s];
        const size_t foff = ((size_t)(b * p.V + v) * T) * p.C;
        if (d_grad_feats[s] != nullptr) {
          // row-major dF[T x C] = Gs[T x Kp] * Wq[Kp x C]   ==   column-major dF^T = Wq^T * Gs^T
          VFA_REQUIRE(cb->sgemm(cb->handle, 0, 0, p.C, T, Kp, &one, wq + s * per_scale, p.C, gs_s, Kp, &zero,
                                d_grad_feats[s] + foff, p.C) == 0,
                      VFA_ERR_CUDA, "cublasSgemm (dFeature) failed");
        }
        if (d_grad_weight[s] != nullptr) {
          // row-major dWr[Kp x C] += Gs^T[Kp x T] * F[T x C]   ==   column-major dWr^T[C x Kp] += F^T[C x T] * Gs[T x Kp]
          VFA_REQUIRE(cb->sgemm(cb->handle, 0, 1, p.C, Kp, T, &one, p.feats[s] + foff, p.C, gs_s, Kp, &one,
                                dwr + s * per_scale, p.C) == 0,
                      VFA_ERR_CUDA, "cublasSgemm (dWeight) failed");
        }
      }
    }
  }
  for (int s = 0; s < p.S; ++s) {
    if (d_grad_weight[s] == nullptr) continue;
    const int blocks = (int)((per_scale + 255) / 256 < 148 * 8 ? (per_scale + 255) / 256 : 148 * 8);
    fs::unprep_dweight_kernel<<<blocks, 256, 0, st>>>(dwr + s * per_scale, d_grad_weight[s], p.C, p.nl);
    VFA_LAUNCH_CHECK("unprep_dweight_kernel");
  }
  return VFA_OK;
}

static bool feature_side_ok(const AggParams& p) { return p.C % 8 == 0 && p.C <= 256; }

// ---- host --------------------------------------------------------------------------------------------------
size_t bwd_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh) {
  const size_t per_scale = (size_t)sh->channels * sh->channels * g->n_layers;
  size_t elems = (size_t)sh->n_scales * per_scale;                 // generic path: prepared weights
  if (sh->channels % 8 == 0 && sh->channels <= 256) {              // feature-side path: + dWr + dOut^T + Gs
    size_t px = 0;
    for (int s = 0; s < sh->n_scales; ++s) px += (size_t)sh->feat_h[s] * sh->feat_w[s];
    elems = 2 * (size_t)sh->n_scales * per_scale + (size_t)g->grid_l * g->grid_w * sh->channels +
            px * g->n_layers * sh->channels;
  }
  size_t bytes = elems * sizeof(float);
  if (sh->channels == 256) {                                       // sized for either C = 256 path
    const size_t g2 = bwd_fside_workspace_bytes(g, sh);
    if (g2 > bytes) bytes = g2;
  }
  return bytes;
}

int launch_bwd(AggParams p, const float* const* d_weight, const float* d_grad_out, float* const* d_grad_feats,
               float* const* d_grad_weight, float* const* d_grad_bias, void* ws, size_t ws_bytes, uint32_t flags,
               cudaStream_t st) {
  VFA_REQUIRE(p.C <= MAXC, VFA_ERR_UNSUPPORTED, "backward supports up to %d channels (got %d)", MAXC, p.C);
  const bool gout_nhwc = (flags & VFA_FLAG_OUT_NHWC) != 0;
  if (gather_backward(p.C))
    return launch_bwd_fside(p, d_weight, d_grad_out, d_grad_feats, d_grad_weight, d_grad_bias, ws, ws_bytes, gout_nhwc, st);
  VFA_REQUIRE(!gout_nhwc, VFA_ERR_UNSUPPORTED, "a channels-last cotangent (VFA_FLAG_OUT_NHWC) is read by the C = 256 gather "
              "backward only");
  if (feature_side_ok(p) && !runtime_config().bwd_generic)
    return launch_bwd_feature_side(p, d_weight, d_grad_out, d_grad_feats, d_grad_weight, d_grad_bias, ws, st);
  BwdParams q;
  q.p = p;
  q.gout = d_grad_out;
  const size_t per_scale = (size_t)p.C * p.C * p.nl;
  bool any_feat = false, any_w = false;
  for (int s = 0; s < VFA_MAX_SCALES; ++s) {
    q.weight[s] = q.wq[s] = nullptr;
    q.gfeat[s] = q.gweight[s] = q.gbias[s] = nullptr;
  }
  for (int s = 0; s < p.S; ++s) {
    q.weight[s] = d_weight[s];
    q.gfeat[s] = d_grad_feats[s];
    q.gweight[s] = d_grad_weight[s];
    q.gbias[s] = d_grad_bias[s];
    VFA_REQUIRE((q.gweight[s] == nullptr) == (q.gbias[s] == nullptr), VFA_ERR_INVALID_ARGUMENT,
                "scale %d: pass both or neither of d_grad_weight / d_grad_bias", s);
    any_feat |= q.gfeat[s] != nullptr;
    any_w |= q.gweight[s] != nullptr;
    if (q.gfeat[s] != nullptr) {
      float* wq = reinterpret_cast<float*>(ws) + s * per_scale;
      const int blocks = (int)((per_scale + 255) / 256 < 148 * 8 ? (per_scale + 255) / 256 : 148 * 8);
      prep_weight_bwd_kernel<<<blocks, 256, 0, st>>>(d_weight[s], wq, p.C, p.nl);
      VFA_LAUNCH_CHECK("prep_weight_bwd_kernel");
      q.wq[s] = wq;
    }
    if (q.gweight[s] != nullptr) {
      VFA_CUDA(cudaMemsetAsync(q.gweight[s], 0, per_scale * sizeof(float), st));
      VFA_CUDA(cudaMemsetAsync(q.gbias[s], 0, (size_t)p.C * sizeof(float), st));
    }
  }
  const int tiles = (p.LW + TM - 1) / TM;
  if (any_feat) {
    dim3 grid(tiles, p.B);
    bwd_feature_kernel<<<grid, THREADS, 0, st>>>(q);
    VFA_LAUNCH_CHECK("bwd_feature_kernel");
  }
  if (any_w) {
    const int chunks = (p.C + KC - 1) / KC;
    const int blocks_x = p.S * p.nl * chunks;
    const long long total = (long long)p.B * p.V * tiles;
    int parts = (2 * 148 + blocks_x - 1) / blocks_x;
    if (parts < 1) parts = 1;
    if (parts > total) parts = (int)total;
    q.cell_parts = parts;
    const size_t smem = (size_t)(TM * GSTRIDE + TM * KC) * sizeof(float) + TM * sizeof(BoxTaps);
    VFA_CUDA(cudaFuncSetAttribute(bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(blocks_x, parts);
    bwd_weight_kernel<<<grid, THREADS, smem, st>>>(q);
    VFA_LAUNCH_CHECK("bwd_weight_kernel");
  }
  return VFA_OK;
}

}  // namespace vfa
