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

// ---- host --------------------------------------------------------------------------------------------------
size_t bwd_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh) {
  return (size_t)sh->n_scales * sh->channels * sh->channels * g->n_layers * sizeof(float);
}

int launch_bwd(AggParams p, const float* const* d_weight, const float* d_grad_out, float* const* d_grad_feats,
               float* const* d_grad_weight, float* const* d_grad_bias, void* ws, cudaStream_t st) {
  VFA_REQUIRE(p.C <= MAXC, VFA_ERR_UNSUPPORTED, "backward supports up to %d channels (got %d)", MAXC, p.C);
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
    static bool attr_set = false;
    if (!attr_set) {
      VFA_CUDA(cudaFuncSetAttribute(bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    dim3 grid(blocks_x, parts);
    bwd_weight_kernel<<<grid, THREADS, smem, st>>>(q);
    VFA_LAUNCH_CHECK("bwd_weight_kernel");
  }
  return VFA_OK;
}

}  // namespace vfa
