// Generic fused aggregation forward, fp32 FFMA (sm_100a).  Any channel count / layer count / scale count.
//
// One CTA owns a tile of 64 consecutive BEV cells of one frame and walks every (view, scale): for each height
// layer n and 32-channel chunk it (1) pools the chunk of every cell's box straight from the channels-last
// feature map into shared memory (lane = channel: each tap is one coalesced 128 B row read), (2) stages the
// matching [32 x 256] slab of the collapse weight, (3) rank-32 updates an 8x8 register tile per thread.  After
// the last chunk: + bias, ReLU, add into the running BEV registers; the [C, L, W] intermediate of the
// reference (vfa_op.py:118-124, 125-354 MB per call) never exists.  This is the path the golden-vector parity
// tests run through for arbitrary C; the tcgen05 kernel (vfa_fwd_umma.cu) replaces it for C = 256.
#include "vfa_common.cuh"

namespace vfa {

constexpr int TM = 64;        // cells per CTA tile
constexpr int KC = 32;        // channels per chunk
constexpr int NB = 256;       // output channels per pass
constexpr int THREADS = 256;

// W[o, c*nl + n]  ->  Wp[n][c][o]   (o contiguous: coalesced slab loads in the main kernel)
__global__ void __launch_bounds__(256) prep_weight_simt_kernel(const float* __restrict__ w, float* __restrict__ wp,
                                                               int C, int nl) {
  const long long total = (long long)C * C * nl;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % C);
    const int c = (int)((idx / C) % C);
    const int n = (int)(idx / ((long long)C * C));
    wp[idx] = w[(long long)o * C * nl + (long long)c * nl + n];
  }
}

__global__ void __launch_bounds__(THREADS, 1) aggregate_fwd_simt_kernel(AggParams p) {
  __shared__ BoxTaps taps[TM];
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Ws[KC][NB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cell0 = blockIdx.x * TM;
  const int b = blockIdx.y;
  const int cg = warp;          // cell group: cells cg*8 .. cg*8+7 in the FFMA phase
  const int og = lane;          // outputs og + 32*j

  for (int ob = 0; ob < p.C; ob += NB) {
    float bev[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) bev[i][j] = 0.f;

    for (int v = 0; v < p.V; ++v) {
      for (int s = 0; s < p.S; ++s) {
        const ScaleConst sc = p.sc[s];
        const float* __restrict__ feat = p.feats[s] + ((long long)(b * p.V + v) * sc.fh * sc.fw) * p.C;
        const float* __restrict__ wp = p.wprep[s];
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        for (int n = 0; n < p.nl; ++n) {
          __syncthreads();                       // previous users of taps[] / As / Ws are done
          if (tid < TM) {
            const int cell = cell0 + tid;
            BoxTaps t;
            if (cell < p.LW) {
              const float4 box = reinterpret_cast<const float4*>(p.boxes)[((long long)v * p.nl + n) * p.LW + cell];
              t = derive_taps(box, sc);
            } else {
              t.x0 = t.y0 = t.nx = t.ny = 0;
              t.wx_first = t.wx_last = t.wy_first = t.wy_last = t.wy_mid = 0.f;
            }
            taps[tid] = t;
          }
          __syncthreads();

          for (int c0 = 0; c0 < p.C; c0 += KC) {
            // (1) pool: warp w handles cells w*8..w*8+7, lane = channel
            const int c = c0 + lane;
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
              const int r = warp * 8 + i;
              const BoxTaps t = taps[r];
              float sum = 0.f;
              if (c < p.C) {
                for (int ty = 0; ty < t.ny; ++ty) {
                  const float wy = tap_wy(t, ty);
                  const float* row = feat + ((long long)(t.y0 + ty) * sc.fw + t.x0) * p.C + c;
                  float rs = 0.f;
                  for (int tx = 0; tx < t.nx; ++tx) rs = fmaf(tap_wx(t, tx), __ldg(row + (long long)tx * p.C), rs);
                  sum = fmaf(wy, rs, sum);
                }
              }
              As[lane][r] = sum;
            }
            // (2) weight slab  Wp[n][c0+kk][ob + o]
            for (int e = tid; e < KC * NB; e += THREADS) {
              const int kk = e / NB, o = e % NB;
              float wv = 0.f;
              if (c0 + kk < p.C && ob + o < p.C) wv = __ldg(wp + ((long long)n * p.C + (c0 + kk)) * p.C + ob + o);
              Ws[kk][o] = wv;
            }
            __syncthreads();
            // (3) rank-KC update
#pragma unroll 4
            for (int kk = 0; kk < KC; ++kk) {
              const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][cg * 8]);
              const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][cg * 8 + 4]);
              const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              float bb[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) bb[j] = Ws[kk][og + 32 * j];
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
          }
        }
        // epilogue of this (view, scale): + bias, ReLU, add            (vfa_op.py:123-124, vfanet.py:79-82)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int o = ob + og + 32 * j;
          const float bs = o < p.C ? __ldg(p.bias[s] + o) : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y = acc[i][j] + bs;
            bev[i][j] += fmaxf(y, 0.f);
            if (p.mask != nullptr) {      // lanes = 32 consecutive output channels of word (ob/32 + j)
              const unsigned bits = __ballot_sync(0xffffffffu, o < p.C && y > 0.f);
              const int cell = cell0 + cg * 8 + i;
              const int word = ob / 32 + j;
              if (lane == 0 && cell < p.LW && word * 32 < p.C)
                p.mask[((((size_t)b * p.V + v) * p.S + s) * ((p.C + 31) / 32) + word) * p.LW + cell] = bits;
            }
          }
        }
      }
    }
    // store [B, C, L*W]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = ob + og + 32 * j;
      if (o >= p.C) continue;
      float* dst = p.out + ((long long)b * p.C + o) * p.LW;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cell = cell0 + cg * 8 + i;
        if (cell < p.LW) dst[cell] = bev[i][j];
      }
    }
  }
}

size_t simt_workspace_bytes(const vfa_geometry_t* g, const vfa_shape_t* sh) {
  return (size_t)sh->n_scales * sh->channels * sh->channels * g->n_layers * sizeof(float);
}

int prep_weights_simt(const AggParams& p, const float* const* d_weight, void* ws, cudaStream_t st) {
  float* wp = reinterpret_cast<float*>(ws);
  const size_t per_scale = (size_t)p.C * p.C * p.nl;
  for (int s = 0; s < p.S; ++s) {
    const long long total = (long long)per_scale;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    prep_weight_simt_kernel<<<blocks, 256, 0, st>>>(d_weight[s], wp + s * per_scale, p.C, p.nl);
    VFA_LAUNCH_CHECK("prep_weight_simt_kernel");
  }
  return VFA_OK;
}

int launch_fwd_simt(AggParams p, const float* const* d_weight, void* ws, uint32_t flags, cudaStream_t st) {
  if (!(flags & VFA_FLAG_WEIGHTS_PREPARED)) {
    if (int rc = prep_weights_simt(p, d_weight, ws, st)) return rc;
  }
  const size_t per_scale = (size_t)p.C * p.C * p.nl;
  for (int s = 0; s < p.S; ++s) p.wprep[s] = reinterpret_cast<const float*>(ws) + s * per_scale;
  dim3 grid((p.LW + TM - 1) / TM, p.B);
  aggregate_fwd_simt_kernel<<<grid, THREADS, 0, st>>>(p);
  VFA_LAUNCH_CHECK("aggregate_fwd_simt_kernel");
  set_path("simt_fp32");
  return VFA_OK;
}

}  // namespace vfa
