// Projection-table kernels (sm_100a): bit-exact restatement of the reference's box construction.
//
// Every arithmetic step is an explicitly rounded fp32 intrinsic (__fmul_rn/__fadd_rn/__fdiv_rn are never
// contracted into FMAs and are IEEE-correct), in the scalar order of SURVEY.md appendix A.2, because the
// reference's torch CPU kernels round every elementwise op separately (reference vfa/model/vfa_op.py:64-88,
// vfa/utils.py:50-59).  This file is additionally compiled with -fmad=false.
#include "vfa_common.cuh"

namespace vfa {

struct TableParams {
  int V, nl, L, W;
  int convert_kind;
  float convert_scale;
  float convert_offset[3];
  float off[8][3];               // cuboid corner offsets, reference order (vfa_op.py:127-133)
  float layer_z[VFA_MAX_LAYERS];
  float image_w, image_h, lo, hi;
};

// torch.clamp semantics: NaN in -> NaN out
__device__ __forceinline__ float clamp_nanprop(float x, float lo, float hi) {
  if (x != x) return x;
  return fminf(fmaxf(x, lo), hi);
}
// torch.min/max over a dim: any NaN wins
__device__ __forceinline__ float min_nanprop(float a, float b) { return (a != a || b != b) ? (a + b) : fminf(a, b); }
__device__ __forceinline__ float max_nanprop(float a, float b) { return (a != a || b != b) ? (a + b) : fmaxf(a, b); }

__global__ void __launch_bounds__(256) table_build_kernel(TableParams p, const float* __restrict__ calibs,
                                                          const float* __restrict__ grid, float4* __restrict__ boxes) {
  const int LW = p.L * p.W;
  const long long total = (long long)p.V * p.nl * LW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(idx % LW);
    const int n = (int)((idx / LW) % p.nl);
    const int v = (int)(idx / ((long long)LW * p.nl));
    const float* P = calibs + v * 12;
    const float gx = grid[cell * 3 + 0], gy = grid[cell * 3 + 1], gz = grid[cell * 3 + 2];
    // grid + (0, 0, z_n)                                                           vfa_op.py:64
    const float bx = __fadd_rn(gx, 0.0f), by = __fadd_rn(gy, 0.0f), bz = __fadd_rn(gz, p.layer_z[n]);
    float xmin = 0.f, ymin = 0.f, xmax = 0.f, ymax = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float X = __fadd_rn(bx, p.off[k][0]);                                       // vfa_op.py:66
      float Y = __fadd_rn(by, p.off[k][1]);
      float Z = __fadd_rn(bz, p.off[k][2]);
      if (p.convert_kind == VFA_CONVERT_DIV) {                                    // vfa_op.py:23-28
        X = __fdiv_rn(X, p.convert_scale);
        Y = __fdiv_rn(Y, p.convert_scale);
        Z = __fdiv_rn(Z, p.convert_scale);
      } else {                                                                    // vfa_op.py:31-35
        X = __fsub_rn(__fmul_rn(X, p.convert_scale), p.convert_offset[0]);
        Y = __fsub_rn(__fmul_rn(Y, p.convert_scale), p.convert_offset[1]);
        Z = __fmul_rn(Z, p.convert_scale);
      }
      float h[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {                                               // utils.py:57
        float acc = __fmul_rn(P[r * 4 + 0], X);
        acc = __fadd_rn(acc, __fmul_rn(P[r * 4 + 1], Y));
        acc = __fadd_rn(acc, __fmul_rn(P[r * 4 + 2], Z));
        h[r] = __fadd_rn(acc, P[r * 4 + 3]);
      }
      const float u = __fdiv_rn(h[0], h[2]);                                      // utils.py:59, no depth test
      const float w = __fdiv_rn(h[1], h[2]);
      const float nx = clamp_nanprop(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, u), p.image_w), 1.0f), p.lo, p.hi);
      const float ny = clamp_nanprop(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, w), p.image_h), 1.0f), p.lo, p.hi);
      if (k == 0) {
        xmin = xmax = nx;
        ymin = ymax = ny;
      } else {                                                                    // vfa_op.py:81-86
        xmin = min_nanprop(xmin, nx);
        xmax = max_nanprop(xmax, nx);
        ymin = min_nanprop(ymin, ny);
        ymax = max_nanprop(ymax, ny);
      }
    }
    boxes[idx] = make_float4(xmin, ymin, xmax, ymax);
  }
}

__global__ void __launch_bounds__(256) table_scale_kernel(const float4* __restrict__ boxes, long long n_boxes,
                                                          ScaleConst sc, float* __restrict__ area_out,
                                                          uint8_t* __restrict__ vis_out, int4* __restrict__ taps_out) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_boxes;
       idx += (long long)gridDim.x * blockDim.x) {
    const float4 b = boxes[idx];
    const float area = box_area_f32(b, sc);
    if (area_out) area_out[idx] = area;
    if (vis_out) vis_out[idx] = box_visible(area, sc) ? 1 : 0;
    if (taps_out) {
      float t[4] = {floorf(unnormalize_f32(b.x, sc.fwf)), floorf(unnormalize_f32(b.y, sc.fhf)),
                    floorf(unnormalize_f32(b.z, sc.fwf)), floorf(unnormalize_f32(b.w, sc.fhf))};
      int4 o;
      o.x = isfinite(t[0]) ? (int)t[0] : -1;
      o.y = isfinite(t[1]) ? (int)t[1] : -1;
      o.z = isfinite(t[2]) ? (int)t[2] : -1;
      o.w = isfinite(t[3]) ? (int)t[3] : -1;
      taps_out[idx] = o;
    }
  }
}

// [n, R, Cc] -> [n, Cc, R] tiled transpose (32x32 tiles through shared memory, both sides coalesced)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int rows, int cols) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = src[base + (long long)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[base + (long long)c * rows + r] = tile[tx][i];
  }
}

int launch_table_build(const vfa_geometry_t* g, int V, const float* d_calibs, const float* d_grid, float* d_boxes,
                       cudaStream_t st) {
  TableParams p;
  p.V = V;
  p.nl = g->n_layers;
  p.L = g->grid_l;
  p.W = g->grid_w;
  p.convert_kind = g->convert_kind;
  p.convert_scale = g->convert_scale;
  for (int a = 0; a < 3; ++a) p.convert_offset[a] = g->convert_offset[a];
  const float l = g->cube[0], w = g->cube[1], h = g->cube[2];
  const float ox[8] = {-l / 2, l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2};
  const float oy[8] = {-w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2, w / 2};
  const float oz[8] = {0, 0, 0, 0, h, h, h, h};
  for (int k = 0; k < 8; ++k) {
    p.off[k][0] = ox[k];
    p.off[k][1] = oy[k];
    p.off[k][2] = oz[k];
  }
  for (int n = 0; n < VFA_MAX_LAYERS; ++n) p.layer_z[n] = n < g->n_layers ? g->layer_z[n] : 0.f;
  p.image_w = g->image_w;
  p.image_h = g->image_h;
  p.lo = g->clamp_lo;
  p.hi = g->clamp_hi;
  const long long total = (long long)V * p.nl * p.L * p.W;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  table_build_kernel<<<blocks, 256, 0, st>>>(p, d_calibs, d_grid, reinterpret_cast<float4*>(d_boxes));
  VFA_LAUNCH_CHECK("table_build_kernel");
  return VFA_OK;
}

int launch_table_scale(const float* d_boxes, long long n_boxes, int fh, int fw, float* d_area, uint8_t* d_vis,
                       int32_t* d_taps, cudaStream_t st) {
  const int blocks = (int)((n_boxes + 255) / 256 < 148 * 16 ? (n_boxes + 255) / 256 : 148 * 16);
  table_scale_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(d_boxes), n_boxes,
                                             make_scale_const(fh, fw), d_area, d_vis, reinterpret_cast<int4*>(d_taps));
  VFA_LAUNCH_CHECK("table_scale_kernel");
  return VFA_OK;
}

// src viewed as [n, rows, cols] -> dst [n, cols, rows]
int launch_transpose(const float* src, float* dst, long long n, int rows, long long cols, cudaStream_t st) {
  VFA_REQUIRE(n <= 65535, VFA_ERR_UNSUPPORTED, "transpose: more than 65535 images in one call (%lld)", n);
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)n);
  VFA_REQUIRE(grid.y <= 65535, VFA_ERR_UNSUPPORTED, "transpose: too many rows (%d)", rows);
  transpose_kernel<<<grid, 256, 0, st>>>(src, dst, rows, (int)cols);
  VFA_LAUNCH_CHECK("transpose_kernel");
  return VFA_OK;
}

}  // namespace vfa
