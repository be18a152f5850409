// Decode tail of the detector: what reference vfa/data/encoder.py:230-273 (decode3d) and
// :275-305 (decode2d) do with ~15 full-map ATen launches -- sigmoid, 5 x 5 max-pool NMS, top-k, gathers of the regression
// heads, argmax of the orientation head -- as two small kernels that touch the regression / orientation maps at the k
// selected cells only:
//
//   nms_candidates_kernel  sigmoid(heatmap), keep a cell iff it equals the maximum of its 5 x 5 neighbourhood (MaxPool2d(5,
//                          stride 1, padding 2) pads with -inf: border windows are simply smaller), append (conf, cell) of the
//                          kept cells with conf > 0 to a per-frame candidate list;
//   topk_decode_kernel     one CTA per frame: bitonic sort of the candidates by (conf descending, cell ascending) in shared
//                          memory, first k -> conf, cell index, centre (cy, cx) from sigmoid(tytx), box size exp(thtwtl) *
//                          class mean, orientation bin = argmax over the angle logits (sigmoid is monotonic).
//
// Arithmetic follows the reference expression by expression in fp32: sigmoid(x) = 1 / (1 + exp(-x)) (ATen's formula),
// (grid + sigmoid(t)) / grid_size * world_size, exp(t) * mean.  Fewer than k candidates: the tail is filled with conf = 0,
// cell = -1 (the reference's top-k returns arbitrary suppressed cells there, all of which its `conf > cls_thresh` mask drops).
#include "vfa_common.cuh"

namespace vfa {
namespace dec {

constexpr int CAND_CAP = 8192;                 // candidates per frame the sort holds (a 5 x 5 NMS keeps at most one cell in 9)

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

struct NmsArgs {
  const float* heat;      // [B, L*W] logits
  int B, L, W;
  unsigned long long* cand;   // [B][CAND_CAP]: conf bits << 32 | ~cell  (sorted descending = conf desc, cell asc)
  int* count;             // [B]
};

__global__ void __launch_bounds__(256) nms_candidates_kernel(const NmsArgs a) {
  const int LW = a.L * a.W;
  const long long total = (long long)a.B * LW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / LW), cell = (int)(idx % LW), y = cell / a.W, x = cell % a.W;
    const float* h = a.heat + (size_t)b * LW;
    const float v = sigmoidf_ref(__ldg(h + cell));
    float m = v;
    for (int dy = -2; dy <= 2; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= a.L) continue;
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= a.W) continue;
        m = fmaxf(m, sigmoidf_ref(__ldg(h + yy * a.W + xx)));
      }
    }
    if (v == m && v > 0.f) {
      const int pos = atomicAdd(a.count + b, 1);
      if (pos < CAND_CAP)
        a.cand[(size_t)b * CAND_CAP + pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(~cell);
    }
  }
}

struct Head {
  const float* p;         // element (b, c, cell) at p[b * sb + c * sc + cell * scell]
  long long sb, sc, scell;
};

struct DecodeArgs {
  const unsigned long long* cand;
  const int* count;
  Head tytx, thtwtl, orient;      // thtwtl.p / orient.p may be NULL (2-D decoding: reference decode2d)
  int B, L, W, topk, n_angles;
  float grid_l, grid_w, world_l, world_w;   // self.grid_size, self.world_size of the reference encoder
  float mean_h, mean_w, mean_l;             // classAverage.get_mean
  float* out_vals;        // [B, topk, 7]: conf, cy, cx, h, w, l, orientation bin
  int* out_cell;          // [B, topk]
};

__device__ __forceinline__ float head_at(const Head& h, int b, int c, int cell) {
  return __ldg(h.p + (long long)b * h.sb + (long long)c * h.sc + (long long)cell * h.scell);
}

__global__ void __launch_bounds__(1024) topk_decode_kernel(const DecodeArgs a) {
  extern __shared__ unsigned long long keys[];               // [CAND_CAP]
  const int b = blockIdx.x;
  const int n = min(a.count[b], CAND_CAP);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) keys[i] = i < n ? a.cand[(size_t)b * CAND_CAP + i] : 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys[i], y = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? x < y : x > y) {
            keys[i] = y;
            keys[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < a.topk; j += nwarps) {            // one warp per detection
    float* o = a.out_vals + ((size_t)b * a.topk + j) * 7;
    if (j >= n) {
      if (lane < 7) o[lane] = 0.f;
      if (lane == 0) a.out_cell[(size_t)b * a.topk + j] = -1;
      continue;
    }
    const unsigned long long key = keys[j];
    const int cell = (int)(~(unsigned)(key & 0xffffffffull));
    const float conf = __uint_as_float((unsigned)(key >> 32));
    // orientation: argmax over the angle logits (first maximum, as torch.max returns for distinct values)
    int best = 0;
    if (a.orient.p != nullptr) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < a.n_angles; c += 32) {
        const float v = head_at(a.orient, b, c, cell);
        if (v > bv) {
          bv = v;
          bi = c;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      best = bi;
    }
    if (lane == 0) {
      const int y = cell / a.W, x = cell % a.W;
      const float ty = sigmoidf_ref(head_at(a.tytx, b, 0, cell)), tx = sigmoidf_ref(head_at(a.tytx, b, 1, cell));
      o[0] = conf;
      o[1] = __fmul_rn(__fdiv_rn(__fadd_rn((float)y, ty), a.grid_l), a.world_l);       // (grid_y + ty) / grid_size[0] * world_size[0]
      o[2] = __fmul_rn(__fdiv_rn(__fadd_rn((float)x, tx), a.grid_w), a.world_w);
      if (a.thtwtl.p != nullptr) {
        o[3] = __fmul_rn(expf(head_at(a.thtwtl, b, 0, cell)), a.mean_h);
        o[4] = __fmul_rn(expf(head_at(a.thtwtl, b, 1, cell)), a.mean_w);
        o[5] = __fmul_rn(expf(head_at(a.thtwtl, b, 2, cell)), a.mean_l);
      } else {
        o[3] = o[4] = o[5] = 0.f;
      }
      o[6] = (float)best;
      a.out_cell[(size_t)b * a.topk + j] = cell;
    }
  }
}

}  // namespace dec

static size_t counts_bytes(int B) { return ((size_t)B * sizeof(int) + 255) & ~(size_t)255; }
size_t decode_workspace_bytes(int B) { return (size_t)B * dec::CAND_CAP * sizeof(unsigned long long) + counts_bytes(B); }

int launch_decode(const vfa_decode_t* d, float* out_vals, int32_t* out_cell, void* ws, cudaStream_t st) {
  using namespace dec;
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  NmsArgs n;
  n.heat = d->heatmap;
  n.B = d->batch;
  n.L = d->grid_l;
  n.W = d->grid_w;
  n.count = reinterpret_cast<int*>(w8);
  n.cand = reinterpret_cast<unsigned long long*>(w8 + counts_bytes(d->batch));
  VFA_CUDA(cudaMemsetAsync(n.count, 0, (size_t)d->batch * sizeof(int), st));
  const long long total = (long long)d->batch * d->grid_l * d->grid_w;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  nms_candidates_kernel<<<blocks, 256, 0, st>>>(n);
  VFA_LAUNCH_CHECK("nms_candidates_kernel");
  DecodeArgs a;
  a.cand = n.cand;
  a.count = n.count;
  a.tytx = {d->loc_offset, d->loc_stride[0], d->loc_stride[1], d->loc_stride[2]};
  a.thtwtl = {d->dim_offset, d->dim_stride[0], d->dim_stride[1], d->dim_stride[2]};
  a.orient = {d->rotation, d->rot_stride[0], d->rot_stride[1], d->rot_stride[2]};
  a.B = d->batch;
  a.L = d->grid_l;
  a.W = d->grid_w;
  a.topk = d->topk;
  a.n_angles = d->n_angles;
  a.grid_l = d->grid_size[0];
  a.grid_w = d->grid_size[1];
  a.world_l = d->world_size[0];
  a.world_w = d->world_size[1];
  a.mean_h = d->dim_mean[0];
  a.mean_w = d->dim_mean[1];
  a.mean_l = d->dim_mean[2];
  a.out_vals = out_vals;
  a.out_cell = out_cell;
  const size_t smem = (size_t)CAND_CAP * sizeof(unsigned long long);
  VFA_CUDA(cudaFuncSetAttribute(topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_decode_kernel<<<d->batch, 1024, smem, st>>>(a);
  VFA_LAUNCH_CHECK("topk_decode_kernel");
  return VFA_OK;
}

}  // namespace vfa
