// Skinny projection  out[M,K] = x[M,F] @ v[K,F]^T  for K <= 32  (see include/gvqa_b200.h).
//
// Computes the collapsed attention-logit terms of the reference's gat.forward
// (gat_skip.py:134-135 a_l/a_r, :150-151 a_e):  <W x, att_h> == x . (W_h^T att_h).
// HBM-bound: x is streamed once with 128-bit loads; v (K*F floats) lives in shared memory and every
// 128-bit read of it is reused for R rows held in registers (R x F/128 independent loads in
// flight per lane), so shared-memory bandwidth does not cap the stream.
#include "common.cuh"

namespace gvqa {

template <int KMAX, int R>
__global__ void __launch_bounds__(256) skinny_matmul_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ v,
                                                            float* __restrict__ out, int64_t M, int F, int K) {
  extern __shared__ __align__(16) float v_s[];  // [K][F]
  for (int i = threadIdx.x * 4; i < K * F; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(v_s + i) = __ldg(reinterpret_cast<const float4*>(v + i));
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  const int F4 = F >> 2;

  for (int64_t row0 = warp0 * R; row0 < M; row0 += nwarps * R) {
    float acc[R][KMAX];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[r][k] = 0.f;
    for (int c4 = lane; c4 < F4; c4 += 32) {
      float4 xv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        xv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) xv[r] = ldg_stream(x + (row0 + r) * ldx + 4 * c4);
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const float4 w = *reinterpret_cast<const float4*>(v_s + k * F + 4 * c4);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][k] = fmaf(xv[r].x, w.x, acc[r][k]);
            acc[r][k] = fmaf(xv[r].y, w.y, acc[r][k]);
            acc[r][k] = fmaf(xv[r].z, w.z, acc[r][k]);
            acc[r][k] = fmaf(xv[r].w, w.w, acc[r][k]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float mine = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const float s = warp_sum(acc[r][k]);
          if (lane == k) mine = s;
        }
      }
      if (lane < K && row0 + r < M) out[(row0 + r) * K + lane] = mine;
    }
  }
}

}  // namespace gvqa

extern "C" GVQA_API int gvqa_skinny_matmul_f32(const float* x, int64_t ldx, const float* v, float* out, int64_t m,
                                               int f, int k, void* stream_) {
  using namespace gvqa;
  if (m < 0 || f <= 0 || k <= 0 || ldx < f) return GVQA_ERR_BAD_SHAPE;
  if (m == 0) return GVQA_OK;
  if (!x || !v || !out) return GVQA_ERR_NULL_POINTER;
  if (k > 32 || (f & 3) || (size_t)k * f * 4 > 200 * 1024) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(v) || (ldx & 3)) return GVQA_ERR_MISALIGNED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (size_t)k * f * sizeof(float);
#define GVQA_SKINNY(KM, RR)                                                                              \
  do {                                                                                                   \
    if (smem > 48 * 1024 &&                                                                              \
        cudaFuncSetAttribute(skinny_matmul_kernel<KM, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem) != cudaSuccess)                                                  \
      return GVQA_ERR_CUDA;                                                                              \
    const int64_t blocks_needed = (m + 8 * RR - 1) / (8 * RR);                                           \
    const unsigned grid = (unsigned)(blocks_needed < 4 * kNumSMs ? blocks_needed : 4 * kNumSMs);         \
    skinny_matmul_kernel<KM, RR><<<grid, 256, smem, stream>>>(x, ldx, v, out, m, f, k);                 \
  } while (0)
  if (k <= 8)
    GVQA_SKINNY(8, 4);
  else if (k <= 16)
    GVQA_SKINNY(16, 4);
  else if (k <= 24)
    GVQA_SKINNY(24, 2);
  else
    GVQA_SKINNY(32, 2);
#undef GVQA_SKINNY
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
