// Per-graph LayerNorm (see include/gvqa_b200.h: gvqa_graph_layernorm_f32).
//
// Restates graph_utils/my_graph_layernorm.py:52-78 of the reference as ONE kernel: one CTA per
// graph stages the graph's n*F floats in shared memory with 128-bit loads (single HBM read),
// computes the mean, then the centred second moment (two-pass, exactly like the reference --
// not E[x^2]-E[x]^2), and writes (x-mean)/(sqrt(var)+eps)*w+b with 128-bit stores.  Graphs too
// large for the staged buffer fall back to re-reading global memory (L2-resident).
#include "common.cuh"

namespace gvqa {

constexpr int kLnThreads = 512;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float t = (lane < kLnThreads / 32) ? scratch[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the block total
}

__global__ void __launch_bounds__(kLnThreads) graph_layernorm_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ graph_ptr, const float* __restrict__ weight,
    const float* __restrict__ bias, float* __restrict__ out, int C, float eps, int smem_floats) {
  extern __shared__ __align__(16) float buf[];
  __shared__ float scratch[32];
  const int g = blockIdx.x;
  const int n0 = graph_ptr[g], n1 = graph_ptr[g + 1];
  const int64_t count = (int64_t)(n1 - n0) * C;
  if (count == 0) return;
  const float* xg = x + (int64_t)n0 * C;
  float* og = out + (int64_t)n0 * C;
  const int64_t count4 = count >> 2;  // C % 4 == 0
  const bool staged = count <= smem_floats;

  float s = 0.f;
  for (int64_t i = threadIdx.x; i < count4; i += kLnThreads) {
    const float4 v = ldg_stream(xg + 4 * i);
    if (staged) *reinterpret_cast<float4*>(buf + 4 * i) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float norm = (float)count;  // degree.clamp(min=1) * F; count > 0 here
  const float mean = block_sum(s, scratch) / norm;

  float q = 0.f;
  for (int64_t i = threadIdx.x; i < count4; i += kLnThreads) {
    const float4 v = staged ? *reinterpret_cast<const float4*>(buf + 4 * i)
                            : __ldg(reinterpret_cast<const float4*>(xg + 4 * i));
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = block_sum(q, scratch) / norm;
  const float denom = sqrtf(var) + eps;
  const bool affine = weight != nullptr && bias != nullptr;
  const float w = affine ? __ldg(weight) : 1.f;
  const float b0 = affine ? __ldg(bias) : 0.f;

  for (int64_t i = threadIdx.x; i < count4; i += kLnThreads) {
    const float4 v = staged ? *reinterpret_cast<const float4*>(buf + 4 * i)
                            : __ldg(reinterpret_cast<const float4*>(xg + 4 * i));
    float4 o;
    o.x = (v.x - mean) / denom; o.y = (v.y - mean) / denom;
    o.z = (v.z - mean) / denom; o.w = (v.w - mean) / denom;
    if (affine) {
      o.x = o.x * w + b0; o.y = o.y * w + b0; o.z = o.z * w + b0; o.w = o.w * w + b0;
    }
    stg_stream(og + 4 * i, o);
  }
}

}  // namespace gvqa

extern "C" GVQA_API int gvqa_graph_layernorm_f32(const float* x, const int32_t* graph_ptr, const float* weight,
                                        const float* bias, float* out, int64_t num_nodes, int64_t num_graphs,
                                        int32_t channels, float eps, int32_t max_nodes_per_graph,
                                        void* stream_) {
  using namespace gvqa;
  if (num_nodes < 0 || num_graphs < 0 || channels <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0 || num_graphs == 0) return GVQA_OK;
  if (!x || !graph_ptr || !out) return GVQA_ERR_NULL_POINTER;
  if (channels & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(out)) return GVQA_ERR_MISALIGNED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // staged buffer: the hinted largest graph if it fits, else a 96 KB default (2 CTAs/SM)
  size_t want = max_nodes_per_graph > 0 ? (size_t)max_nodes_per_graph * channels * 4 : 96 * 1024;
  if (want > 200 * 1024) want = 96 * 1024;
  want = (want + 15) & ~(size_t)15;
  if (want > 48 * 1024 &&
      cudaFuncSetAttribute(graph_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want) !=
          cudaSuccess)
    return GVQA_ERR_CUDA;
  graph_layernorm_kernel<<<(unsigned)num_graphs, kLnThreads, want, stream>>>(x, graph_ptr, weight, bias, out,
                                                                               channels, eps, (int)(want / 4));
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
