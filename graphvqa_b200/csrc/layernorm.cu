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

  // the graph is one dependent chain (load -> mean -> variance -> store), so the load phase must be ONE round trip:
  // eight independent 128-bit loads per thread are requested before the first is consumed (a GQA graph of 30 x 512
  // floats is 7.5 per thread; one load per iteration left 8 KB in flight per CTA, i.e. ~8 serial DRAM latencies --
  // ncu: 11.7 us, No Eligible 55 %, DRAM 16 %, profiles/r02/ncu_k2_layernorm.txt)
  constexpr int kLnBatch = 8;
  float s = 0.f;
  for (int64_t base = threadIdx.x; base < count4; base += (int64_t)kLnThreads * kLnBatch) {
    float4 v[kLnBatch];
#pragma unroll
    for (int u = 0; u < kLnBatch; ++u) {
      const int64_t i = base + (int64_t)u * kLnThreads;
      v[u] = i < count4 ? ldg_stream(xg + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kLnBatch; ++u) {
      const int64_t i = base + (int64_t)u * kLnThreads;
      if (i < count4) {
        if (staged) *reinterpret_cast<float4*>(buf + 4 * i) = v[u];
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      }
    }
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


// ------------------------------------------------------------------------------------------
// Cluster version (default): a thread-block CLUSTER of S CTAs owns one graph, each CTA stages 1/S of the graph's rows in
// its shared memory (single HBM read also for graphs far too large for one CTA: 8 x ~200 KB) and the two statistics
// cross the cluster through distributed shared memory (mapa + ld.shared::cluster).  With one CTA per graph, 256 GQA
// graphs are 1.7 CTAs per SM -- an unbalanced, latency-bound launch (34 % of the HBM peak at cfg2); S = 2..8 gives the
// scheduler 512..2048 smaller CTAs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_cluster_f32(const float* local_smem, uint32_t rank) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(local_smem)), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

constexpr int kLnClThreads = 256;

__device__ __forceinline__ float block_sum_n(float v, float* scratch, int nwarps) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float t = lane < nwarps ? scratch[lane] : 0.f;
  return warp_sum(t);
}

__global__ void __launch_bounds__(kLnClThreads) graph_layernorm_cluster_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ graph_ptr, const float* __restrict__ weight,
    const float* __restrict__ bias, float* __restrict__ out, int C, float eps, int smem_floats, int S) {
  extern __shared__ __align__(16) float buf[];
  __shared__ float scratch[32];
  __shared__ float part[2];              // this CTA's partial sum / partial centred square sum (read by its peers)
  const uint32_t rank = cluster_ctarank();
  const int g = blockIdx.x / S;
  const int n0 = graph_ptr[g], n1 = graph_ptr[g + 1], n = n1 - n0;
  // rows [r0, r1) of the graph belong to this CTA (balanced split; empty for tiny graphs)
  const int per = (n + S - 1) / S;
  const int r0 = min(n, (int)rank * per), r1 = min(n, r0 + per);
  const int64_t count4 = (int64_t)(r1 - r0) * (C >> 2);
  const float* xg = x + (int64_t)(n0 + r0) * C;
  float* og = out + (int64_t)(n0 + r0) * C;
  const bool staged = count4 * 4 <= smem_floats;

  float s = 0.f;
  for (int64_t i = threadIdx.x; i < count4; i += kLnClThreads) {
    const float4 v = ldg_stream(xg + 4 * i);
    if (staged) *reinterpret_cast<float4*>(buf + 4 * i) = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  s = block_sum_n(s, scratch, kLnClThreads / 32);
  if (threadIdx.x == 0) part[0] = s;
  cluster_barrier();
  float total = 0.f;
  for (int r = 0; r < S; ++r) total += ld_cluster_f32(&part[0], (uint32_t)r);     // same order in every CTA
  const float norm = (float)max(n, 1) * (float)C;     // degree.clamp(min=1) * F
  const float mean = total / norm;

  float q = 0.f;
  for (int64_t i = threadIdx.x; i < count4; i += kLnClThreads) {
    const float4 v = staged ? *reinterpret_cast<const float4*>(buf + 4 * i)
                            : __ldg(reinterpret_cast<const float4*>(xg + 4 * i));
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  q = block_sum_n(q, scratch, kLnClThreads / 32);
  if (threadIdx.x == 0) part[1] = q;
  cluster_barrier();
  float qt = 0.f;
  for (int r = 0; r < S; ++r) qt += ld_cluster_f32(&part[1], (uint32_t)r);
  cluster_barrier();                     // nobody leaves (and frees its shared memory) while a peer may still read it
  const float denom = sqrtf(qt / norm) + eps;
  const bool affine = weight != nullptr && bias != nullptr;
  const float w = affine ? __ldg(weight) : 1.f;
  const float b0 = affine ? __ldg(bias) : 0.f;
  for (int64_t i = threadIdx.x; i < count4; i += kLnClThreads) {
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
  // GVQA_LN_CLUSTER = S > 0 selects the cluster kernel with S CTAs per graph (-1: automatic S).  Default off: at cfg2
  // (256 graphs of 30 x 512 floats) it measured 16.4 / 16.4 / 18.5 / 26.6 us for S = 1 / 2 / 4 / 8 against 14.3 us of
  // the one-CTA-per-graph kernel below (profiles/r02/kernel_roofline_block_vs_warp.txt); it is the single-HBM-read
  // path for graphs beyond one CTA's shared memory (> ~100 nodes at F = 512), where it is selected automatically.
  static const int env_cluster = [] {
    const char* e = getenv("GVQA_LN_CLUSTER");
    return e ? atoi(e) : 0;
  }();
  const bool big_graphs = max_nodes_per_graph > 0 && (size_t)max_nodes_per_graph * channels * 4 > 200 * 1024;
  if (env_cluster != 0 || big_graphs) {
    // cluster size: enough CTAs for ~4 per SM, at most 8 (the portable maximum), and no more than the rows allow
    const int64_t rows = max_nodes_per_graph > 0 ? max_nodes_per_graph : (num_nodes + num_graphs - 1) / num_graphs;
    int S = 1;
    while (S < 8 && num_graphs * S < 4 * kNumSMs && rows >= 4 * S) S *= 2;
    // a graph that does not fit S x 200 KB needs the largest cluster
    while (S < 8 && (size_t)((rows + S - 1) / S) * channels * 4 > 200 * 1024) S *= 2;
    if (env_cluster > 0) S = env_cluster;
    size_t want = (size_t)((rows + S - 1) / S) * channels * 4;
    if (want > 200 * 1024) want = 96 * 1024;        // slices that do not fit are re-read from L2
    if (want < 16) want = 16;
    want = (want + 15) & ~(size_t)15;
    static size_t configured = 0;
    if (want > 48 * 1024 && want > configured) {
      if (cudaFuncSetAttribute(graph_layernorm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
          cudaSuccess) {
        (void)cudaGetLastError();
        return GVQA_ERR_CUDA;
      }
      configured = 200 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_graphs * S));
    cfg.blockDim = dim3(kLnClThreads);
    cfg.dynamicSmemBytes = want;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, graph_layernorm_cluster_kernel, x, graph_ptr, weight, bias, out, (int)channels, eps,
                           (int)(want / 4), S) != cudaSuccess) {
      (void)cudaGetLastError();
      return GVQA_ERR_CUDA;
    }
    GVQA_LAUNCH_CHECK();
    return GVQA_OK;
  }
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
