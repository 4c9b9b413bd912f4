// Segment / gather kernels for the steps right before and after the hop stack:
//   gvqa_gather_add_relu_f32   : act(a[src_k] + b[dst_k] + c[k] + bias)  -- the first Linear of the
//       scene-graph encoder's edge / node MLPs after splitting the concatenated input
//       (pipeline_model_gat.py:75-77 cat([src, dest, edge_attr]); :94 cat([x[row], edge_attr])),
//       so the [E,900] / [E,600] concatenations are never materialised
//   gvqa_segment_mean_rows_f32 : torch_scatter.scatter_mean(out, col, dim=0, dim_size=N) of the node
//       model (pipeline_model_gat.py:96) over the batch's destination-CSR (no atomics, edge order)
//   gvqa_attention_pool_f32    : MyConditionalGlobalAttention's per-graph softmax + weighted sum
//       (pipeline_model_gat.py:178-179), PyG softmax semantics
// All HBM-bound: warp per output row, 128-bit lanes.
#include "common.cuh"

namespace gvqa {

template <typename Index>
__global__ void __launch_bounds__(256) gather_add_relu_kernel(
    const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, const float* __restrict__ c,
    int64_t ldc, const float* __restrict__ bias, const Index* __restrict__ edge_index, float* __restrict__ out, int64_t E,
    int F, int relu) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= E) return;
  const int64_t src = edge_index ? (int64_t)edge_index[k] : k, dst = edge_index ? (int64_t)edge_index[E + k] : k;
  const int F4 = F >> 2;
  for (int c4 = lane; c4 < F4; c4 += 32) {
    float4 v = ldg_cached(a + src * lda + 4 * c4);
    if (b) {
      const float4 u = ldg_cached(b + dst * ldb + 4 * c4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (c) {
      const float4 u = ldg_stream(c + k * ldc + 4 * c4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (bias) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    stg_stream(out + k * F + 4 * c4, v);
  }
}

__global__ void __launch_bounds__(256) segment_mean_rows_kernel(
    const float* __restrict__ values, const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
    float* __restrict__ out, int S, int F, int mean) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= S) return;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int F4 = F >> 2;
  for (int c4 = lane; c4 < F4; c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = e0; k < e1; ++k) {
      const int64_t e = perm ? perm[k] : k;
      const float4 v = ldg_stream(values + e * F + 4 * c4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (mean) {   // sum / count.clamp(min=1), like torch_scatter (a true division)
      const float cnt = (float)max(e1 - e0, 1);
      acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
    }
    stg_stream(out + (int64_t)i * F + 4 * c4, acc);
  }
}

// out[n,:] = x[n,:] * q[g(n),:]  -- the `ques_nn(u)[batch] * node_nn(x)` product of MyConditionalGlobalAttention
// (pipeline_model_gat.py:166-167) without materialising q[batch]
__global__ void __launch_bounds__(256) graph_scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                                               const int32_t* __restrict__ node_graph,
                                                               float* __restrict__ out, int64_t N, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N) return;
  const float* qr = q + (int64_t)node_graph[i] * C;
  for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
    float4 v = ldg_stream(x + i * C + 4 * c4);
    const float4 u = ldg_cached(qr + 4 * c4);
    v.x *= u.x; v.y *= u.y; v.z *= u.z; v.w *= u.w;
    stg_stream(out + i * C + 4 * c4, v);
  }
}

// out[n,:] = sign[n] * sum_t table[tokens[n,t],:]  -- the token-embedding sums that open the scene-graph encoder
// (pipeline_model_gat.py:583-594: sg_vocab_embedding(x).sum(-2); the `added_sym_edge` negation of edge rows as a
// per-row sign).  The reference materialises [N, T, F] (189 MB at cfg2) and reduces it; here one warp per row reads
// the T table rows (the table is L2-resident) and writes F floats.
template <typename Index>
__global__ void __launch_bounds__(256) embedding_sum_kernel(const float* __restrict__ table, const Index* __restrict__ tokens,
                                                            const float* __restrict__ sign, float* __restrict__ out,
                                                            int64_t N, int T, int F, int64_t vocab) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N) return;
  const Index* tk = tokens + i * T;
  const float sg = sign ? sign[i] : 1.0f;
  for (int c4 = lane; c4 < (F >> 2); c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
      int64_t id = (int64_t)tk[t];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);      // out-of-range ids are a caller bug: clamp, never fault
      const float4 v = ldg_cached(table + id * F + 4 * c4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (sign) { acc.x *= sg; acc.y *= sg; acc.z *= sg; acc.w *= sg; }
    stg_stream(out + i * F + 4 * c4, acc);
  }
}

// out[n,c] = act(x[n,c] * scale[c] + shift[c])  -- BatchNorm1d(eval) folded to an affine map + ReLU, the whole
// bug-faithful GCN / GINE hop (the conv result is discarded by the reference, pipeline_model_gcn.py:660-668)
__global__ void __launch_bounds__(256) affine_relu_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ out,
                                                          int64_t total4, int C4, int relu) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += stride) {
    const int c4 = (int)(t % C4);
    float4 v = ldg_stream(x + 4 * t);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    stg_stream(out + 4 * t, v);
  }
}

// One CTA per graph: softmax of the per-node gate over the graph's nodes, then sum_n w_n x[n,:].
// With `hid` given, the gate itself is computed here first: gate[n] = <hid[n,:], w_gate> + b_gate (the last Linear
// of gate_nn, pipeline_model_gat.py:128 / :168), one warp per node, and parked in `gate` (scratch, [N]).
__global__ void __launch_bounds__(256) attention_pool_kernel(float* __restrict__ gate,
                                                             const float* __restrict__ hid, int Ch,
                                                             const float* __restrict__ w_gate,
                                                             const float* __restrict__ b_gate,
                                                             const float* __restrict__ x,
                                                             const int32_t* __restrict__ graph_ptr,
                                                             float* __restrict__ out, int C) {
  __shared__ float red[8];
  __shared__ float w_s[256];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n0 = graph_ptr[g], n1 = graph_ptr[g + 1], n = n1 - n0;
  if (hid != nullptr) {
    const float b0 = b_gate ? __ldg(b_gate) : 0.f;
    for (int i = wid; i < n; i += 8) {
      const float* hr = hid + (int64_t)(n0 + i) * Ch;
      float acc = 0.f;
      for (int c4 = lane; c4 < (Ch >> 2); c4 += 32) {
        const float4 v = ldg_stream(hr + 4 * c4);
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_gate) + c4);
        acc = fmaf(v.x, w.x, acc); acc = fmaf(v.y, w.y, acc); acc = fmaf(v.z, w.z, acc); acc = fmaf(v.w, w.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) gate[n0 + i] = acc + b0;
    }
    __syncthreads();     // the CTA's own global writes are visible to its threads after the barrier
  }
  // pass 1: segment max (empty segment -> 0, and the output row is 0)
  float mx = -INFINITY;
  for (int i = tid; i < n; i += 256) mx = fmaxf(mx, __ldcg(gate + n0 + i));
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  if (n == 0) mx = 0.f;
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < n; i += 256) sum += expf(__ldcg(gate + n0 + i) - mx);
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float denom = sum + 1e-16f;
  // pass 2: weighted sum over the graph's nodes.  The 256 threads form G = 256 / C4 groups (2 at C = 512, 3 at 300);
  // group r sums nodes r, r+G, ... of every 256-node chunk with eight independent row loads in flight per thread, the
  // G partial rows are added in group order through shared memory (deterministic).  One group walking all nodes with
  // four loads in flight was ~8 serial round trips for a 30-node graph.
  __shared__ __align__(16) float part_s[256 * 4];
  const int C4 = C >> 2;
  for (int c_base = 0; c_base < C4; c_base += 256) {            // all threads take part in the barriers
    const int w4 = min(256, C4 - c_base);
    const int G = max(1, min(8, 256 / w4));
    const int grp = tid / w4, c4 = c_base + (tid - grp * w4);
    const bool active = grp < G;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = 0; base < n; base += 256) {
      __syncthreads();
      if (base + tid < n) w_s[tid] = expf(__ldcg(gate + n0 + base + tid) - mx) / denom;
      __syncthreads();
      if (active) {
        const int lim = min(256, n - base);
#pragma unroll 8
        for (int i = grp; i < lim; i += G) fma4(acc, w_s[i], ldg_stream(x + (int64_t)(n0 + base + i) * C + 4 * c4));
      }
    }
    __syncthreads();
    if (active) *reinterpret_cast<float4*>(part_s + 4 * tid) = acc;
    __syncthreads();
    if (grp == 0) {
      for (int r = 1; r < G; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(part_s + 4 * (r * w4 + tid));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      stg_stream(out + (int64_t)g * C + 4 * c4, acc);
    }
  }
}

}  // namespace gvqa

using namespace gvqa;

extern "C" GVQA_API int gvqa_gather_add_relu_strided_f32(const float* a, int64_t lda, const float* b, int64_t ldb,
                                                         const float* c, int64_t ldc, const float* bias,
                                                         const void* edge_index, int32_t index_bytes, float* out,
                                                         int64_t num_edges, int32_t feat, int32_t relu, void* stream_) {
  if (num_edges < 0 || feat <= 0 || lda < feat || (b && ldb < feat) || (c && ldc < feat)) return GVQA_ERR_BAD_SHAPE;
  if (num_edges == 0) return GVQA_OK;
  if (!a || !out) return GVQA_ERR_NULL_POINTER;      // edge_index == NULL: no gather, out[k] = act(a[k] + b[k] + c[k] + bias)
  if ((feat & 3) || (index_bytes != 4 && index_bytes != 8)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(a) || !aligned16(out) || (b && !aligned16(b)) || (c && !aligned16(c)) || (bias && !aligned16(bias)) ||
      (lda & 3) || (ldb & 3) || (ldc & 3))
    return GVQA_ERR_MISALIGNED;
  const unsigned grid = (unsigned)((num_edges + 7) / 8);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (index_bytes == 8)
    gather_add_relu_kernel<int64_t><<<grid, 256, 0, stream>>>(a, lda, b, ldb, c, ldc, bias,
                                                             static_cast<const int64_t*>(edge_index), out, num_edges, feat, relu);
  else
    gather_add_relu_kernel<int32_t><<<grid, 256, 0, stream>>>(a, lda, b, ldb, c, ldc, bias,
                                                             static_cast<const int32_t*>(edge_index), out, num_edges, feat, relu);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gather_add_relu_f32(const float* a, const float* b, const float* c, const float* bias,
                                                 const int64_t* edge_index, float* out, int64_t num_edges,
                                                 int32_t feat, int32_t relu, void* stream_) {
  return gvqa_gather_add_relu_strided_f32(a, feat, b, feat, c, feat, bias, edge_index, 8, out, num_edges, feat, relu, stream_);
}

extern "C" GVQA_API int gvqa_gather_add_relu_i32_f32(const float* a, const float* b, const float* c, const float* bias,
                                                     const int32_t* edge_index, float* out, int64_t num_edges,
                                                     int32_t feat, int32_t relu, void* stream_) {
  return gvqa_gather_add_relu_strided_f32(a, feat, b, feat, c, feat, bias, edge_index, 4, out, num_edges, feat, relu, stream_);
}

extern "C" GVQA_API int gvqa_embedding_sum_f32(const float* table, int64_t vocab, const void* tokens, int32_t token_bytes,
                                               const float* sign, float* out, int64_t num_rows, int32_t tokens_per_row,
                                               int32_t feat, void* stream_) {
  if (num_rows < 0 || tokens_per_row <= 0 || feat <= 0 || vocab <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_rows == 0) return GVQA_OK;
  if (!table || !tokens || !out) return GVQA_ERR_NULL_POINTER;
  if ((feat & 3) || (token_bytes != 4 && token_bytes != 8)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(table) || !aligned16(out)) return GVQA_ERR_MISALIGNED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = (unsigned)((num_rows + 7) / 8);
  if (token_bytes == 8)
    embedding_sum_kernel<int64_t><<<grid, 256, 0, stream>>>(table, static_cast<const int64_t*>(tokens), sign, out, num_rows,
                                                            tokens_per_row, feat, vocab);
  else
    embedding_sum_kernel<int32_t><<<grid, 256, 0, stream>>>(table, static_cast<const int32_t*>(tokens), sign, out, num_rows,
                                                            tokens_per_row, feat, vocab);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_affine_relu_f32(const float* x, const float* scale, const float* shift, float* out,
                                             int64_t num_rows, int32_t channels, int32_t relu, void* stream_) {
  if (num_rows < 0 || channels <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_rows == 0) return GVQA_OK;
  if (!x || !scale || !shift || !out) return GVQA_ERR_NULL_POINTER;
  if (channels & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(out) || !aligned16(scale) || !aligned16(shift)) return GVQA_ERR_MISALIGNED;
  const int64_t total4 = num_rows * (channels >> 2);
  int64_t blocks = (total4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;     // grid-stride over a resident grid
  affine_relu_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, scale, shift, out, total4,
                                                                                        channels >> 2, relu);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_graph_scale_rows_f32(const float* x, const float* q, const int32_t* node_graph, float* out,
                                                  int64_t num_nodes, int32_t channels, void* stream_) {
  if (num_nodes < 0 || channels <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!x || !q || !node_graph || !out) return GVQA_ERR_NULL_POINTER;
  if (channels & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(q) || !aligned16(out)) return GVQA_ERR_MISALIGNED;
  graph_scale_rows_kernel<<<(unsigned)((num_nodes + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, q, node_graph, out, num_nodes, channels);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_segment_mean_rows_f32(const float* values, const int32_t* perm, const int32_t* rowptr,
                                                   float* out, int64_t num_segments, int32_t feat, int32_t mean,
                                                   void* stream_) {
  if (num_segments < 0 || feat <= 0 || num_segments >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_segments == 0) return GVQA_OK;
  if (!values || !rowptr || !out) return GVQA_ERR_NULL_POINTER;
  if (feat & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(values) || !aligned16(out)) return GVQA_ERR_MISALIGNED;
  segment_mean_rows_kernel<<<(unsigned)((num_segments + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      values, perm, rowptr, out, (int)num_segments, feat, mean);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_attention_pool_f32(const float* gate, const float* x, const int32_t* graph_ptr, float* out,
                                                int64_t num_graphs, int32_t channels, void* stream_) {
  if (num_graphs < 0 || channels <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_graphs == 0) return GVQA_OK;
  if (!gate || !x || !graph_ptr || !out) return GVQA_ERR_NULL_POINTER;
  if (channels & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(out)) return GVQA_ERR_MISALIGNED;
  attention_pool_kernel<<<(unsigned)num_graphs, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      const_cast<float*>(gate), nullptr, 0, nullptr, nullptr, x, graph_ptr, out, channels);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_attention_pool_gate_f32(const float* hid, int32_t hid_channels, const float* w_gate,
                                                     const float* b_gate, float* gate_scratch, const float* x,
                                                     const int32_t* graph_ptr, float* out, int64_t num_graphs,
                                                     int32_t channels, void* stream_) {
  if (num_graphs < 0 || channels <= 0 || hid_channels <= 0) return GVQA_ERR_BAD_SHAPE;
  if (num_graphs == 0) return GVQA_OK;
  if (!hid || !w_gate || !gate_scratch || !x || !graph_ptr || !out) return GVQA_ERR_NULL_POINTER;
  if ((channels & 3) || (hid_channels & 3)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(out) || !aligned16(hid) || !aligned16(w_gate)) return GVQA_ERR_MISALIGNED;
  attention_pool_kernel<<<(unsigned)num_graphs, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      gate_scratch, hid, hid_channels, w_gate, b_gate, x, graph_ptr, out, channels);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
