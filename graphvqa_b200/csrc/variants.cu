// Message-passing kernels of the GCN / GINE / LCGN variants (see include/gvqa_b200.h).
//
// Same skeleton as the GAT hop: int32 destination-CSR, one warp per destination node, lanes own
// 128-bit column slices, in-edges summed in the caller's edge order, no atomics.
//   gvqa_gine_aggregate_f32 : torch_geometric GINEConv's propagate + (1+eps)*x_i
//                             (baseline_and_test_models/pipeline_model_gine.py:665)
//   gvqa_gcn_degree_f32 / gvqa_gcn_aggregate_f32 : GCNConv's gcn_norm + propagate + bias
//                             (baseline_and_test_models/pipeline_model_gcn.py:660)
//   gvqa_lcgn_hop_f32       : gat_lcgn.forward/message (baseline_and_test_models/lcgn.py:120-238)
#include "common.cuh"

namespace gvqa {

constexpr int kVarThreads = 128;  // 4 warps, one destination node per warp at a time

// ------------------------------------------------------------------------------------------
// GINE:  z[i, :F]     = (1+eps) h[i] + sum_k relu(h[src_k] + edge_attr[e_k])
//        z[i, F:F+D]  = (1+eps) ins[g] + deg(i) * relu(2 ins[g])      (instruction halves of
//        x_cat / edge_cat are the same per-graph vector, pipeline_model_gine.py:652-661)
// ------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(kVarThreads) gine_aggregate_kernel(
    const float* __restrict__ h, const float* __restrict__ edge_attr, const float* __restrict__ ins,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ node_graph, float* __restrict__ z, int N, int F, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kVarThreads / 32) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int F4 = F >> 2, D4 = D >> 2;
  const int64_t ldz = (int64_t)F + D;
  const float self_scale = 1.0f + eps;
  // lane owns float4 columns lane, lane+32, ... (J of them): the 2 x J loads of an edge are issued together and two
  // edges are unrolled, instead of one dependent round trip per column group
  float4 acc[J], self[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    self[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < F4) self[j] = ldg_cached(h + (int64_t)i * F + 4 * c4);
  }
#pragma unroll 2
  for (int k = e0; k < e1; ++k) {
    const int src = col_src[k];
    const int64_t e = perm ? perm[k] : k;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < F4) {
        const float4 a = ldg_cached(h + (int64_t)src * F + 4 * c4);
        const float4 b = ldg_stream(edge_attr + e * F + 4 * c4);
        acc[j].x += fmaxf(a.x + b.x, 0.f); acc[j].y += fmaxf(a.y + b.y, 0.f);
        acc[j].z += fmaxf(a.z + b.z, 0.f); acc[j].w += fmaxf(a.w + b.w, 0.f);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < F4) {
      // PyG adds (1+eps)*x_i AFTER the aggregation (out += (1 + eps) * x_r)
      float4 o = acc[j];
      o.x += self_scale * self[j].x; o.y += self_scale * self[j].y;
      o.z += self_scale * self[j].z; o.w += self_scale * self[j].w;
      stg_stream(z + (int64_t)i * ldz + 4 * c4, o);
    }
  }
  if (D4 > 0) {
    const float deg = (float)(e1 - e0);
    const float* iv = ins + (int64_t)node_graph[i] * D;
    for (int c4 = lane; c4 < D4; c4 += 32) {
      const float4 v = ldg_cached(iv + 4 * c4);
      float4 o;
      // sum of deg identical terms relu(v+v), accumulated like the reference (repeated addition
      // of equal values == deg * value up to rounding), then the self term
      o.x = deg * fmaxf(v.x + v.x, 0.f) + self_scale * v.x; o.y = deg * fmaxf(v.y + v.y, 0.f) + self_scale * v.y;
      o.z = deg * fmaxf(v.z + v.z, 0.f) + self_scale * v.z; o.w = deg * fmaxf(v.w + v.w, 0.f) + self_scale * v.w;
      stg_stream(z + (int64_t)i * ldz + F + 4 * c4, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GCN.  add_remaining_self_loops: existing self-loop edges are dropped, one loop per node is
// appended; deg(i) = 1 + #non-loop in-edges (duplicates counted); norm_k = dinv[src] dinv[dst].
// ------------------------------------------------------------------------------------------
__global__ void gcn_degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
                                  float* __restrict__ dinv, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int deg = 1;
  for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) deg += col_src[k] != i;
  dinv[i] = 1.0f / sqrtf((float)deg);
}

//   out[i] = sum_{k: src_k != i} dinv[src_k] dinv[i] (xw[src_k] + P[g]) + dinv[i]^2 (xw[i] + P[g]) + b
// with xw = h @ W[:F] and P = ins @ W[F:] (x_cat @ W split by rows of W).
template <int J>
__global__ void __launch_bounds__(kVarThreads) gcn_aggregate_kernel(
    const float* __restrict__ xw, const float* __restrict__ graph_term, const float* __restrict__ dinv,
    const float* __restrict__ bias, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
    const int32_t* __restrict__ node_graph, float* __restrict__ out, int N, int C) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kVarThreads / 32) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int C4 = C >> 2;
  const float di = dinv[i];
  const float* gt = graph_term ? graph_term + (int64_t)node_graph[i] * C : nullptr;
  float4 acc[J], p4[J], self[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    p4[j] = acc[j];
    self[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      if (gt) p4[j] = ldg_cached(gt + 4 * c4);
      self[j] = ldg_cached(xw + (int64_t)i * C + 4 * c4);
    }
  }
#pragma unroll 2
  for (int k = e0; k < e1; ++k) {
    const int src = col_src[k];
    if (src == i) continue;                 // pre-existing self-loops are removed by gcn_norm
    const float w = dinv[src] * di;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        const float4 v = ldg_cached(xw + (int64_t)src * C + 4 * c4);
        acc[j].x += w * (v.x + p4[j].x); acc[j].y += w * (v.y + p4[j].y);
        acc[j].z += w * (v.z + p4[j].z); acc[j].w += w * (v.w + p4[j].w);
      }
    }
  }
  const float ws = di * di;                 // the appended loop comes last in PyG's edge list
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      float4 o = acc[j];
      o.x += ws * (self[j].x + p4[j].x); o.y += ws * (self[j].y + p4[j].y);
      o.z += ws * (self[j].z + p4[j].z); o.w += ws * (self[j].w + p4[j].w);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      stg_stream(out + (int64_t)i * C + 4 * c4, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// LCGN (heads = 1):
//   logit_k = sum_c xl[src_k,c] * (proj_cmd[g,c] * xr[i,c])          (lcgn.py:154, :209)
//   alpha   = segment softmax of leaky_relu(logit)                    (:211-212)
//   out[i]  = cal_cmd[g] * sum_k alpha_k xv[src_k] + bias             (:229-238, :183-186)
// One pass with a running (max, sum) so each source row is read once per in-edge.
// ------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(kVarThreads) lcgn_hop_kernel(
    const float* __restrict__ xl, const float* __restrict__ xr, const float* __restrict__ xv, int64_t ld,
    const float* __restrict__ proj_cmd, const float* __restrict__ cal_cmd, const float* __restrict__ bias,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src, const int32_t* __restrict__ node_graph,
    float* __restrict__ out, int N, int C, float slope) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kVarThreads / 32) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int C4 = C >> 2;
  const int g = node_graph[i];
  float4 q[J], acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc[j] = q[j];
    if (c4 < C4) {
      const float4 a = ldg_cached(proj_cmd + (int64_t)g * C + 4 * c4);
      const float4 b = ldg_stream(xr + (int64_t)i * ld + 4 * c4);
      q[j] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    }
  }
  float m = -INFINITY, s = 0.f;
  for (int k = e0; k < e1; ++k) {
    const int src = col_src[k];
    float4 lv[J], vv[J];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      lv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      vv[j] = lv[j];
      if (c4 < C4) {
        lv[j] = ldg_cached(xl + (int64_t)src * ld + 4 * c4);
        vv[j] = ldg_cached(xv + (int64_t)src * ld + 4 * c4);
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
      dot += (lv[j].x * q[j].x + lv[j].y * q[j].y) + (lv[j].z * q[j].z + lv[j].w * q[j].w);
    const float l = leaky_relu(warp_sum(dot), slope);
    const float m_new = fmaxf(m, l);
    const float rescale = expf(m - m_new);     // exp(-inf) = 0 on the first edge
    const float w = expf(l - m_new);
    s = s * rescale + w;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      acc[j].x = acc[j].x * rescale + w * vv[j].x; acc[j].y = acc[j].y * rescale + w * vv[j].y;
      acc[j].z = acc[j].z * rescale + w * vv[j].z; acc[j].w = acc[j].w * rescale + w * vv[j].w;
    }
    m = m_new;
  }
  const float inv = 1.0f / (s + 1e-16f);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      const float4 cc = ldg_cached(cal_cmd + (int64_t)g * C + 4 * c4);
      float4 o = make_float4(acc[j].x * inv * cc.x, acc[j].y * inv * cc.y, acc[j].z * inv * cc.z, acc[j].w * inv * cc.w);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      stg_stream(out + (int64_t)i * C + 4 * c4, o);
    }
  }
}


// ==========================================================================================
// Block-phase versions (default).  The warp-per-node kernels above pay 3-5 DEPENDENT round trips per node
// (row pointers -> sources -> per-source scalars -> rows), which at GQA sizes (N ~ 4-8 k nodes, 2-4 in-edges each)
// is pure latency: 30-35 % of the HBM peak (profiles/r01/kernel_roofline.txt).  Like the fused GAT hop, a 128-thread
// CTA now owns `npc` consecutive destination nodes (npc = ceil(N / (148 SMs x 4 CTAs)), one resident wave), loads its
// CSR slice and the per-edge scalars edge-parallel for the whole CTA, and only then lets every warp stream its nodes
// with all row loads of a node in flight together.  CTAs whose in-edges exceed the staging capacity run the
// warp-per-node code path (same arithmetic, same order).
// ==========================================================================================
constexpr int kVbNodes = 16, kVbThreads = 128, kVbEdgeCap = 256;

struct VbTopo {
  int i0, nn, eA, eC;
};

__host__ __device__ inline int vb_nodes_per_cta(int64_t N) {
  int npc = (int)((N + kNumSMs * 4 - 1) / (kNumSMs * 4));
  return npc < 4 ? 4 : (npc > kVbNodes ? kVbNodes : npc);
}

// round trips 1 + 2 of the block: row pointers, then sources (and original edge ids).  Returns false when the block's
// in-edges do not fit (the caller then takes the per-warp path).  The caller synchronises after its own edge loads.
__device__ __forceinline__ bool vb_load(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
                                        const int32_t* __restrict__ perm, int N, int npc, int32_t* rp_s, int32_t* src_s,
                                        int32_t* eid_s, VbTopo& t) {
  const int tid = threadIdx.x;
  t.i0 = blockIdx.x * npc;
  t.nn = min(npc, N - t.i0);
  for (int q = tid; q <= t.nn; q += kVbThreads) rp_s[q] = rowptr[t.i0 + q];
  __syncthreads();
  t.eA = rp_s[0];
  t.eC = rp_s[t.nn] - t.eA;
  if (t.eC > kVbEdgeCap) return false;
  for (int k = tid; k < t.eC; k += kVbThreads) {
    src_s[k] = col_src[t.eA + k];
    if (eid_s) eid_s[k] = perm ? perm[t.eA + k] : (t.eA + k);
  }
  return true;
}

template <int J>
__device__ __forceinline__ void gcn_node(const float* __restrict__ xw, const float* __restrict__ gt, const float* __restrict__ bias,
                                         float* __restrict__ out, int i, int C, int lane, float di, int e0, int e1,
                                         const int32_t* src_of, const float* dinv_of, bool staged,
                                         const int32_t* __restrict__ col_src, const float* __restrict__ dinv) {
  const int C4 = C >> 2;
  float4 acc[J], p4[J], self[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    p4[j] = acc[j];
    self[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      if (gt) p4[j] = ldg_cached(gt + 4 * c4);
      self[j] = ldg_cached(xw + (int64_t)i * C + 4 * c4);
    }
  }
#pragma unroll 2
  for (int k = e0; k < e1; ++k) {
    const int src = staged ? src_of[k] : col_src[k];
    if (src == i) continue;                 // pre-existing self-loops are removed by gcn_norm
    const float w = (staged ? dinv_of[k] : dinv[src]) * di;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        const float4 v = ldg_cached(xw + (int64_t)src * C + 4 * c4);
        acc[j].x += w * (v.x + p4[j].x); acc[j].y += w * (v.y + p4[j].y);
        acc[j].z += w * (v.z + p4[j].z); acc[j].w += w * (v.w + p4[j].w);
      }
    }
  }
  const float ws = di * di;                 // the appended loop comes last in PyG's edge list
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      float4 o = acc[j];
      o.x += ws * (self[j].x + p4[j].x); o.y += ws * (self[j].y + p4[j].y);
      o.z += ws * (self[j].z + p4[j].z); o.w += ws * (self[j].w + p4[j].w);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      stg_stream(out + (int64_t)i * C + 4 * c4, o);
    }
  }
}

template <int J>
__global__ void __launch_bounds__(kVbThreads) gcn_aggregate_block_kernel(
    const float* __restrict__ xw, const float* __restrict__ graph_term, const float* __restrict__ dinv,
    const float* __restrict__ bias, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
    const int32_t* __restrict__ node_graph, float* __restrict__ out, int N, int C, int npc) {
  __shared__ int32_t rp_s[kVbNodes + 1], src_s[kVbEdgeCap], gid_s[kVbNodes];
  __shared__ float dsrc_s[kVbEdgeCap], di_s[kVbNodes];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  VbTopo t;
  const bool staged = vb_load(rowptr, col_src, nullptr, N, npc, rp_s, src_s, nullptr, t);
  if (tid < t.nn) {
    di_s[tid] = dinv[t.i0 + tid];
    gid_s[tid] = graph_term ? node_graph[t.i0 + tid] : 0;
  }
  __syncthreads();
  if (staged) {
    for (int k = tid; k < t.eC; k += kVbThreads) dsrc_s[k] = dinv[src_s[k]];     // round trip 3, edge-parallel
    __syncthreads();
  }
  for (int node = wid; node < t.nn; node += kVbThreads / 32) {
    const int i = t.i0 + node;
    const float* gt = graph_term ? graph_term + (int64_t)gid_s[node] * C : nullptr;
    if (staged)
      gcn_node<J>(xw, gt, bias, out, i, C, lane, di_s[node], rp_s[node] - t.eA, rp_s[node + 1] - t.eA, src_s, dsrc_s, true,
                  col_src, dinv);
    else
      gcn_node<J>(xw, gt, bias, out, i, C, lane, di_s[node], rp_s[node], rp_s[node + 1], nullptr, nullptr, false, col_src, dinv);
  }
}

template <int J>
__global__ void __launch_bounds__(kVbThreads) gine_aggregate_block_kernel(
    const float* __restrict__ h, const float* __restrict__ edge_attr, const float* __restrict__ ins,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ node_graph, float* __restrict__ z, int N, int F, int D, float eps, int npc) {
  __shared__ int32_t rp_s[kVbNodes + 1], src_s[kVbEdgeCap], eid_s[kVbEdgeCap], gid_s[kVbNodes];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  VbTopo t;
  const bool staged = vb_load(rowptr, col_src, perm, N, npc, rp_s, src_s, eid_s, t);
  if (tid < t.nn) gid_s[tid] = D > 0 ? node_graph[t.i0 + tid] : 0;
  __syncthreads();
  const int F4 = F >> 2, D4 = D >> 2;
  const int64_t ldz = (int64_t)F + D;
  const float self_scale = 1.0f + eps;
  for (int node = wid; node < t.nn; node += kVbThreads / 32) {
    const int i = t.i0 + node;
    const int e0 = rp_s[node], e1 = rp_s[node + 1];
    float4 acc[J], self[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      self[j] = acc[j];
      const int c4 = lane + 32 * j;
      if (c4 < F4) self[j] = ldg_cached(h + (int64_t)i * F + 4 * c4);
    }
#pragma unroll 2
    for (int k = e0; k < e1; ++k) {
      const int src = staged ? src_s[k - t.eA] : col_src[k];
      const int64_t e = staged ? eid_s[k - t.eA] : (perm ? perm[k] : k);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < F4) {
          const float4 a = ldg_cached(h + (int64_t)src * F + 4 * c4);
          const float4 b = ldg_stream(edge_attr + e * F + 4 * c4);
          acc[j].x += fmaxf(a.x + b.x, 0.f); acc[j].y += fmaxf(a.y + b.y, 0.f);
          acc[j].z += fmaxf(a.z + b.z, 0.f); acc[j].w += fmaxf(a.w + b.w, 0.f);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < F4) {
        float4 o = acc[j];           // PyG adds (1+eps)*x_i AFTER the aggregation
        o.x += self_scale * self[j].x; o.y += self_scale * self[j].y;
        o.z += self_scale * self[j].z; o.w += self_scale * self[j].w;
        stg_stream(z + (int64_t)i * ldz + 4 * c4, o);
      }
    }
    if (D4 > 0) {
      const float deg = (float)(e1 - e0);
      const float* iv = ins + (int64_t)gid_s[node] * D;
      for (int c4 = lane; c4 < D4; c4 += 32) {
        const float4 v = ldg_cached(iv + 4 * c4);
        float4 o;
        o.x = deg * fmaxf(v.x + v.x, 0.f) + self_scale * v.x; o.y = deg * fmaxf(v.y + v.y, 0.f) + self_scale * v.y;
        o.z = deg * fmaxf(v.z + v.z, 0.f) + self_scale * v.z; o.w = deg * fmaxf(v.w + v.w, 0.f) + self_scale * v.w;
        stg_stream(z + (int64_t)i * ldz + F + 4 * c4, o);
      }
    }
  }
}

// LCGN, block-phase and two-pass: pass 1 reads the xl rows of a node's in-edges (all loads independent) and leaves the
// logits in shared memory, the segment softmax is evaluated exactly like PyG's (max, exp(l - max), / (sum + 1e-16)),
// pass 2 streams the xv rows with their weights -- xl and xv are different arrays, so every byte is still read once per
// in-edge, but no load waits for the running (max, sum) chain any more.
template <int J>
__global__ void __launch_bounds__(kVbThreads) lcgn_hop_block_kernel(
    const float* __restrict__ xl, const float* __restrict__ xr, const float* __restrict__ xv, int64_t ld,
    const float* __restrict__ proj_cmd, const float* __restrict__ cal_cmd, const float* __restrict__ bias,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src, const int32_t* __restrict__ node_graph,
    float* __restrict__ out, int N, int C, float slope, int npc) {
  __shared__ int32_t rp_s[kVbNodes + 1], src_s[kVbEdgeCap], gid_s[kVbNodes];
  __shared__ float logit_s[kVbEdgeCap];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  VbTopo t;
  const bool staged = vb_load(rowptr, col_src, nullptr, N, npc, rp_s, src_s, nullptr, t);
  if (tid < t.nn) gid_s[tid] = node_graph[t.i0 + tid];
  __syncthreads();
  const int C4 = C >> 2;
  for (int node = wid; node < t.nn; node += kVbThreads / 32) {
    const int i = t.i0 + node;
    const int g = gid_s[node];
    const int e0 = rp_s[node], e1 = rp_s[node + 1];
    float4 q[J], acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = q[j];
      if (c4 < C4) {
        const float4 a = ldg_cached(proj_cmd + (int64_t)g * C + 4 * c4);
        const float4 b = ldg_stream(xr + (int64_t)i * ld + 4 * c4);
        q[j] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
      }
    }
    float inv = 0.f, m = 0.f;
    if (staged) {
      // pass 1: logits of all in-edges (two edges' loads in flight), kept in shared memory
#pragma unroll 2
      for (int k = e0; k < e1; ++k) {
        const int src = src_s[k - t.eA];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < C4) {
            const float4 lv = ldg_cached(xl + (int64_t)src * ld + 4 * c4);
            dot += (lv.x * q[j].x + lv.y * q[j].y) + (lv.z * q[j].z + lv.w * q[j].w);
          }
        }
        const float l = leaky_relu(warp_sum(dot), slope);
        if (lane == 0) logit_s[k - t.eA] = l;
      }
      __syncwarp();
      float mx = -INFINITY;
      for (int k = e0 + lane; k < e1; k += 32) mx = fmaxf(mx, logit_s[k - t.eA]);
      m = warp_max(mx);
      float s = 0.f;
      for (int k = e0; k < e1; ++k) s += expf(logit_s[k - t.eA] - m);      // edge order, like a sequential scatter
      inv = 1.0f / (s + 1e-16f);
      // pass 2: weighted sum of the xv rows
#pragma unroll 2
      for (int k = e0; k < e1; ++k) {
        const int src = src_s[k - t.eA];
        const float w = expf(logit_s[k - t.eA] - m) * inv;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < C4) fma4(acc[j], w, ldg_cached(xv + (int64_t)src * ld + 4 * c4));
        }
      }
      inv = 1.0f;      // the weights are already normalised
    } else {
      // oversize block: one pass with a running (max, sum), rows from global memory
      float s = 0.f;
      m = -INFINITY;
      for (int k = e0; k < e1; ++k) {
        const int src = col_src[k];
        float4 lv[J], vv[J];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          lv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          vv[j] = lv[j];
          if (c4 < C4) {
            lv[j] = ldg_cached(xl + (int64_t)src * ld + 4 * c4);
            vv[j] = ldg_cached(xv + (int64_t)src * ld + 4 * c4);
          }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) dot += (lv[j].x * q[j].x + lv[j].y * q[j].y) + (lv[j].z * q[j].z + lv[j].w * q[j].w);
        const float l = leaky_relu(warp_sum(dot), slope);
        const float m_new = fmaxf(m, l);
        const float rescale = expf(m - m_new);
        const float w = expf(l - m_new);
        s = s * rescale + w;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[j].x = acc[j].x * rescale + w * vv[j].x; acc[j].y = acc[j].y * rescale + w * vv[j].y;
          acc[j].z = acc[j].z * rescale + w * vv[j].z; acc[j].w = acc[j].w * rescale + w * vv[j].w;
        }
        m = m_new;
      }
      inv = 1.0f / (s + 1e-16f);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        const float4 cc = ldg_cached(cal_cmd + (int64_t)g * C + 4 * c4);
        float4 o = make_float4(acc[j].x * inv * cc.x, acc[j].y * inv * cc.y, acc[j].z * inv * cc.z, acc[j].w * inv * cc.w);
        if (bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        stg_stream(out + (int64_t)i * C + 4 * c4, o);
      }
    }
  }
}

}  // namespace gvqa

using namespace gvqa;

// GVQA_VARIANT_BLOCK: bit 0 GCN, bit 1 GINE, bit 2 LCGN run the block-phase kernel.  Default 1: measured on B200 at
// the BASELINE shapes (profiles/r02/kernel_roofline_block_vs_warp.txt) the block-phase mapping gains only for GCN
// (16.4 -> 15.4 us); GINE (24.6 -> 27.6 us) and LCGN (14.4 -> 18.5 us) move too few bytes per node for 16 warps per
// SM walking their nodes serially -- the warp-per-node kernels keep up to 64 independent warps per SM in flight.
static int variants_block_mask() {
  static const int mask = [] {
    const char* e = getenv("GVQA_VARIANT_BLOCK");
    return e ? atoi(e) : 1;
  }();
  return mask;
}

extern "C" GVQA_API int gvqa_gine_aggregate_f32(const float* h, const float* edge_attr, const float* ins,
                                                const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                                const int32_t* node_graph, float* z, int64_t num_nodes,
                                                int32_t feat, int32_t ins_dim, float eps, void* stream_) {
  if (num_nodes < 0 || feat <= 0 || ins_dim < 0 || num_nodes >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!h || !edge_attr || !rowptr || !col_src || !z || (ins_dim > 0 && (!ins || !node_graph)))
    return GVQA_ERR_NULL_POINTER;
  if ((feat & 3) || (ins_dim & 3)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(h) || !aligned16(edge_attr) || !aligned16(z) || (ins && !aligned16(ins))) return GVQA_ERR_MISALIGNED;
  if (feat > 1024) return GVQA_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((num_nodes + 3) / 4);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int j = (feat / 4 + 31) / 32;
  const int npc = vb_nodes_per_cta(num_nodes);
  const unsigned bgrid = (unsigned)((num_nodes + npc - 1) / npc);
#define GVQA_GINE(JJ)                                                                                          \
  if (variants_block_mask() & 2)                                                                               \
    gine_aggregate_block_kernel<JJ><<<bgrid, kVbThreads, 0, stream>>>(h, edge_attr, ins, rowptr, col_src, perm, node_graph, \
                                                                      z, (int)num_nodes, feat, ins_dim, eps, npc); \
  else                                                                                                         \
    gine_aggregate_kernel<JJ><<<grid, kVarThreads, 0, stream>>>(h, edge_attr, ins, rowptr, col_src, perm, node_graph, z, \
                                                               (int)num_nodes, feat, ins_dim, eps)
  if (j <= 1) GVQA_GINE(1);
  else if (j == 2) GVQA_GINE(2);
  else if (j <= 4) GVQA_GINE(4);
  else GVQA_GINE(8);
#undef GVQA_GINE
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gcn_degree_f32(const int32_t* rowptr, const int32_t* col_src, float* dinv,
                                            int64_t num_nodes, void* stream_) {
  if (num_nodes < 0 || num_nodes >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!rowptr || !col_src || !dinv) return GVQA_ERR_NULL_POINTER;
  gcn_degree_kernel<<<(unsigned)((num_nodes + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      rowptr, col_src, dinv, (int)num_nodes);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gcn_aggregate_f32(const float* xw, const float* graph_term, const float* dinv,
                                               const float* bias, const int32_t* rowptr, const int32_t* col_src,
                                               const int32_t* node_graph, float* out, int64_t num_nodes,
                                               int32_t channels, void* stream_) {
  if (num_nodes < 0 || channels <= 0 || num_nodes >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!xw || !dinv || !rowptr || !col_src || !out || (graph_term && !node_graph)) return GVQA_ERR_NULL_POINTER;
  if (channels & 3) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(xw) || !aligned16(out) || (graph_term && !aligned16(graph_term)) || (bias && !aligned16(bias)))
    return GVQA_ERR_MISALIGNED;
  if (channels > 1024) return GVQA_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((num_nodes + 3) / 4);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int j = (channels / 4 + 31) / 32;
  const int npc = vb_nodes_per_cta(num_nodes);
  const unsigned bgrid = (unsigned)((num_nodes + npc - 1) / npc);
#define GVQA_GCN(JJ)                                                                                        \
  if (variants_block_mask() & 1)                                                                            \
    gcn_aggregate_block_kernel<JJ><<<bgrid, kVbThreads, 0, stream>>>(xw, graph_term, dinv, bias, rowptr, col_src, node_graph, \
                                                                     out, (int)num_nodes, channels, npc);   \
  else                                                                                                      \
    gcn_aggregate_kernel<JJ><<<grid, kVarThreads, 0, stream>>>(xw, graph_term, dinv, bias, rowptr, col_src, node_graph, \
                                                              out, (int)num_nodes, channels)
  if (j <= 1) GVQA_GCN(1);
  else if (j == 2) GVQA_GCN(2);
  else if (j <= 4) GVQA_GCN(4);
  else GVQA_GCN(8);
#undef GVQA_GCN
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_lcgn_hop_f32(const float* xl, const float* xr, const float* xv, int64_t ld,
                                          const float* proj_cmd, const float* cal_cmd, const float* bias,
                                          const int32_t* rowptr, const int32_t* col_src, const int32_t* node_graph,
                                          float* out, int64_t num_nodes, int32_t channels, float negative_slope,
                                          void* stream_) {
  if (num_nodes < 0 || channels <= 0 || ld < channels || num_nodes >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!xl || !xr || !xv || !proj_cmd || !cal_cmd || !rowptr || !col_src || !node_graph || !out)
    return GVQA_ERR_NULL_POINTER;
  if ((channels & 3) || channels > 1024) return GVQA_ERR_UNSUPPORTED;
  if ((ld & 3) || !aligned16(xl) || !aligned16(xr) || !aligned16(xv) || !aligned16(proj_cmd) || !aligned16(cal_cmd) ||
      !aligned16(out) || (bias && !aligned16(bias)))
    return GVQA_ERR_MISALIGNED;
  const unsigned grid = (unsigned)((num_nodes + 3) / 4);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int j = (channels / 4 + 31) / 32;
  const int npc = vb_nodes_per_cta(num_nodes);
  const unsigned bgrid = (unsigned)((num_nodes + npc - 1) / npc);
#define GVQA_LCGN(JJ)                                                                                         \
  if (variants_block_mask() & 4)                                                                              \
    lcgn_hop_block_kernel<JJ><<<bgrid, kVbThreads, 0, stream>>>(xl, xr, xv, ld, proj_cmd, cal_cmd, bias, rowptr, col_src, \
                                                                node_graph, out, (int)num_nodes, channels, negative_slope, npc); \
  else                                                                                                        \
    lcgn_hop_kernel<JJ><<<grid, kVarThreads, 0, stream>>>(xl, xr, xv, ld, proj_cmd, cal_cmd, bias, rowptr, col_src, \
                                                          node_graph, out, (int)num_nodes, channels, negative_slope)
  if (j <= 1) GVQA_LCGN(1);
  else if (j == 2) GVQA_LCGN(2);
  else if (j <= 4) GVQA_LCGN(4);
  else GVQA_LCGN(8);
#undef GVQA_LCGN
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
