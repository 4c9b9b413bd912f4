// Destination-CSR build for batched disjoint scene graphs (see include/gvqa_b200.h).
//
// Replaces the per-call edge bookkeeping of torch_geometric's MessagePassing.propagate
// (reference gat_skip.py:155) with a once-per-batch integer pre-pass:
//   count in-degrees -> single-CTA exclusive scan -> slot fill -> per-row rank sort (stable).
// Everything is int32 on the device; the reference's int64 COO is only read.
#include "common.cuh"

namespace gvqa {

// stats layout (int32[8]): 0 max nodes/graph, 1 max in-edges/graph, 2 max in-degree, 3 bad edges
__global__ void csr_count_kernel(const int64_t* __restrict__ edge_index, int64_t E,
                                 const int64_t* __restrict__ batch, int64_t N, int64_t B,
                                 int32_t* __restrict__ deg, int32_t* __restrict__ graph_ptr,
                                 int32_t* __restrict__ node_graph, int32_t* __restrict__ stats) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < E) {
    const int64_t s = edge_index[t], d = edge_index[E + t];
    if (s < 0 || s >= N || d < 0 || d >= N) {
      atomicAdd(&stats[3], 1);
    } else {
      atomicAdd(&deg[d], 1);
      if (batch[s] != batch[d]) atomicAdd(&stats[3], 1);
    }
  }
  if (t < N) {
    // graph boundaries from the non-decreasing batch vector (empty graphs get empty ranges).  Ids outside
    // [0, B) or a decreasing step would make the hop kernels index graph_bias / a_graph out of bounds: they are
    // clamped here and counted with the bad edges (stats[3]), which GraphCSR.read_stats() exposes.
    int64_t b = batch[t];
    int64_t prev = t > 0 ? batch[t - 1] : -1;
    if (b < 0 || b >= B || b < prev) atomicAdd(&stats[3], 1);
    b = b < 0 ? 0 : (b >= B ? (B > 0 ? B - 1 : 0) : b);
    prev = prev < -1 ? -1 : (prev >= B ? B - 1 : prev);
    node_graph[t] = (int32_t)b;
    for (int64_t g = prev + 1; g <= b && g <= B; ++g) graph_ptr[g] = (int32_t)t;
    if (t == N - 1)
      for (int64_t g = b + 1; g <= B; ++g) graph_ptr[g] = (int32_t)N;
  }
  if (N == 0 && t == 0)
    for (int64_t g = 0; g <= B; ++g) graph_ptr[g] = 0;
}

// Single-CTA exclusive scan deg[N] -> rowptr[N+1]; also resets deg to 0 for reuse as fill cursor
// and records the max in-degree.  Each thread owns kScanItems consecutive entries (serial scan in
// registers) so N <= 1024*kScanItems needs a single block-wide scan step.
constexpr int kScanItems = 16;

__global__ void __launch_bounds__(1024) csr_scan_kernel(int32_t* __restrict__ deg, int64_t N,
                                                        int32_t* __restrict__ rowptr,
                                                        int32_t* __restrict__ stats) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  int32_t local_max = 0;
  __syncthreads();
  for (int64_t base = 0; base < N; base += 1024 * kScanItems) {
    const int64_t i0 = base + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int32_t tsum = 0;
    // all loads first, then the zero stores: interleaving them serialises 16 dependent global round trips
    // (the compiler must assume the store may alias the next load) -- measured 12.8 us for N = 7680.  128-bit
    // accesses: a thread's 16 consecutive entries are 64 contiguous bytes, and one warp instruction touches 32
    // different sectors whatever its width, so four wide accesses cost a quarter of sixteen narrow ones
    const bool full = i0 + kScanItems <= N;
    if (full) {
#pragma unroll
      for (int j = 0; j < kScanItems; j += 4) {
        const int4 q = __ldcg(reinterpret_cast<const int4*>(deg + i0 + j));
        v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < kScanItems; j += 4) *reinterpret_cast<int4*>(deg + i0 + j) = make_int4(0, 0, 0, 0);
    } else {
#pragma unroll
      for (int j = 0; j < kScanItems; ++j) v[j] = (i0 + j < N) ? __ldcg(deg + i0 + j) : 0;
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
        if (i0 + j < N) deg[i0 + j] = 0;
    }
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      local_max = max(local_max, v[j]);
      tsum += v[j];
    }
    int32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(kFull, w, o);
        if (lane >= o) w += u;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t warp_off = wid > 0 ? warp_tot[wid - 1] : 0;
    int32_t run = carry + warp_off + incl - tsum;
    if (full) {
#pragma unroll
      for (int j = 0; j < kScanItems; j += 4) {
        int4 q;
        q.x = run; run += v[j];
        q.y = run; run += v[j + 1];
        q.z = run; run += v[j + 2];
        q.w = run; run += v[j + 3];
        *reinterpret_cast<int4*>(rowptr + i0 + j) = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kScanItems; ++j) {
        if (i0 + j < N) rowptr[i0 + j] = run;
        run += v[j];
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[N] = carry_s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(kFull, local_max, o));
  if (lane == 0) atomicMax(&stats[2], local_max);
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ slots) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= E) return;
  const int64_t s = edge_index[t], d = edge_index[E + t];
  if (s < 0 || s >= N || d < 0 || d >= N) return;
  const int32_t pos = rowptr[d] + atomicAdd(&cursor[d], 1);
  slots[pos] = (int32_t)t;
}

// One warp per destination node: rank-sort the (distinct) edge ids of the row so the row keeps
// the caller's edge order, then emit perm / col_src.  Also per-graph maxima for the stats.
__global__ void csr_sort_kernel(const int64_t* __restrict__ edge_index, int64_t N,
                                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ slots,
                                int32_t* __restrict__ perm, int32_t* __restrict__ col_src,
                                const int32_t* __restrict__ graph_ptr, int64_t B, int32_t* __restrict__ stats) {
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t node = gtid >> 5;
  const int lane = threadIdx.x & 31;
  if (gtid < ((B + 31) & ~(int64_t)31)) {       // (whole warps) per-graph maxima for the stats: thread = graph
    int32_t nn = 0, ne = 0;
    if (gtid < B) {
      const int32_t n0 = graph_ptr[gtid], n1 = graph_ptr[gtid + 1];
      nn = n1 - n0;
      ne = rowptr[n1] - rowptr[n0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nn = max(nn, __shfl_xor_sync(kFull, nn, o));
      ne = max(ne, __shfl_xor_sync(kFull, ne, o));
    }
    if (lane == 0) {
      atomicMax(&stats[0], nn);
      atomicMax(&stats[1], ne);
    }
  }
  if (node >= N) return;
  const int32_t e0 = rowptr[node], e1 = rowptr[node + 1];
  for (int32_t t = e0 + lane; t < e1; t += 32) {
    const int32_t v = slots[t];
    int32_t rank = 0;
    for (int32_t u = e0; u < e1; ++u) rank += slots[u] < v;
    perm[e0 + rank] = v;
    col_src[e0 + rank] = (int32_t)edge_index[v];
  }
}

__global__ void csr_graph_stats_kernel(const int32_t* __restrict__ graph_ptr, const int32_t* __restrict__ rowptr,
                                       int64_t B, int32_t* __restrict__ stats) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int32_t nn = 0, ne = 0;
  if (g < B) {
    const int32_t n0 = graph_ptr[g], n1 = graph_ptr[g + 1];
    nn = n1 - n0;
    ne = rowptr[n1] - rowptr[n0];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nn = max(nn, __shfl_xor_sync(kFull, nn, o));
    ne = max(ne, __shfl_xor_sync(kFull, ne, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&stats[0], nn);
    atomicMax(&stats[1], ne);
  }
}

}  // namespace gvqa

extern "C" GVQA_API size_t gvqa_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
  if (num_nodes < 0 || num_edges < 0) return 0;
  // cursor[N] + slots[E], padded to 16 bytes each
  const size_t a = ((size_t)num_nodes * 4 + 15) & ~(size_t)15;
  const size_t b = ((size_t)num_edges * 4 + 15) & ~(size_t)15;
  return a + b + 16;
}

extern "C" GVQA_API int gvqa_build_csr(const int64_t* edge_index, int64_t E, const int64_t* batch, int64_t N,
                              int64_t B, int32_t* rowptr, int32_t* col_src, int32_t* perm,
                              int32_t* graph_ptr, int32_t* node_graph, int32_t* stats, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  using namespace gvqa;
  if (N < 0 || E < 0 || B < 0 || N >= (1ll << 31) || E >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (!rowptr || !graph_ptr || !stats || !workspace) return GVQA_ERR_NULL_POINTER;
  if (N > 0 && (!batch || !node_graph)) return GVQA_ERR_NULL_POINTER;
  if (E > 0 && (!edge_index || !col_src || !perm)) return GVQA_ERR_NULL_POINTER;
  if (workspace_bytes < gvqa_csr_workspace_bytes(N, E)) return GVQA_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int32_t* cursor = static_cast<int32_t*>(workspace);
  int32_t* slots = cursor + (((size_t)N * 4 + 15) & ~(size_t)15) / 4;

  if (cudaMemsetAsync(cursor, 0, (size_t)N * 4, stream) != cudaSuccess) return GVQA_ERR_CUDA;
  if (cudaMemsetAsync(stats, 0, 8 * sizeof(int32_t), stream) != cudaSuccess) return GVQA_ERR_CUDA;
  const int64_t work = (N > E ? N : E) > 0 ? (N > E ? N : E) : 1;
  const int threads = 256;
  csr_count_kernel<<<(unsigned)((work + threads - 1) / threads), threads, 0, stream>>>(
      edge_index, E, batch, N, B, cursor, graph_ptr, node_graph, stats);
  GVQA_LAUNCH_CHECK();
  csr_scan_kernel<<<1, 1024, 0, stream>>>(cursor, N, rowptr, stats);
  GVQA_LAUNCH_CHECK();
  if (E > 0) {
    csr_fill_kernel<<<(unsigned)((E + threads - 1) / threads), threads, 0, stream>>>(edge_index, E, N, rowptr,
                                                                                      cursor, slots);
    GVQA_LAUNCH_CHECK();
    const int64_t sort_threads = (N * 32 > ((B + 31) & ~(int64_t)31)) ? N * 32 : ((B + 31) & ~(int64_t)31);
    csr_sort_kernel<<<(unsigned)((sort_threads + threads - 1) / threads), threads, 0, stream>>>(
        edge_index, N, rowptr, slots, perm, col_src, graph_ptr, B, stats);
    GVQA_LAUNCH_CHECK();
  } else if (B > 0) {
    csr_graph_stats_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, stream>>>(graph_ptr, rowptr, B,
                                                                                            stats);
    GVQA_LAUNCH_CHECK();
  }
  return GVQA_OK;
}
