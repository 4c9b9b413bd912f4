// Host-side destination-CSR build (see include/gvqa_b200.h: gvqa_build_csr_host).
//
// The loader-side twin of gvqa_build_csr: a data-loader worker builds the int32 destination-CSR of a collated
// batch straight into (pinned) host buffers, so the device receives the topology in its final form -- no int64 COO
// on the wire, no gvqa_build_csr launches on the GPU (the reference collates with
// torch_geometric's Batch.from_data_list, gqa_dataset_entry.py:631-675, and ships int64 COO).
// Stable counting sort by target: the slots of one node keep the caller's edge order, exactly like the device
// build, so both produce identical arrays (tests/test_host_csr.py).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/gvqa_b200.h"

namespace {

template <typename T>
inline int64_t at(const void* p, int64_t i) {
  return (int64_t) static_cast<const T*>(p)[i];
}

template <typename EI, typename BT>
int build(const void* ei, int64_t E, const void* batch, int64_t N, int64_t B, int32_t* rowptr, int32_t* col_src,
          int32_t* perm, int32_t* graph_ptr, int32_t* node_graph, int32_t* stats) {
  int32_t bad = 0, max_deg = 0, max_nodes = 0, max_edges = 0;
  // graph boundaries from the non-decreasing batch vector (same clamping rules as csr_count_kernel)
  int64_t prev = -1;
  for (int64_t g = 0; g <= B; ++g) graph_ptr[g] = (int32_t)N;
  for (int64_t t = 0; t < N; ++t) {
    int64_t b = at<BT>(batch, t);
    if (b < 0 || b >= B || b < prev) ++bad;
    b = b < 0 ? 0 : (b >= B ? (B > 0 ? B - 1 : 0) : b);
    node_graph[t] = (int32_t)b;
    for (int64_t g = prev + 1; g <= b && g <= B; ++g) graph_ptr[g] = (int32_t)t;
    if (b > prev) prev = b;
  }
  if (N == 0)
    for (int64_t g = 0; g <= B; ++g) graph_ptr[g] = 0;
  // in-degrees -> exclusive scan
  memset(rowptr, 0, sizeof(int32_t) * (size_t)(N + 1));
  for (int64_t k = 0; k < E; ++k) {
    const int64_t s = at<EI>(ei, k), d = at<EI>(ei, E + k);
    if (s < 0 || s >= N || d < 0 || d >= N) {
      ++bad;
      continue;
    }
    if (node_graph[s] != node_graph[d]) ++bad;
    ++rowptr[d + 1];
  }
  for (int64_t i = 0; i < N; ++i) {
    if (rowptr[i + 1] > max_deg) max_deg = rowptr[i + 1];
    rowptr[i + 1] += rowptr[i];
  }
  // stable fill in the caller's edge order
  std::vector<int32_t> cursor(rowptr, rowptr + N);
  for (int64_t k = 0; k < E; ++k) {
    const int64_t s = at<EI>(ei, k), d = at<EI>(ei, E + k);
    if (s < 0 || s >= N || d < 0 || d >= N) continue;
    const int32_t pos = cursor[d]++;
    perm[pos] = (int32_t)k;
    col_src[pos] = (int32_t)s;
  }
  for (int64_t g = 0; g < B; ++g) {
    const int32_t n0 = graph_ptr[g], n1 = graph_ptr[g + 1];
    if (n1 - n0 > max_nodes) max_nodes = n1 - n0;
    if (rowptr[n1] - rowptr[n0] > max_edges) max_edges = rowptr[n1] - rowptr[n0];
  }
  if (stats) {
    memset(stats, 0, 8 * sizeof(int32_t));
    stats[0] = max_nodes; stats[1] = max_edges; stats[2] = max_deg; stats[3] = bad;
  }
  return GVQA_OK;
}

}  // namespace

extern "C" GVQA_API int gvqa_build_csr_host(const void* edge_index_host, int32_t index_bytes, int64_t E,
                                            const void* batch_host, int32_t batch_bytes, int64_t N, int64_t B,
                                            int32_t* rowptr_host, int32_t* col_src_host, int32_t* perm_host,
                                            int32_t* graph_ptr_host, int32_t* node_graph_host, int32_t* stats_host) {
  if (N < 0 || E < 0 || B < 0 || N >= (1ll << 31) || E >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (!rowptr_host || !graph_ptr_host) return GVQA_ERR_NULL_POINTER;
  if (N > 0 && (!batch_host || !node_graph_host)) return GVQA_ERR_NULL_POINTER;
  if (E > 0 && (!edge_index_host || !col_src_host || !perm_host)) return GVQA_ERR_NULL_POINTER;
  if ((index_bytes != 4 && index_bytes != 8) || (batch_bytes != 4 && batch_bytes != 8)) return GVQA_ERR_UNSUPPORTED;
#define GVQA_HOST_CSR(EI, BT)                                                                                       \
  return build<EI, BT>(edge_index_host, E, batch_host, N, B, rowptr_host, col_src_host, perm_host, graph_ptr_host, \
                       node_graph_host, stats_host)
  if (index_bytes == 8 && batch_bytes == 8) GVQA_HOST_CSR(int64_t, int64_t);
  if (index_bytes == 8 && batch_bytes == 4) GVQA_HOST_CSR(int64_t, int32_t);
  if (index_bytes == 4 && batch_bytes == 8) GVQA_HOST_CSR(int32_t, int64_t);
  GVQA_HOST_CSR(int32_t, int32_t);
#undef GVQA_HOST_CSR
}
