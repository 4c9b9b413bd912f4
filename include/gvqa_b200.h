/*
 * gvqa_b200.h -- C ABI of the B200-native GraphVQA scene-graph message-passing engine.
 *
 * The reference (codexxxl/GraphVQA) is pure Python and has no FFI of its own; each entry point
 * below therefore cites the reference *Python interface* whose device work it replaces
 * (paths relative to the reference root).  The Python host mirror in graphvqa_b200/ binds these
 * with ctypes (see INTEGRATION.md for the stub a maintainer would add to the reference).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer (cudaMalloc'ed or a torch tensor's data_ptr) unless the
 *    name ends in _host.  The caller owns every buffer; nothing here allocates, frees,
 *    synchronises or touches a stream other than the one passed in.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *    enqueued asynchronously and is CUDA-graph capturable.
 *  - Matrices are row-major fp32.  Graph topology is int32 destination-CSR produced by
 *    gvqa_build_csr from the reference's int64 COO `edge_index` ([0]=source j, [1]=target i).
 *  - Return value: GVQA_OK (0) or a negative gvqa_status; gvqa_error_string() describes it.
 *    Argument errors are detected on the host before anything is enqueued.
 *  - Threading: the entry points keep no per-call state; concurrent calls from several host threads on
 *    different streams are safe.  Process-wide state exists in three places, all write-once or debug-only:
 *    (1) one-time cudaFuncSetAttribute / occupancy queries cached in function-local statics (thread-safe
 *    initialisation), (2) environment switches for experiments (GVQA_PDL, GVQA_HOP_NPC, ...) latched at first
 *    use, (3) the debug hooks declared in gvqa_b200_debug.h (trace buffers, stage-skipping flags), which are
 *    plain globals: set them only while no other thread is launching.
 */
#ifndef GVQA_B200_H_
#define GVQA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVQA_ABI_VERSION 5

#if defined(__GNUC__)
#define GVQA_API __attribute__((visibility("default")))
#else
#define GVQA_API
#endif

typedef enum gvqa_status {
  GVQA_OK = 0,
  GVQA_ERR_NULL_POINTER = -1,   /* a required pointer is NULL */
  GVQA_ERR_BAD_SHAPE = -2,      /* negative / inconsistent sizes */
  GVQA_ERR_UNSUPPORTED = -3,    /* channel count not a multiple of 4, > 1024, heads > 8, ... */
  GVQA_ERR_MISALIGNED = -4,     /* pointer or leading dimension not 16-byte aligned */
  GVQA_ERR_WORKSPACE = -5,      /* workspace too small */
  GVQA_ERR_CUDA = -6            /* a CUDA runtime call failed (launch error) */
} gvqa_status;

/* Epilogue applied by the hop kernels after  out = conv + skip. */
typedef enum gvqa_epilogue {
  GVQA_EPI_NONE = 0,         /* last hop of gat_seq (gat_skip.py:273: no BN after the final conv) */
  GVQA_EPI_AFFINE = 1,       /* y = out*scale[c] + shift[c]            (BatchNorm1d in eval mode)  */
  GVQA_EPI_AFFINE_RELU = 2,  /* y = relu(out*scale[c] + shift[c])      (gat_skip.py:274-275)       */
  GVQA_EPI_GRAPH_LN = 3      /* y = my_graph_layernorm.LayerNorm(out, batch): per-graph statistics over all
                                nodes x channels, two-pass variance, eps added to the std, scalar affine
                                (graph_utils/my_graph_layernorm.py:52-78); one CTA per graph keeps the rows in
                                shared memory, so the normalisation costs no extra HBM pass                 */
} gvqa_epilogue;

GVQA_API int gvqa_abi_version(void);
GVQA_API const char* gvqa_error_string(int status);

/* L2 residency for producer -> consumer tensors (x_l between the projection GEMM and the hop kernel).
 * gvqa_device_set_l2_persist_limit: set aside up to `bytes` of L2 for persisting lines; returns the
 * MiB granted (>= 0) or a negative status.  gvqa_stream_set_l2_window: access-policy window of the
 * stream (ptr == NULL clears it); captured into CUDA-graph kernel nodes. */
GVQA_API int gvqa_device_set_l2_persist_limit(size_t bytes);
GVQA_API int gvqa_stream_set_l2_window(const void* ptr, size_t bytes, float hit_ratio, void* stream);

/* ------------------------------------------------------------------------------------------
 * Destination-CSR build.  Replaces what torch_geometric's MessagePassing.propagate derives
 * from `edge_index` on every call (gat_skip.py:155) and the
 * `int(batch.max())` / `batch[-1].item()` host syncs (my_graph_layernorm.py:59,
 * pipeline_model_gat.py:152).
 *
 *   edge_index [2,E] int64 (row 0 = source, row 1 = target), batch [N] int64 non-decreasing.
 * Outputs (int32):
 *   rowptr[N+1]      in-edge range of node i is [rowptr[i], rowptr[i+1])
 *   col_src[E]       source node of CSR slot k
 *   perm[E]          original edge id of CSR slot k; slots of one node keep the original
 *                    edge order (stable), so sums run in the order of a sequential index_add
 *   graph_ptr[B+1]   node range of graph g is [graph_ptr[g], graph_ptr[g+1])
 *   node_graph[N]    int32 copy of batch
 *   stats[8]         {max nodes/graph, max in-edges/graph, max in-degree, #edges whose
 *                    endpoints lie in different graphs or out of range, 0...}
 * workspace: gvqa_csr_workspace_bytes(N, E) bytes of scratch.
 */
GVQA_API size_t gvqa_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
GVQA_API int gvqa_build_csr(const int64_t* edge_index, int64_t num_edges, const int64_t* batch,
                   int64_t num_nodes, int64_t num_graphs, int32_t* rowptr, int32_t* col_src,
                   int32_t* perm, int32_t* graph_ptr, int32_t* node_graph, int32_t* stats,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The same build on the HOST, for the data-loader side (the reference collates with
 * torch_geometric's Batch.from_data_list and ships int64 COO, gqa_dataset_entry.py:631-675): every pointer is a
 * HOST pointer (ideally pinned), `index_bytes` / `batch_bytes` give the width of the inputs (4 = int32, 8 = int64).
 * Outputs equal gvqa_build_csr's bit for bit (stable counting sort, same clamping, same stats); the device then
 * receives rowptr / col_src / perm / graph_ptr / node_graph in their final int32 form and runs no CSR kernel.
 * No workspace argument: scratch of N int32 is allocated internally (host code). */
GVQA_API int gvqa_build_csr_host(const void* edge_index_host, int32_t index_bytes, int64_t num_edges,
                                 const void* batch_host, int32_t batch_bytes, int64_t num_nodes, int64_t num_graphs,
                                 int32_t* rowptr_host, int32_t* col_src_host, int32_t* perm_host,
                                 int32_t* graph_ptr_host, int32_t* node_graph_host, int32_t* stats_host);

/* ------------------------------------------------------------------------------------------
 * Skinny projection  out[M,K] = x[M,F] @ v[K,F]^T,  K <= 32, F % 4 == 0  (v laid out like an
 * nn.Linear weight, [out, in]).  Computes the attention-logit terms of gat.forward
 * (gat_skip.py:134-135 a_l/a_r, :150-151 a_e) after the algebraic collapse
 * <W x, att_h>  =  x . (W_h^T att_h)  (SURVEY.md section 8a).  x rows have stride ldx floats.
 */
GVQA_API int gvqa_skinny_matmul_f32(const float* x, int64_t ldx, const float* v, float* out, int64_t m,
                           int f, int k, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused GAT hop.  One launch replaces, for one hop of gat_seq.forward (gat_skip.py:254-276):
 * index_select x3, add, leaky_relu, torch_geometric.utils.softmax (scatter_max, exp,
 * scatter_add, div), mul, scatter-add aggregate, head mean, +bias, +skip, BatchNorm1d(eval),
 * ReLU  (gat_skip.py:155-168, 183-208, 270-275).
 *
 *   xl[n,h,:]      = x_l[n*ldx + h*C + :]
 *   logit[k,h]     = a_node[src_k, h] + a_node[i, H+h] + (a_graph ? a_graph[g(i), h] : 0)
 *                    + a_edge[e_k*lde + h],   e_k = perm ? perm[k] : k        (k in CSR order)
 *   alpha          = softmax over the in-edges of i of leaky_relu(logit, negative_slope),
 *                    with PyG's  exp(l - max) / (sum + 1e-16)
 *   out[i,:]       = (1/H) sum_h sum_k alpha[k,h] * xl[src_k, h, :]
 *                    + (deg(i) > 0 ? graph_bias[g(i), :] : 0) + bias + h_prev[i,:]
 *   h_out[i,:]     = epilogue(out[i,:])
 * graph_bias carries the instruction half of the reference's concatenated input: every source
 * row of graph g contains the same term P[g,h,:] = ins[g] @ W_l[h, :, F:]^T, and the softmax
 * weights of a destination sum to 1, so  sum_k alpha[k,h] P[g,h,:] = P[g,h,:]  and the head mean
 * of P is added once per destination that has in-edges (graph_bias = mean_h P).
 * alpha_out (optional) receives alpha[E,H] in ORIGINAL edge order (return_attention_weights,
 * gat_skip.py:170-175).  Nodes without in-edges get out = bias + h_prev (empty softmax).
 */
typedef struct gvqa_gat_hop_args {
  const float* x_l;        /* [N, ldx] projected node features, head-major columns h*C+c     */
  int64_t ldx;             /* row stride of x_l in floats (>= H*C, multiple of 4)            */
  const float* graph_bias; /* [B, C] per-graph additive output term (see above), or NULL     */
  const float* a_node;     /* [N, ld_a_node]: columns 0..H-1 = source term a_l, H..2H-1 = target a_r */
  int64_t ld_a_node;       /* row stride of a_node in floats; 0 = dense (2H).  The logit columns may
                              live inside the x_l buffer (extra GEMM output columns)               */
  const float* a_graph;    /* [B, H] per-graph additive logit term, or NULL                  */
  const float* a_edge;     /* [E, lde] per-edge logit term                                    */
  int64_t lde;             /* row stride of a_edge in floats (>= H)                           */
  const int32_t* rowptr;   /* [N+1] */
  const int32_t* col_src;  /* [E]   */
  const int32_t* perm;     /* [E] or NULL when a_edge is already in CSR order                 */
  const int32_t* graph_ptr;  /* [B+1] */
  const int32_t* node_graph; /* [N]   */
  const float* h_prev;     /* [N, C] skip input, or NULL                                      */
  const float* bias;       /* [C] or NULL                                                     */
  const float* ep_scale;   /* [C] (epilogue != NONE)                                          */
  const float* ep_shift;   /* [C] (epilogue != NONE)                                          */
  float* h_out;            /* [N, C]                                                          */
  float* alpha_out;        /* [E, H] original edge order, or NULL                             */
  int64_t num_nodes, num_edges, num_graphs;
  int32_t heads, channels; /* H <= 8; C multiple of 4, <= 1024                                */
  float negative_slope;
  int32_t epilogue;        /* gvqa_epilogue                                                   */
  int32_t max_nodes_per_graph;    /* loader hints (0 = unknown) that size the shared-memory staged  */
  int32_t max_in_edges_per_graph; /* kernel; never a correctness input (oversize graphs fall back)   */
  int32_t variant;         /* 0 = auto (5 when slabs are given, else 3), 1 = warp-per-node gather, 2 = TMA smem-staged
                              rows, 3 = block-phase gather, 4 = persistent warp-specialised (producer warp prepares
                              chunks, needs `sched`), 5 = block-phase gather with the one-round-trip slab prologue
                              (needs slab_idx / slab_f from gvqa_gat_hop_build_slabs_f32)                        */
  int64_t ld_graph_bias;   /* row stride of graph_bias in floats; 0 = dense (C); multiple of 4            */
  int64_t ld_a_graph;      /* row stride of a_graph in floats; 0 = dense (H).  Both terms may be column
                              blocks of one pre-pass GEMM output                                          */
  int32_t flags;           /* GVQA_HOP_* bits                                                             */
  float ln_eps;            /* GVQA_EPI_GRAPH_LN: eps (added to the standard deviation)                            */
  const float* ln_weight;  /* GVQA_EPI_GRAPH_LN: ONE float each on the device (the reference's parameters have    */
  const float* ln_bias;    /*   shape [1]), or both NULL for no affine                                            */
  const int32_t* slab_idx; /* variant 5: the per-batch index slabs and THIS hop's logit-term slabs (slab_f + hop *  */
  const float* slab_f;     /*   f_words_per_hop), see gvqa_gat_hop_build_slabs_f32; NULL otherwise.  max_nodes_per_graph
                              sizes the a_node window the kernel stages (sources farther away are loaded from global
                              memory: a hint, never a correctness input)                                          */
  int32_t* sched;          /* variant 4: two int32 of device scratch, zero before the first launch; the kernel
                              leaves them zero again (dynamic chunk scheduler); may be NULL otherwise.  Two
                              launches that may run concurrently need separate words                              */
} gvqa_gat_hop_args;

/* The block kernel is launched with programmatic stream serialization: its CTAs may start while the previous
 * kernel of the stream drains, and block (griddepcontrol.wait) before touching anything that kernel wrote.
 * Set this bit when rowptr / col_src / perm / node_graph / a_edge / a_graph were NOT written by the kernel
 * launched immediately before this one (in gat_seq they come from the per-batch pre-pass, the predecessor is the
 * projection GEMM): the index round trips then overlap the predecessor's tail.  x_l, a_node, h_prev and all
 * outputs are always accessed after the wait.  Leaving the bit clear is always safe. */
#define GVQA_HOP_INPUTS_OLDER_THAN_PREDECESSOR 1

GVQA_API int gvqa_gat_hop_f32(const gvqa_gat_hop_args* args, void* stream);

/* Per-batch slabs for hop variant 5.  Everything the hop kernel's index prologue gathers except a_node is
 * hop-invariant per batch (topology, per-edge logit terms of all hops, per-graph logit terms), and the CTA partition
 * depends on the node count only; this pass writes it once per batch in a per-CTA layout whose address depends on the
 * block index alone, so each hop CTA fetches it with two TMA bulk copies (cp.async.bulk + mbarrier) in ONE round trip
 * instead of three dependent ones (row pointers -> sources / edge ids -> logit terms).
 *   a_edge [E, lde]: hop j's per-edge term at columns [j*H, (j+1)*H), original edge order (gathered through perm);
 *   a_graph: hop j's per-graph term at a_graph + j*hop_stride_a_graph + g*ld_a_graph + h, or NULL.
 * Buffers: slab_idx int32[plan.idx_words], slab_f float[hops * plan.f_words_per_hop], both 16-byte aligned. */
typedef struct gvqa_gat_slab_plan {
  int32_t nodes_per_cta, edge_capacity, num_ctas;
  int64_t idx_words, f_words_per_hop;
} gvqa_gat_slab_plan;
GVQA_API int gvqa_gat_hop_slab_plan(int64_t num_nodes, int64_t num_edges, int32_t heads, gvqa_gat_slab_plan* plan);
GVQA_API int gvqa_gat_hop_build_slabs_f32(const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                          const int32_t* node_graph, const float* a_edge, int64_t lde,
                                          const float* a_graph, int64_t ld_a_graph, int64_t hop_stride_a_graph,
                                          int32_t hops, int64_t num_nodes, int64_t num_edges, int32_t heads,
                                          int32_t* slab_idx, float* slab_f, void* stream);

/* ------------------------------------------------------------------------------------------
 * One GAT hop as ONE tensor-core kernel ("aggregate, then project"; gat_skip.py:133-168 + :270-275).
 * The projection lin_l is linear, so  sum_k alpha[k,h] (W_h h[src_k]) = W_h (sum_k alpha[k,h] h[src_k]):
 * the kernel builds z[i,h,:] = sum_k alpha[k,h] h[src_k,:] tile by tile from a TMA-staged window of h rows
 * (never stored), runs  Z[N, H*F] @ W'[C, H*F]^T  on tcgen05 with the fp16 split of gvqa_proj_gemm_3xf16, and
 * applies the hop epilogue (head mean, graph_bias for rows with in-edges, bias, skip, affine / ReLU) to the
 * accumulators.  x_l[N, H*C] is never materialised.  Pieces:
 *   gvqa_gat_fused_pack_f16   once per checkpoint: scale * lin_l.weight[:, :F] ([H*C, ldw] fp32, row h*C + c) -> packed
 *                             fp16 [C, Fp/16, H, (16 hi | 16 lo)] (Fp = F rounded up to 16; lo = x - hi, unscaled);
 *                             `scale`: a power of two that lifts the low parts into fp16's normal range (2^9 / max|W|
 *                             rounded down to a power of two is what gat_seq uses), passed again as w_scale to the
 *                             hop, which divides it out; gvqa_gat_fused_pack_halves gives the element count;
 *   gvqa_gat_fused_plan       once per batch: row tiles {first row, rows, first window row, 0} (int32 x 4 each) from
 *                             graph_ptr: whole graphs packed greedily into <= 128 rows; larger graphs are cut into
 *                             128-row chunks.  `window` = gvqa_gat_fused_window(max nodes per graph) (128 or 256 rows
 *                             of h staged per tile; sources outside it are read from global memory: never a
 *                             correctness input).  tiles: 16-byte aligned, gvqa_gat_fused_max_tiles entries; count: one
 *                             int32 on the device.  _host: the same plan built by the loader (wire format);
 *   gvqa_gat_alpha_f32        per hop: softmax weights of all in-edges in CSR order, alpha[k, h] (PyG semantics,
 *                             gat_skip.py:183-192); a_node [N, >= 2H] = (a_l | a_r) -- or a_node_parts partial sums
 *                             a_node_part_stride floats apart (the fused hop's a_part), added in order --, a_edge
 *                             gathered through perm, a_graph per graph or NULL; alpha_out (optional): the same in
 *                             original edge order;
 *   gvqa_gat_fused_hop_f32    per hop.  heads in {2, 4}, in_channels and channels multiples of 4
 *                             (gvqa_gat_fused_supported), epilogue NONE / AFFINE / AFFINE_RELU.  `overflow` as in
 *                             gvqa_proj_gemm_3xf16 (OR-ed with 1 when an aggregated input does not fit fp16). */
typedef struct gvqa_gat_fused_args {
  const float* h_in;        /* [N, in_channels] node states, row stride ld_h (0 = dense)                 */
  int64_t ld_h;
  const void* w_pack;       /* gvqa_gat_fused_pack_f16 output                                             */
  const int32_t* tiles;     /* gvqa_gat_fused_plan                                                        */
  const int32_t* tile_count;
  const int32_t* rowptr;    /* destination-CSR (gvqa_build_csr)                                           */
  const int32_t* col_src;
  const int32_t* node_graph;
  float* alpha;             /* [E, heads]: softmax weights from gvqa_gat_alpha_f32 (read only) -- or, with
                               logit_terms, scratch for tiles that do not fit the kernel's staging          */
  const float* logit_terms; /* optional: THIS hop's [E, heads] block of gvqa_gat_fused_logit_terms_f32.  The kernel
                               then computes the softmax weights itself in each tile's prologue (no
                               gvqa_gat_alpha_f32 launch) from these and a_node                            */
  const float* a_node;      /* with logit_terms: [a_node_parts][N][2*heads] node logits a_l | a_r as partial
                               sums a_node_part_stride floats apart (the previous hop's a_part, or one block) */
  int64_t a_node_part_stride;
  int32_t a_node_parts;
  float negative_slope;
  int64_t ld_a_node;        /* row stride of a_node in floats (0 = dense 2*heads; a multiple of 4): the logits may be
                               the leading columns of a wider GEMM output                                   */
  int32_t flags;            /* GVQA_HOP_INPUTS_OLDER_THAN_PREDECESSOR (with logit_terms): tiles, topology, logit_terms,
                               graph_bias, bias, ep_*, v_next and w_pack were NOT written by the kernel launched
                               right before this one (in gat_seq that is the previous hop, which writes h_in, a_node):
                               the kernel reads them while its predecessor drains                          */
  const float* skip;        /* [N, channels] added to every row (gat_skip.py:270), row stride ld_skip; or NULL */
  int64_t ld_skip;
  const float* graph_bias;  /* [B, channels] per-graph instruction term (rows with in-edges only) or NULL */
  int64_t ld_graph_bias;
  const float* bias;        /* [channels] or NULL                                                         */
  const float* ep_scale;    /* [channels] for GVQA_EPI_AFFINE(_RELU)                                      */
  const float* ep_shift;
  float* h_out;             /* [N, channels] dense                                                        */
  int32_t* overflow;        /* device int32 or NULL                                                       */
  int64_t num_nodes;
  int32_t in_channels, channels, heads, epilogue, window;
  float w_scale;            /* the scale given to gvqa_gat_fused_pack_f16                                 */
  const float* v_next;      /* optional: [2*heads, channels] collapsed logit vectors (V_l ; V_r) of the NEXT hop:
                               the epilogue also emits that hop's node logits a_l | a_r = h_out . V
                               (gat_skip.py:134-135) as partial sums over 128-column blocks,
                               a_part[block][node][2*heads]; gvqa_gat_alpha_f32 adds the blocks in fixed order */
  float* a_part;
  int32_t a_part_blocks;    /* gvqa_gat_fused_part_blocks(channels) when v_next is given                  */
} gvqa_gat_fused_args;
GVQA_API int32_t gvqa_gat_fused_part_blocks(int64_t num_nodes, int32_t channels);
GVQA_API int gvqa_gat_fused_supported(int32_t heads, int32_t in_channels, int32_t channels);
GVQA_API int64_t gvqa_gat_fused_pack_halves(int32_t heads, int32_t channels, int32_t in_channels);
GVQA_API int gvqa_gat_fused_pack_f16(const float* w, int64_t ldw, int32_t heads, int32_t channels,
                                     int32_t in_channels, float scale, void* packed, void* stream);
GVQA_API int64_t gvqa_gat_fused_max_tiles(int64_t num_nodes, int64_t num_graphs);
GVQA_API int32_t gvqa_gat_fused_window(int32_t max_nodes_per_graph);
GVQA_API int gvqa_gat_fused_plan(const int32_t* graph_ptr, int64_t num_graphs, int32_t window, int32_t* tiles,
                                 int32_t* count, int64_t max_tiles, void* stream);
/* the same plan straight from the reference's non-decreasing int64 `batch` vector (device): does not wait for the CSR */
GVQA_API int gvqa_gat_fused_plan_from_batch(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t window,
                                            int32_t* tiles, int32_t* count, int64_t max_tiles, void* stream);
GVQA_API int gvqa_gat_fused_plan_host(const int32_t* graph_ptr_host, int64_t num_graphs, int32_t window,
                                      int32_t* tiles_host, int32_t* count_host, int64_t max_tiles);
GVQA_API int gvqa_gat_alpha_f32(const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                const int32_t* node_graph, const float* a_node, int64_t ld_a_node,
                                int32_t a_node_parts, int64_t a_node_part_stride, const float* a_edge, int64_t lde, const float* a_graph, int64_t ld_a_graph,
                                float negative_slope, int64_t num_nodes, int32_t heads, float* alpha,
                                float* alpha_out, void* stream);
/* Hop-invariant logit terms of all hops in CSR order, once per batch:
 *   terms[hop][k][h] = a_edge[perm[k]][hop*heads + h] + a_graph[hop][graph of the destination][h]
 * a_edge [E, lde], a_graph at a_graph + hop*hop_stride_a_graph + g*ld_a_graph + h (or NULL); hop j's block starts at
 * terms + j*terms_hop_stride (floats; >= E*heads, a multiple of 4 so that every block is 16-byte aligned). */
GVQA_API int gvqa_gat_fused_logit_terms_f32(const int32_t* rowptr, const int32_t* perm, const int32_t* node_graph,
                                            const float* a_edge, int64_t lde, const float* a_graph, int64_t ld_a_graph,
                                            int64_t hop_stride_a_graph, int32_t hops, int64_t num_nodes,
                                            int64_t num_edges, int32_t heads, float* terms,
                                            int64_t terms_hop_stride, void* stream);
GVQA_API int gvqa_gat_fused_hop_f32(const gvqa_gat_fused_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-graph LayerNorm (graph_utils/my_graph_layernorm.py:52-78): statistics over all
 * nodes x channels of each graph, two-pass variance, eps added to the std, scalar affine.
 * weight/bias point to ONE float each on the device (the reference's parameters have shape
 * [1], my_graph_layernorm.py:40-41) or are NULL.  In-place (out == x) is allowed.
 * max_nodes_per_graph is a loader hint (0 = unknown) that sizes the shared-memory staging.
 */
GVQA_API int gvqa_graph_layernorm_f32(const float* x, const int32_t* graph_ptr, const float* weight,
                             const float* bias, float* out, int64_t num_nodes,
                             int64_t num_graphs, int32_t channels, float eps,
                             int32_t max_nodes_per_graph, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-accurate projection GEMM on the tcgen05 tensor cores ("3xTF32"):
 *   C[M,N] = A[M,K] @ B[N,K]^T      (x_l = lin_l(x_cat), gat_skip.py:133; a cuBLAS SGEMM there)
 * Every operand is split x = hi + lo (hi = tf32(x), lo = tf32(x - hi)); A_lo*B_hi + A_hi*B_lo +
 * A_hi*B_hi is accumulated in fp32 in tensor memory.  B must be pre-split with gvqa_split_tf32
 * (weights: once per checkpoint); A is split on the fly.  k, lda, ldb, ldc multiples of 4.
 */
GVQA_API int gvqa_split_tf32(const float* w, float* hi, float* lo, int64_t count, void* stream);
GVQA_API int gvqa_proj_gemm_3xtf32(const float* a, int64_t lda, const float* b_hi, const float* b_lo,
                                   int64_t ldb, float* c, int64_t ldc, int64_t m, int32_t n, int32_t k,
                                   void* stream);

/* Same projection with fp16 tensor-core operands: x = hi + 2^-11 * lo', hi = fp16(x), lo' = fp16((x - hi) * 2^11);
 * C = A_hi*B_hi + 2^-11 (A_hi*B_lo' + A_lo'*B_hi), fp32 accumulation -- the accuracy of the 3xTF32 kernel at twice
 * the tensor-core rate and 2/3 of its shared-memory traffic.  Requires |A| < 65504 (fp16 range): when `overflow`
 * (device int32, may be NULL) is given, it is OR-ed with 1 if an element of A is outside the range or not finite;
 * callers then redo the product with gvqa_proj_gemm_3xtf32.
 * gvqa_split_f16 prepacks the weights: w [rows, cols] fp32 (row stride ld_in) -> hi, lo' [rows, ld_out] fp16,
 * ld_out a multiple of 8 (zero padded).  k and lda multiples of 4, ldb a multiple of 8, ldc a multiple of 4. */
GVQA_API int gvqa_split_f16(const float* w, int64_t ld_in, void* hi, void* lo, int64_t ld_out, int64_t rows,
                            int64_t cols, void* stream);
GVQA_API int gvqa_proj_gemm_3xf16(const float* a, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                                  float* c, int64_t ldc, int64_t m, int32_t n, int32_t k, int32_t* overflow,
                                  void* stream);
/* `batch` independent products C[z] = A[z] @ B[z]^T in one launch (element strides between consecutive z; stride_a
 * and stride_c multiples of 4, stride_b a multiple of 8): gat_seq's per-hop instruction terms
 * [hops, B, D] x [hops, C+H, D]^T (gat_skip.py:256-264, split cat). */
/* Several independent products in ONE persistent launch (at most 3; tiles of problem 0 are scheduled first).
 * gat_seq issues hop 0's projection together with its two small pre-pass products this way. */
typedef struct gvqa_gemm_problem {
  const float* a;          /* [batch, m, k] fp32, row stride lda, batch stride stride_a (elements) */
  int64_t lda, stride_a;
  const void* b_hi;        /* [batch, n, k] fp16 from gvqa_split_f16, row stride ldb, batch stride stride_b */
  const void* b_lo;
  int64_t ldb, stride_b;
  float* c;                /* [batch, m, n] fp32, row stride ldc, batch stride stride_c */
  int64_t ldc, stride_c;
  int64_t m;
  int32_t n, k, batch;     /* batch >= 1; the strides are ignored when batch == 1 */
  int32_t relu;            /* != 0: C = max(C, 0) after the bias                                      */
  const float* bias;       /* [n] added to every row of C (nn.Linear's bias), or NULL                 */
} gvqa_gemm_problem;
GVQA_API int gvqa_proj_gemm_3xf16_grouped(const gvqa_gemm_problem* problems, int32_t count, int32_t* overflow,
                                          void* stream);
/* nn.Linear (+ ReLU) in one launch: C = act(A @ B^T + bias), bias [n] or NULL, relu 0/1 -- the dense layers of the
 * scene-graph encoder, the attention pooling and the answer head (pipeline_model_gat.py:65-98, 126-128, 722-728). */
GVQA_API int gvqa_linear_3xf16(const float* a, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                               const float* bias, int32_t relu, float* c, int64_t ldc, int64_t m, int32_t n, int32_t k,
                               int32_t* overflow, void* stream);
GVQA_API int gvqa_proj_gemm_3xf16_batched(const float* a, int64_t lda, int64_t stride_a, const void* b_hi,
                                          const void* b_lo, int64_t ldb, int64_t stride_b, float* c, int64_t ldc,
                                          int64_t stride_c, int64_t m, int32_t n, int32_t k, int32_t batch,
                                          int32_t* overflow, void* stream);

/* ------------------------------------------------------------------------------------------
 * GINE message passing: the propagate + self term of torch_geometric's GINEConv as called by
 * gine_seq.forward (baseline_and_test_models/pipeline_model_gine.py:652-665) on the concatenated
 * inputs x_cat = [h | ins[batch]], edge_cat = [edge_attr | ins[batch[src]]]:
 *   z[i, :F]    = (1+eps) h[i] + sum_k relu(h[src_k] + edge_attr[e_k]),  e_k = perm ? perm[k] : k
 *   z[i, F:F+D] = (1+eps) ins[g(i)] + deg(i) relu(2 ins[g(i)])
 * z [N, F+D] then feeds the conv's MLP (two library GEMMs on the host side).
 */
GVQA_API int gvqa_gine_aggregate_f32(const float* h, const float* edge_attr, const float* ins,
                                     const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                     const int32_t* node_graph, float* z, int64_t num_nodes,
                                     int32_t feat, int32_t ins_dim, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * GCN message passing: gcn_norm + propagate + bias of torch_geometric's GCNConv
 * (pipeline_model_gcn.py:660): existing self-loops dropped, one loop per
 * node appended, deg over targets, norm = deg^-1/2[src] deg^-1/2[dst].
 *   gvqa_gcn_degree_f32:    dinv[i] = (1 + #non-loop in-edges of i)^-1/2
 *   gvqa_gcn_aggregate_f32: out[i] = sum_{k: src_k != i} dinv[src_k] dinv[i] (xw[src_k] + P[g])
 *                                    + dinv[i]^2 (xw[i] + P[g]) + bias
 * with xw = h @ W[:F] [N,C] and the optional per-graph term P = ins @ W[F:] [B,C].
 */
GVQA_API int gvqa_gcn_degree_f32(const int32_t* rowptr, const int32_t* col_src, float* dinv,
                                 int64_t num_nodes, void* stream);
GVQA_API int gvqa_gcn_aggregate_f32(const float* xw, const float* graph_term, const float* dinv,
                                    const float* bias, const int32_t* rowptr, const int32_t* col_src,
                                    const int32_t* node_graph, float* out, int64_t num_nodes,
                                    int32_t channels, void* stream);

/* ------------------------------------------------------------------------------------------
 * LCGN hop, heads = 1: gat_lcgn.forward + message (baseline_and_test_models/lcgn.py:120-238)
 * after the node projections, with the per-edge Linear cal_x(x_j) hoisted to node level
 * (xv = cal_x(x)) and the one-hot matmuls replaced by a gather on node_graph:
 *   logit_k = sum_c xl[src_k,c] * proj_cmd[g,c] * xr[i,c];   alpha = segment_softmax(leaky_relu)
 *   out[i]  = cal_cmd[g] * sum_k alpha_k xv[src_k] + bias
 * xl / xr / xv are [N, C] views with a common row stride ld (e.g. slices of one [N, 3C] GEMM).
 */
GVQA_API int gvqa_lcgn_hop_f32(const float* xl, const float* xr, const float* xv, int64_t ld,
                               const float* proj_cmd, const float* cal_cmd, const float* bias,
                               const int32_t* rowptr, const int32_t* col_src, const int32_t* node_graph,
                               float* out, int64_t num_nodes, int32_t channels, float negative_slope,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * Steps right before / after the hop stack.
 *  gvqa_gather_add_relu_f32: out[k,:] = act(a[src_k,:] + (b ? b[dst_k,:] : 0) + (c ? c[k,:] : 0) + bias)
 *     -- first Linear of the scene-graph encoder's edge / node MLPs on the split concatenation
 *     (pipeline_model_gat.py:75-77, :94); edge_index is the reference's int64 [2,E].
 *  gvqa_segment_mean_rows_f32: out[i,:] = sum (mean != 0: / max(count,1)) of values[perm[k],:] over
 *     k in [rowptr[i], rowptr[i+1])  -- torch_scatter.scatter_mean by target (pipeline_model_gat.py:96).
 *  gvqa_attention_pool_f32: per graph g, w = softmax(gate[n0:n1]) (exp(x-max)/(sum+1e-16)),
 *     out[g,:] = sum_n w_n x[n,:]  -- MyConditionalGlobalAttention (pipeline_model_gat.py:178-179).
 */
GVQA_API int gvqa_gather_add_relu_f32(const float* a, const float* b, const float* c, const float* bias,
                                      const int64_t* edge_index, float* out, int64_t num_edges,
                                      int32_t feat, int32_t relu, void* stream);
/* the same with the loader-side int32 wire format: edge_index int32 [2,E] */
GVQA_API int gvqa_gather_add_relu_i32_f32(const float* a, const float* b, const float* c, const float* bias,
                                          const int32_t* edge_index, float* out, int64_t num_edges,
                                          int32_t feat, int32_t relu, void* stream);
/* general form: a / b / c are row-strided views (lda, ldb, ldc floats, multiples of 4), e.g. column blocks of one
 * stacked GEMM output; index_bytes 4 or 8.  edge_index == NULL: no gather, out[k] = act(a[k] + b[k] + c[k] + bias)
 * (the node model's  relu(W1 [x | agg] + b1)  with the concatenation split over W1's columns) */
GVQA_API int gvqa_gather_add_relu_strided_f32(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c,
                                              int64_t ldc, const float* bias, const void* edge_index,
                                              int32_t index_bytes, float* out, int64_t num_edges, int32_t feat,
                                              int32_t relu, void* stream);
GVQA_API int gvqa_segment_mean_rows_f32(const float* values, const int32_t* perm, const int32_t* rowptr,
                                        float* out, int64_t num_segments, int32_t feat, int32_t mean,
                                        void* stream);
GVQA_API int gvqa_attention_pool_f32(const float* gate, const float* x, const int32_t* graph_ptr, float* out,
                                     int64_t num_graphs, int32_t channels, void* stream);

/* MyConditionalGlobalAttention's tail in two launches instead of six (pipeline_model_gat.py:166-179):
 *  gvqa_graph_scale_rows_f32: out[n,:] = x[n,:] * q[node_graph[n],:]   (ques_nn(u)[batch] * node_nn(x), q[batch]
 *     never materialised);
 *  gvqa_attention_pool_gate_f32: gate[n] = <hid[n,:], w_gate> + *b_gate (gate_nn's last Linear(C,1)), then the
 *     per-graph softmax and the weighted sum of gvqa_attention_pool_f32 in the same kernel.  gate_scratch: [N]
 *     floats of device scratch (holds the gates afterwards); b_gate: ONE device float or NULL. */
/* Token-embedding sum that opens the scene-graph encoder (pipeline_model_gat.py:583-594):
 *   out[n,:] = (sign ? sign[n] : 1) * sum_t table[tokens[n*T + t], :]
 * tokens: int32 (token_bytes 4, wire format) or int64 (8, the reference's layout); sign: the `added_sym_edge`
 * negation of edge rows (:590) as a per-row +-1 vector, or NULL.  The reference materialises [N, T, F]. */
GVQA_API int gvqa_embedding_sum_f32(const float* table, int64_t vocab, const void* tokens, int32_t token_bytes,
                                    const float* sign, float* out, int64_t num_rows, int32_t tokens_per_row,
                                    int32_t feat, void* stream);
/* out[n,c] = x[n,c]*scale[c] + shift[c], then ReLU when relu != 0: BatchNorm1d(eval)+ReLU of the GCN / GINE
 * sequences (pipeline_model_gcn.py:666-668), whose conv results the reference discards.  In-place allowed. */
GVQA_API int gvqa_affine_relu_f32(const float* x, const float* scale, const float* shift, float* out,
                                  int64_t num_rows, int32_t channels, int32_t relu, void* stream);
GVQA_API int gvqa_graph_scale_rows_f32(const float* x, const float* q, const int32_t* node_graph, float* out,
                                       int64_t num_nodes, int32_t channels, void* stream);
GVQA_API int gvqa_attention_pool_gate_f32(const float* hid, int32_t hid_channels, const float* w_gate,
                                          const float* b_gate, float* gate_scratch, const float* x,
                                          const int32_t* graph_ptr, float* out, int64_t num_graphs,
                                          int32_t channels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVQA_B200_H_ */
