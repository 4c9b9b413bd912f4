// Fused GAT hop for batched disjoint scene graphs (see include/gvqa_b200.h: gvqa_gat_hop_f32).
//
// One launch = one hop of the reference's gat_seq.forward (gat_skip.py:254-276) after the node
// projection: gather of source rows, additive edge logits, LeakyReLU, per-destination softmax
// (PyG semantics: exp(l-max)/(sum+1e-16)), weighted aggregation, head mean, +bias, +skip,
// BatchNorm(eval)+ReLU epilogue.  No atomics: destination-CSR, one warp owns one output row,
// in-edges are summed in the caller's edge order -> bitwise deterministic.
//
// Three kernels (gvqa_gat_hop_args.variant; 0 = auto):
//  1 gat_hop_gather_kernel -- one warp per destination node, everything from global/L2.
//  3 gat_hop_block_kernel  -- one 128-thread CTA per 16 consecutive destination nodes: the CTA
//    loads its CSR slice and logit terms edge-parallel (3 dependent round trips for the whole
//    CTA instead of 5 per node), runs the softmax with one thread per (node, head) in shared
//    memory, then every warp streams its nodes' source rows with 32 independent 128-bit loads in
//    flight per lane.  The whole grid is resident in ONE wave (no tail), so the kernel behaves
//    like a streaming copy.  Default.
//  4 gat_hop_ws_kernel     -- persistent, warp-specialised: one producer warp per CTA claims chunks of 8 destination
//    nodes from a device counter and prepares them (index round trips + softmax) into a 3-slot shared-memory ring
//    while 4 consumer warps stream the previous chunks; only the first chunk's index latency is exposed and fast
//    SMs take more chunks than slow ones (the block kernel's per-CTA stream phase spreads 7.9-17.2 us for equal
//    work, profiles/r01/hop_timeline_block_kernel.txt).
//  5 gat_hop_graphln_kernel -- epilogue GVQA_EPI_GRAPH_LN: one CTA per graph keeps the graph's output rows in shared
//    memory, so the per-graph LayerNorm of my_graph_layernorm.py:52-78 (two-pass variance) runs before the single
//    HBM write.  Selected by the epilogue mode, not by `variant`.
//  2 gat_hop_staged_kernel -- one CTA per (graph, channel-window): softmax once for all heads,
//    then one stage per head (the window of every node row of the graph) streamed by the TMA
//    engine (cp.async.bulk + mbarrier) through a 3-deep shared-memory ring, accumulators in
//    registers across heads; each x_l byte crosses L2->SM exactly once.  Measured slower than 3 on
//    B200 for GQA-sized graphs (per-copy TMA issue cost, profiles/microbench), kept selectable.
#include "common.cuh"

namespace gvqa {

constexpr int kEdgeChunk = 32;  // in-edges whose alpha are staged per warp at a time

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct HopParams {
  unsigned long long* trace;   // debug only (gvqa_debug_set_hop_trace): [grid][8] globaltimer stamps per CTA
  const float* __restrict__ x_l;
  const float* __restrict__ graph_bias;
  const float* __restrict__ a_node;
  const float* __restrict__ a_graph;
  const float* __restrict__ a_edge;
  const int32_t* __restrict__ rowptr;
  const int32_t* __restrict__ col_src;
  const int32_t* __restrict__ perm;
  const int32_t* __restrict__ graph_ptr;
  const int32_t* __restrict__ node_graph;
  const float* __restrict__ h_prev;
  const float* __restrict__ bias;
  const float* __restrict__ ep_scale;
  const float* __restrict__ ep_shift;
  float* __restrict__ h_out;
  float* __restrict__ alpha_out;
  int64_t ldx, lde, lda, ldgb, ldag;   // row strides of x_l, a_edge, a_node, graph_bias, a_graph (floats)
  int32_t N, E, B, C;
  float slope;
  int32_t epilogue;
  int32_t early;   // GVQA_HOP_INPUTS_OLDER_THAN_PREDECESSOR: topology / a_edge / a_graph may be read before pdl_wait()
  int32_t prefetch;   // GVQA_HOP_PREFETCH (experiments): 1 = the CTA's h_prev rows are pulled into L2 during the index prologue
  const int32_t* slab_idx;   // variant 5 (slab kernel): per-CTA index slabs / this hop's logit-term slabs
  const float* slab_f;
  int32_t win;               // rows of a_node staged either side of the CTA's node range (max nodes per graph)
  int32_t* sched;     // variant 4: {next chunk, finished CTAs}, zero before the first launch, self-resetting
  const float* ln_weight;   // GVQA_EPI_GRAPH_LN: one float each (or NULL), my_graph_layernorm.py:40-41
  const float* ln_bias;
  float ln_eps;
};

// Softmax weights of the in-edges [e0,e1) of node `i` for all H heads, computed by one warp.
// Lane l serves head (l % H); the 32/H lanes of a head stride over the edges.  On return the
// per-head (max, 1/(sum+1e-16)) live in every lane of that head.
template <int H>
struct WarpSoftmax {
  float m, inv;
  const HopParams& p;
  int i, g, e0, e1, lane, head, slot;
  float target_term;

  __device__ __forceinline__ WarpSoftmax(const HopParams& p_, int i_, int g_, int e0_, int e1_, int lane_)
      : p(p_), i(i_), g(g_), e0(e0_), e1(e1_), lane(lane_) {
    head = lane % H;
    slot = lane / H;
    target_term = p.a_node[(int64_t)i * p.lda + H + head];
    if (p.a_graph) target_term += p.a_graph[(int64_t)g * p.ldag + head];
  }

  __device__ __forceinline__ float logit(int k) const {
    const int src = p.col_src[k];
    const int64_t e = p.perm ? p.perm[k] : k;
    const float v = p.a_node[(int64_t)src * p.lda + head] + target_term + p.a_edge[e * p.lde + head];
    return leaky_relu(v, p.slope);
  }

  // first_logit: logit of edge e0+slot if it exists (kept by the caller for reuse)
  __device__ __forceinline__ void run(float& first_logit) {
    constexpr int per = 32 / H;
    float mx = -INFINITY;
    first_logit = 0.f;
    if (e0 + slot < e1) {
      first_logit = logit(e0 + slot);
      mx = first_logit;
    }
    for (int k = e0 + slot + per; k < e1; k += per) mx = fmaxf(mx, logit(k));
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    m = mx;
    float s = 0.f;
    if (e0 + slot < e1) s = expf(first_logit - m);
    for (int k = e0 + slot + per; k < e1; k += per) s += expf(logit(k) - m);
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    inv = 1.0f / (s + 1e-16f);
  }
};

// Epilogue for one float4 column (absolute float4 index c4) of output row i: head mean,
// + per-graph instruction term (rows with in-edges only), +bias, +skip, BatchNorm(eval) affine,
// ReLU; 128-bit streaming store.
__device__ __forceinline__ float4 epilogue_value4(const HopParams& p, int c4, float4 o, float inv_heads, bool add_gb,
                                                  const float4& gb, bool have_skip, const float4& skip,
                                                  int epilogue);

__device__ __forceinline__ void epilogue_store4(const HopParams& p, int i, int c4, float4 o, float inv_heads,
                                                bool add_gb, const float4& gb, bool have_skip,
                                                const float4& skip) {
  stg_stream(p.h_out + (int64_t)i * p.C + 4 * c4,
             epilogue_value4(p, c4, o, inv_heads, add_gb, gb, have_skip, skip, p.epilogue));
}

__device__ __forceinline__ float4 epilogue_value4(const HopParams& p, int c4, float4 o, float inv_heads, bool add_gb,
                                                  const float4& gb, bool have_skip, const float4& skip,
                                                  int epilogue) {
  o.x *= inv_heads; o.y *= inv_heads; o.z *= inv_heads; o.w *= inv_heads;
  if (add_gb) { o.x += gb.x; o.y += gb.y; o.z += gb.z; o.w += gb.w; }
  if (p.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + c4);
    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
  }
  if (have_skip) { o.x += skip.x; o.y += skip.y; o.z += skip.z; o.w += skip.w; }
  if (epilogue == GVQA_EPI_AFFINE || epilogue == GVQA_EPI_AFFINE_RELU) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.ep_scale) + c4);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.ep_shift) + c4);
    o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y);
    o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
    if (epilogue == GVQA_EPI_AFFINE_RELU) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
  }
  return o;     // GVQA_EPI_GRAPH_LN: the per-graph normalisation follows in gat_hop_graphln_kernel
}

// One warp computes output row i for the float4 column window [c4_lo, c4_lo + c4_n), gathering
// source rows from global memory.  alpha_w / src_w: per-warp shared scratch (32*H floats, 32 ints).
template <int J, int H>
__device__ __forceinline__ void gather_node(const HopParams& p, int i, int lane, int c4_lo, int c4_n,
                                            float* alpha_w, int32_t* src_w, bool write_alpha) {
  const int e0 = p.rowptr[i], e1 = p.rowptr[i + 1];
  const int g = p.node_graph[i];
  constexpr int per = 32 / H;

  float4 acc[J], skip[J], gb[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    skip[j] = acc[j];
    gb[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < c4_n) {
      if (p.h_prev) skip[j] = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * (c4_lo + c4));
      if (p.graph_bias && e1 > e0) gb[j] = ldg_cached(p.graph_bias + (int64_t)g * p.ldgb + 4 * (c4_lo + c4));
    }
  }

  if (e1 > e0) {
    WarpSoftmax<H> sm(p, i, g, e0, e1, lane);
    float l0;
    sm.run(l0);
    for (int c0 = e0; c0 < e1; c0 += kEdgeChunk) {
      const int cn = min(kEdgeChunk, e1 - c0);
      // stage alpha (and sources) of this chunk; the first chunk reuses the cached logit
      for (int kk = sm.slot; kk < cn; kk += per) {
        const int k = c0 + kk;
        const float l = (c0 == e0 && kk == sm.slot) ? l0 : sm.logit(k);
        const float a = expf(l - sm.m) * sm.inv;
        alpha_w[kk * H + sm.head] = a;
        if (sm.head == 0) src_w[kk] = p.col_src[k];
        if (write_alpha && p.alpha_out) {
          const int64_t e = p.perm ? p.perm[k] : k;
          p.alpha_out[e * H + sm.head] = a;
        }
      }
      __syncwarp();
#pragma unroll 2
      for (int kk = 0; kk < cn; ++kk) {
        const float* row = p.x_l + (int64_t)src_w[kk] * p.ldx + 4 * c4_lo;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float a = alpha_w[kk * H + h];
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int c4 = lane + 32 * j;
            if (c4 < c4_n) fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
          }
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < c4_n)
      epilogue_store4(p, i, c4_lo + c4, acc[j], 1.0f / H, p.graph_bias != nullptr && e1 > e0, gb[j],
                      p.h_prev != nullptr, skip[j]);
  }
}

// ------------------------------------------------------------------------------------------
// Kernel 1: gather from global/L2.  One warp per destination node; lane owns float4 columns
// lane, lane+32, ... (J of them) of the C output channels.
// ------------------------------------------------------------------------------------------
template <int J, int H>
__global__ void __launch_bounds__(256) gat_hop_gather_kernel(const HopParams p) {
  __shared__ float alpha_s[8][kEdgeChunk * H];
  __shared__ int32_t src_s[8][kEdgeChunk];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + wid;
  if (i >= p.N) return;
  gather_node<J, H>(p, i, lane, 0, p.C >> 2, alpha_s[wid], src_s[wid], true);
}

// ------------------------------------------------------------------------------------------
// Kernel 3: block-phase gather.  CTA = 4 warps = 16 consecutive destination nodes.
// ------------------------------------------------------------------------------------------
constexpr int kBlkNodes = 16;     // max destination nodes per CTA (the launch picks npc <= 16 so that the
                                  // grid is one balanced wave: npc = ceil(N / (148 SMs x 4 resident CTAs)))
constexpr int kBlkThreads = 128;  // 4 warps; warp w owns nodes w, w+4, w+8, w+12 of the CTA
constexpr int kBlkEdgeCap = 256;  // in-edges of the CTA's nodes staged in shared memory

template <int J, int H, int kBlkThreads = 128, int kBlkNodes = 16, int kBlkEdgeCap = 256>
__global__ void __launch_bounds__(kBlkThreads, 512 / kBlkThreads) gat_hop_block_kernel(const HopParams p, const int npc) {
  constexpr int kWarps = kBlkThreads / 32;
  static_assert(kBlkNodes * H <= kBlkThreads || H > 8, "one (node, head) pair per thread in the prologue");
  __shared__ int32_t rp_s[kBlkNodes + 1];
  __shared__ int32_t gid_s[kBlkNodes];
  __shared__ float tgt_s[kBlkNodes * H];
  __shared__ int32_t src_s[kBlkEdgeCap];
  __shared__ __align__(16) float alpha_s[kBlkEdgeCap * H];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = blockIdx.x * npc;
  const int nn = min(npc, p.N - i0);
  const int C4 = p.C >> 2;
#define GVQA_HOP_TRACE(col) do { if (p.trace && tid == 0) p.trace[(size_t)blockIdx.x * 8 + (col)] = gtime_ns(); } while (0)
  GVQA_HOP_TRACE(0);
  if (p.prefetch && p.h_prev) {
    // the skip rows of the CTA's own nodes are contiguous, older than the predecessor kernel and known without any
    // index: request them while the three dependent index round trips keep the memory system idle
    const char* base = reinterpret_cast<const char*>(p.h_prev + (int64_t)i0 * p.C);
    const int bytes = nn * p.C * 4;
    for (int off = tid * 128; off < bytes; off += kBlkThreads * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }

  // ---- round trips 1+2 touch only the per-batch topology and the pre-pass outputs, none of which the
  // projection GEMM right before this kernel writes: under programmatic dependent launch they run while that
  // GEMM is still draining on other SMs.  Row pointers first (the only thing the first barrier waits for);
  // graph ids and their logit terms are requested now and consumed after the second barrier.
  if (!p.early) {            // the caller made no promise about the previous kernel: wait first
    pdl_wait();
    pdl_launch_dependents();
  }
  for (int t = tid; t <= nn; t += kBlkThreads) rp_s[t] = p.rowptr[i0 + t];
  float tg_reg = 0.f;
  int g_reg = 0;
  const int my_node = tid / H, my_head = tid - my_node * H;
  if (tid < nn * H) {
    g_reg = p.node_graph[i0 + my_node];
    if (p.a_graph) tg_reg = p.a_graph[(int64_t)g_reg * p.ldag + my_head];
  }
  __syncthreads();
  GVQA_HOP_TRACE(1);
  const int eA = rp_s[0], eC = rp_s[nn] - eA;

  if (eC > kBlkEdgeCap) {
    // hub-heavy block: per-warp chunked path (scratch carved from alpha_s / src_s)
    if (p.early) {
      pdl_wait();
      pdl_launch_dependents();
    }
    for (int node = wid; node < nn; node += kWarps)
      gather_node<J, H>(p, i0 + node, lane, 0, C4, alpha_s + wid * kEdgeChunk * H, src_s + wid * kEdgeChunk, true);
    return;
  }

  // sources, then edge and source logit terms, edge-parallel.  Without the early-start promise both terms are
  // requested together (round trip 3); with it the edge term is staged first and the source term -- an output of
  // the previous kernel -- is added after the dependency wait.
  for (int k = tid; k < eC; k += kBlkThreads) {
    const int src = p.col_src[eA + k];
    const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
    src_s[k] = src;
    const float* an = p.a_node + (int64_t)src * p.lda;
#pragma unroll
    for (int h = 0; h < H; ++h) alpha_s[k * H + h] = p.a_edge[e * p.lde + h] + (p.early ? 0.f : an[h]);
  }

  // ---- everything below reads what the previous kernel wrote (a_node and x_l come out of the GEMM) --------
  if (p.early) {
    pdl_wait();
    pdl_launch_dependents();
    for (int k = tid; k < eC; k += kBlkThreads) {   // each thread revisits the edges it staged above
      const float* an = p.a_node + (int64_t)src_s[k] * p.lda;
#pragma unroll
      for (int h = 0; h < H; ++h) alpha_s[k * H + h] += an[h];
    }
  }
  if (tid < nn * H) {
    tg_reg += p.a_node[(int64_t)(i0 + my_node) * p.lda + H + my_head];
    tgt_s[tid] = tg_reg;
    if (my_head == 0) gid_s[my_node] = g_reg;
  }
  __syncthreads();
  GVQA_HOP_TRACE(2);

  // ---- softmax: one thread per (node, head).  The passes over the node's in-edges only READ shared memory
  // (the logit is recomputed instead of being written back), so the unrolled iterations overlap instead of
  // forming one long load-store chain ----------------------------------------------------------
  if (tid < nn * H) {
    const int r0 = rp_s[my_node] - eA, r1 = rp_s[my_node + 1] - eA;
    const float tg = tg_reg;
    const int h = my_head;
    float mx = -INFINITY;
#pragma unroll 4
    for (int k = r0; k < r1; ++k) mx = fmaxf(mx, leaky_relu(alpha_s[k * H + h] + tg, p.slope));
    float sum = 0.f;
#pragma unroll 4
    for (int k = r0; k < r1; ++k) sum += expf(leaky_relu(alpha_s[k * H + h] + tg, p.slope) - mx);
    const float inv = 1.0f / (sum + 1e-16f);
#pragma unroll 4
    for (int k = r0; k < r1; ++k) {
      const float a = expf(leaky_relu(alpha_s[k * H + h] + tg, p.slope) - mx) * inv;
      alpha_s[k * H + h] = a;     // this thread is the only reader and writer of column h of these rows
      if (p.alpha_out) {
        const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
        p.alpha_out[e * H + h] = a;
      }
    }
  }
  __syncthreads();
  GVQA_HOP_TRACE(3);

  // ---- weighted gather: warp w owns nodes w, w+4, ...; 2 edges x H x J 128-bit loads in flight
#pragma unroll 1
  for (int node = wid; node < nn; node += kWarps) {
    const int i = i0 + node;
    const int r0 = rp_s[node] - eA, r1 = rp_s[node + 1] - eA;
    float4 acc[J], skip[J], gb[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      skip[j] = acc[j];
      gb[j] = acc[j];
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        if (p.h_prev) skip[j] = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * c4);
        if (p.graph_bias && r1 > r0) gb[j] = ldg_cached(p.graph_bias + (int64_t)gid_s[node] * p.ldgb + 4 * c4);
      }
    }
#pragma unroll 2
    for (int k = r0; k < r1; ++k) {
      const float* row = p.x_l + (int64_t)src_s[k] * p.ldx;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float a = alpha_s[k * H + h];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < C4) fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4)
        epilogue_store4(p, i, c4, acc[j], 1.0f / H, p.graph_bias != nullptr && r1 > r0, gb[j],
                        p.h_prev != nullptr, skip[j]);
    }
  }
  if (p.trace && lane == 0 && wid < 4) p.trace[(size_t)blockIdx.x * 8 + 4 + wid] = gtime_ns();   // per-warp finish
}


// ------------------------------------------------------------------------------------------
// Shared streaming step of kernels 4 and 5: output row i from staged (src, alpha) lists in shared
// memory.  2 edges x H x J independent 128-bit loads in flight per lane; the epilogue value goes to
// `sink(c4, value)`.
// ------------------------------------------------------------------------------------------
template <int J, int H, typename Sink>
__device__ __forceinline__ void stream_node(const HopParams& p, int i, int lane, int C4, const int32_t* src_s,
                                            const float* alpha_s, int r0, int r1, int gid, int epilogue, Sink&& sink) {
  float4 acc[J], skip[J], gb[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    skip[j] = acc[j];
    gb[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      if (p.h_prev) skip[j] = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * c4);
      if (p.graph_bias && r1 > r0) gb[j] = ldg_cached(p.graph_bias + (int64_t)gid * p.ldgb + 4 * c4);
    }
  }
#pragma unroll 2
  for (int k = r0; k < r1; ++k) {
    const float* row = p.x_l + (int64_t)src_s[k] * p.ldx;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = alpha_s[k * H + h];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4)
      sink(c4, epilogue_value4(p, c4, acc[j], 1.0f / H, p.graph_bias != nullptr && r1 > r0, gb[j], p.h_prev != nullptr,
                               skip[j], epilogue));
  }
}

// 128-bit read-only load (L1-allocating like __ldg) that the compiler must issue where it stands: used to start the
// first node's row traffic BEFORE the softmax phase, so that the phase hides behind the memory latency.
__device__ __forceinline__ float4 ldg_issue_now(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// stream_node for a node whose skip row, graph term and FIRST in-edge's source rows were requested earlier
// (`pre[h][j]`): same accumulation order as stream_node (edge r0 first, heads in order), so the results are identical.
template <int J, int H, int HP, typename Sink>
__device__ __forceinline__ void stream_node_pre(const HopParams& p, int i, int lane, int C4, const int32_t* src_s,
                                                const float* alpha_s, int r0, int r1, int gid, const float4 (&pre)[HP][J],
                                                int epilogue, Sink&& sink) {
  float4 acc[J], skip[J], gb[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    skip[j] = acc[j];
    gb[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      if (p.h_prev) skip[j] = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * c4);
      if (p.graph_bias && r1 > r0) gb[j] = ldg_cached(p.graph_bias + (int64_t)gid * p.ldgb + 4 * c4);
    }
  }
  if (r1 > r0) {
    const float* row = p.x_l + (int64_t)src_s[r0] * p.ldx;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = alpha_s[r0 * H + h];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) {
          if (h < HP) fma4(acc[j], a, pre[h < HP ? h : 0][j]);
          else fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
        }
      }
    }
  }
#pragma unroll 2
  for (int k = r0 + 1; k < r1; ++k) {
    const float* row = p.x_l + (int64_t)src_s[k] * p.ldx;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = alpha_s[k * H + h];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4)
      sink(c4, epilogue_value4(p, c4, acc[j], 1.0f / H, p.graph_bias != nullptr && r1 > r0, gb[j], p.h_prev != nullptr,
                               skip[j], epilogue));
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------
// Kernel 4: persistent warp-specialised hop.  Warp 0 = producer (claims chunks, index round trips,
// softmax into a ring slot), warps 1..4 = consumers (weighted gather + epilogue).
// ------------------------------------------------------------------------------------------
constexpr int kWsSlots = 3;
constexpr int kWsEdgeCap = 128;   // in-edges of one chunk staged in a slot; larger chunks take the per-warp path

template <int H>
struct WsGeom {
  // kChunk * H <= 32 (one producer lane per (node, head)); one consumer warp per node of a chunk, so a chunk is ONE
  // round of the consumers; 7 + 1 warps = 256 threads leave 128 registers per thread at two CTAs per SM
  static constexpr int kChunk = H >= 8 ? 3 : 7;
  static constexpr int kConsumers = kChunk;
  static constexpr int kThreads = 32 * (kConsumers + 1);
};

template <int H>
struct __align__(16) WsSlot {
  float alpha[kWsEdgeCap * H];
  int32_t src[kWsEdgeCap];
  int32_t rp[WsGeom<H>::kChunk + 1];
  int32_t gid[WsGeom<H>::kChunk];
  int32_t i0, nn, overflow, pad;
};

template <int J, int H>
__global__ void __launch_bounds__(WsGeom<H>::kThreads, H >= 8 ? 4 : 2) gat_hop_ws_kernel(const HopParams p, const int nchunks) {
  constexpr int kChunk = WsGeom<H>::kChunk;
  constexpr int kWsConsumers = WsGeom<H>::kConsumers;
  __shared__ WsSlot<H> slots[kWsSlots];
  __shared__ __align__(8) uint64_t full_bar[kWsSlots], empty_bar[kWsSlots];
  __shared__ float gather_alpha[kWsConsumers][kEdgeChunk * H];   // scratch of the oversize-chunk path
  __shared__ int32_t gather_src[kWsConsumers][kEdgeChunk];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int C4 = p.C >> 2;
  if (p.trace && tid == 0) {
    p.trace[(size_t)blockIdx.x * 8 + 0] = gtime_ns();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[(size_t)blockIdx.x * 8 + 3] = smid;
  }
  if (tid == 0) {
    for (int s = 0; s < kWsSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kWsConsumers);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();

  if (wid == 0) {
    // ---------------- producer ----------------
    const int node = lane / H, head = lane - node * H;
    // chunk claims: the first is static (blockIdx.x: no round trip before the first index load), every later one is
    // requested one iteration ahead, so the atomic's latency hides behind the preparation of the current chunk
    int c = blockIdx.x, c_next = 0;
#pragma unroll 1
    for (int it = 0;; ++it) {
      const int s = it % kWsSlots;
      if (it > 0) c = __shfl_sync(kFull, c_next, 0);
      if (lane == 0 && c < nchunks) c_next = (int)gridDim.x + atomicAdd(p.sched, 1);
      mbar_wait(&empty_bar[s], (uint32_t)(((it / kWsSlots) & 1) ^ 1));     // consumers have left the slot
      WsSlot<H>& sl = slots[s];
      if (c >= nchunks) {
        if (lane == 0) {
          sl.nn = 0;
          // the CTA whose final claim is the last one resets the scheduler words for the next launch
          if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 8 + 2] = (unsigned long long)it;
        break;
      }
      const int i0 = c * kChunk;
      const int nn = min(kChunk, p.N - i0);
      // round trip 1: row pointers, graph ids, target-side logit terms
      const int rpv = lane <= nn ? p.rowptr[i0 + lane] : 0;
      float tg = 0.f;
      int g = 0;
      const bool owner = lane < nn * H;
      if (owner) {
        g = p.node_graph[i0 + node];
        if (p.a_graph) tg = p.a_graph[(int64_t)g * p.ldag + head];
        tg += p.a_node[(int64_t)(i0 + node) * p.lda + H + head];
      }
      const int eA = __shfl_sync(kFull, rpv, 0);
      const int eC = __shfl_sync(kFull, rpv, nn) - eA;
      if (lane <= nn) sl.rp[lane] = rpv - eA;
      if (owner && head == 0) sl.gid[node] = g;
      const bool over = eC > kWsEdgeCap;
      if (!over) {
        // round trips 2 + 3: sources / edge ids, then the edge and source logit terms
        for (int k = lane; k < eC; k += 32) {
          const int src = p.col_src[eA + k];
          const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
          sl.src[k] = src;
          const float* an = p.a_node + (int64_t)src * p.lda;
#pragma unroll
          for (int h = 0; h < H; ++h) sl.alpha[k * H + h] = p.a_edge[e * p.lde + h] + an[h];
        }
      }
      __syncwarp();
      if (!over && owner) {
        const int r0 = sl.rp[node], r1 = sl.rp[node + 1];
        float mx = -INFINITY;
#pragma unroll 4
        for (int k = r0; k < r1; ++k) mx = fmaxf(mx, leaky_relu(sl.alpha[k * H + head] + tg, p.slope));
        float sum = 0.f;
#pragma unroll 4
        for (int k = r0; k < r1; ++k) sum += expf(leaky_relu(sl.alpha[k * H + head] + tg, p.slope) - mx);
        const float inv = 1.0f / (sum + 1e-16f);
#pragma unroll 4
        for (int k = r0; k < r1; ++k) {
          const float a = expf(leaky_relu(sl.alpha[k * H + head] + tg, p.slope) - mx) * inv;
          sl.alpha[k * H + head] = a;
          if (p.alpha_out) {
            const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
            p.alpha_out[e * H + head] = a;
          }
        }
      }
      if (lane == 0) {
        sl.i0 = i0;
        sl.nn = nn;
        sl.overflow = over ? 1 : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (p.trace && lane == 0 && it == 0) p.trace[(size_t)blockIdx.x * 8 + 1] = gtime_ns();
    }
  } else {
    // ---------------- consumers ----------------
    const int cw = wid - 1;
#pragma unroll 1
    for (int it = 0;; ++it) {
      const int s = it % kWsSlots;
      mbar_wait(&full_bar[s], (uint32_t)((it / kWsSlots) & 1));
      const WsSlot<H>& sl = slots[s];
      const int nn = sl.nn;
      if (nn == 0) break;
      const int i0 = sl.i0;
      if (sl.overflow) {
        for (int node = cw; node < nn; node += kWsConsumers)
          gather_node<J, H>(p, i0 + node, lane, 0, C4, gather_alpha[cw], gather_src[cw], true);
      } else {
#pragma unroll 1
        for (int node = cw; node < nn; node += kWsConsumers) {
          const int i = i0 + node;
          stream_node<J, H>(p, i, lane, C4, sl.src, sl.alpha, sl.rp[node], sl.rp[node + 1], sl.gid[node], p.epilogue,
                            [&](int c4, const float4& v) { stg_stream(p.h_out + (int64_t)i * p.C + 4 * c4, v); });
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 8 + 4 + cw] = gtime_ns();
  }
}

// ------------------------------------------------------------------------------------------
// Kernel 5: graph-LayerNorm epilogue (my_graph_layernorm.py:52-78 applied to conv + skip).  One CTA per
// graph; the graph's pre-normalisation rows live in shared memory when they fit (n <= n_cap rows), else
// they make one trip through h_out (written, then re-read by the same CTA).  Two-pass variance like the
// reference: mean first, then the centred second moment; eps is added to the standard deviation.
// ------------------------------------------------------------------------------------------
constexpr int kLnHopThreads = 256;
constexpr int kLnHopEdgeCap = 512;     // staged in-edges per graph; larger graphs take the per-warp path

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float t = lane < kLnHopThreads / 32 ? scratch[lane] : 0.f;
  return warp_sum(t);
}

template <int J, int H>
__global__ void __launch_bounds__(kLnHopThreads) gat_hop_graphln_kernel(const HopParams p, const int n_cap) {
  extern __shared__ __align__(16) float rows_s[];                 // [n_cap][C]
  __shared__ __align__(16) float alpha_s[kLnHopEdgeCap * H];
  __shared__ int32_t src_s[kLnHopEdgeCap];
  __shared__ float scratch[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kWarps = kLnHopThreads / 32;
  const int g = blockIdx.x;
  const int n0 = p.graph_ptr[g], n1 = p.graph_ptr[g + 1], n = n1 - n0;
  if (n <= 0) return;
  const int C = p.C, C4 = C >> 2;
  const int eA = p.rowptr[n0], eC = p.rowptr[n1] - eA;
  const bool rows_fit = n <= n_cap;
  const bool edges_fit = eC <= kLnHopEdgeCap;
  float* rows = rows_fit ? rows_s : p.h_out + (int64_t)n0 * C;

  float local = 0.f;     // sum of this thread's output elements
  if (edges_fit) {
    for (int k = tid; k < eC; k += kLnHopThreads) {
      const int src = p.col_src[eA + k];
      const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
      src_s[k] = src;
      const float* an = p.a_node + (int64_t)src * p.lda;
#pragma unroll
      for (int h = 0; h < H; ++h) alpha_s[k * H + h] = p.a_edge[e * p.lde + h] + an[h];
    }
    __syncthreads();
    for (int t = tid; t < n * H; t += kLnHopThreads) {
      const int node = t / H, h = t - node * H;
      const int r0 = p.rowptr[n0 + node] - eA, r1 = p.rowptr[n0 + node + 1] - eA;
      float tg = p.a_node[(int64_t)(n0 + node) * p.lda + H + h];
      if (p.a_graph) tg = p.a_graph[(int64_t)g * p.ldag + h] + tg;
      float mx = -INFINITY;
      for (int k = r0; k < r1; ++k) mx = fmaxf(mx, leaky_relu(alpha_s[k * H + h] + tg, p.slope));
      float sum = 0.f;
      for (int k = r0; k < r1; ++k) sum += expf(leaky_relu(alpha_s[k * H + h] + tg, p.slope) - mx);
      const float inv = 1.0f / (sum + 1e-16f);
      for (int k = r0; k < r1; ++k) {
        const float a = expf(leaky_relu(alpha_s[k * H + h] + tg, p.slope) - mx) * inv;
        alpha_s[k * H + h] = a;
        if (p.alpha_out) {
          const int64_t e = p.perm ? p.perm[eA + k] : (eA + k);
          p.alpha_out[e * H + h] = a;
        }
      }
    }
    __syncthreads();
    for (int node = wid; node < n; node += kWarps) {
      const int i = n0 + node;
      const int r0 = p.rowptr[i] - eA, r1 = p.rowptr[i + 1] - eA;
      float* dst = rows + (int64_t)node * C;
      stream_node<J, H>(p, i, lane, C4, src_s, alpha_s, r0, r1, g, GVQA_EPI_NONE, [&](int c4, const float4& v) {
        *reinterpret_cast<float4*>(dst + 4 * c4) = v;
        local += (v.x + v.y) + (v.z + v.w);
      });
    }
  } else {
    // oversize graph: per-warp gather straight to h_out without the normalisation, then normalise in place
    HopParams q = p;
    q.epilogue = GVQA_EPI_NONE;
    rows = p.h_out + (int64_t)n0 * C;
    float* alpha_w = alpha_s + wid * (kEdgeChunk * H);
    int32_t* src_w = src_s + wid * kEdgeChunk;
    for (int node = wid; node < n; node += kWarps) gather_node<J, H>(q, n0 + node, lane, 0, C4, alpha_w, src_w, true);
    __syncthreads();
    const int64_t cnt4 = (int64_t)n * C4;
    for (int64_t t = tid; t < cnt4; t += kLnHopThreads) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(rows) + t);
      local += (v.x + v.y) + (v.z + v.w);
    }
  }
  const float norm = (float)n * (float)C;
  const float mean = block_sum_256(local, scratch) / norm;
  const int64_t cnt4 = (int64_t)n * C4;
  const bool in_smem = rows == rows_s;
  float q2 = 0.f;
  for (int64_t t = tid; t < cnt4; t += kLnHopThreads) {
    const float4 v = in_smem ? reinterpret_cast<const float4*>(rows)[t] : __ldcg(reinterpret_cast<const float4*>(rows) + t);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q2 += (a * a + b * b) + (c * c + d * d);
  }
  const float var = block_sum_256(q2, scratch) / norm;
  const float denom = sqrtf(var) + p.ln_eps;
  const bool affine = p.ln_weight != nullptr && p.ln_bias != nullptr;
  const float w = affine ? __ldg(p.ln_weight) : 1.f;
  const float b0 = affine ? __ldg(p.ln_bias) : 0.f;
  float* out = p.h_out + (int64_t)n0 * C;
  for (int64_t t = tid; t < cnt4; t += kLnHopThreads) {
    const float4 v = in_smem ? reinterpret_cast<const float4*>(rows)[t] : __ldcg(reinterpret_cast<const float4*>(rows) + t);
    float4 o;
    o.x = (v.x - mean) / denom; o.y = (v.y - mean) / denom; o.z = (v.z - mean) / denom; o.w = (v.w - mean) / denom;
    if (affine) { o.x = o.x * w + b0; o.y = o.y * w + b0; o.z = o.z * w + b0; o.w = o.w * w + b0; }
    stg_stream(out + 4 * t, o);
  }
}


// ------------------------------------------------------------------------------------------
// Kernel 6: block-phase gather with a ONE-round-trip prologue ("slab" kernel).
//
// The block kernel's prologue is three dependent trips (row pointers -> sources / edge ids -> logit terms) during
// which no bulk traffic flows (3.6 us of a ~20 us kernel at cfg2, profiles/r01/hop_timeline_block_kernel.txt).
// Everything in it except a_node is hop-invariant per batch, and the CTA partition depends on N only.  So a
// per-batch pass (gat_slab_build_kernel, beside the CSR build) writes, per hop CTA, a fixed-stride slab
//     idx  : { eC | -1, eA, rp_rel[npc+1], gid[npc], src[cap], local target[cap] }       (int32, once per batch)
//     f[j] : { a_graph[g(i), h] per node, a_edge_j[perm[k], h] per staged in-edge }       (fp32, per hop j)
// whose address depends on blockIdx only: the hop CTA pulls both with two TMA bulk copies (cp.async.bulk +
// mbarrier) while its threads load the a_node rows of a WINDOW of nodes around its own range (sources live in the
// same graph, i.e. within max_nodes_per_graph rows) -- one trip.  Sources outside the window and CTAs whose in-edges
// exceed the slab capacity fall back to global loads / the block kernel's path, so the hints stay hints.
// ------------------------------------------------------------------------------------------
struct SlabGeom {
  int32_t npc, cap, grid, idx_stride, f_stride;    // strides in 4-byte words (multiples of 4 -> 16-byte aligned)
};

__host__ __device__ inline SlabGeom slab_geom(int64_t N, int64_t E, int H) {
  SlabGeom g;
  int npc = (int)((N + kNumSMs * 4 - 1) / (kNumSMs * 4));
  g.npc = npc < 4 ? 4 : (npc > kBlkNodes ? kBlkNodes : npc);
  const int64_t avg2 = N > 0 ? (2 * E * g.npc + N - 1) / N : 0;     // twice the mean in-edges of a CTA
  int cap = (int)((avg2 + 31) / 32 * 32);
  g.cap = cap < 32 ? 32 : (cap > kBlkEdgeCap ? kBlkEdgeCap : cap);
  g.grid = (int)((N + g.npc - 1) / g.npc);
  g.idx_stride = (2 + (g.npc + 1) + g.npc + 3) / 4 * 4 + 2 * g.cap;      // header | src[cap] | local target[cap]
  g.f_stride = ((g.npc + g.cap) * H + 3) / 4 * 4;
  return g;
}

template <int H>
__global__ void __launch_bounds__(128) gat_slab_build_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ node_graph, const float* __restrict__ a_edge, int64_t lde,
    const float* __restrict__ a_graph, int64_t ldag, int64_t hop_stride_ag, int hops, int N, SlabGeom g,
    int32_t* __restrict__ slab_idx, float* __restrict__ slab_f) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int i0 = b * g.npc, nn = min(g.npc, N - i0);
  int32_t* idx = slab_idx + (size_t)b * g.idx_stride;
  const int hdr = g.idx_stride - 2 * g.cap;
  const int eA = rowptr[i0], eC = rowptr[i0 + nn] - eA;
  const bool fits = eC <= g.cap;
  if (tid == 0) { idx[0] = fits ? eC : -1; idx[1] = eA; }
  for (int t = tid; t <= g.npc; t += 128) idx[2 + t] = t <= nn ? rowptr[i0 + t] - eA : eC;
  for (int t = tid; t < g.npc; t += 128) idx[2 + g.npc + 1 + t] = t < nn ? node_graph[i0 + t] : 0;
  if (!fits) return;
  for (int k = tid; k < g.cap; k += 128) idx[hdr + k] = k < eC ? col_src[eA + k] : 0;
  for (int node = 0; node < nn; ++node)           // local target node of every staged in-edge
    for (int k = rowptr[i0 + node] - eA + tid; k < rowptr[i0 + node + 1] - eA; k += 128) idx[hdr + g.cap + k] = node;
  for (int j = 0; j < hops; ++j) {
    float* f = slab_f + ((size_t)j * g.grid + b) * g.f_stride;
    for (int t = tid; t < g.npc * H; t += 128) {
      const int node = t / H, h = t - node * H;
      f[t] = (a_graph != nullptr && node < nn) ? a_graph[(size_t)j * hop_stride_ag + (size_t)node_graph[i0 + node] * ldag + h] : 0.f;
    }
    for (int t = tid; t < g.cap * H; t += 128) {
      const int k = t / H, h = t - k * H;
      float v = 0.f;
      if (k < eC) {
        const int64_t e = perm ? perm[eA + k] : (eA + k);
        v = a_edge[e * lde + (size_t)j * H + h];
      }
      f[g.npc * H + t] = v;
    }
  }
}

constexpr int kSlabWinRows = 512;      // a_node rows staged around the CTA's node range (2H floats each)

template <int J, int H, int HP>
__global__ void __launch_bounds__(kBlkThreads, 512 / kBlkThreads) gat_hop_slab_kernel(
    const HopParams p, const SlabGeom g, const int32_t* __restrict__ slab_idx, const float* __restrict__ slab_f,
    const int win) {
  extern __shared__ __align__(128) unsigned char slab_smem[];
  // layout: [idx: idx_stride ints][f: f_stride floats][a_node window: win_rows * 2H floats][mbarrier]
  int32_t* idx_s = reinterpret_cast<int32_t*>(slab_smem);
  float* f_s = reinterpret_cast<float*>(idx_s + g.idx_stride);
  const int win_rows = g.npc + 2 * win;
  float* an_s = f_s + g.f_stride;
  uint64_t* bar = reinterpret_cast<uint64_t*>(an_s + (size_t)win_rows * 2 * H);
  __shared__ float gscratch[4][kEdgeChunk * H];     // per-warp scratch of the oversize path
  __shared__ int32_t gsrc[4][kEdgeChunk];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = blockIdx.x * g.npc;
  const int nn = min(g.npc, p.N - i0);
  const int C4 = p.C >> 2;
  GVQA_HOP_TRACE(0);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // the slabs are per-batch data: under programmatic dependent launch (GVQA_PDL bit 1) with the caller's promise that
  // they are older than the predecessor kernel, their fetch starts while that kernel (the projection GEMM) drains
  if (!p.early) {
    pdl_wait();
    pdl_launch_dependents();
  }
  if (tid == 0) {
    const uint32_t bytes_i = (uint32_t)g.idx_stride * 4u, bytes_f = (uint32_t)g.f_stride * 4u;
    mbar_expect_tx(bar, bytes_i + bytes_f);
    bulk_g2s(idx_s, slab_idx + (size_t)blockIdx.x * g.idx_stride, bytes_i, bar);
    bulk_g2s(f_s, slab_f + (size_t)blockIdx.x * g.f_stride, bytes_f, bar);
  }
  if (p.early) {
    pdl_wait();              // a_node (and x_l, h_prev below) come out of the predecessor
    pdl_launch_dependents();
  }
  // the same trip: a_node rows of the window [w0, w1) (source logit terms a_l of every node a source can be, and the
  // target terms a_r of the CTA's own nodes)
  const int w0 = max(0, i0 - win), w1 = min(p.N, i0 + g.npc + win);
  for (int t = tid; t < (w1 - w0) * H; t += kBlkThreads) {        // 2H floats per row = H float2 (8-byte aligned rows)
    const int r = t / H, q = t - r * H;
    reinterpret_cast<float2*>(an_s)[t] = __ldg(reinterpret_cast<const float2*>(p.a_node + (int64_t)(w0 + r) * p.lda) + q);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  GVQA_HOP_TRACE(1);
  const int eC = idx_s[0];
  if (eC < 0) {
    // more in-edges than the slab holds: per-warp chunked path from global memory
    for (int node = wid; node < nn; node += kBlkThreads / 32)
      gather_node<J, H>(p, i0 + node, lane, 0, C4, gscratch[wid], gsrc[wid], true);
    return;
  }
  const int32_t* rp_s = idx_s + 2;
  const int32_t* gid_s = idx_s + 2 + g.npc + 1;
  const int32_t* src_s = idx_s + (g.idx_stride - 2 * g.cap);
  const int32_t* dst_s = src_s + g.cap;
  float* alpha_s = f_s + g.npc * H;
  // ---- the sources are known: request this warp's first node now (skip row, graph term, the source rows of its
  // first in-edge); the logit / softmax phase below then runs under that memory latency instead of before it ----
  // HP = heads of the first in-edge requested early (template parameter, chosen by launch_slab): limited by the
  // 128-register budget -- a spill store of such a register would wait for its load right there and serialise the
  // phase it is meant to hide
  constexpr bool kPre = HP > 0;
  float4 pre[kPre ? HP : 1][kPre ? J : 1];
  const int node0 = wid;
  const int r0_0 = node0 < nn ? rp_s[node0] : 0, r1_0 = node0 < nn ? rp_s[node0 + 1] : 0;
  if constexpr (kPre) {
    // unconditional loads from always-valid addresses (a node without in-edges or a lane beyond the row reads a
    // row / column it ignores later): no zero-initialisation, so nothing has to wait for the data before its use
    const int src0 = (node0 < nn && r1_0 > r0_0) ? src_s[r0_0] : i0;
    const float* row = p.x_l + (int64_t)src0 * p.ldx;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c4 = min(lane + 32 * j, C4 - 1);
#pragma unroll
      for (int h = 0; h < HP; ++h) pre[h][j] = ldg_issue_now(row + h * p.C + 4 * c4);
    }
  }
  // ---- logits: one thread per (in-edge, head), assembled from shared memory only ----
  for (int t = tid; t < eC * H; t += kBlkThreads) {
    const int k = t / H, h = t - k * H;
    const int src = src_s[k], node = dst_s[k];
    const float al = (src >= w0 && src < w1) ? an_s[(size_t)(src - w0) * 2 * H + h] : p.a_node[(int64_t)src * p.lda + h];
    const float tg = f_s[node * H + h] + an_s[(size_t)(i0 + node - w0) * 2 * H + H + h];
    alpha_s[t] = leaky_relu((alpha_s[t] + al) + tg, p.slope);
  }
  __syncthreads();
  // ---- softmax: one thread per (node, head) over its few logits (PyG: exp(l - max) / (sum + 1e-16)) ----
  if (tid < nn * H) {
    const int node = tid / H, h = tid - node * H;
    const int r0 = rp_s[node], r1 = rp_s[node + 1];
    float mx = -INFINITY;
#pragma unroll 4
    for (int k = r0; k < r1; ++k) mx = fmaxf(mx, alpha_s[k * H + h]);
    float sum = 0.f;
#pragma unroll 4
    for (int k = r0; k < r1; ++k) {
      const float ex = expf(alpha_s[k * H + h] - mx);
      alpha_s[k * H + h] = ex;
      sum += ex;
    }
    const float inv = 1.0f / (sum + 1e-16f);
#pragma unroll 4
    for (int k = r0; k < r1; ++k) alpha_s[k * H + h] *= inv;      // column h of these rows is this thread's alone
  }
  __syncthreads();
  GVQA_HOP_TRACE(3);
  if constexpr (kPre) {
    if (node0 < nn) {
      const int i = i0 + node0;
      stream_node_pre<J, H, kPre ? HP : 1>(p, i, lane, C4, src_s, alpha_s, r0_0, r1_0, gid_s[node0], pre, p.epilogue,
                            [&](int c4, const float4& v) { stg_stream(p.h_out + (int64_t)i * p.C + 4 * c4, v); });
    }
  }
#pragma unroll 1
  for (int node = wid + (kPre ? kBlkThreads / 32 : 0); node < nn; node += kBlkThreads / 32) {
    const int i = i0 + node;
    stream_node<J, H>(p, i, lane, C4, src_s, alpha_s, rp_s[node], rp_s[node + 1], gid_s[node], p.epilogue,
                      [&](int c4, const float4& v) { stg_stream(p.h_out + (int64_t)i * p.C + 4 * c4, v); });
  }
  if (p.trace && lane == 0 && wid < 4) p.trace[(size_t)blockIdx.x * 8 + 4 + wid] = gtime_ns();
}

template <int J, int H>
static int launch_flat(const HopParams& p, int variant, cudaStream_t stream) {
  if (variant == 1) {
    gat_hop_gather_kernel<J, H><<<(unsigned)((p.N + 7) / 8), 256, 0, stream>>>(p);
  } else {
    int npc = (p.N + kNumSMs * 4 - 1) / (kNumSMs * 4);
    npc = npc < 4 ? 4 : (npc > kBlkNodes ? kBlkNodes : npc);
    if (const char* e = getenv("GVQA_HOP_NPC")) { const int v = atoi(e); if (v >= 1 && v <= kBlkNodes) npc = v; }   // experiments
    if (launch_pdl(2, gat_hop_block_kernel<J, H>, dim3((unsigned)((p.N + npc - 1) / npc)), dim3(kBlkThreads), 0, stream, p,
                   npc) != cudaSuccess) {
      (void)cudaGetLastError();
      return GVQA_ERR_CUDA;
    }
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int J, int H>
static int launch_ws(const HopParams& p, cudaStream_t stream) {
  constexpr int kWsThreads = WsGeom<H>::kThreads;
  static const int per_sm = [] {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gat_hop_ws_kernel<J, H>, kWsThreads, 0) != cudaSuccess || nb < 1) {
      (void)cudaGetLastError();
      nb = 1;
    }
    if (const char* e = getenv("GVQA_HOP_WS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= nb) nb = v; }   // experiments
    return nb;
  }();
  const int chunk = WsGeom<H>::kChunk;
  const int nchunks = (p.N + chunk - 1) / chunk;
  const int grid = nchunks < kNumSMs * per_sm ? nchunks : kNumSMs * per_sm;
  if (launch_pdl(2, gat_hop_ws_kernel<J, H>, dim3((unsigned)grid), dim3(kWsThreads), 0, stream, p, nchunks) != cudaSuccess) {
    (void)cudaGetLastError();
    return GVQA_ERR_CUDA;
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int J, int H>
static int launch_graphln(const HopParams& p, int max_nodes_hint, cudaStream_t stream) {
  // shared-memory rows for the hinted largest graph (unknown: 64 rows), capped so that two CTAs fit per SM when
  // the graphs are small and one CTA can still stage ~200 KB
  const size_t row = (size_t)p.C * 4;
  size_t want = (size_t)(max_nodes_hint > 0 ? max_nodes_hint : 64) * row;
  const size_t cap = 200 * 1024;
  if (want > cap) want = cap - cap % row;
  const int n_cap = (int)(want / row);
  static size_t configured = 0;     // largest dynamic shared-memory size the attribute was raised to for this instance
  if (want > 48 * 1024 && want > configured) {
    if (cudaFuncSetAttribute(gat_hop_graphln_kernel<J, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap) != cudaSuccess) {
      (void)cudaGetLastError();
      return GVQA_ERR_CUDA;
    }
    configured = cap;
  }
  gat_hop_graphln_kernel<J, H><<<(unsigned)p.B, kLnHopThreads, want, stream>>>(p, n_cap);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int J, int H>
static int launch_slab(const HopParams& p, cudaStream_t stream) {
  const SlabGeom g = slab_geom(p.N, p.E, H);
  int win = p.win;
  if (g.npc + 2 * win > kSlabWinRows) win = (kSlabWinRows - g.npc) / 2;   // farther sources: global loads
  const size_t smem = (size_t)g.idx_stride * 4 + (size_t)g.f_stride * 4 + (size_t)(g.npc + 2 * win) * 2 * H * 4 + 16;
  if (smem > 48 * 1024) return GVQA_ERR_UNSUPPORTED;
  // early-issued heads: default = what fits the register budget without spills (measured, profiles/r02/); the
  // GVQA_HOP_PRE environment variable overrides it for experiments (0, 1, 2 or H)
  static const int env_pre = [] {
    const char* e = getenv("GVQA_HOP_PRE");
    return e ? atoi(e) : -1;
  }();
  constexpr int kDefault = (J * H <= 12) ? H : 0;
  const int hp = env_pre < 0 ? kDefault : env_pre;
  cudaError_t err;
#define GVQA_SLAB_LAUNCH(HPV)                                                                                       \
  err = launch_pdl(2, gat_hop_slab_kernel<J, H, HPV>, dim3((unsigned)g.grid), dim3(kBlkThreads), smem, stream, p, g, \
                   p.slab_idx, p.slab_f, win)
  if (hp >= H && J * H <= 16) GVQA_SLAB_LAUNCH(H);
  else if (hp >= 2 && H >= 2 && J * 2 <= 16) GVQA_SLAB_LAUNCH((H >= 2 ? 2 : 1));
  else if (hp >= 1 && J <= 8) GVQA_SLAB_LAUNCH(1);
  else GVQA_SLAB_LAUNCH(0);
#undef GVQA_SLAB_LAUNCH
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    return GVQA_ERR_CUDA;
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int H>
static int dispatch_flat(const HopParams& p, int variant, cudaStream_t stream, int max_nodes_hint = 0) {
  const int jj = (p.C / 4 + 31) / 32;
  if (p.epilogue == GVQA_EPI_GRAPH_LN) {
    switch (jj) {
      case 1: return launch_graphln<1, H>(p, max_nodes_hint, stream);
      case 2: return launch_graphln<2, H>(p, max_nodes_hint, stream);
      case 3: return launch_graphln<3, H>(p, max_nodes_hint, stream);
      case 4: return launch_graphln<4, H>(p, max_nodes_hint, stream);
      case 5: case 6: return launch_graphln<6, H>(p, max_nodes_hint, stream);
      case 7: case 8: return launch_graphln<8, H>(p, max_nodes_hint, stream);
      default: return GVQA_ERR_UNSUPPORTED;
    }
  }
  if (variant == 5) {
    switch (jj) {
      case 1: return launch_slab<1, H>(p, stream);
      case 2: return launch_slab<2, H>(p, stream);
      case 3: return launch_slab<3, H>(p, stream);
      case 4: return launch_slab<4, H>(p, stream);
      case 5: case 6: return launch_slab<6, H>(p, stream);
      case 7: case 8: return launch_slab<8, H>(p, stream);
      default: return GVQA_ERR_UNSUPPORTED;
    }
  }
  if (variant == 4) {
    switch (jj) {
      case 1: return launch_ws<1, H>(p, stream);
      case 2: return launch_ws<2, H>(p, stream);
      case 3: return launch_ws<3, H>(p, stream);
      case 4: return launch_ws<4, H>(p, stream);
      case 5: case 6: return launch_ws<6, H>(p, stream);
      case 7: case 8: return launch_ws<8, H>(p, stream);
      default: return GVQA_ERR_UNSUPPORTED;
    }
  }
  switch (jj) {
    case 1: return launch_flat<1, H>(p, variant, stream);
    case 2: return launch_flat<2, H>(p, variant, stream);
    case 3: return launch_flat<3, H>(p, variant, stream);
    case 4: return launch_flat<4, H>(p, variant, stream);
    case 5: case 6: return launch_flat<6, H>(p, variant, stream);
    case 7: case 8: return launch_flat<8, H>(p, variant, stream);
    default: return GVQA_ERR_UNSUPPORTED;
  }
}

// ------------------------------------------------------------------------------------------
// Kernel 2: shared-memory staged (TMA).  One CTA = one work unit (graph g, channel window w of
// W4 float4 columns).  The CTA loads the graph's CSR slice + logit terms and runs the
// per-destination softmax ONCE for all heads, while the TMA engine (cp.async.bulk + mbarrier)
// streams one stage per head -- the window of every node row of the graph for that head -- through
// a 3-deep ring of shared-memory buffers.  Each warp keeps the output accumulators of its
// destination nodes in registers across the H stages.  Units that do not fit the compiled
// capacity (n > n_cap or in-edges > e_cap) fall back to the global gather path inside the same
// kernel, so the loader hints are never a correctness input.
// ------------------------------------------------------------------------------------------
constexpr int kStagedBufs = 3;

struct StagedCfg {
  int nwin, W4, n_cap, e_cap, nbuf;
};

__host__ __device__ inline size_t staged_smem_bytes(const StagedCfg& c, int H) {
  size_t b = (size_t)c.nbuf * c.n_cap * c.W4 * 16;    // stage ring
  b += (size_t)c.e_cap * H * 4;                       // alpha
  b += (size_t)c.n_cap * H * 4;                       // target term
  b += (size_t)c.e_cap * 4;                           // sources (graph-local)
  b += (size_t)(c.n_cap + 1) * 4;                     // rowptr
  b = (b + 15) & ~(size_t)15;
  return b + 8 * kStagedBufs;                         // mbarriers
}

template <int H, int J, int NPW>
__global__ void __launch_bounds__(512) gat_hop_staged_kernel(const HopParams p, const StagedCfg cfg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t stage_floats = (size_t)cfg.n_cap * cfg.W4 * 4;
  float* stage0 = reinterpret_cast<float*>(smem_raw);
  float* alpha_s = stage0 + cfg.nbuf * stage_floats;
  float* tgt_s = alpha_s + (size_t)cfg.e_cap * H;
  int32_t* src_s = reinterpret_cast<int32_t*>(tgt_s + (size_t)cfg.n_cap * H);
  int32_t* rp_s = src_s + cfg.e_cap;
  uint64_t* full = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(rp_s + cfg.n_cap + 1) + 15) & ~(uintptr_t)15);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int g = blockIdx.x / cfg.nwin, w = blockIdx.x - g * cfg.nwin;
  const int n0 = p.graph_ptr[g], n1 = p.graph_ptr[g + 1], n = n1 - n0;
  const int C4 = p.C >> 2;
  const int c4_lo = w * cfg.W4;
  const int w4 = min(cfg.W4, C4 - c4_lo);
  if (n <= 0 || w4 <= 0) return;
  const int e0 = p.rowptr[n0], e1 = p.rowptr[n1], eg = e1 - e0;

  if (n > cfg.n_cap || eg > cfg.e_cap || n > nwarps * NPW) {
    // oversize unit: global gather path, scratch carved from the (unused) stage ring
    float* alpha_w = stage0 + wid * (kEdgeChunk * H + kEdgeChunk);
    int32_t* src_w = reinterpret_cast<int32_t*>(alpha_w + kEdgeChunk * H);
    for (int i = n0 + wid; i < n1; i += nwarps) gather_node<J, H>(p, i, lane, c4_lo, w4, alpha_w, src_w, w == 0);
    return;
  }

  // ---- 1. barriers, then kick off the first stages (one stage = one head) -------------------
  if (tid == 0) {
    for (int b = 0; b < cfg.nbuf; ++b) mbar_init(&full[b], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const float* win_base = p.x_l + (int64_t)n0 * p.ldx + 4 * c4_lo;
  auto issue_stage = [&](int q) {  // called by warp 0 only
    const int b = q % cfg.nbuf;
    if (lane == 0) mbar_expect_tx(&full[b], (uint32_t)(n * w4 * 16));
    __syncwarp();
    float* dst = stage0 + b * stage_floats;
    for (int r = lane; r < n; r += 32)
      bulk_g2s(dst + (size_t)r * cfg.W4 * 4, win_base + (int64_t)r * p.ldx + q * p.C, (uint32_t)(w4 * 16), &full[b]);
  };
  if (wid == 0)
    for (int q = 0; q < cfg.nbuf && q < H; ++q) issue_stage(q);

  // ---- 2. CSR slice + per-node target terms (one round trip) --------------------------------
  for (int i = tid; i <= n; i += nthreads) rp_s[i] = p.rowptr[n0 + i] - e0;
  for (int t = tid; t < n * H; t += nthreads) {
    const int node = t / H, h = t - node * H;
    float v = p.a_node[(int64_t)(n0 + node) * p.lda + H + h];
    if (p.a_graph) v += p.a_graph[(int64_t)g * p.ldag + h];
    tgt_s[t] = v;
  }
  for (int k = tid; k < eg; k += nthreads) src_s[k] = p.col_src[e0 + k] - n0;
  __syncthreads();

  // ---- 3. source + edge logit terms, edge-parallel (second round trip) ----------------------
  for (int t = tid; t < eg * H; t += nthreads) {
    const int k = t / H, h = t - k * H;
    const int64_t e = p.perm ? p.perm[e0 + k] : (e0 + k);
    alpha_s[t] = p.a_node[(int64_t)(src_s[k] + n0) * p.lda + h] + p.a_edge[e * p.lde + h];
  }
  __syncthreads();

  // ---- 4. per-destination softmax in shared memory (warp per node, all heads at once) -------
  constexpr int per = 32 / H;
  const int head = lane % H, slot = lane / H;
  for (int node = wid; node < n; node += nwarps) {
    const int r0 = rp_s[node], r1 = rp_s[node + 1];
    const float tg = tgt_s[node * H + head];
    float mx = -INFINITY;
    for (int k = r0 + slot; k < r1; k += per) {
      const float l = leaky_relu(alpha_s[k * H + head] + tg, p.slope);
      alpha_s[k * H + head] = l;
      mx = fmaxf(mx, l);
    }
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    float sum = 0.f;
    for (int k = r0 + slot; k < r1; k += per) {
      const float ex = expf(alpha_s[k * H + head] - mx);
      alpha_s[k * H + head] = ex;
      sum += ex;
    }
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    const float inv = 1.0f / (sum + 1e-16f);
    for (int k = r0 + slot; k < r1; k += per) {
      const float a = alpha_s[k * H + head] * inv;
      alpha_s[k * H + head] = a;
      if (p.alpha_out && w == 0) {
        const int64_t e = p.perm ? p.perm[e0 + k] : (e0 + k);
        p.alpha_out[e * H + head] = a;
      }
    }
  }
  __syncwarp();  // a warp consumes only the alpha rows it produced (same node -> warp mapping below)

  // ---- 5. stream the H stages; accumulators of this warp's nodes stay in registers ----------
  float4 acc[NPW][J];
#pragma unroll
  for (int t = 0; t < NPW; ++t)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[t][j] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
  for (int q = 0; q < H; ++q) {
    const int b = q % cfg.nbuf;
    mbar_wait(&full[b], (uint32_t)((q / cfg.nbuf) & 1));
    const float* st = stage0 + b * stage_floats;
#pragma unroll
    for (int t = 0; t < NPW; ++t) {
      const int node = wid + t * nwarps;
      if (node < n) {
        const int r0 = rp_s[node], r1 = rp_s[node + 1];
#pragma unroll 2
        for (int k = r0; k < r1; ++k) {
          const int src = src_s[k];
          const float a = alpha_s[k * H + q];
          if (src >= 0 && src < n) {
            const float* row = st + (size_t)src * cfg.W4 * 4;
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const int c4 = lane + 32 * j;
              if (c4 < w4) fma4(acc[t][j], a, *reinterpret_cast<const float4*>(row + 4 * c4));
            }
          } else {  // source outside this graph (flagged by gvqa_build_csr): read it from global
            const float* row = p.x_l + (int64_t)(src + n0) * p.ldx + q * p.C + 4 * c4_lo;
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const int c4 = lane + 32 * j;
              if (c4 < w4) fma4(acc[t][j], a, ldg_cached(row + 4 * c4));
            }
          }
        }
      }
    }
    if (q + cfg.nbuf < H) {      // refill this buffer with stage q + nbuf once every warp is done with it
      __syncthreads();
      if (wid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_stage(q + cfg.nbuf);
      }
    }
  }

  // ---- 6. epilogue: head mean, +graph term, +bias, +skip, BatchNorm(eval)+ReLU ---------------
#pragma unroll
  for (int t = 0; t < NPW; ++t) {
    const int node = wid + t * nwarps;
    if (node < n) {
      const bool has_in = rp_s[node + 1] > rp_s[node];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < w4) {
          float4 skip = make_float4(0.f, 0.f, 0.f, 0.f), gb = skip;
          if (p.h_prev) skip = ldg_stream(p.h_prev + (int64_t)(n0 + node) * p.C + 4 * (c4_lo + c4));
          if (p.graph_bias && has_in) gb = ldg_cached(p.graph_bias + (int64_t)g * p.ldgb + 4 * (c4_lo + c4));
          epilogue_store4(p, n0 + node, c4_lo + c4, acc[t][j], 1.0f / H, p.graph_bias != nullptr && has_in, gb,
                          p.h_prev != nullptr, skip);
        }
      }
    }
  }
}

struct StagedPlan {
  StagedCfg cfg;
  size_t smem;
  int J, NPW, threads;
};

// Window geometry for the staged kernel; returns false when the graph slab cannot be staged
// (the caller then uses the block kernel).
static bool plan_staged(int C, int H, int max_nodes, int max_in_edges, StagedPlan* out) {
  if (max_nodes <= 0) return false;
  const int C4 = C / 4;
  StagedPlan pl;
  StagedCfg& c = pl.cfg;
  c.n_cap = max_nodes;
  c.e_cap = max_in_edges > 0 ? max_in_edges : 8 * max_nodes;
  if (c.e_cap < 64) c.e_cap = 64;
  c.nbuf = H < kStagedBufs ? H : kStagedBufs;
  pl.threads = 512;  // node capacity: 16 warps x NPW nodes
  if (c.n_cap <= 32) pl.NPW = 2;
  else if (c.n_cap <= 64) pl.NPW = 4;
  else return false;
  const size_t budget = 112 * 1024;  // >= 2 CTAs per SM so units overlap each other's set-up
  for (c.nwin = 1; c.nwin <= C4; ++c.nwin) {
    c.W4 = (C4 + c.nwin - 1) / c.nwin;
    pl.J = (c.W4 + 31) / 32;
    if (pl.J > 2) continue;
    if (c.W4 < 32 && c.nwin > 1) return false;  // narrow windows: tiny TMA segments
    // the gather fallback needs nwarps x (32*H floats + 32 ints) of scratch inside the ring
    if ((size_t)c.nbuf * c.n_cap * c.W4 * 16 < (size_t)16 * (kEdgeChunk * H + kEdgeChunk) * 4) return false;
    pl.smem = staged_smem_bytes(c, H);
    if (pl.smem <= budget) {
      *out = pl;
      return true;
    }
  }
  return false;
}

template <int H, int J, int NPW>
static int launch_staged3(const HopParams& p, const StagedPlan& pl, cudaStream_t stream) {
  if (cudaFuncSetAttribute(gat_hop_staged_kernel<H, J, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)pl.smem) != cudaSuccess)
    return GVQA_ERR_CUDA;
  const unsigned grid = (unsigned)((int64_t)p.B * pl.cfg.nwin);
  gat_hop_staged_kernel<H, J, NPW><<<grid, pl.threads, pl.smem, stream>>>(p, pl.cfg);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int H>
static int launch_staged(const HopParams& p, const StagedPlan& pl, cudaStream_t stream) {
  if (pl.J == 1 && pl.NPW == 2) return launch_staged3<H, 1, 2>(p, pl, stream);
  if (pl.J == 1 && pl.NPW == 4) return launch_staged3<H, 1, 4>(p, pl, stream);
  if (pl.J == 2 && pl.NPW == 2) return launch_staged3<H, 2, 2>(p, pl, stream);
  if (pl.J == 2 && pl.NPW == 4) return launch_staged3<H, 2, 4>(p, pl, stream);
  return GVQA_ERR_UNSUPPORTED;
}

}  // namespace gvqa

static unsigned long long* g_hop_trace = nullptr;   // debug only
extern "C" GVQA_API void gvqa_debug_set_hop_trace(unsigned long long* device_buffer) { g_hop_trace = device_buffer; }

extern "C" GVQA_API int gvqa_gat_hop_f32(const gvqa_gat_hop_args* a, void* stream_) {
  using namespace gvqa;
  if (!a) return GVQA_ERR_NULL_POINTER;
  const int H = a->heads, C = a->channels;
  if (a->num_nodes < 0 || a->num_edges < 0 || a->num_graphs < 0 || H <= 0 || C <= 0) return GVQA_ERR_BAD_SHAPE;
  if (a->num_nodes >= (1ll << 31) || a->num_edges >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (a->ldx < (int64_t)H * C || a->lde < H || (a->ld_a_node != 0 && a->ld_a_node < 2 * H)) return GVQA_ERR_BAD_SHAPE;
  if ((a->ld_graph_bias != 0 && a->ld_graph_bias < C) || (a->ld_a_graph != 0 && a->ld_a_graph < H)) return GVQA_ERR_BAD_SHAPE;
  if (a->num_nodes == 0) return GVQA_OK;
  if (!a->x_l || !a->a_node || !a->rowptr || !a->node_graph || !a->h_out) return GVQA_ERR_NULL_POINTER;
  if (a->num_edges > 0 && (!a->a_edge || !a->col_src)) return GVQA_ERR_NULL_POINTER;
  if (a->epilogue < GVQA_EPI_NONE || a->epilogue > GVQA_EPI_GRAPH_LN) return GVQA_ERR_UNSUPPORTED;
  if ((a->epilogue == GVQA_EPI_AFFINE || a->epilogue == GVQA_EPI_AFFINE_RELU) && (!a->ep_scale || !a->ep_shift))
    return GVQA_ERR_NULL_POINTER;
  if (a->epilogue == GVQA_EPI_GRAPH_LN && (!a->graph_ptr || a->num_graphs <= 0)) return GVQA_ERR_NULL_POINTER;
  if (a->variant == 4 && !a->sched) return GVQA_ERR_NULL_POINTER;
  if ((C & 3) || C > 1024 || !(H == 1 || H == 2 || H == 4 || H == 8)) return GVQA_ERR_UNSUPPORTED;
  if (a->variant < 0 || a->variant > 5) return GVQA_ERR_UNSUPPORTED;
  if (a->variant == 5 && (!a->slab_idx || !a->slab_f)) return GVQA_ERR_NULL_POINTER;
  if ((a->slab_idx && (reinterpret_cast<uintptr_t>(a->slab_idx) & 15)) || (a->slab_f && !aligned16(a->slab_f))) return GVQA_ERR_MISALIGNED;
  if ((a->ldx & 3) || (a->ld_graph_bias & 3) || !aligned16(a->x_l) || !aligned16(a->h_out) || (a->h_prev && !aligned16(a->h_prev)) ||
      (a->graph_bias && !aligned16(a->graph_bias)) || (a->bias && !aligned16(a->bias)) ||
      (a->ep_scale && !aligned16(a->ep_scale)) || (a->ep_shift && !aligned16(a->ep_shift)))
    return GVQA_ERR_MISALIGNED;

  HopParams p;
  p.trace = g_hop_trace;
  p.x_l = a->x_l; p.graph_bias = a->graph_bias; p.a_node = a->a_node; p.a_graph = a->a_graph; p.a_edge = a->a_edge;
  p.rowptr = a->rowptr; p.col_src = a->col_src; p.perm = a->perm; p.graph_ptr = a->graph_ptr;
  p.node_graph = a->node_graph; p.h_prev = a->h_prev; p.bias = a->bias; p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift; p.h_out = a->h_out; p.alpha_out = a->alpha_out;
  p.ldx = a->ldx; p.lde = a->lde; p.lda = a->ld_a_node > 0 ? a->ld_a_node : 2 * H;
  p.ldgb = a->ld_graph_bias > 0 ? a->ld_graph_bias : C; p.ldag = a->ld_a_graph > 0 ? a->ld_a_graph : H;
  p.N = (int32_t)a->num_nodes; p.E = (int32_t)a->num_edges; p.B = (int32_t)a->num_graphs; p.C = C;
  p.slope = a->negative_slope; p.epilogue = a->epilogue;
  // the early-start layout only pays when the launch carries the PDL attribute (off by default, common.cuh)
  p.early = ((a->flags & GVQA_HOP_INPUTS_OLDER_THAN_PREDECESSOR) && (pdl_mask() & 2)) ? 1 : 0;
  static const int env_prefetch = [] {
    const char* e = getenv("GVQA_HOP_PREFETCH");
    return e ? atoi(e) : 0;
  }();
  p.prefetch = env_prefetch;
  p.sched = a->sched; p.ln_weight = a->ln_weight; p.ln_bias = a->ln_bias; p.ln_eps = a->ln_eps;
  p.slab_idx = a->slab_idx; p.slab_f = a->slab_f; p.win = a->max_nodes_per_graph > 0 ? a->max_nodes_per_graph : 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  if (a->variant == 2 && a->epilogue != GVQA_EPI_GRAPH_LN) {
    StagedPlan plan;
    if (!a->graph_ptr || a->num_graphs <= 0 ||
        !plan_staged(C, H, a->max_nodes_per_graph, a->max_in_edges_per_graph, &plan))
      return GVQA_ERR_UNSUPPORTED;
    switch (H) {
      case 1: return launch_staged<1>(p, plan, stream);
      case 2: return launch_staged<2>(p, plan, stream);
      case 4: return launch_staged<4>(p, plan, stream);
      case 8: return launch_staged<8>(p, plan, stream);
    }
  }
  // auto: the slab kernel when the caller prepared slabs and does not ask for the attention weights (the slabs
  // carry no edge ids), else the block kernel
  int variant = a->variant == 1 ? 1 : (a->variant == 4 ? 4 : 3);
  if ((a->variant == 0 || a->variant == 5) && a->slab_idx && a->slab_f && !a->alpha_out && (a->ld_a_node == 0 || (a->ld_a_node & 1) == 0))
    variant = 5;
  else if (a->variant == 5)
    return GVQA_ERR_UNSUPPORTED;
  switch (H) {
    case 1: return dispatch_flat<1>(p, variant, stream, a->max_nodes_per_graph);
    case 2: return dispatch_flat<2>(p, variant, stream, a->max_nodes_per_graph);
    case 4: return dispatch_flat<4>(p, variant, stream, a->max_nodes_per_graph);
    case 8: return dispatch_flat<8>(p, variant, stream, a->max_nodes_per_graph);
  }
  return GVQA_ERR_UNSUPPORTED;
}


extern "C" GVQA_API int gvqa_gat_hop_slab_plan(int64_t num_nodes, int64_t num_edges, int32_t heads, gvqa_gat_slab_plan* plan) {
  if (!plan) return GVQA_ERR_NULL_POINTER;
  if (num_nodes <= 0 || num_edges < 0 || num_nodes >= (1ll << 31) || num_edges >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (!(heads == 1 || heads == 2 || heads == 4 || heads == 8)) return GVQA_ERR_UNSUPPORTED;
  const gvqa::SlabGeom g = gvqa::slab_geom(num_nodes, num_edges, heads);
  plan->nodes_per_cta = g.npc;
  plan->edge_capacity = g.cap;
  plan->num_ctas = g.grid;
  plan->idx_words = (int64_t)g.grid * g.idx_stride;
  plan->f_words_per_hop = (int64_t)g.grid * g.f_stride;
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_hop_build_slabs_f32(const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                                     const int32_t* node_graph, const float* a_edge, int64_t lde,
                                                     const float* a_graph, int64_t ld_a_graph, int64_t hop_stride_a_graph,
                                                     int32_t hops, int64_t num_nodes, int64_t num_edges, int32_t heads,
                                                     int32_t* slab_idx, float* slab_f, void* stream_) {
  using namespace gvqa;
  if (num_nodes < 0 || num_edges < 0 || hops <= 0 || num_nodes >= (1ll << 31) || num_edges >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!rowptr || !node_graph || !slab_idx || !slab_f || (num_edges > 0 && (!col_src || !a_edge))) return GVQA_ERR_NULL_POINTER;
  if (lde < (int64_t)hops * heads || (a_graph && ld_a_graph < heads)) return GVQA_ERR_BAD_SHAPE;
  if (!(heads == 1 || heads == 2 || heads == 4 || heads == 8)) return GVQA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(slab_idx) & 15) || !aligned16(slab_f)) return GVQA_ERR_MISALIGNED;
  const SlabGeom g = slab_geom(num_nodes, num_edges, heads);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
#define GVQA_SLAB(HH)                                                                                                \
  gat_slab_build_kernel<HH><<<(unsigned)g.grid, 128, 0, stream>>>(rowptr, col_src, perm, node_graph, a_edge, lde, a_graph, \
                                                                  ld_a_graph, hop_stride_a_graph, hops, (int)num_nodes, g, \
                                                                  slab_idx, slab_f)
  switch (heads) {
    case 1: GVQA_SLAB(1); break;
    case 2: GVQA_SLAB(2); break;
    case 4: GVQA_SLAB(4); break;
    default: GVQA_SLAB(8); break;
  }
#undef GVQA_SLAB
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
