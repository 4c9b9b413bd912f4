// Fused GAT hop for batched disjoint scene graphs (see include/gvqa_b200.h: gvqa_gat_hop_f32).
//
// One launch = one hop of the reference's gat_seq.forward (gat_skip.py:254-276) after the node
// projection: gather of source rows, additive edge logits, LeakyReLU, per-destination softmax
// (PyG semantics: exp(l-max)/(sum+1e-16)), weighted aggregation, head mean, +bias, +skip,
// BatchNorm(eval)+ReLU epilogue.  No atomics: destination-CSR, one warp owns one output row,
// in-edges are summed in the caller's edge order -> bitwise deterministic.
//
// Two kernels:
//  * gat_hop_gather_kernel  -- source rows are gathered straight from global memory (L2);
//    works for any graph size.
//  * gat_hop_staged_kernel  -- one CTA owns a (graph, channel-slice) work unit: the slice of
//    every node row of the graph (all heads) is staged ONCE into shared memory by the TMA engine
//    (cp.async.bulk + mbarrier, double buffered across units), so each x_l byte crosses
//    L2->SM exactly once; the gathers then hit shared memory.
#include "common.cuh"

namespace gvqa {

constexpr int kEdgeChunk = 32;  // in-edges whose alpha are staged per warp at a time

struct HopParams {
  const float* __restrict__ x_l;
  const float* __restrict__ x_graph;
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
  int64_t ldx, lde;
  int32_t N, E, B, C;
  float slope;
  int32_t epilogue;
};

// Softmax weights of the in-edges [e0,e1) of node `i` for all H heads, computed by one warp.
// Lane l serves head (l % H); the 32/H lanes of a head stride over the edges.  On return the
// per-head (max, 1/(sum+1e-16), sum/(sum+1e-16)) live in every lane of that head.
template <int H>
struct WarpSoftmax {
  float m, inv, total;
  const HopParams& p;
  int i, g, e0, e1, lane, head, slot;
  float target_term;

  __device__ __forceinline__ WarpSoftmax(const HopParams& p_, int i_, int g_, int e0_, int e1_, int lane_)
      : p(p_), i(i_), g(g_), e0(e0_), e1(e1_), lane(lane_) {
    head = lane % H;
    slot = lane / H;
    target_term = p.a_node[(int64_t)i * 2 * H + H + head];
    if (p.a_graph) target_term += p.a_graph[(int64_t)g * H + head];
  }

  __device__ __forceinline__ float logit(int k) const {
    const int src = p.col_src[k];
    const int64_t e = p.perm ? p.perm[k] : k;
    const float v = p.a_node[(int64_t)src * 2 * H + head] + target_term + p.a_edge[e * p.lde + head];
    return leaky_relu(v, p.slope);
  }

  __device__ __forceinline__ float reduce_max(float v) const {
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
  }
  __device__ __forceinline__ float reduce_sum(float v) const {
#pragma unroll
    for (int o = 16; o >= H; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
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
    m = reduce_max(mx);
    float s = 0.f;
    if (e0 + slot < e1) s = expf(first_logit - m);
    for (int k = e0 + slot + per; k < e1; k += per) s += expf(logit(k) - m);
    s = reduce_sum(s);
    inv = 1.0f / (s + 1e-16f);
    total = s * inv;
  }
};

// Epilogue for one float4 column (absolute float4 index c4) of output row i:
// head mean, +bias, +skip, BatchNorm(eval) affine, ReLU; 128-bit streaming store.
__device__ __forceinline__ void epilogue_store4(const HopParams& p, int i, int c4, float4 o, float inv_heads,
                                                bool have_skip, const float4& skip) {
  o.x *= inv_heads; o.y *= inv_heads; o.z *= inv_heads; o.w *= inv_heads;
  if (p.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + c4);
    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
  }
  if (have_skip) { o.x += skip.x; o.y += skip.y; o.z += skip.z; o.w += skip.w; }
  if (p.epilogue != GVQA_EPI_NONE) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.ep_scale) + c4);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.ep_shift) + c4);
    o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y);
    o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
    if (p.epilogue == GVQA_EPI_AFFINE_RELU) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
  }
  stg_stream(p.h_out + (int64_t)i * p.C + 4 * c4, o);
}

// One warp computes output row i for the float4 column window [c4_lo, c4_lo + c4_n), gathering
// source rows from global memory.  alpha_w / src_w: per-warp shared scratch (32*H floats, 32 ints).
template <int J, int H>
__device__ __forceinline__ void gather_node(const HopParams& p, int i, int lane, int c4_lo, int c4_n,
                                            float* alpha_w, int32_t* src_w, bool write_alpha) {
  const int e0 = p.rowptr[i], e1 = p.rowptr[i + 1];
  const int g = p.node_graph[i];
  constexpr int per = 32 / H;

  float4 acc[J], skip[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    skip[j] = acc[j];
    const int c4 = lane + 32 * j;
    if (p.h_prev && c4 < c4_n) skip[j] = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * (c4_lo + c4));
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
    if (p.x_graph) {
      // sum_k alpha[k,h] * x_graph[g,h,:] = (sum_k alpha[k,h]) * x_graph[g,h,:]
      const float* row = p.x_graph + (int64_t)g * H * p.C + 4 * c4_lo;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float t = __shfl_sync(kFull, sm.total, h);  // lane h serves head h
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < c4_n) fma4(acc[j], t, ldg_cached(row + h * p.C + 4 * c4));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < c4_n) epilogue_store4(p, i, c4_lo + c4, acc[j], 1.0f / H, p.h_prev != nullptr, skip[j]);
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

template <int J, int H>
static int launch_gather(const HopParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)((p.N + 7) / 8);
  gat_hop_gather_kernel<J, H><<<grid, 256, 0, stream>>>(p);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

template <int H>
static int dispatch_gather(const HopParams& p, cudaStream_t stream) {
  const int j = (p.C / 4 + 31) / 32;
  switch (j) {
    case 1: return launch_gather<1, H>(p, stream);
    case 2: return launch_gather<2, H>(p, stream);
    case 3: return launch_gather<3, H>(p, stream);
    case 4: return launch_gather<4, H>(p, stream);
    case 5: case 6: return launch_gather<6, H>(p, stream);
    case 7: case 8: return launch_gather<8, H>(p, stream);
    default: return GVQA_ERR_UNSUPPORTED;
  }
}

// ------------------------------------------------------------------------------------------
// Kernel 2: shared-memory staged.  One CTA = one work unit (graph g, channel slice s of CW4
// float4 columns).  The slice of every node row of the graph, all H heads, is copied ONCE into
// shared memory by the TMA engine (cp.async.bulk, one 16*cw4-byte row segment per (node, head),
// mbarrier completion) while the CTA loads the graph's CSR slice and logit terms and runs the
// per-destination softmax; the weighted gathers then read shared memory only.  ~63 KB per CTA at
// n=30, H=4, CW4=32 -> 3 CTAs per SM overlap each other's loads.
// Units that do not fit the compiled capacity (n > n_cap or in-edges > e_cap) fall back to the
// global gather path inside the same kernel, so the loader hints are never a correctness input.
// ------------------------------------------------------------------------------------------
struct StagedCfg {
  int S, CW4, n_cap, e_cap;
};

__host__ __device__ inline size_t staged_smem_bytes(const StagedCfg& c, int H) {
  size_t b = (size_t)c.n_cap * H * c.CW4 * 16;        // slab
  b += (size_t)c.e_cap * H * 4;                       // alpha
  b += (size_t)c.n_cap * H * 4;                       // target term / alpha row sums
  b += (size_t)c.e_cap * 4;                           // sources
  b += (size_t)(c.n_cap + 1) * 4;                     // rowptr
  b = (b + 15) & ~(size_t)15;
  return b + 16;                                      // mbarrier
}

template <int H>
__global__ void __launch_bounds__(256) gat_hop_staged_kernel(const HopParams p, const StagedCfg cfg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  float* alpha_s = slab + (size_t)cfg.n_cap * H * cfg.CW4 * 4;
  float* tgt_s = alpha_s + (size_t)cfg.e_cap * H;
  int32_t* src_s = reinterpret_cast<int32_t*>(tgt_s + (size_t)cfg.n_cap * H);
  int32_t* rp_s = src_s + cfg.e_cap;
  uint64_t* bar = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(rp_s + cfg.n_cap + 1) + 15) & ~(uintptr_t)15);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = blockIdx.x / cfg.S, s = blockIdx.x - g * cfg.S;
  const int n0 = p.graph_ptr[g], n1 = p.graph_ptr[g + 1], n = n1 - n0;
  const int C4 = p.C >> 2;
  const int c4_lo = s * cfg.CW4;
  const int cw4 = min(cfg.CW4, C4 - c4_lo);
  if (n <= 0 || cw4 <= 0) return;
  const int e0 = p.rowptr[n0], e1 = p.rowptr[n1], eg = e1 - e0;

  if (n > cfg.n_cap || eg > cfg.e_cap) {
    // oversize unit: global gather path, scratch carved from the (unused) slab
    float* alpha_w = slab + wid * (kEdgeChunk * H + kEdgeChunk);
    int32_t* src_w = reinterpret_cast<int32_t*>(alpha_w + kEdgeChunk * H);
    for (int i = n0 + wid; i < n1; i += 8) gather_node<1, H>(p, i, lane, c4_lo, cw4, alpha_w, src_w, s == 0);
    return;
  }

  // ---- 1. kick off the slab copy (TMA engine) ----------------------------------------------
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) mbar_expect_tx(bar, (uint32_t)(n * H * cw4 * 16));
  for (int r = tid; r < n * H; r += 256) {
    const int node = r / H, h = r - node * H;
    bulk_g2s(slab + (size_t)r * cfg.CW4 * 4, p.x_l + (int64_t)(n0 + node) * p.ldx + h * p.C + 4 * c4_lo,
             (uint32_t)(cw4 * 16), bar);
  }

  // ---- 2. CSR slice + per-node target terms (one round trip) --------------------------------
  for (int i = tid; i <= n; i += 256) rp_s[i] = p.rowptr[n0 + i] - e0;
  for (int t = tid; t < n * H; t += 256) {
    const int node = t / H, h = t - node * H;
    float v = p.a_node[(int64_t)(n0 + node) * 2 * H + H + h];
    if (p.a_graph) v += p.a_graph[(int64_t)g * H + h];
    tgt_s[t] = v;
  }
  for (int k = tid; k < eg; k += 256) src_s[k] = p.col_src[e0 + k];
  // per-lane constants of this slice
  float4 xg[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    xg[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.x_graph && lane < cw4) xg[h] = ldg_cached(p.x_graph + ((int64_t)g * H + h) * p.C + 4 * (c4_lo + lane));
  }
  __syncthreads();

  // ---- 3. source + edge logit terms, edge-parallel (second round trip) ----------------------
  for (int t = tid; t < eg * H; t += 256) {
    const int k = t / H, h = t - k * H;
    const int64_t e = p.perm ? p.perm[e0 + k] : (e0 + k);
    alpha_s[t] = p.a_node[(int64_t)src_s[k] * 2 * H + h] + p.a_edge[e * p.lde + h];
  }
  __syncthreads();

  // ---- 4. per-destination softmax in shared memory (warp per node) --------------------------
  constexpr int per = 32 / H;
  const int head = lane % H, slot = lane / H;
  for (int node = wid; node < n; node += 8) {
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
      if (p.alpha_out && s == 0) {
        const int64_t e = p.perm ? p.perm[e0 + k] : (e0 + k);
        p.alpha_out[e * H + head] = a;
      }
    }
    __syncwarp();
    if (slot == 0) tgt_s[node * H + head] = r1 > r0 ? sum * inv : 0.f;  // row sum of alpha (x_graph term)
  }
  __syncwarp();

  // ---- 5. weighted gather from the staged slab + epilogue -----------------------------------
  mbar_wait(bar, 0);
  const bool active = lane < cw4;
  for (int node = wid; node < n; node += 8) {
    const int i = n0 + node;
    const int r0 = rp_s[node], r1 = rp_s[node + 1];
    float4 skip = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.h_prev && active) skip = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * (c4_lo + lane));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
#pragma unroll 2
      for (int k = r0; k < r1; ++k) {
        const int src = src_s[k] - n0;
        if (src >= 0 && src < n) {
          const float* row = slab + ((size_t)src * H * cfg.CW4 + lane) * 4;
#pragma unroll
          for (int h = 0; h < H; ++h)
            fma4(acc, alpha_s[k * H + h], *reinterpret_cast<const float4*>(row + (size_t)h * cfg.CW4 * 4));
        } else {  // source outside this graph (flagged by gvqa_build_csr): read it from global
          const float* row = p.x_l + (int64_t)src_s[k] * p.ldx + 4 * (c4_lo + lane);
#pragma unroll
          for (int h = 0; h < H; ++h) fma4(acc, alpha_s[k * H + h], ldg_cached(row + h * p.C));
        }
      }
      if (p.x_graph) {
#pragma unroll
        for (int h = 0; h < H; ++h) fma4(acc, tgt_s[node * H + h], xg[h]);
      }
      epilogue_store4(p, i, c4_lo + lane, acc, 1.0f / H, p.h_prev != nullptr, skip);
    }
  }
}

// Slice geometry for the staged kernel; returns false when the graph slab cannot be staged.
static bool plan_staged(int C, int H, int max_nodes, int max_in_edges, StagedCfg* out, size_t* smem) {
  if (max_nodes <= 0) return false;
  const int C4 = C / 4;
  StagedCfg c;
  c.S = (C4 + 31) / 32;
  c.CW4 = (C4 + c.S - 1) / c.S;
  c.n_cap = max_nodes;
  c.e_cap = max_in_edges > 0 ? max_in_edges : 8 * max_nodes;
  if (c.e_cap < 64) c.e_cap = 64;
  // the gather fallback needs 8 x (32*H floats + 32 ints) of scratch inside the slab
  while ((size_t)c.n_cap * H * c.CW4 * 16 < (size_t)8 * (kEdgeChunk * H + kEdgeChunk) * 4) ++c.n_cap;
  const size_t b = staged_smem_bytes(c, H);
  if (b > 112 * 1024) return false;  // keep >= 2 CTAs per SM so units overlap each other's loads
  *out = c;
  *smem = b;
  return true;
}

template <int H>
static int launch_staged(const HopParams& p, const StagedCfg& cfg, size_t smem, cudaStream_t stream) {
  if (cudaFuncSetAttribute(gat_hop_staged_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return GVQA_ERR_CUDA;
  const unsigned grid = (unsigned)((int64_t)p.B * cfg.S);
  gat_hop_staged_kernel<H><<<grid, 256, smem, stream>>>(p, cfg);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

}  // namespace gvqa

extern "C" GVQA_API int gvqa_gat_hop_f32(const gvqa_gat_hop_args* a, void* stream_) {
  using namespace gvqa;
  if (!a) return GVQA_ERR_NULL_POINTER;
  const int H = a->heads, C = a->channels;
  if (a->num_nodes < 0 || a->num_edges < 0 || a->num_graphs < 0 || H <= 0 || C <= 0) return GVQA_ERR_BAD_SHAPE;
  if (a->num_nodes >= (1ll << 31) || a->num_edges >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (a->ldx < (int64_t)H * C || a->lde < H) return GVQA_ERR_BAD_SHAPE;
  if (a->num_nodes == 0) return GVQA_OK;
  if (!a->x_l || !a->a_node || !a->rowptr || !a->node_graph || !a->h_out) return GVQA_ERR_NULL_POINTER;
  if (a->num_edges > 0 && (!a->a_edge || !a->col_src)) return GVQA_ERR_NULL_POINTER;
  if (a->epilogue != GVQA_EPI_NONE && (!a->ep_scale || !a->ep_shift)) return GVQA_ERR_NULL_POINTER;
  if (a->epilogue < GVQA_EPI_NONE || a->epilogue > GVQA_EPI_AFFINE_RELU) return GVQA_ERR_UNSUPPORTED;
  if ((C & 3) || C > 1024 || !(H == 1 || H == 2 || H == 4 || H == 8)) return GVQA_ERR_UNSUPPORTED;
  if ((a->ldx & 3) || !aligned16(a->x_l) || !aligned16(a->h_out) || (a->h_prev && !aligned16(a->h_prev)) ||
      (a->x_graph && !aligned16(a->x_graph)) || (a->bias && !aligned16(a->bias)) ||
      (a->ep_scale && !aligned16(a->ep_scale)) || (a->ep_shift && !aligned16(a->ep_shift)))
    return GVQA_ERR_MISALIGNED;

  HopParams p;
  p.x_l = a->x_l; p.x_graph = a->x_graph; p.a_node = a->a_node; p.a_graph = a->a_graph; p.a_edge = a->a_edge;
  p.rowptr = a->rowptr; p.col_src = a->col_src; p.perm = a->perm; p.graph_ptr = a->graph_ptr;
  p.node_graph = a->node_graph; p.h_prev = a->h_prev; p.bias = a->bias; p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift; p.h_out = a->h_out; p.alpha_out = a->alpha_out;
  p.ldx = a->ldx; p.lde = a->lde;
  p.N = (int32_t)a->num_nodes; p.E = (int32_t)a->num_edges; p.B = (int32_t)a->num_graphs; p.C = C;
  p.slope = a->negative_slope; p.epilogue = a->epilogue;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  StagedCfg cfg;
  size_t smem = 0;
  bool staged = false;
  if (a->variant != 1 && a->graph_ptr && a->num_graphs > 0)
    staged = plan_staged(C, H, a->max_nodes_per_graph, a->max_in_edges_per_graph, &cfg, &smem);
  if (a->variant == 2 && !staged) return GVQA_ERR_UNSUPPORTED;
  if (staged) {
    switch (H) {
      case 1: return launch_staged<1>(p, cfg, smem, stream);
      case 2: return launch_staged<2>(p, cfg, smem, stream);
      case 4: return launch_staged<4>(p, cfg, smem, stream);
      case 8: return launch_staged<8>(p, cfg, smem, stream);
    }
  }
  switch (H) {
    case 1: return dispatch_gather<1>(p, stream);
    case 2: return dispatch_gather<2>(p, stream);
    case 4: return dispatch_gather<4>(p, stream);
    case 8: return dispatch_gather<8>(p, stream);
  }
  return GVQA_ERR_UNSUPPORTED;
}
