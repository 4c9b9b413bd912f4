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

template <int J>
__device__ __forceinline__ void epilogue_store(const HopParams& p, int i, int lane, float4 (&acc)[J], float inv_heads) {
  const int C4 = p.C >> 2;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      float4 o = acc[j];
      o.x *= inv_heads; o.y *= inv_heads; o.z *= inv_heads; o.w *= inv_heads;
      if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + c4);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (p.h_prev) {
        const float4 r = ldg_stream(p.h_prev + (int64_t)i * p.C + 4 * c4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
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
  const int e0 = p.rowptr[i], e1 = p.rowptr[i + 1];
  const int g = p.node_graph[i];
  const int C4 = p.C >> 2;
  constexpr int per = 32 / H;

  float4 acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

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
        alpha_s[wid][kk * H + sm.head] = a;
        if (sm.head == 0) src_s[wid][kk] = p.col_src[k];
        if (p.alpha_out) {
          const int64_t e = p.perm ? p.perm[k] : k;
          p.alpha_out[e * H + sm.head] = a;
        }
      }
      __syncwarp();
#pragma unroll 2
      for (int kk = 0; kk < cn; ++kk) {
        const float* row = p.x_l + (int64_t)src_s[wid][kk] * p.ldx;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float a = alpha_s[wid][kk * H + h];
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int c4 = lane + 32 * j;
            if (c4 < C4) fma4(acc[j], a, ldg_cached(row + h * p.C + 4 * c4));
          }
        }
      }
      __syncwarp();
    }
    if (p.x_graph) {
      // sum_k alpha[k,h] * x_graph[g,h,:] = (sum_k alpha[k,h]) * x_graph[g,h,:]
      const float* row = p.x_graph + (int64_t)g * H * p.C;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float t = __shfl_sync(kFull, sm.total, h);  // lane h serves head h
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < C4) fma4(acc[j], t, ldg_cached(row + h * p.C + 4 * c4));
        }
      }
    }
  }
  epilogue_store<J>(p, i, lane, acc, 1.0f / H);
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

  switch (H) {
    case 1: return dispatch_gather<1>(p, stream);
    case 2: return dispatch_gather<2>(p, stream);
    case 4: return dispatch_gather<4>(p, stream);
    case 8: return dispatch_gather<8>(p, stream);
  }
  return GVQA_ERR_UNSUPPORTED;
}
