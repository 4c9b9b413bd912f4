// One GAT hop as ONE tensor-core kernel: "aggregate, then project" (see include/gvqa_b200.h: gvqa_gat_fused_hop_f32).
//
// The reference projects every node with all H heads, x_l = lin_l(cat(h, ins)) [N, H*C] (gat_skip.py:133), and then
// sums x_l rows over the in-edges with the softmax weights (gat_skip.py:155-168, 203-208).  The projection is linear,
// so the two steps commute:
//
//   out[i, :] = 1/H * sum_h sum_k alpha[k,h] (W_h h[src_k])  =  1/H * sum_h W_h z[i,h,:],   z[i,h,:] = sum_k alpha[k,h] h[src_k,:]
//
// i.e. ONE GEMM  Z[N, H*F] @ W'[C, H*F]^T  whose A operand is never stored anywhere: the converter warps of the GEMM
// build the A tile on the fly from a TMA-staged window of h rows (graphs are contiguous row ranges, and a tile is a
// set of whole graphs, so every source row of a tile lies in the tile's own window), split it into fp16 hi / lo'
// and hand it to the tensor core through tensor memory.  x_l[N, H*C] -- 63 MB written and 63 MB read back per hop at
// BASELINE cfg2 -- does not exist, the output of the GEMM is h_out[N, C] itself, and the hop's epilogue (head mean,
// per-graph instruction term, bias, skip, BatchNorm(eval) affine, ReLU; gat_skip.py:270-275) runs on the accumulators.
// HBM traffic per hop: read h (15.7 MB) once per column block (the second read hits L2), write h_out (15.7 MB).
//
// Split-precision product: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) UNSCALED (fp16 subnormals keep 2^-24 absolute
// resolution; the weights are pre-scaled by a power of two so that their low parts stay normal), three fp16 MMAs
// (lo*hi, hi*lo, hi*hi) accumulated in fp32 in ONE tensor-memory accumulator -- proj_gemm_f16.cu scales lo by 2^11 and
// needs a separate accumulator for it; here the accumulator budget goes into 256-column items instead.
//
// Kernel structure (two-CTA pairs, tcgen05.mma.cta_group::2, M = 256: 128 rows per CTA; N <= 256 columns per item):
//   warp 0       TMA producer: per slice of 16 input channels one stage = the [WIN x 16] fp32 window of h rows
//                (64B swizzle) + for each head pair the CTA's half of the [N x 2 x (16 hi | 16 lo)] fp16 weight tile
//   warp 1       MMA issuer (leader CTA, one elected lane): per stage 3 x H MMAs of M256 x N x K16
//   warps 4-11   converters, thread = one destination row: sum of alpha[k,h] * (16 channels of source row k) over
//                the in-edges for all H heads at once (packed FFMA2 from shared memory), fp16 split, tcgen05.st into
//                a four-slot tensor-memory ring; two groups of four warps alternate stages.  At the start of an item
//                they also compute the tile's softmax weights (PyG semantics) from the hop-invariant logit terms and
//                the node logits the previous hop's epilogue emitted
//   warps 12-15  epilogue: tcgen05.ld, per-warp transpose through shared memory, then the hop epilogue on full
//                128-byte row segments (skip rows prefetched by cp.async), and the NEXT hop's node logits
// Work items = (pair of row tiles, column tile), dealt round-robin to the 74 pairs; the row tiles come from a
// per-batch plan (gvqa_gat_fused_plan: greedy packing of whole graphs into <= 128 rows; graphs larger than 128 nodes
// are cut into chunks that stage the whole graph as their window).
#include <cuda_fp16.h>
#include <string.h>

#include "tcgen05_utils.cuh"

namespace gvqa {
namespace fused {

constexpr int kBM = 128;
constexpr int kKB = 16;                            // input channels per stage (x H heads)
constexpr int kMaxNB = 256;                        // output columns per work item (one MMA column block)
constexpr int kASlots = 4;                         // tensor-memory ring of split A sub-blocks
constexpr int kConvWarps = 8, kEpiWarps = 4;
constexpr int kConvThreads = kConvWarps * 32, kEpiThreads = kEpiWarps * 32;
// warps 0-3 = {TMA producer, MMA issuer, two idle warps}, 4-11 = converters, 12-15 = epilogue: 16 warps = four per SM
// sub-partition, i.e. 128 registers per thread (the register file is per sub-partition; the converters' 64 packed
// accumulators need ~120)
constexpr int kFirstConvWarp = 4;
constexpr int kFirstEpiWarp = kFirstConvWarp + kConvWarps;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);     // 512
constexpr int kEdgeCap = 640;                      // in-edges of a tile staged in shared memory (more: read from global)
constexpr uint32_t kBBoxBytes = 128 * 128;         // one head pair's half tile: <= 128 weight rows x 2 x (16 hi | 16 lo) fp16
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemA = kMaxNB;                // accumulator [0,256), A ring [256,512)
constexpr uint32_t kEpiStageBytes = 2 * 32 * 128;  // per epilogue warp: 32 rows x 32 columns fp32, accumulators + skip rows

template <int WIN, int H>
struct Cfg {
  static_assert(H == 2 || H == 4, "heads are staged in pairs");
  static constexpr int kStages = WIN == 128 ? 4 : 3;                    // shared-memory ring
  static constexpr uint32_t kABytes = WIN * 64;                         // [WIN x 16] fp32
  static constexpr uint32_t kStageBytes = kABytes + (H / 2) * kBBoxBytes;
  static constexpr uint32_t kListBytes = kEdgeCap * (4 * H + 4) + (kBM + 4) * 4;
  static constexpr uint32_t kNodeLogitBytes = WIN * 2 * H * 4;          // a_l | a_r of the window rows
  static constexpr uint32_t kConstBytes = (3 + 2 * H) * kMaxNB * 4;     // bias, scale, shift, next hop's logit vectors
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kListBytes +
                                  kNodeLogitBytes + kConstBytes + 256;
  static_assert(kSmem <= 232448, "shared memory budget of one CTA");
  static_assert(kTmemA + kASlots * 16 * H <= kTmemCols, "tensor-memory budget");
};

struct Params {
  CUtensorMap map_a, map_b;
  const int4* tiles;               // {first row, rows, first window row, 0}
  const int32_t* tile_count;
  const int32_t* rowptr;
  const int32_t* col_src;
  float* alpha;                    // [E, H] softmax weights, CSR order: input (logit_terms == NULL) or scratch
  const float* logit_terms;        // [E, H] hop-invariant logit terms, CSR order: softmax in the tile prologue
  const float* a_node;             // [parts][N][ld_an >= 2H] node logits a_l | a_r (partial sums)
  int64_t a_node_part_stride;
  int32_t ld_an;
  int32_t early;                   // per-batch inputs (plan, topology, logit terms, constants, weights) are older than the
                                   // predecessor kernel: they may be read before griddepcontrol.wait
  int32_t a_node_parts;
  float slope;
  const int32_t* node_graph;
  const float* h_in;               // [N, F] (also reached through map_a); global path for sources outside the window
  int64_t ld_h;
  const float* skip;               // [N, C] or NULL
  int64_t ld_skip;
  const float* graph_bias;         // [B, .] or NULL
  int64_t ldgb;
  const float* bias;
  const float* ep_scale;
  const float* ep_shift;
  float* h_out;                    // [N, C]
  int32_t* overflow;
  int32_t N, F, C, n_ct, nb, ks, epilogue;
  float out_scale;                 // 1 / (heads * weight scale)
  const float* v_next;             // [2H, C] collapsed logit vectors of the NEXT hop, or NULL
  float* a_part;                   // [3 * n_ct][N][2H] partial node logits of the next hop (three blocks per column block)
  unsigned long long* trace;       // debug only
};

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 fsub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, u64& a, u64& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {   // a -> low half (lower k)
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  const __half2 h = *reinterpret_cast<const __half2*>(&v);
  return __half22float2(h);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {   // completion of all MMAs so far -> this barrier in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mma_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void conv_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory"); }

// The staged window holds [WIN x 16] fp32 = 64-byte rows under the 64-byte TMA swizzle (address bits 4-5 ^= bits 7-8):
// 16-byte chunk c of row L lives at chunk c ^ ((L / 2) & 3) of its row.
__device__ __forceinline__ uint32_t win_addr(uint32_t stage_a, int L, int c) {
  return stage_a + (uint32_t)L * 64u + ((((uint32_t)c) ^ (((uint32_t)L >> 1) & 3u)) << 4);
}

// One destination row's share of a [16-channel x H-head] A sub-block: acc[h][i] = channels (2i, 2i+1) of
// sum_k alpha[k,h] * h[src_k, 16 j + ...].
// FAST: the tile's in-edge lists are staged in shared memory and every source row lies in the staged window.
template <int WIN, int H>
__device__ __forceinline__ void aggregate_fast(u64 (&acc)[H][8], int eb, int ee, const int32_t* src_s, const float* alpha_s,
                                               uint32_t stage_a) {
#pragma unroll 2
  for (int e = eb; e < ee; ++e) {
    const int L = src_s[e];
    float a[H];
    if constexpr (H == 4) {
      const float4 t = *reinterpret_cast<const float4*>(alpha_s + 4 * e);
      a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) a[h] = alpha_s[e * H + h];
    }
    u64 a2[H];
#pragma unroll
    for (int h = 0; h < H; ++h) a2[h] = pack2(a[h], a[h]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      u64 v01, v23;
      lds_2x64(win_addr(stage_a, L, q), v01, v23);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        ffma2(acc[h][2 * q], a2[h], v01);
        ffma2(acc[h][2 * q + 1], a2[h], v23);
      }
    }
  }
}

// generic: lists from global memory, sources outside the window (graphs larger than the window, cross-graph edges)
// read from global memory
template <int WIN, int H>
__device__ __forceinline__ void aggregate_any(u64 (&acc)[H][8], int eb, int ee, const Params& p, int e0, int win0,
                                              uint32_t stage_a, int j) {
#pragma unroll 1
  for (int e = eb; e < ee; ++e) {
    const int L = __ldg(p.col_src + e0 + e) - win0;
    u64 a2[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = p.alpha[(int64_t)(e0 + e) * H + h];     // (may have been written by this CTA: no ld.global.nc)
      a2[h] = pack2(a, a);
    }
    const bool inside = (unsigned)L < (unsigned)WIN;
    const int col = kKB * j;
    const float* gp = p.h_in + (int64_t)(win0 + L) * p.ld_h + col;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      u64 v01 = 0ull, v23 = 0ull;
      if (inside) {
        lds_2x64(win_addr(stage_a, L, q), v01, v23);
      } else if (col + 4 * q < p.F) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(gp + 4 * q));
        v01 = pack2(v.x, v.y);
        v23 = pack2(v.z, v.w);
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        ffma2(acc[h][2 * q], a2[h], v01);
        ffma2(acc[h][2 * q + 1], a2[h], v23);
      }
    }
  }
}

// per-CTA start / end (globaltimer, ns) in rows 100 + blockIdx.x / 4 of the same buffer
#define GVQA_FUSED_SPAN(which) \
  do { if (p.trace && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
       p.trace[(100 + (blockIdx.x >> 2)) * 8 + (blockIdx.x & 3) * 2 + (which)] = t_; } } while (0)
#define GVQA_FUSED_TRACE(slot, col) \
  do { if (p.trace && blockIdx.x == 0 && (slot) < 1100) p.trace[(slot) * 8 + (col)] = clock64(); } while (0)

// One warp's share of an item's epilogue: accumulator rows [32 * quarter, +32), column passes [pass_begin, pass_end)
// of 32 columns each.  tcgen05.ld (thread = row) -> transpose through shared memory -> lane = (row sub-index, float4
// column): every global access of a warp covers four full 128-byte row segments.  The skip rows of pass n+1 travel
// global -> shared by cp.async while pass n is finished; per-column constants come from cst_s (staged once per item:
// the L1 left beside ~220 KB of shared memory does not hold them).  With v_next the final values go back through the
// staging block and thread = row accumulates the NEXT hop's collapsed node logits a_l | a_r = h_out . V
// (gat_skip.py:134-135) into a_part[blk][row][2H]; zero_other: the sibling blocks blk + 1, blk + 2 get zeros (nobody
// else covers them).
template <int H>
__device__ __forceinline__ void run_epilogue(const Params& p, uint32_t accbuf, const float* cst_s, uint32_t tmem_base,
                                             int quarter, int lane, int row0, int nrows, int ct, int pass_begin,
                                             int pass_end, int blk, uint64_t* acc_full, uint32_t full_parity,
                                             uint64_t* acc_empty, bool zero_other, bool own_buffers, bool trace_on) {
  const uint32_t skipbuf = accbuf + 4096u;
  const int rsub = lane >> 3, f4 = lane & 7;
  const float* __restrict__ skip = p.skip;
  const float* __restrict__ gbias = p.graph_bias;
  float* __restrict__ h_out = p.h_out;
  const bool relu = p.epilogue == GVQA_EPI_AFFINE_RELU;
  const int colb = ct * p.nb;                              // first column of the item
  const int col_end = min(p.C, colb + p.nb);
  const int nr_w = nrows - quarter * 32;                   // rows of the tile in this warp's quarter
  const float* skip_base = skip ? skip + (int64_t)(row0 + quarter * 32 + rsub) * p.ld_skip + colb + 4 * f4 : p.h_in;
  auto issue_skip = [&](int pass) {                        // skip rows of `pass` -> skipbuf (zero-filled where there is none)
    const bool cok = skip != nullptr && colb + pass * 32 + 4 * f4 < col_end;
    const float* src0 = skip_base + pass * 32;
#pragma unroll
    for (int it8 = 0; it8 < 8; ++it8) {
      const int row = it8 * 4 + rsub;
      const bool ok = cok && row < nr_w;
      const float* src = ok ? src0 + (int64_t)it8 * 4 * p.ld_skip : p.h_in;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(skipbuf + (uint32_t)row * 128u +
                                                                        (uint32_t)((f4 ^ (row & 7)) << 4)),
                   "l"(src), "r"(ok ? 16 : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // (a helper warp stages through the operand ring: nothing may be written there before every MMA has completed)
  if (own_buffers && pass_begin < pass_end) issue_skip(pass_begin);
  int gid = -1;                                            // graph of the thread's own row when it takes a graph_bias
  const int r = quarter * 32 + lane;
  if (r < nrows && gbias && __ldg(p.rowptr + row0 + r + 1) > __ldg(p.rowptr + row0 + r)) gid = __ldg(p.node_graph + row0 + r);
  int grow8[8];                                            // graph (or -1) of the eight rows this lane finishes per pass
#pragma unroll
  for (int it8 = 0; it8 < 8; ++it8) grow8[it8] = __shfl_sync(kFull, gid, it8 * 4 + rsub);
  float4 gbv[8];
  auto request_gb = [&](int pass) {                        // graph_bias segments of `pass` (rows with in-edges only)
    const int col = colb + pass * 32 + 4 * f4;
#pragma unroll
    for (int it8 = 0; it8 < 8; ++it8) {
      gbv[it8] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow8[it8] >= 0 && col < col_end) gbv[it8] = __ldg(reinterpret_cast<const float4*>(gbias + (int64_t)grow8[it8] * p.ldgb + col));
    }
  };
  if (pass_begin < pass_end) request_gb(pass_begin);
  u64 part2[2 * H];                                        // (sum over even, over odd columns) of the logit dot products
#pragma unroll
  for (int v = 0; v < 2 * H; ++v) part2[v] = 0ull;
  mbar_wait(acc_full, full_parity);
  if (trace_on && lane == 0) GVQA_FUSED_TRACE(1024, 2);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!own_buffers && pass_begin < pass_end) issue_skip(pass_begin);
  const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
  for (int pass = pass_begin; pass < pass_end; ++pass) {
    if (trace_on && lane == 0) GVQA_FUSED_TRACE(1040 + pass, 0);
    {
      uint32_t v32[32];
      GVQA_TMEM_LD32(v32, tcol + (uint32_t)(pass * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (pass == pass_end - 1 && acc_empty != nullptr) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_cta(acc_empty, 0);                     // TMEM is free: the next item's MMAs may start
        if (trace_on && lane == 0) GVQA_FUSED_TRACE(1024, 3);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(accbuf + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4), __uint_as_float(v32[4 * c]),
               __uint_as_float(v32[4 * c + 1]), __uint_as_float(v32[4 * c + 2]), __uint_as_float(v32[4 * c + 3]));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const int cl = pass * 32 + 4 * f4;                     // column inside the item
    const int col = colb + cl;
    const bool cvalid = col < col_end;
    const float4 bi = *reinterpret_cast<const float4*>(cst_s + cl);
    const float4 sc = *reinterpret_cast<const float4*>(cst_s + kMaxNB + cl);
    const float4 sh = *reinterpret_cast<const float4*>(cst_s + 2 * kMaxNB + cl);
    float* outp = h_out + (int64_t)(row0 + quarter * 32 + rsub) * p.C + col;
#pragma unroll
    for (int it8 = 0; it8 < 8; ++it8) {
      const int row = it8 * 4 + rsub;
      const uint32_t off = (uint32_t)row * 128u + (uint32_t)((f4 ^ (row & 7)) << 4);
      float4 t = lds128(accbuf + off);
      const float4 sk = lds128(skipbuf + off);
      t.x = fmaf(t.x, p.out_scale, gbv[it8].x) + bi.x + sk.x;
      t.y = fmaf(t.y, p.out_scale, gbv[it8].y) + bi.y + sk.y;
      t.z = fmaf(t.z, p.out_scale, gbv[it8].z) + bi.z + sk.z;
      t.w = fmaf(t.w, p.out_scale, gbv[it8].w) + bi.w + sk.w;
      t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y);          // (scale 1, shift 0 without an affine epilogue)
      t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
      if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      if (row < nr_w && cvalid) *reinterpret_cast<float4*>(outp + (int64_t)it8 * 4 * p.C) = t;
      if (p.v_next) sts128(accbuf + off, t.x, t.y, t.z, t.w);            // final values for the logit dot products
    }
    __syncwarp();
    if (trace_on && lane == 0) GVQA_FUSED_TRACE(1040 + pass, 3);
    if (pass + 1 < pass_end) {                             // skipbuf is free; both land while the logits are accumulated
      issue_skip(pass + 1);
      request_gb(pass + 1);
    }
    if (trace_on && lane == 0) GVQA_FUSED_TRACE(1040 + pass, 4);
    if (p.v_next) {
      u64 x[8][2];
#pragma unroll
      for (int c = 0; c < 8; ++c) lds_2x64(accbuf + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4), x[c][0], x[c][1]);
      const uint32_t vb = smem_u32(cst_s) + (uint32_t)(3 * kMaxNB + pass * 32) * 4u;
#pragma unroll
      for (int v = 0; v < 2 * H; ++v) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          u64 w0, w1;
          lds_2x64(vb + (uint32_t)(v * kMaxNB + 4 * c) * 4u, w0, w1);                              // 0 past C
          ffma2(part2[v], x[c][0], w0);
          ffma2(part2[v], x[c][1], w1);
        }
      }
      __syncwarp();                                        // the staging block may be overwritten by the next pass
    }
    if (trace_on && lane == 0) GVQA_FUSED_TRACE(1040 + pass, 5);
  }
  if (p.v_next && r < nrows) {
    float part[2 * H];
#pragma unroll
    for (int v = 0; v < 2 * H; ++v) {
      float lo, hi;
      unpack2(part2[v], lo, hi);
      part[v] = lo + hi;
    }
    float* dst = p.a_part + ((int64_t)blk * p.N + row0 + r) * (2 * H);
#pragma unroll
    for (int v4 = 0; v4 < 2 * H / 4; ++v4) {
      *reinterpret_cast<float4*>(dst + 4 * v4) = make_float4(part[4 * v4], part[4 * v4 + 1], part[4 * v4 + 2], part[4 * v4 + 3]);
      if (zero_other) {
        *reinterpret_cast<float4*>(dst + (int64_t)p.N * (2 * H) + 4 * v4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + (int64_t)2 * p.N * (2 * H) + 4 * v4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (trace_on && lane == 0) GVQA_FUSED_TRACE(1024, 4);
}

template <int WIN, int H>
__global__ void __launch_bounds__(kThreads, 1) gat_fused_hop_kernel(const __grid_constant__ Params p) {
  using C_ = Cfg<WIN, H>;
  constexpr uint32_t kABytes = C_::kABytes, kStageBytes = C_::kStageBytes;
  constexpr int kStages = C_::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi_stage = smem + (size_t)kStages * kStageBytes;
  float* alpha_s = reinterpret_cast<float*>(epi_stage + kEpiWarps * kEpiStageBytes);       // [kEdgeCap][H]
  int32_t* src_s = reinterpret_cast<int32_t*>(alpha_s + kEdgeCap * H);                     // [kEdgeCap] window-local
  int32_t* rp_s = src_s + kEdgeCap;                                                        // [kBM + 1], then a flag word
  int32_t* far_s = rp_s + kBM + 2;                        // [2] != 0: some source of the tile lies outside the window
                                                          // (one word per item parity: cleared an item ahead)
  float* an_s = reinterpret_cast<float*>(rp_s + kBM + 4);                                  // [WIN][2H] a_l | a_r
  float* cst_s = an_s + WIN * 2 * H;                      // [3 + 2H][kMaxNB]: bias, scale, shift, next hop's V rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(cst_s + (3 + 2 * H) * kMaxNB);
  uint64_t* tma_full = bars;                    // [kStages]
  uint64_t* smem_empty = bars + kStages;        // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;       // [kASlots]
  uint64_t* a_empty = a_ready + kASlots;        // [kASlots]
  uint64_t* acc_full = a_empty + kASlots;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* cst_full = acc_empty + 1;           // the item's epilogue constants are staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cst_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  GVQA_FUSED_SPAN(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&smem_empty[s], 1);
    }
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&a_ready[s], 128 * 2);            // four warps of each CTA of the pair
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiThreads * 2);
    mbar_init(cst_full, kEpiThreads);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch: with `early` everything up to the first read of the predecessor's outputs (h_in,
  // a_node) -- set-up, the plan, the tile's topology and logit terms, the epilogue constants -- overlaps the
  // predecessor's tail; each role waits right before its first dependent read.
  const bool early = p.early != 0;
  if (!early) pdl_wait();
  pdl_launch_dependents();

  const int T = __ldg(p.tile_count);
  const int n_ct = p.n_ct, ks = p.ks;
  const int items = ((T + 1) >> 1) * n_ct;

  if (warp < kFirstConvWarp) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        uint32_t it = 0;
        const uint32_t stage_tx = kABytes + (uint32_t)(H / 2) * (uint32_t)(p.nb >> 1) * 128u;
        if (early) pdl_wait();
        for (int item = pair_id; item < items; item += npairs) {
          const int prt = item / n_ct, ct = item - prt * n_ct;
          const int te = 2 * prt + rank;
          const int win0 = te < T ? __ldg(&p.tiles[te].z) : 0;
          const int n0 = ct * p.nb;
          const int ncols = min(p.nb, (p.C - n0 + 31) & ~31);
          const int nb0 = n0 + rank * (ncols >> 1);        // this CTA stages its half of the weight tile's rows
          for (int j = 0; j < ks; ++j, ++it) {
            const int s = it % kStages;
            mbar_wait(&smem_empty[s], ((it / kStages) & 1) ^ 1);
            unsigned char* st = smem + (size_t)s * kStageBytes;
            GVQA_FUSED_TRACE(it, 0);
            mbar_expect_tx(&tma_full[s], stage_tx);
            tma_load_2d(st, &p.map_a, &tma_full[s], kKB * j, win0);
#pragma unroll
            for (int hp = 0; hp < H / 2; ++hp)
              tma_load_2d(st + kABytes + hp * kBBoxBytes, &p.map_b, &tma_full[s], (j * (H / 2) + hp) * 64, nb0);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA) =====================
      if (rank == 0 && elect_one()) {
        uint32_t it = 0, item_it = 0;
        const uint64_t desc0 = umma_desc(smem_u32(smem));
        for (int item = pair_id; item < items; item += npairs, ++item_it) {
          const int ct = item % n_ct;
          const int ncols = min(p.nb, (p.C - ct * p.nb + 31) & ~31);
          const uint32_t idesc = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)((kBM * 2) >> 4) << 24);
          GVQA_FUSED_TRACE(1024 + item_it, 0);
          mbar_wait(acc_empty, (item_it & 1) ^ 1);
          GVQA_FUSED_TRACE(1024 + item_it, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int j = 0; j < ks; ++j, ++it) {
            const uint32_t s = it % kStages, slot = it % kASlots;
            const uint64_t bstage = desc0 + (uint64_t)((s * kStageBytes + kABytes) >> 4);
            mbar_wait(&a_ready[slot], (it / kASlots) & 1);     // implies tma_full[s] in both CTAs
            GVQA_FUSED_TRACE(it, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_slot = tmem_base + kTmemA + slot * (uint32_t)(16 * H);
#pragma unroll
            for (int h = 0; h < H; ++h) {
              // head pair box h/2, within a 128-byte row: head (h & 1) at +64 bytes, hi at +0, lo at +32
              const uint64_t b_hi = bstage + (uint64_t)(((h >> 1) * kBBoxBytes) >> 4) + 4 * (h & 1), b_lo = b_hi + 2;
              const uint32_t a_hi = a_slot + 16 * h, a_lo = a_hi + 8;
              mma_pair(tmem_base, a_lo, b_hi, idesc, (j | h) != 0);
              mma_pair(tmem_base, a_hi, b_lo, idesc, 1);
              mma_pair(tmem_base, a_hi, b_hi, idesc, 1);
            }
            commit_pair(&a_empty[slot]);
            commit_pair(&smem_empty[s]);
            if (j == ks - 1) commit_pair(acc_full);
            GVQA_FUSED_TRACE(it, 7);
          }
        }
      }
    }                                                      // (two idle warps complete warpgroup 0)
  } else if (warp < kFirstEpiWarp) {
    // ===================== converters: thread = one destination row of the tile =====================
    const int quarter = warp & 3;                          // TMEM lane quarter a warp may touch = warp id % 4
    const int grp = (warp - kFirstConvWarp) >> 2;          // group g converts the stages with (stage & 1) == g
    const int ctid = threadIdx.x - kFirstConvWarp * 32;
    const int r = quarter * 32 + lane;
    const uint32_t ta0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + kTmemA;
    uint32_t it = 0;
    float amax = 0.f;
    if (ctid == 0) far_s[0] = far_s[1] = 0;
    int item_par = 0;
    for (int item = pair_id; item < items; item += npairs, item_par ^= 1) {
      const int prt = item / n_ct;
      const int te = 2 * prt + rank;
      int4 tile = make_int4(0, 0, 0, 0);
      if (te < T) tile = __ldg(&p.tiles[te]);
      const int row0 = tile.x, nrows = tile.y, win0 = tile.z;
      if (ctid == 0) GVQA_FUSED_TRACE(1060, 0);
      conv_bar_sync();                                     // the previous item's lists are no longer read
      int e0 = 0, ne = 0;
      if (nrows > 0) {
        e0 = __ldg(p.rowptr + row0);
        ne = __ldg(p.rowptr + row0 + nrows) - e0;
      }
      for (int t = ctid; t <= nrows; t += kConvThreads) rp_s[t] = __ldg(p.rowptr + row0 + t) - e0;
      const bool staged = ne <= kEdgeCap;
      const bool softmax_here = p.logit_terms != nullptr;
      const float* lsrc = softmax_here ? p.logit_terms : p.alpha;   // what the staged list starts from
      if (staged) {
        int far = 0;
        for (int k = ctid; k < ne; k += kConvThreads) {
          const int L = __ldg(p.col_src + e0 + k) - win0;
          src_s[k] = L;
          far |= (unsigned)L >= (unsigned)WIN;
          if constexpr (H == 4) {
            *reinterpret_cast<float4*>(alpha_s + 4 * k) = __ldg(reinterpret_cast<const float4*>(lsrc) + e0 + k);
          } else {
            *reinterpret_cast<float2*>(alpha_s + 2 * k) = __ldg(reinterpret_cast<const float2*>(lsrc) + e0 + k);
          }
        }
        if (far) far_s[item_par] = 1;                      // (every writer stores 1)
      }
      if (early && item == pair_id) pdl_wait();            // a_node is the predecessor's output
      if (softmax_here && nrows > 0) {
        // node logits of the window rows, partial sums added in fixed order
        for (int t = ctid; t < WIN * 2 * H / 4; t += kConvThreads) {
          const int L = t / (2 * H / 4), part4 = t - L * (2 * H / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (win0 + L < p.N) {
            const float* src = p.a_node + (int64_t)(win0 + L) * p.ld_an + 4 * part4;
            v = __ldg(reinterpret_cast<const float4*>(src));
            for (int pt = 1; pt < p.a_node_parts; ++pt) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(src + pt * p.a_node_part_stride));
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
          }
          *reinterpret_cast<float4*>(an_s + 4 * t) = v;
        }
      }
      conv_bar_sync();
      if (ctid == 0) GVQA_FUSED_TRACE(1060, 1);
      const bool fast = staged && far_s[item_par] == 0;
      if (ctid == 0) far_s[item_par ^ 1] = 0;              // the other word was last read an item ago, behind two barriers
      if (softmax_here) {
        if (fast) {
          // PyG softmax per (row, head) over the row's in-edges (gat_skip.py:183-192), in place in alpha_s
          for (int pr = ctid; pr < nrows * H; pr += kConvThreads) {
            const int row = pr / H, h = pr - row * H;
            const int kb = rp_s[row], ke = rp_s[row + 1];
            const float ar = an_s[(row0 - win0 + row) * (2 * H) + H + h];
            auto logit = [&](int k) {
              return leaky_relu((alpha_s[k * H + h] + an_s[src_s[k] * (2 * H) + h]) + ar, p.slope);
            };
            if (ke - kb <= 6) {                            // the common case: every logit and exponential once
              float l[6];
              float mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                l[i] = kb + i < ke ? logit(kb + i) : -INFINITY;
                mx = fmaxf(mx, l[i]);
              }
              float sum = 0.f;
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                l[i] = kb + i < ke ? expf(l[i] - mx) : 0.f;
                sum += l[i];
              }
              const float inv = 1.0f / (sum + 1e-16f);
#pragma unroll
              for (int i = 0; i < 6; ++i)
                if (kb + i < ke) alpha_s[(kb + i) * H + h] = l[i] * inv;
            } else {
              float mx = -INFINITY;
              for (int k = kb; k < ke; ++k) mx = fmaxf(mx, logit(k));
              float sum = 0.f;
              for (int k = kb; k < ke; ++k) sum += expf(logit(k) - mx);
              const float inv = 1.0f / (sum + 1e-16f);
              for (int k = kb; k < ke; ++k) alpha_s[k * H + h] = expf(logit(k) - mx) * inv;
            }
          }
        } else if (ctid < nrows) {
          // generic path: one thread per row, everything from global memory, weights to the scratch array
          auto node_term = [&](int64_t node, int col) {
            float t = __ldg(p.a_node + node * p.ld_an + col);
            for (int pt = 1; pt < p.a_node_parts; ++pt) t += __ldg(p.a_node + pt * p.a_node_part_stride + node * p.ld_an + col);
            return t;
          };
          const int kb = e0 + rp_s[ctid], ke = e0 + rp_s[ctid + 1];
          for (int h = 0; h < H; ++h) {
            const float ar = node_term(row0 + ctid, H + h);
            auto logit = [&](int k) {
              return leaky_relu((__ldg(p.logit_terms + (int64_t)k * H + h) + node_term(__ldg(p.col_src + k), h)) + ar, p.slope);
            };
            float mx = -INFINITY;
            for (int k = kb; k < ke; ++k) mx = fmaxf(mx, logit(k));
            float sum = 0.f;
            for (int k = kb; k < ke; ++k) sum += expf(logit(k) - mx);
            const float inv = 1.0f / (sum + 1e-16f);
            for (int k = kb; k < ke; ++k) p.alpha[(int64_t)k * H + h] = expf(logit(k) - mx) * inv;
          }
        }
        conv_bar_sync();
      }
      if (ctid == 0) GVQA_FUSED_TRACE(1060, 2);
      int eb = 0, ee = 0;
      if (r < nrows) {
        eb = rp_s[r];
        ee = rp_s[r + 1];
      }
      for (int j = 0; j < ks; ++j, ++it) {
        if ((int)(it & 1) != grp) continue;
        const uint32_t s = it % kStages, slot = it % kASlots;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 1);
        const uint32_t stage_a = smem_u32(smem + (size_t)s * kStageBytes);
        u64 acc[H][8];
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[h][i] = 0ull;
        if (fast) aggregate_fast<WIN, H>(acc, eb, ee, src_s, alpha_s, stage_a);
        else aggregate_any<WIN, H>(acc, eb, ee, p, e0, win0, stage_a, j);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 2);
        mbar_wait(&a_empty[slot], ((it / kASlots) & 1) ^ 1);   // the MMAs that read this slot four stages ago are done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = ta0 + slot * (uint32_t)(16 * H);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          uint32_t pk[16];                                 // [0,8) hi, [8,16) lo; two k elements per word
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x0, x1, d0, d1;
            unpack2(acc[h][i], x0, x1);
            amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
            const uint32_t h2 = pack_half2(x0, x1);
            const float2 f = unpack_half2(h2);
            unpack2(fsub2(acc[h][i], pack2(f.x, f.y)), d0, d1);
            pk[i] = h2;
            pk[8 + i] = pack_half2(d0, d1);                // unscaled: fp16 subnormals keep 2^-24 absolute resolution
          }
          GVQA_TMEM_ST16(ta + 16 * h, pk, 0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_cta(&a_ready[slot], 0);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 3);
      }
    }
    if (p.overflow != nullptr && !(amax <= 65000.0f)) atomicOr(p.overflow, 1);   // also catches NaN / inf
    if (items > pair_id) {
      // helper epilogue of the CTA's last item: group 1 takes the middle third of the column passes, group 0 the last
      // third, staged through the first bytes of the operand ring (no load is in flight or will be issued any more,
      // and acc_full says every MMA has read it)
      int last_item = pair_id, n_it = 0;
      while (last_item + npairs < items) {
        last_item += npairs;
        ++n_it;
      }
      const int prt = last_item / n_ct, ct = last_item - prt * n_ct;
      const int te = 2 * prt + rank;
      int4 tile = make_int4(0, 0, 0, 0);
      if (te < T) tile = __ldg(&p.tiles[te]);
      const int npass = (min(p.C, ct * p.nb + p.nb) - ct * p.nb + 31) >> 5;
      const int third = (npass + 2) / 3;
      const int pb = grp == 1 ? third : min(2 * third, npass), pe = grp == 1 ? min(2 * third, npass) : npass;
      mbar_wait(cst_full, n_it & 1);
      run_epilogue<H>(p, smem_u32(smem) + (uint32_t)((1 - grp) * 4 + quarter) * kEpiStageBytes, cst_s, tmem_base, quarter, lane, tile.x,
                      tile.y, ct, pb, pe, 3 * ct + 2 - grp, acc_full, n_it & 1, nullptr, false, false, false);
    }
  } else {
    // ===================== epilogue: warp = 32 accumulator rows x the columns of the item =====================
    const int quarter = warp & 3;
    const int etid = threadIdx.x - kFirstEpiWarp * 32;
    const uint32_t accbuf = smem_u32(epi_stage + (size_t)(warp - kFirstEpiWarp) * kEpiStageBytes);
    const bool affine = p.epilogue == GVQA_EPI_AFFINE || p.epilogue == GVQA_EPI_AFFINE_RELU;
    uint32_t item_it = 0;
    for (int item = pair_id; item < items; item += npairs, ++item_it) {
      const int prt = item / n_ct, ct = item - prt * n_ct;
      const int te = 2 * prt + rank;
      int4 tile = make_int4(0, 0, 0, 0);
      if (te < T) tile = __ldg(&p.tiles[te]);
      const int colb = ct * p.nb;
      epi_bar_sync();                                      // the previous item's constants are no longer read
      for (int c = etid; c < p.nb; c += kEpiThreads) {
        const int col = colb + c;
        const bool ok = col < p.C;
        cst_s[c] = (ok && p.bias) ? __ldg(p.bias + col) : 0.f;
        cst_s[kMaxNB + c] = (ok && affine) ? __ldg(p.ep_scale + col) : 1.f;
        cst_s[2 * kMaxNB + c] = (ok && affine) ? __ldg(p.ep_shift + col) : 0.f;
#pragma unroll
        for (int v = 0; v < 2 * H; ++v) cst_s[(3 + v) * kMaxNB + c] = (ok && p.v_next) ? __ldg(p.v_next + (int64_t)v * p.C + col) : 0.f;
      }
      epi_bar_sync();
      mbar_arrive(cst_full);                               // (the helper warps of the last item wait for this)
      if (early && item_it == 0) pdl_wait();               // the skip rows are the predecessor's output
      // on the CTA's last item the two converter groups, idle by then, take the upper two thirds of the column passes
      const bool last = item + npairs >= items;
      const int npass = (min(p.C, colb + p.nb) - colb + 31) >> 5;
      const int split = last ? (npass + 2) / 3 : npass;
      run_epilogue<H>(p, accbuf, cst_s, tmem_base, quarter, lane, tile.x, tile.y, ct, 0, split, 3 * ct, acc_full, item_it & 1,
                      acc_empty, !last, true, item_it == 0 && warp == kFirstEpiWarp);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();         // neither CTA leaves (or frees tensor memory) while its peer may still signal or read it
  GVQA_FUSED_SPAN(1);
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// ---- per-batch row-tile plan --------------------------------------------------------------------------------------
// Greedy: consecutive whole graphs are packed while they fit 128 rows (window = the tile's own rows); a graph with
// more than 128 nodes is cut into 128-row chunks whose window starts at the graph (or, for graphs larger than the
// window, is centred on the chunk: sources outside it take the kernel's global-memory path).
__host__ __device__ inline int plan_tiles(const int32_t* graph_ptr, int B, int win, int4* tiles, int max_tiles) {
  int t = 0, g = 0;
  while (g < B) {
    const int r0 = graph_ptr[g], n_g = graph_ptr[g + 1] - r0;
    if (n_g <= 0) {
      ++g;
      continue;
    }
    if (n_g > kBM) {
      for (int s = r0; s < r0 + n_g; s += kBM) {
        const int nr = (r0 + n_g - s) < kBM ? (r0 + n_g - s) : kBM;
        int w0 = r0;
        if (n_g > win) {
          w0 = s - (win - kBM) / 2;
          if (w0 > r0 + n_g - win) w0 = r0 + n_g - win;
          if (w0 < r0) w0 = r0;
        }
        if (t < max_tiles) tiles[t] = make_int4(s, nr, w0, 0);
        ++t;
      }
      ++g;
      continue;
    }
    int end = g + 1;
    while (end < B && graph_ptr[end + 1] - r0 <= kBM) ++end;
    if (t < max_tiles) tiles[t] = make_int4(r0, graph_ptr[end] - r0, r0, 0);
    ++t;
    g = end;
  }
  return t < max_tiles ? t : max_tiles;
}

// One CTA.  The greedy packing is sequential only in WHICH graphs start a tile: every graph computes in parallel where
// a tile starting at it would end (binary search in graph_ptr) and how many tiles it would emit, one thread then walks
// the chain of starts (two shared-memory loads per tile), and the tiles are written in parallel.  Same result as
// plan_tiles() on the host.
__global__ void __launch_bounds__(1024) fused_plan_kernel(const int32_t* __restrict__ graph_ptr, const int64_t* __restrict__ batch,
                                                         int N, int B, int win, int4* tiles, int32_t* count, int max_tiles) {
  extern __shared__ int32_t plan_s[];
  int32_t* gp_s = plan_s;                 // [B + 1]
  int32_t* nxt_s = gp_s + (B + 1);        // [B] first graph after a tile that starts at g
  int32_t* cnt_s = nxt_s + B;             // [B] tiles such a start emits
  int32_t* off_s = cnt_s + B;             // [B] index of its first tile, -1 when g does not start a tile
  if (graph_ptr != nullptr) {
    for (int i = threadIdx.x; i <= B; i += blockDim.x) gp_s[i] = graph_ptr[i];
  } else {
    // graph boundaries straight from the non-decreasing batch vector (what gvqa_build_csr derives as well, ids clamped
    // the same way): the plan then does not wait for the CSR build
    for (int base = 0; base < N; base += 4 * (int)blockDim.x) {      // eight independent loads in flight per thread
      int64_t bv[4], pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = base + i * (int)blockDim.x + (int)threadIdx.x;
        bv[i] = t < N ? batch[t] : 0;
        pv[i] = (t < N && t > 0) ? batch[t - 1] : -1;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = base + i * (int)blockDim.x + (int)threadIdx.x;
        if (t >= N) continue;
        int64_t b = bv[i], prev = pv[i];
        b = b < 0 ? 0 : (b >= B ? (B > 0 ? B - 1 : 0) : b);
        prev = prev < -1 ? -1 : (prev >= B ? B - 1 : prev);
        for (int64_t g = prev + 1; g <= b && g <= B; ++g) gp_s[g] = t;
        if (t == N - 1)
          for (int64_t g = b + 1; g <= B; ++g) gp_s[g] = N;
      }
    }
    if (N == 0)
      for (int i = threadIdx.x; i <= B; i += blockDim.x) gp_s[i] = 0;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < B; g += blockDim.x) {
    const int r0 = gp_s[g], n_g = gp_s[g + 1] - r0;
    int nxt = g + 1, cnt = 0;
    if (n_g > kBM) {
      cnt = (n_g + kBM - 1) / kBM;
    } else if (n_g > 0) {
      int lo = g + 1, hi = B;             // largest e in [g + 1, B] with graph_ptr[e] <= r0 + 128
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (gp_s[mid] - r0 <= kBM) lo = mid;
        else hi = mid - 1;
      }
      nxt = lo;
      cnt = 1;
    }
    nxt_s[g] = nxt;
    cnt_s[g] = cnt;
    off_s[g] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int g = 0, t = 0;
    while (g < B) {
      const int c = cnt_s[g];
      if (c > 0) off_s[g] = t;
      t += c;
      g = nxt_s[g];
    }
    *count = t < max_tiles ? t : max_tiles;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < B; g += blockDim.x) {
    const int t0 = off_s[g];
    if (t0 < 0) continue;
    const int r0 = gp_s[g], n_g = gp_s[g + 1] - r0;
    if (n_g > kBM) {
      int t = t0;
      for (int s = r0; s < r0 + n_g; s += kBM, ++t) {
        const int nr = (r0 + n_g - s) < kBM ? (r0 + n_g - s) : kBM;
        int w0 = r0;
        if (n_g > win) {
          w0 = s - (win - kBM) / 2;
          if (w0 > r0 + n_g - win) w0 = r0 + n_g - win;
          if (w0 < r0) w0 = r0;
        }
        if (t < max_tiles) tiles[t] = make_int4(s, nr, w0, 0);
      }
    } else if (t0 < max_tiles) {
      tiles[t0] = make_int4(r0, gp_s[nxt_s[g]] - r0, r0, 0);
    }
  }
}

// ---- softmax weights of all in-edges (gat_skip.py:183-192 + PyG utils.softmax), CSR order ---------------------------
template <int H>
__global__ void __launch_bounds__(256) gat_alpha_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
                                                        const int32_t* __restrict__ perm, const int32_t* __restrict__ node_graph,
                                                        const float* __restrict__ a_node, int64_t lda, int parts,
                                                        int64_t part_stride,
                                                        const float* __restrict__ a_edge, int64_t lde,
                                                        const float* __restrict__ a_graph, int64_t ldag, float slope, int N,
                                                        float* __restrict__ alpha, float* __restrict__ alpha_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t / H, h = t - i * H;
  if (i >= N) return;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  if (e1 <= e0) return;
  auto node_term = [&](int64_t node, int col) {           // a_node may arrive as partial sums (fixed order)
    float t = a_node[node * lda + col];
    for (int pt = 1; pt < parts; ++pt) t += a_node[pt * part_stride + node * lda + col];
    return t;
  };
  float tg = a_graph ? a_graph[(int64_t)node_graph[i] * ldag + h] : 0.f;
  tg += node_term(i, H + h);
  auto logit = [&](int k) {
    const int64_t e = perm ? perm[k] : k;
    const float v = (a_edge[e * lde + h] + node_term(col_src[k], h)) + tg;
    return leaky_relu(v, slope);
  };
  float mx = -INFINITY;
  for (int k = e0; k < e1; ++k) mx = fmaxf(mx, logit(k));
  float sum = 0.f;
  for (int k = e0; k < e1; ++k) sum += expf(logit(k) - mx);
  const float inv = 1.0f / (sum + 1e-16f);
  for (int k = e0; k < e1; ++k) {
    const float a = expf(logit(k) - mx) * inv;
    alpha[(int64_t)k * H + h] = a;
    if (alpha_out) alpha_out[(int64_t)(perm ? perm[k] : k) * H + h] = a;
  }
}

// ---- hop-invariant logit terms of all hops in CSR order (once per batch):
//   terms[hop][k][h] = a_edge[perm[k]][hop*H + h] + a_graph[hop][graph of the edge's destination][h]
// One thread per in-edge (CSR position k): the destination row by binary search in rowptr, then per hop one H-wide
// vector load / store (coalesced stores, 16-byte gathers through perm).
template <int H>
__global__ void __launch_bounds__(256) logit_terms_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ node_graph,
                                                          const float* __restrict__ a_edge, int64_t lde,
                                                          const float* __restrict__ a_graph, int64_t ldag, int64_t hop_stride,
                                                          int hops, int N, int64_t E, int64_t terms_hop_stride,
                                                          float* __restrict__ terms) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= E) return;
  int lo = 0, hi = N - 1;                      // largest row i with rowptr[i] <= k
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(rowptr + mid) <= k) lo = mid;
    else hi = mid - 1;
  }
  const int g = a_graph ? __ldg(node_graph + lo) : 0;
  const int64_t e = perm ? __ldg(perm + k) : k;
  const float* ae = a_edge + e * lde;
  for (int hop = 0; hop < hops; ++hop) {
    float v[H];
#pragma unroll
    for (int h = 0; h < H; ++h) v[h] = __ldg(ae + hop * H + h) + (a_graph ? __ldg(a_graph + hop * hop_stride + (int64_t)g * ldag + h) : 0.f);
    float* dst = terms + hop * terms_hop_stride + k * H;
    if constexpr (H == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  }
}

// ---- weight prepack: W [H*C, >= F] fp32 (row h*C + c) scaled by `scale` (a power of two that brings the low parts into
// fp16's normal range) -> [C, Fp/16, H, (16 hi | 16 lo)] fp16, Fp = F rounded up to 16 ------------------------------
__global__ void fused_pack_kernel(const float* __restrict__ w, int64_t ldw, int H, int C, int F, int Fp, float scale,
                                  __half* __restrict__ out) {
  const int64_t total = (int64_t)C * H * Fp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ((int64_t)H * Fp);
    const int rem = (int)(i - c * H * Fp), h = rem / Fp, k = rem - h * Fp;
    const float x = k < F ? w[((int64_t)h * C + c) * ldw + k] * scale : 0.f;
    const __half hi = __float2half_rn(x);
    const int64_t o = c * ((int64_t)H * Fp * 2) + (int64_t)((k >> 4) * H + h) * 32 + (k & 15);
    out[o] = hi;
    out[o + 16] = __float2half_rn(x - __half2float(hi));
  }
}

unsigned long long* g_fused_trace = nullptr;

// column blocks of a launch: as few as fit the 256-column accumulator, equal widths (multiples of 32: each CTA of a
// pair stages half a block); one more split when the batch would otherwise leave more than half of the pairs idle
inline void column_blocks(int64_t num_nodes, int channels, int* n_ct, int* nb) {
  int n = (channels + kMaxNB - 1) / kMaxNB;
  const int64_t pair_tiles = (num_nodes + 2 * kBM - 1) / (2 * kBM);
  if (pair_tiles * n * 2 <= kNumSMs / 2 && channels / (2 * n) >= 64) n *= 2;
  *nb = (((channels + n - 1) / n) + 31) & ~31;
  *n_ct = (channels + *nb - 1) / *nb;
}

}  // namespace fused
}  // namespace gvqa

using namespace gvqa;

extern "C" GVQA_API void gvqa_debug_set_fused_trace(unsigned long long* buf) { fused::g_fused_trace = buf; }

extern "C" GVQA_API int64_t gvqa_gat_fused_max_tiles(int64_t num_nodes, int64_t num_graphs) {
  return num_graphs + (num_nodes + fused::kBM - 1) / fused::kBM + 2;
}

extern "C" GVQA_API int32_t gvqa_gat_fused_part_blocks(int64_t num_nodes, int32_t channels) {
  int n_ct, nb;
  fused::column_blocks(num_nodes, channels, &n_ct, &nb);
  return 3 * n_ct;
}

extern "C" GVQA_API int32_t gvqa_gat_fused_window(int32_t max_nodes_per_graph) {
  return max_nodes_per_graph > fused::kBM ? 256 : 128;
}

static int launch_plan(const int32_t* graph_ptr, const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t window,
                       int32_t* tiles, int32_t* count, int64_t max_tiles, void* stream_) {
  if (num_graphs < 0 || max_tiles < 0 || num_nodes < 0 || num_nodes >= (1ll << 31) || (window != 128 && window != 256))
    return GVQA_ERR_BAD_SHAPE;
  if ((!graph_ptr && !batch && num_nodes > 0) || !tiles || !count) return GVQA_ERR_NULL_POINTER;
  if (!aligned16(tiles)) return GVQA_ERR_MISALIGNED;
  if ((4 * num_graphs + 1) * 4 > 200 * 1024) return GVQA_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(4 * num_graphs + 1) * 4;
  if (smem > 48 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      if (cudaFuncSetAttribute(fused::fused_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        (void)cudaGetLastError();
        return GVQA_ERR_CUDA;
      }
      attr_done = true;
    }
  }
  fused::fused_plan_kernel<<<1, batch ? 1024 : 256, smem, static_cast<cudaStream_t>(stream_)>>>(
      graph_ptr, batch, (int)num_nodes, (int)num_graphs, window, reinterpret_cast<int4*>(tiles), count, (int)max_tiles);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_fused_plan(const int32_t* graph_ptr, int64_t num_graphs, int32_t window, int32_t* tiles,
                                            int32_t* count, int64_t max_tiles, void* stream_) {
  if (!graph_ptr) return GVQA_ERR_NULL_POINTER;
  return launch_plan(graph_ptr, nullptr, 0, num_graphs, window, tiles, count, max_tiles, stream_);
}

extern "C" GVQA_API int gvqa_gat_fused_plan_from_batch(const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                                                       int32_t window, int32_t* tiles, int32_t* count, int64_t max_tiles,
                                                       void* stream_) {
  if (!batch && num_nodes > 0) return GVQA_ERR_NULL_POINTER;
  return launch_plan(nullptr, batch, num_nodes, num_graphs, window, tiles, count, max_tiles, stream_);
}

extern "C" GVQA_API int gvqa_gat_fused_plan_host(const int32_t* graph_ptr_host, int64_t num_graphs, int32_t window,
                                                 int32_t* tiles_host, int32_t* count_host, int64_t max_tiles) {
  if (num_graphs < 0 || max_tiles < 0 || (window != 128 && window != 256)) return GVQA_ERR_BAD_SHAPE;
  if (!graph_ptr_host || !tiles_host || !count_host) return GVQA_ERR_NULL_POINTER;
  *count_host = fused::plan_tiles(graph_ptr_host, (int)num_graphs, window, reinterpret_cast<int4*>(tiles_host), (int)max_tiles);
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_alpha_f32(const int32_t* rowptr, const int32_t* col_src, const int32_t* perm,
                                           const int32_t* node_graph, const float* a_node, int64_t ld_a_node,
                                           int32_t a_node_parts, int64_t a_node_part_stride, const float* a_edge, int64_t lde, const float* a_graph, int64_t ld_a_graph,
                                           float negative_slope, int64_t num_nodes, int32_t heads, float* alpha,
                                           float* alpha_out, void* stream_) {
  if (num_nodes < 0 || num_nodes >= (1ll << 28) || a_node_parts < 1) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!rowptr || !col_src || !a_node || !a_edge || !alpha || (a_graph && !node_graph)) return GVQA_ERR_NULL_POINTER;
  const int64_t threads = num_nodes * heads;
  const dim3 grid((unsigned)((threads + 255) / 256)), block(256);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
#define GVQA_ALPHA(HH)                                                                                                  \
  if (launch_pdl(2, fused::gat_alpha_kernel<HH>, grid, block, 0, st, rowptr, col_src, perm, node_graph, a_node, ld_a_node, \
                 (int)a_node_parts, a_node_part_stride, a_edge, lde, a_graph, ld_a_graph, negative_slope, (int)num_nodes, alpha, alpha_out) != cudaSuccess) {  \
    (void)cudaGetLastError();                                                                                           \
    return GVQA_ERR_CUDA;                                                                                               \
  }
  switch (heads) {
    case 1: GVQA_ALPHA(1) break;
    case 2: GVQA_ALPHA(2) break;
    case 4: GVQA_ALPHA(4) break;
    case 8: GVQA_ALPHA(8) break;
    default: return GVQA_ERR_UNSUPPORTED;
  }
#undef GVQA_ALPHA
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_fused_logit_terms_f32(const int32_t* rowptr, const int32_t* perm, const int32_t* node_graph,
                                                       const float* a_edge, int64_t lde, const float* a_graph,
                                                       int64_t ld_a_graph, int64_t hop_stride_a_graph, int32_t hops,
                                                       int64_t num_nodes, int64_t num_edges, int32_t heads, float* terms,
                                                       int64_t terms_hop_stride, void* stream_) {
  if (num_nodes < 0 || num_nodes >= (1ll << 28) || num_edges < 0 || hops < 1 || terms_hop_stride < num_edges * heads)
    return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0 || num_edges == 0) return GVQA_OK;
  if (!rowptr || !a_edge || !terms || (a_graph && !node_graph)) return GVQA_ERR_NULL_POINTER;
  if (!aligned16(terms) || (terms_hop_stride & 3)) return GVQA_ERR_MISALIGNED;
  const dim3 grid((unsigned)((num_edges + 255) / 256)), block(256);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (heads == 4)
    fused::logit_terms_kernel<4><<<grid, block, 0, st>>>(rowptr, perm, node_graph, a_edge, lde, a_graph, ld_a_graph,
                                                         hop_stride_a_graph, hops, (int)num_nodes, num_edges,
                                                         terms_hop_stride, terms);
  else if (heads == 2)
    fused::logit_terms_kernel<2><<<grid, block, 0, st>>>(rowptr, perm, node_graph, a_edge, lde, a_graph, ld_a_graph,
                                                         hop_stride_a_graph, hops, (int)num_nodes, num_edges,
                                                         terms_hop_stride, terms);
  else
    return GVQA_ERR_UNSUPPORTED;
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int64_t gvqa_gat_fused_pack_halves(int32_t heads, int32_t channels, int32_t in_channels) {
  const int64_t fp = (in_channels + 15) & ~15;
  return (int64_t)channels * heads * fp * 2;
}

extern "C" GVQA_API int gvqa_gat_fused_pack_f16(const float* w, int64_t ldw, int32_t heads, int32_t channels,
                                                int32_t in_channels, float scale, void* packed, void* stream_) {
  if (heads <= 0 || channels <= 0 || in_channels <= 0 || ldw < in_channels || !(scale > 0.f)) return GVQA_ERR_BAD_SHAPE;
  if (!w || !packed) return GVQA_ERR_NULL_POINTER;
  const int fp = (in_channels + 15) & ~15;
  const int64_t total = (int64_t)channels * heads * fp;
  const int64_t blocks = (total + 255) / 256;
  fused::fused_pack_kernel<<<(unsigned)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      w, ldw, heads, channels, in_channels, fp, scale, static_cast<__half*>(packed));
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_fused_supported(int32_t heads, int32_t in_channels, int32_t channels) {
  return (heads == 4 || heads == 2) && in_channels > 0 && (in_channels & 3) == 0 && channels > 0 &&
         (channels & 3) == 0;
}

extern "C" GVQA_API int gvqa_gat_fused_hop_f32(const gvqa_gat_fused_args* a, void* stream_) {
  using namespace fused;
  if (!a) return GVQA_ERR_NULL_POINTER;
  if (a->num_nodes < 0 || a->num_nodes >= (1ll << 31) || a->in_channels <= 0 || a->channels <= 0 || a->heads <= 0)
    return GVQA_ERR_BAD_SHAPE;
  if (!gvqa_gat_fused_supported(a->heads, a->in_channels, a->channels) || (a->window != 128 && a->window != 256))
    return GVQA_ERR_UNSUPPORTED;
  if (a->num_nodes == 0) return GVQA_OK;
  if (!a->h_in || !a->w_pack || !a->tiles || !a->tile_count || !a->rowptr || !a->col_src || !a->alpha || !a->node_graph ||
      !a->h_out || (a->logit_terms && !a->a_node))
    return GVQA_ERR_NULL_POINTER;
  if (a->logit_terms && (a->a_node_parts < 1 || !aligned16(a->logit_terms) || !aligned16(a->a_node) ||
                         (a->a_node_part_stride & 3)))
    return GVQA_ERR_MISALIGNED;
  if ((a->epilogue == GVQA_EPI_AFFINE || a->epilogue == GVQA_EPI_AFFINE_RELU) && (!a->ep_scale || !a->ep_shift))
    return GVQA_ERR_NULL_POINTER;
  if (a->epilogue == GVQA_EPI_GRAPH_LN) return GVQA_ERR_UNSUPPORTED;
  const int64_t ld_h = a->ld_h ? a->ld_h : a->in_channels, ld_skip = a->ld_skip ? a->ld_skip : a->channels;
  const int64_t ldgb = a->ld_graph_bias ? a->ld_graph_bias : a->channels;
  if ((ld_h & 3) || (ld_skip & 3) || (ldgb & 3)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(a->h_in) || !aligned16(a->w_pack) || !aligned16(a->tiles) || !aligned16(a->h_out) || !aligned16(a->alpha) ||
      (a->skip && !aligned16(a->skip)) || (a->graph_bias && !aligned16(a->graph_bias)) || (a->bias && !aligned16(a->bias)) ||
      (a->ep_scale && !aligned16(a->ep_scale)) || (a->ep_shift && !aligned16(a->ep_shift)))
    return GVQA_ERR_MISALIGNED;
  if (!(a->w_scale > 0.f)) return GVQA_ERR_BAD_SHAPE;
  Params p;
  memset(&p, 0, sizeof(p));
  const int fp = (a->in_channels + 15) & ~15;
  int n_ct, nb;
  column_blocks(a->num_nodes, a->channels, &n_ct, &nb);
  if (!make_map(&p.map_a, a->h_in, a->num_nodes, a->in_channels, ld_h, a->window, kKB, CU_TENSOR_MAP_SWIZZLE_64B) ||
      !make_map_f16(&p.map_b, a->w_pack, a->channels, (int64_t)a->heads * fp * 2, (int64_t)a->heads * fp * 2, nb / 2))
    return GVQA_ERR_CUDA;
  p.tiles = reinterpret_cast<const int4*>(a->tiles);
  p.tile_count = a->tile_count;
  p.rowptr = a->rowptr; p.col_src = a->col_src; p.alpha = a->alpha; p.node_graph = a->node_graph;
  p.logit_terms = a->logit_terms; p.a_node = a->a_node; p.a_node_parts = a->a_node_parts;
  p.a_node_part_stride = a->a_node_part_stride; p.slope = a->negative_slope;
  p.early = (a->flags & GVQA_HOP_INPUTS_OLDER_THAN_PREDECESSOR) && a->logit_terms != nullptr;
  p.ld_an = a->ld_a_node ? (int)a->ld_a_node : 2 * a->heads;
  if (a->logit_terms && (p.ld_an < 2 * a->heads || (p.ld_an & 3))) return GVQA_ERR_UNSUPPORTED;
  p.h_in = a->h_in; p.ld_h = ld_h;
  p.skip = a->skip; p.ld_skip = ld_skip;
  p.graph_bias = a->graph_bias; p.ldgb = ldgb;
  p.bias = a->bias; p.ep_scale = a->ep_scale; p.ep_shift = a->ep_shift;
  p.h_out = a->h_out;
  p.overflow = a->overflow;
  p.N = (int)a->num_nodes; p.F = a->in_channels; p.C = a->channels;
  p.n_ct = n_ct;
  p.nb = nb;
  p.ks = fp / kKB;
  p.epilogue = a->epilogue;
  p.out_scale = 1.0f / ((float)a->heads * a->w_scale);
  p.v_next = a->v_next;
  p.a_part = a->a_part;
  if (a->v_next && (!a->a_part || !aligned16(a->v_next) || !aligned16(a->a_part))) return GVQA_ERR_MISALIGNED;
  if (a->v_next && a->a_part_blocks != 3 * n_ct) return GVQA_ERR_BAD_SHAPE;
  p.trace = g_fused_trace;

  void (*kernel)(const Params) = nullptr;
  size_t smem = 0;
  int slot = 0;
#define GVQA_PICK(W, HH, S) { kernel = gat_fused_hop_kernel<W, HH>; smem = Cfg<W, HH>::kSmem; slot = S; }
  if (a->window == 128) {
    if (a->heads == 4) GVQA_PICK(128, 4, 0) else GVQA_PICK(128, 2, 1)
  } else {
    if (a->heads == 4) GVQA_PICK(256, 4, 3) else GVQA_PICK(256, 2, 4)
  }
#undef GVQA_PICK
  static bool attr_done[6] = {false, false, false, false, false, false};
  if (!attr_done[slot]) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return GVQA_ERR_CUDA;
    }
    attr_done[slot] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)kNumSMs);         // 74 pairs; the item count is device data (the plan's tile count)
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_mask() & 1) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (cudaLaunchKernelEx(&cfg, kernel, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return GVQA_ERR_CUDA;
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
