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
// HBM traffic per hop: read h (15.7 MB) once per column tile from L2, write h_out (15.7 MB).
//
// Split-precision product as in proj_gemm_f16.cu: x = hi + 2^-11 lo', three fp16 MMAs with fp32 accumulation in
// tensor memory (hi*hi alternating between two accumulators per k-slice, both lo terms in a third).
//
// Kernel structure (two-CTA pairs, tcgen05.mma.cta_group::2, M = 256: 128 rows per CTA; N = 128 columns per item):
//   warp 0       TMA producer: per k-slice of 32 input channels one stage = the [WIN x 32] fp32 window of h rows
//                (128B swizzle) + for each head the CTA's half of the [128 x (32 hi | 32 lo')] fp16 weight tile
//   warp 1       MMA issuer (leader CTA, one elected lane): per 16-channel sub-block 3 x H MMAs
//   warps 2-9    converters, thread = one destination row: sum of alpha[k,h] * (16 channels of source row k) over
//                the in-edges for all H heads at once (packed FFMA2 from shared memory), fp16 split, tcgen05.st into
//                a two-slot tensor-memory ring; two groups of four warps own the two sub-blocks of a stage
//   warps 10-17  epilogue: tcgen05.ld, TMEM released, per-warp transpose through shared memory, then the hop epilogue
//                with coalesced loads of skip / graph_bias rows and coalesced stores of h_out
// Work items = (pair of row tiles, column tile), dealt round-robin to the 74 pairs; the row tiles come from a
// per-batch plan (gvqa_gat_fused_plan: greedy packing of whole graphs into <= 128 rows; graphs larger than 128 nodes
// are cut into chunks that stage the whole graph as their window).
#include <cuda_fp16.h>
#include <string.h>

#include "tcgen05_utils.cuh"

namespace gvqa {
namespace fused {

constexpr int kBM = 128, kBN = 128, kKS = 32;
constexpr int kStages = 3;
constexpr int kConvWarps = 8, kEpiWarps = 8;
constexpr int kConvThreads = kConvWarps * 32, kEpiThreads = kEpiWarps * 32;
// Warpgroup 0 = {TMA producer, MMA issuer, two idle warps}, warpgroups 1-2 = converters, 3-4 = epilogue.  The
// register file is per SM sub-partition (5 warps each here: 96 registers per thread at launch); warpgroup 0 hands
// 64 registers per thread to the converters (setmaxnreg), whose 64 packed accumulators do not fit 96.
constexpr int kFirstConvWarp = 4;
constexpr int kFirstEpiWarp = kFirstConvWarp + kConvWarps;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);     // 640
constexpr int kRegsLean = 32, kRegsConv = 128;
constexpr int kEdgeCap = 768;                      // in-edges of a tile staged in shared memory (more: read from global)
constexpr uint32_t kBHeadBytes = 64 * 128;         // one head's half tile: 64 weight rows x (32 hi | 32 lo') fp16
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemLo = 2 * kBN, kTmemA = 3 * kBN;     // accumulators [0,128) [128,256) hi*hi, [256,384) lo terms
constexpr uint32_t kEpiStageBytes = 32 * 64;       // per epilogue warp: 32 rows x 16 columns fp32
constexpr float kLoScale = 2048.0f, kLoUnscale = 1.0f / 2048.0f;

template <int WIN, int H>
struct Cfg {
  static constexpr uint32_t kABytes = WIN * 128;                        // [WIN x 32] fp32
  static constexpr uint32_t kStageBytes = kABytes + H * kBHeadBytes;
  static constexpr uint32_t kListBytes = kEdgeCap * (4 * H + 4) + (kBM + 4) * 4;
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kListBytes + 256;
  static_assert(kSmem <= 232448, "shared memory budget of one CTA");
  static_assert(kTmemA + 2 * 16 * H <= kTmemCols, "tensor-memory budget");
};

struct Params {
  CUtensorMap map_a, map_b;
  const int4* tiles;               // {first row, rows, first window row, 0}
  const int32_t* tile_count;
  const int32_t* rowptr;
  const int32_t* col_src;
  const float* alpha;              // [E, H] softmax weights, CSR order
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
  int32_t N, F, C, Fp, n_ct, ks, epilogue;
  float inv_heads;
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
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
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

// One destination row's share of a [16-channel x H-head] A sub-block: acc[h][i] = channels (2i, 2i+1) of
// sum_k alpha[k,h] * h[src_k, 32 j + 16 c + ...].
// FAST: the tile's in-edge lists are staged in shared memory and every source row lies in the staged window.
template <int WIN, int H>
__device__ __forceinline__ void aggregate_fast(u64 (&acc)[H][8], int eb, int ee, const int32_t* src_s, const float* alpha_s,
                                               uint32_t stage_a, int c) {
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
    const uint32_t ra = stage_a + (uint32_t)L * 128u, sw = (uint32_t)L & 7u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      u64 v01, v23;
      lds_2x64(ra + ((((uint32_t)(4 * c + q)) ^ sw) << 4), v01, v23);
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
                                              uint32_t stage_a, int c, int j) {
#pragma unroll 1
  for (int e = eb; e < ee; ++e) {
    const int L = __ldg(p.col_src + e0 + e) - win0;
    u64 a2[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = __ldg(p.alpha + (int64_t)(e0 + e) * H + h);
      a2[h] = pack2(a, a);
    }
    const bool inside = (unsigned)L < (unsigned)WIN;
    const uint32_t ra = stage_a + (uint32_t)L * 128u, sw = (uint32_t)L & 7u;
    const int col = kKS * j + 16 * c;
    const float* gp = p.h_in + (int64_t)(win0 + L) * p.ld_h + col;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      u64 v01 = 0ull, v23 = 0ull;
      if (inside) {
        lds_2x64(ra + ((((uint32_t)(4 * c + q)) ^ sw) << 4), v01, v23);
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

#define GVQA_FUSED_TRACE(slot, col) \
  do { if (p.trace && blockIdx.x == 0 && (slot) < 1100) p.trace[(slot) * 8 + (col)] = clock64(); } while (0)

template <int WIN, int H>
__global__ void __launch_bounds__(kThreads, 1) gat_fused_hop_kernel(const __grid_constant__ Params p) {
  using C_ = Cfg<WIN, H>;
  constexpr uint32_t kABytes = C_::kABytes, kStageBytes = C_::kStageBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi_stage = smem + (size_t)kStages * kStageBytes;
  float* alpha_s = reinterpret_cast<float*>(epi_stage + kEpiWarps * kEpiStageBytes);       // [kEdgeCap][H]
  int32_t* src_s = reinterpret_cast<int32_t*>(alpha_s + kEdgeCap * H);                     // [kEdgeCap] window-local
  int32_t* rp_s = src_s + kEdgeCap;                                                        // [kBM + 1], then a flag word
  int32_t* far_s = rp_s + kBM + 2;                        // != 0: some source of the tile lies outside the window
  uint64_t* bars = reinterpret_cast<uint64_t*>(rp_s + kBM + 4);
  uint64_t* tma_full = bars;                    // [kStages]
  uint64_t* smem_empty = bars + kStages;        // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;       // [2]
  uint64_t* a_empty = a_ready + 2;              // [2]
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&smem_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_ready[s], 128 * 2);            // four warps of each CTA of the pair
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiThreads * 2);
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

  pdl_wait();
  pdl_launch_dependents();

  const int T = __ldg(p.tile_count);
  const int n_ct = p.n_ct, ks = p.ks;
  const int items = ((T + 1) >> 1) * n_ct;

  if (warp < kFirstConvWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLean));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int item = pair_id; item < items; item += npairs) {
        const int prt = item / n_ct, ct = item - prt * n_ct;
        const int te = 2 * prt + rank;
        const int win0 = te < T ? __ldg(&p.tiles[te].z) : 0;
        const int n0 = ct * kBN;
        const int ncols = min(kBN, (p.C - n0 + 31) & ~31);
        const int nb0 = n0 + rank * (ncols >> 1);        // this CTA stages its half of the weight tile's rows
        for (int j = 0; j < ks; ++j, ++it) {
          const int s = it % kStages;
          mbar_wait(&smem_empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* st = smem + (size_t)s * kStageBytes;
          GVQA_FUSED_TRACE(it, 0);
          mbar_expect_tx(&tma_full[s], kStageBytes);
          tma_load_2d(st, &p.map_a, &tma_full[s], kKS * j, win0);
#pragma unroll
          for (int h = 0; h < H; ++h)
            tma_load_2d(st + kABytes + h * kBHeadBytes, &p.map_b, &tma_full[s], (h * p.Fp + kKS * j) * 2, nb0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0 && elect_one()) {
      uint32_t it = 0, item_it = 0;
      const uint64_t desc0 = umma_desc(smem_u32(smem));
      const uint32_t d_lo = tmem_base + kTmemLo;
      for (int item = pair_id; item < items; item += npairs, ++item_it) {
        const int ct = item % n_ct;
        const int ncols = min(kBN, (p.C - ct * kBN + 31) & ~31);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)((kBM * 2) >> 4) << 24);
        GVQA_FUSED_TRACE(1024 + item_it, 0);
        mbar_wait(acc_empty, (item_it & 1) ^ 1);
        GVQA_FUSED_TRACE(1024 + item_it, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < ks; ++j, ++it) {
          const uint32_t s = it % kStages;
          const uint32_t d_big = tmem_base + ((j & 1) ? (uint32_t)kBN : 0u);
          const uint64_t bstage = desc0 + (uint64_t)((s * kStageBytes + kABytes) >> 4);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_wait(&a_ready[c], it & 1);        // implies tma_full[s] in both CTAs: the converters waited on it
            GVQA_FUSED_TRACE(it, 4 + c);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_slot = tmem_base + kTmemA + (uint32_t)(c * 16 * H);
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const uint64_t b_hi = bstage + (uint64_t)((h * kBHeadBytes) >> 4) + 2 * c, b_lo = b_hi + 4;
              const uint32_t a_hi = a_slot + 16 * h, a_lo = a_hi + 8;
              mma_pair(d_lo, a_lo, b_hi, idesc, (j | c | h) != 0);
              mma_pair(d_lo, a_hi, b_lo, idesc, 1);
              mma_pair(d_big, a_hi, b_hi, idesc, !(j < 2 && c == 0 && h == 0));
            }
            commit_pair(&a_empty[c]);
          }
          commit_pair(&smem_empty[s]);
          if (j == ks - 1) commit_pair(acc_full);
          GVQA_FUSED_TRACE(it, 7);
        }
      }
    }
  }                                                        // (two idle warps complete warpgroup 0)
  } else if (warp < kFirstEpiWarp) {
    // ===================== converters: thread = one destination row of the tile =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConv));
    const int quarter = warp & 3;                          // TMEM lane quarter a warp may touch = warp id % 4
    const int grp = (warp - kFirstConvWarp) >> 2;          // which 16-channel sub-block of every stage
    const int ctid = threadIdx.x - kFirstConvWarp * 32;
    const int r = quarter * 32 + lane;
    const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + kTmemA + (uint32_t)(grp * 16 * H);
    const u64 scale2 = pack2(kLoScale, kLoScale);
    uint32_t it = 0;
    float amax = 0.f;
    for (int item = pair_id; item < items; item += npairs) {
      const int prt = item / n_ct;
      const int te = 2 * prt + rank;
      int4 tile = make_int4(0, 0, 0, 0);
      if (te < T) tile = __ldg(&p.tiles[te]);
      const int row0 = tile.x, nrows = tile.y, win0 = tile.z;
      conv_bar_sync();                                     // the previous item's lists are no longer read
      if (ctid == 0) *far_s = 0;
      int e0 = 0, ne = 0;
      if (nrows > 0) {
        e0 = __ldg(p.rowptr + row0);
        ne = __ldg(p.rowptr + row0 + nrows) - e0;
      }
      for (int t = ctid; t <= nrows; t += kConvThreads) rp_s[t] = __ldg(p.rowptr + row0 + t) - e0;
      const bool staged = ne <= kEdgeCap;
      if (staged) {
        int far = 0;
        for (int k = ctid; k < ne; k += kConvThreads) {
          const int L = __ldg(p.col_src + e0 + k) - win0;
          src_s[k] = L;
          far |= (unsigned)L >= (unsigned)WIN;
#pragma unroll
          for (int h = 0; h < H; ++h) alpha_s[k * H + h] = __ldg(p.alpha + (int64_t)(e0 + k) * H + h);
        }
        if (far) *far_s = 1;                               // (benign race: every writer stores 1)
      }
      conv_bar_sync();
      const bool fast = staged && *far_s == 0;
      int eb = 0, ee = 0;
      if (r < nrows) {
        eb = rp_s[r];
        ee = rp_s[r + 1];
      }
      for (int j = 0; j < ks; ++j, ++it) {
        const int s = it % kStages;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 1);
        const uint32_t stage_a = smem_u32(smem + (size_t)s * kStageBytes);
        u64 acc[H][8];
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[h][i] = 0ull;
        if (fast) aggregate_fast<WIN, H>(acc, eb, ee, src_s, alpha_s, stage_a, grp);
        else aggregate_any<WIN, H>(acc, eb, ee, p, e0, win0, stage_a, grp, j);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 2);
        mbar_wait(&a_empty[grp], (it & 1) ^ 1);            // the MMAs of the previous stage's sub-block have read the slot
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int h = 0; h < H; ++h) {
          uint32_t pk[16];                                 // [0,8) hi, [8,16) lo'; two k elements per word
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x0, x1, d0, d1;
            unpack2(acc[h][i], x0, x1);
            amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
            const uint32_t h2 = pack_half2(x0, x1);
            const float2 f = unpack_half2(h2);
            const u64 d = fmul2(fsub2(acc[h][i], pack2(f.x, f.y)), scale2);
            unpack2(d, d0, d1);
            pk[i] = h2;
            pk[8 + i] = pack_half2(d0, d1);
          }
          GVQA_TMEM_ST16(ta + 16 * h, pk, 0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_cta(&a_ready[grp], 0);
        if (threadIdx.x == kFirstConvWarp * 32) GVQA_FUSED_TRACE(it, 3);
      }
    }
    if (p.overflow != nullptr && !(amax <= 65000.0f)) atomicOr(p.overflow, 1);   // also catches NaN / inf
  } else {
    // ===================== epilogue: thread = one accumulator row, 64 of the item's 128 columns =====================
    const int quarter = warp & 3;
    const int chalf = (warp - kFirstEpiWarp) >> 2;
    const uint32_t stage = smem_u32(epi_stage + (size_t)(warp - kFirstEpiWarp) * kEpiStageBytes);
    uint32_t item_it = 0;
    for (int item = pair_id; item < items; item += npairs, ++item_it) {
      const int prt = item / n_ct, ct = item - prt * n_ct;
      const int te = 2 * prt + rank;
      int4 tile = make_int4(0, 0, 0, 0);
      if (te < T) tile = __ldg(&p.tiles[te]);
      const int row0 = tile.x, nrows = tile.y;
      const int r = quarter * 32 + lane;
      int gid = 0, has_in = 0;
      if (r < nrows) {                                     // requested before the accumulators are ready
        gid = __ldg(p.node_graph + row0 + r);
        has_in = __ldg(p.rowptr + row0 + r + 1) > __ldg(p.rowptr + row0 + r);
      }
      const int col0 = ct * kBN + chalf * 64;
      const bool live = col0 < p.C;
      mbar_wait(acc_full, item_it & 1);
      if (threadIdx.x == kFirstEpiWarp * 32) GVQA_FUSED_TRACE(1024 + item_it, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float acc[64];
      if (live) {
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64);
        const bool two = ks >= 2;
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          uint32_t rs[16], r0[16];
          GVQA_TMEM_LD16(rs, tcol + kTmemLo + (uint32_t)(pc * 16));
          GVQA_TMEM_LD16(r0, tcol + (uint32_t)(pc * 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[pc * 16 + e] = __uint_as_float(r0[e]);
          if (two) {
            GVQA_TMEM_LD16(r0, tcol + kBN + (uint32_t)(pc * 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[pc * 16 + e] += __uint_as_float(r0[e]);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[pc * 16 + e] = fmaf(__uint_as_float(rs[e]), kLoUnscale, acc[pc * 16 + e]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive_cta(acc_empty, 0);                       // TMEM is free: the next item's MMAs may start
      if (threadIdx.x == kFirstEpiWarp * 32) GVQA_FUSED_TRACE(1024 + item_it, 3);
      if (live) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int cpass = col0 + pass * 16;
          if (cpass < p.C) {                               // warp-uniform
#pragma unroll
            for (int q = 0; q < 4; ++q)
              sts128(stage + (uint32_t)lane * 64u + (uint32_t)((q ^ ((lane >> 1) & 3)) << 4), acc[pass * 16 + 4 * q],
                     acc[pass * 16 + 4 * q + 1], acc[pass * 16 + 4 * q + 2], acc[pass * 16 + 4 * q + 3]);
            __syncwarp();
            const int q = lane & 3, col = cpass + 4 * q;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int row = i4 * 8 + (lane >> 2);
              float4 o = lds128(stage + (uint32_t)row * 64u + (uint32_t)((q ^ ((row >> 1) & 3)) << 4));
              const int g = __shfl_sync(kFull, gid, row), hin = __shfl_sync(kFull, has_in, row);
              const int tr = quarter * 32 + row;
              if (tr < nrows && col < p.C) {
                const int64_t grow = row0 + tr;
                o.x *= p.inv_heads; o.y *= p.inv_heads; o.z *= p.inv_heads; o.w *= p.inv_heads;
                if (p.graph_bias && hin) {
                  const float4 gb = __ldg(reinterpret_cast<const float4*>(p.graph_bias + (int64_t)g * p.ldgb + col));
                  o.x += gb.x; o.y += gb.y; o.z += gb.z; o.w += gb.w;
                }
                if (p.bias) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                  o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (p.skip) {
                  const float4 sk = ldg_stream(p.skip + grow * p.ld_skip + col);
                  o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;
                }
                if (p.epilogue == GVQA_EPI_AFFINE || p.epilogue == GVQA_EPI_AFFINE_RELU) {
                  const float4 sc = __ldg(reinterpret_cast<const float4*>(p.ep_scale + col));
                  const float4 sh = __ldg(reinterpret_cast<const float4*>(p.ep_shift + col));
                  o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y);
                  o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
                  if (p.epilogue == GVQA_EPI_AFFINE_RELU) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                  }
                }
                *reinterpret_cast<float4*>(p.h_out + grow * p.C + col) = o;
              }
            }
            __syncwarp();
          }
        }
      }
      if (threadIdx.x == kFirstEpiWarp * 32) GVQA_FUSED_TRACE(1024 + item_it, 4);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();         // neither CTA leaves (or frees tensor memory) while its peer may still signal or read it
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

__global__ void fused_plan_kernel(const int32_t* __restrict__ graph_ptr, int B, int win, int4* tiles, int32_t* count,
                                  int max_tiles) {
  extern __shared__ int32_t gp_s[];
  for (int i = threadIdx.x; i <= B; i += blockDim.x) gp_s[i] = graph_ptr[i];
  __syncthreads();
  if (threadIdx.x == 0) *count = plan_tiles(gp_s, B, win, tiles, max_tiles);
}

// ---- softmax weights of all in-edges (gat_skip.py:183-192 + PyG utils.softmax), CSR order ---------------------------
template <int H>
__global__ void __launch_bounds__(256) gat_alpha_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col_src,
                                                        const int32_t* __restrict__ perm, const int32_t* __restrict__ node_graph,
                                                        const float* __restrict__ a_node, int64_t lda,
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
  float tg = a_graph ? a_graph[(int64_t)node_graph[i] * ldag + h] : 0.f;
  tg += a_node[(int64_t)i * lda + H + h];
  auto logit = [&](int k) {
    const int64_t e = perm ? perm[k] : k;
    const float v = (a_edge[e * lde + h] + a_node[(int64_t)col_src[k] * lda + h]) + tg;
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

// ---- weight prepack: W [H*C, >= F] fp32 (row h*C + c) -> [C, H * Fp * 2] fp16, per 32 input channels 32 hi | 32 lo' ----
__global__ void fused_pack_kernel(const float* __restrict__ w, int64_t ldw, int H, int C, int F, int Fp,
                                  __half* __restrict__ out) {
  const int64_t total = (int64_t)C * H * Fp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ((int64_t)H * Fp);
    const int rem = (int)(i - c * H * Fp), h = rem / Fp, k = rem - h * Fp;
    const float x = k < F ? w[((int64_t)h * C + c) * ldw + k] : 0.f;
    const __half hi = __float2half_rn(x);
    const int64_t o = c * ((int64_t)H * Fp * 2) + (int64_t)((h * Fp + k) >> 5) * 64 + (k & 31);
    out[o] = hi;
    out[o + 32] = __float2half_rn((x - __half2float(hi)) * kLoScale);
  }
}

unsigned long long* g_fused_trace = nullptr;

}  // namespace fused
}  // namespace gvqa

using namespace gvqa;

extern "C" GVQA_API void gvqa_debug_set_fused_trace(unsigned long long* buf) { fused::g_fused_trace = buf; }

extern "C" GVQA_API int64_t gvqa_gat_fused_max_tiles(int64_t num_nodes, int64_t num_graphs) {
  return num_graphs + (num_nodes + fused::kBM - 1) / fused::kBM + 2;
}

extern "C" GVQA_API int32_t gvqa_gat_fused_window(int32_t max_nodes_per_graph) {
  return max_nodes_per_graph > fused::kBM ? 256 : 128;
}

extern "C" GVQA_API int gvqa_gat_fused_plan(const int32_t* graph_ptr, int64_t num_graphs, int32_t window, int32_t* tiles,
                                            int32_t* count, int64_t max_tiles, void* stream_) {
  if (num_graphs < 0 || max_tiles < 0 || (window != 128 && window != 256)) return GVQA_ERR_BAD_SHAPE;
  if (!graph_ptr || !tiles || !count) return GVQA_ERR_NULL_POINTER;
  if (!aligned16(tiles)) return GVQA_ERR_MISALIGNED;
  if ((num_graphs + 1) * 4 > 200 * 1024) return GVQA_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(num_graphs + 1) * 4;
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
  fused::fused_plan_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream_)>>>(
      graph_ptr, (int)num_graphs, window, reinterpret_cast<int4*>(tiles), count, (int)max_tiles);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
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
                                           const float* a_edge, int64_t lde, const float* a_graph, int64_t ld_a_graph,
                                           float negative_slope, int64_t num_nodes, int32_t heads, float* alpha,
                                           float* alpha_out, void* stream_) {
  if (num_nodes < 0 || num_nodes >= (1ll << 28)) return GVQA_ERR_BAD_SHAPE;
  if (num_nodes == 0) return GVQA_OK;
  if (!rowptr || !col_src || !a_node || !a_edge || !alpha || (a_graph && !node_graph)) return GVQA_ERR_NULL_POINTER;
  const int64_t threads = num_nodes * heads;
  const dim3 grid((unsigned)((threads + 255) / 256)), block(256);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
#define GVQA_ALPHA(HH)                                                                                                  \
  if (launch_pdl(2, fused::gat_alpha_kernel<HH>, grid, block, 0, st, rowptr, col_src, perm, node_graph, a_node, ld_a_node, \
                 a_edge, lde, a_graph, ld_a_graph, negative_slope, (int)num_nodes, alpha, alpha_out) != cudaSuccess) {  \
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

extern "C" GVQA_API int64_t gvqa_gat_fused_pack_halves(int32_t heads, int32_t channels, int32_t in_channels) {
  const int64_t fp = (in_channels + 31) & ~31;
  return (int64_t)channels * heads * fp * 2;
}

extern "C" GVQA_API int gvqa_gat_fused_pack_f16(const float* w, int64_t ldw, int32_t heads, int32_t channels,
                                                int32_t in_channels, void* packed, void* stream_) {
  if (heads <= 0 || channels <= 0 || in_channels <= 0 || ldw < in_channels) return GVQA_ERR_BAD_SHAPE;
  if (!w || !packed) return GVQA_ERR_NULL_POINTER;
  const int fp = (in_channels + 31) & ~31;
  const int64_t total = (int64_t)channels * heads * fp;
  const int64_t blocks = (total + 255) / 256;
  fused::fused_pack_kernel<<<(unsigned)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      w, ldw, heads, channels, in_channels, fp, static_cast<__half*>(packed));
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_gat_fused_supported(int32_t heads, int32_t in_channels, int32_t channels) {
  return (heads == 4 || heads == 2 || heads == 1) && in_channels > 0 && (in_channels & 3) == 0 && channels > 0 &&
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
      !a->h_out)
    return GVQA_ERR_NULL_POINTER;
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
  Params p;
  memset(&p, 0, sizeof(p));
  const int fp = (a->in_channels + 31) & ~31;
  if (!make_map(&p.map_a, a->h_in, a->num_nodes, a->in_channels, ld_h, a->window, 32) ||
      !make_map_f16(&p.map_b, a->w_pack, a->channels, (int64_t)a->heads * fp * 2, (int64_t)a->heads * fp * 2, 64))
    return GVQA_ERR_CUDA;
  p.tiles = reinterpret_cast<const int4*>(a->tiles);
  p.tile_count = a->tile_count;
  p.rowptr = a->rowptr; p.col_src = a->col_src; p.alpha = a->alpha; p.node_graph = a->node_graph;
  p.h_in = a->h_in; p.ld_h = ld_h;
  p.skip = a->skip; p.ld_skip = ld_skip;
  p.graph_bias = a->graph_bias; p.ldgb = ldgb;
  p.bias = a->bias; p.ep_scale = a->ep_scale; p.ep_shift = a->ep_shift;
  p.h_out = a->h_out;
  p.overflow = a->overflow;
  p.N = (int)a->num_nodes; p.F = a->in_channels; p.C = a->channels; p.Fp = fp;
  p.n_ct = (a->channels + kBN - 1) / kBN;
  p.ks = fp / kKS;
  p.epilogue = a->epilogue;
  p.inv_heads = 1.0f / (float)a->heads;
  p.trace = g_fused_trace;

  void (*kernel)(const Params) = nullptr;
  size_t smem = 0;
  int slot = 0;
#define GVQA_PICK(W, HH, S) { kernel = gat_fused_hop_kernel<W, HH>; smem = Cfg<W, HH>::kSmem; slot = S; }
  if (a->window == 128) {
    if (a->heads == 4) GVQA_PICK(128, 4, 0) else if (a->heads == 2) GVQA_PICK(128, 2, 1) else GVQA_PICK(128, 1, 2)
  } else {
    if (a->heads == 4) GVQA_PICK(256, 4, 3) else if (a->heads == 2) GVQA_PICK(256, 2, 4) else GVQA_PICK(256, 1, 5)
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
