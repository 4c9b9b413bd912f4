// fp32-accurate node projection with fp16 tensor-core operands (see include/gvqa_b200.h:
// gvqa_proj_gemm_3xf16, gvqa_split_f16).  Same job and same structure as proj_gemm.cu
// (C[M,N] = A[M,K] @ B[N,K]^T for the reference's x_l = lin_l(x_cat), gat_skip.py:133), different split:
//
//   x = hi + lo,  hi = fp16(x),  lo = x - hi (exact in fp32),  lo' = fp16(lo * 2^11)
//   C = A_hi*B_hi  +  2^-11 * (A_hi*B_lo' + A_lo'*B_hi)              (the lo*lo term is ~2^-22 relative, dropped)
//
// fp16 carries the same 11 significant bits as tf32, so the three-product result has the accuracy of the
// 3xTF32 kernel, but tcgen05.mma.kind::f16 consumes K = 16 per instruction where kind::tf32 consumes K = 8
// (profiles/r01/mma_rate_elect.txt: 4096 vs 2048 FMA/clk/SM), and the fp16 B tiles are half the bytes of the
// tf32 ones, which matters because the tf32 kernel runs close to the SM's shared-memory bandwidth.  The lo
// parts are scaled by 2^11 so they stay in fp16's normal range; the scale is undone once, in the epilogue.
//
// Range: fp16 tops out at 65504.  The converters track max|A| and raise a sticky flag (`overflow`, optional)
// when an element would not fit; callers fall back to gvqa_proj_gemm_3xtf32 for such inputs.  GraphVQA's node
// states are graph-LayerNorm / BatchNorm outputs of order 1-10.
//
// One k-block = 64 K-elements: A raw [128 x 64] fp32 (two 128B-swizzled [128 x 32] boxes), B_hi and B_lo'
// [128 x 64] fp16 (one 128B-swizzled box each) = 64 KB per stage, 3 stages.  Warp roles, TMEM map, the early
// TMEM release and the TMA-store epilogue are those of proj_gemm.cu.
#include <cuda_fp16.h>
#include <string.h>

#include "tcgen05_utils.cuh"

namespace gvqa {
namespace f16gemm {

constexpr int kBM = 128, kBN = 128, kBK = 64;       // tile: rows of A, rows of B, k elements per k-block
constexpr int kAStages = 4;                         // tensor-memory ring of split A sub-blocks (32 k-elements each)                         // tensor-memory ring of split A tiles
// warp 0 TMA, warp 1 MMA, then 4 x CH converter warps (CH = 1: a thread converts a whole 64-float row of the
// k-block; CH = 2: two warps per TMEM lane quarter, one [128 x 32] box each), then 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr uint32_t kAHalfBytes = kBM * 32 * 4;      // one [128 x 32] fp32 box = 16 KB
constexpr uint32_t kABytes = 2 * kAHalfBytes;       // 32 KB
// PAIR = 1: one CTA per 128 x 128 tile.  PAIR = 2: a two-CTA cluster (the two SMs of a TPC) owns a 256 x 128 tile;
// each CTA loads and converts its own 128 rows of A but only HALF of the B tile, and the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), which reads both halves.  That removes 24 of the 144 KB of shared-memory
// traffic per k-block and a quarter of the L2 -> SM traffic, and the smaller stages allow a 4-deep ring
// (profiles/r01/gemm_f16_stage_attribution.txt: 49.2 -> 46.9 us at cfg2; step 0.381 -> 0.369 ms).
template <int PAIR>
struct Cfg {
  static constexpr int kBRows = kBN / PAIR;                              // B rows a CTA stages per k-block
  static constexpr uint32_t kBBytes = kBRows * kBK * 2;                  // 16 KB / 8 KB
  static constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;         // 64 KB / 48 KB
  static constexpr int kStages = PAIR == 2 ? 4 : 3;                      // shared-memory ring (192 KB either way)
};
constexpr uint32_t kTmemCols = 512;
// accumulators: [0,128) and [128,256) hi*hi, [256,384) the lo terms (scaled by 2^11).
// DB = 0: the two hi*hi accumulators hold the two K-halves of one tile (a shorter chain halves the bias of the tensor
//         core's truncating accumulate); the epilogue of a tile and the MMAs of the next one alternate.
// DB = 1: they hold CONSECUTIVE TILES: the MMAs of tile t+1 start on the other accumulator while the epilogue still
//         reads tile t (only the lo accumulator is shared: the epilogue reads it first, and the first k-block of
//         a tile issues its hi*hi products before it waits for it).  One hi*hi chain per tile: used for K <= 512,
//         where it stays inside the accuracy bar (profiles/r01/gemm_f16_merge_probe.txt).
constexpr uint32_t kTmemSmall = 2 * kBN;
// A ring: 4 sub-blocks x (16 columns hi | 16 columns lo'), two fp16 per 32-bit column.  A sub-block is one
// [128 x 32] fp32 box of a stage (the MMAs of a sub-block start as soon as its box is converted).  Measured equal
// to a ring of two whole k-blocks (46.9 vs 46.8 us at cfg2, profiles/r01/gemm_f16_stage_attribution.txt).
constexpr uint32_t kTmemA = 3 * kBN;
constexpr uint32_t kASubCols = 32;                  // TMEM columns per sub-block (hi at +0, lo' at +16)
constexpr uint32_t kEpiStageBytes = 32 * 32 * 4;
constexpr size_t kGemmSmem = (size_t)3 * 65536 + (kEpiThreads / 32) * kEpiStageBytes + 1024 + 256;
static_assert(Cfg<1>::kStages * Cfg<1>::kStageBytes == 3 * 65536 && Cfg<2>::kStages * Cfg<2>::kStageBytes == 3 * 65536,
              "both configurations use the same 192 KB operand ring");
constexpr float kLoScale = 2048.0f, kLoUnscale = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {   // a -> low half (lower k), b -> high half
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

__device__ __forceinline__ void cluster_sync_all() {      // every thread of both CTAs; also a CTA-wide barrier
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// arrive on the mbarrier that sits at the same shared-memory offset in CTA `cta` of the cluster
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

template <int PAIR>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  if constexpr (PAIR == 2) mbar_arrive_cta(bar, 0);
  else mbar_arrive(bar);
}

// completion of all MMAs issued so far -> the barrier at this offset (PAIR = 2: in BOTH CTAs of the pair)
template <int PAIR>
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
  if constexpr (PAIR == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  } else {
    umma_commit(bar);
  }
}

template <int PAIR>
__device__ __forceinline__ void umma_f16_ts_x(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  if constexpr (PAIR == 2) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_f16_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
  }
}

// Up to kMaxProblems independent products share one persistent launch (tiles of problem 0 first, then 1, ...):
// gat_seq issues hop 0's projection together with the two small pre-pass products, whose ~170 short tiles then
// fill the last, partially empty wave instead of costing two more launches.
constexpr int kMaxProblems = 3;

struct alignas(64) GroupedParams {
  CUtensorMap map_a[kMaxProblems], map_bhi[kMaxProblems], map_blo[kMaxProblems], map_c[kMaxProblems];
  int M[kMaxProblems], N[kMaxProblems], K[kMaxProblems];
  int tiles_per_batch[kMaxProblems], n_tiles[kMaxProblems];
  int tile_end[kMaxProblems];      // exclusive prefix of tiles
  const float* bias[kMaxProblems]; // optional epilogue: C[m, n] += bias[n], then ReLU when relu != 0
  int relu[kMaxProblems];
  int count;
  long long* trace;   // debug only (gvqa_debug_set_gemm_trace): CTA 0 records clock64() per k-block and role
  int dbg;   // debug only (gvqa_debug_set_gemm_flags): bit0 no TMA loads, bit1 converters idle, bit2 no epilogue, bit3 no MMAs
};

struct TileInfo {
  int p, z, m0, n0, kblocks;
};

template <int PAIR>
__device__ __forceinline__ TileInfo decode_tile(const GroupedParams& g, int tile, int rank) {
  TileInfo t;
  t.p = tile < g.tile_end[0] ? 0 : (tile < g.tile_end[1] ? 1 : 2);
  const int local = tile - (t.p == 0 ? 0 : g.tile_end[t.p - 1]);
  t.z = local / g.tiles_per_batch[t.p];
  const int t2 = local - t.z * g.tiles_per_batch[t.p];
  t.m0 = (t2 / g.n_tiles[t.p]) * (kBM * PAIR) + rank * kBM;   // PAIR = 2: the pair's tile is 256 rows, 128 per CTA
  t.n0 = (t2 % g.n_tiles[t.p]) * kBN;
  t.kblocks = (g.K[t.p] + kBK - 1) / kBK;          // the ragged last k-block is zero-filled by TMA
  return t;
}

// k-blocks of a tile without the divisions of decode_tile (converters only need this)
__device__ __forceinline__ int tile_kblocks(const GroupedParams& g, int tile) {
  const int p = tile < g.tile_end[0] ? 0 : (tile < g.tile_end[1] ? 1 : 2);
  return (g.K[p] + kBK - 1) / kBK;
}

#define GVQA_MAP(field, p) ((p) == 0 ? &g.field[0] : ((p) == 1 ? &g.field[1] : &g.field[2]))

template <int CH, int PAIR, int DB>
__global__ void __launch_bounds__(64 + 128 * CH + kEpiThreads, 1)
proj_gemm_3xf16_kernel(const __grid_constant__ GroupedParams g, int32_t* __restrict__ overflow) {
  // all tensor maps are 3-D [batch, rows, K]; a tile index decomposes into (problem, batch z, row tile, column tile)
  // trace rows: [it][0] producer past smem_empty, [1] converter past tma_full, [2] past a_empty (first box), [3] done,
  // [4] MMA loop top, [5] past a_ready (first box), [6] past a_ready (second box), [7] committed;
  // [1024 + tile][0] MMA before acc_empty, [1] after, [2] epilogue past acc_full, [3] TMEM released, [4] stores issued
  // (profiles/microbench/trace_gemm_f16.py)
#define GVQA_F16_TRACE(slot, col) \
  do { if (g.trace && blockIdx.x == 0 && (slot) < 1100) g.trace[(slot) * 8 + (col)] = clock64(); } while (0)
  constexpr int kStages = Cfg<PAIR>::kStages;
  constexpr uint32_t kStageBytes = Cfg<PAIR>::kStageBytes, kBBytes = Cfg<PAIR>::kBBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi_stage = smem + (size_t)kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (kEpiThreads / 32) * kEpiStageBytes);
  uint64_t* tma_full = bars;                        // [kStages]
  uint64_t* smem_empty = bars + kStages;            // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;           // [kAStages]
  uint64_t* a_empty = a_ready + kAStages;           // [kAStages]
  uint64_t* acc_full = a_empty + kAStages;
  uint64_t* acc_empty = acc_full + 1;               // DB = 0: all three accumulators read
  uint64_t* small_empty = acc_empty + 1;            // DB = 1: the lo accumulator read
  uint64_t* big_empty = small_empty + 1;            // [2] DB = 1: hi*hi accumulator b read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(big_empty + 2);

  constexpr int kFirstEpiWarp = 2 + 4 * CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = g.tile_end[g.count - 1];
  const int dbg = g.dbg;
  // PAIR = 2: `rank` is the CTA's place in its pair (0 = leader: owns a_ready / acc_empty and issues the MMAs);
  // tiles are dealt to pairs, not CTAs
  const int rank = PAIR == 2 ? (int)cluster_ctarank() : 0;
  const int first_tile = PAIR == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&smem_empty[s], 1);
    }
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_ready[s], 128 * PAIR);             // a sub-block is converted by one warp per lane quarter (per CTA)
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiThreads * PAIR);
    mbar_init(small_empty, kEpiThreads * PAIR);
    mbar_init(&big_empty[0], kEpiThreads * PAIR);
    mbar_init(&big_empty[1], kEpiThreads * PAIR);
    mbar_fence_init();
  }
  if (warp == 1) {                                         // PAIR = 2: the same warp of both CTAs allocates collectively
    if constexpr (PAIR == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR == 2) cluster_sync_all();             // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();                 // set-up above overlaps the previous kernel's tail (see common.cuh)
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer: one elected lane runs the whole loop =====================
    if (elect_one()) {
      uint32_t it = 0;
      // the next tile is decoded (two integer divisions, ~700 cycles with the parameter loads) right after the first
      // k-block of the current one has been issued, not at the tile boundary where every role would wait for it
      TileInfo t = decode_tile<PAIR>(g, first_tile < num_tiles ? first_tile : 0, rank), tn = t;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, t = tn) {
        const CUtensorMap *ma = GVQA_MAP(map_a, t.p), *mh = GVQA_MAP(map_bhi, t.p), *ml = GVQA_MAP(map_blo, t.p);
        // PAIR = 2: this CTA stages rows [rank * ncols / 2, +ncols / 2) of the B tile (the box is 64 rows; the MMA reads
        // only the first ncols / 2 of them)
        const int nb0 = PAIR == 2 ? t.n0 + rank * (min(kBN, (g.N[t.p] - t.n0 + 31) & ~31) >> 1) : t.n0;
        for (int kb = 0; kb < t.kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&smem_empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* st = smem + (size_t)s * kStageBytes;
          GVQA_F16_TRACE(it, 0);
          if (dbg & 1) {
            mbar_arrive(&tma_full[s]);
          } else {
            mbar_expect_tx(&tma_full[s], kStageBytes);
            tma_load_3d(st, ma, &tma_full[s], kb * kBK, t.m0, t.z);
            tma_load_3d(st + kAHalfBytes, ma, &tma_full[s], kb * kBK + 32, t.m0, t.z);
            tma_load_3d(st + kABytes, mh, &tma_full[s], kb * kBK, nb0, t.z);
            tma_load_3d(st + kABytes + kBBytes, ml, &tma_full[s], kb * kBK, nb0, t.z);
          }
          if (kb == 0 && tile + tile_step < num_tiles) tn = decode_tile<PAIR>(g, tile + tile_step, rank);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected lane runs the whole loop =====================
    if ((PAIR == 1 || rank == 0) && elect_one()) {
      uint32_t it = 0, tile_it = 0;
      const uint64_t desc0 = umma_desc(smem_u32(smem));   // descriptor of (base + c) == desc0 + (c >> 4)
      TileInfo t = decode_tile<PAIR>(g, first_tile < num_tiles ? first_tile : 0, rank), tn = t;   // pipelined like the producer's
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++tile_it, t = tn) {
        // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, M = 128, N = tile width (x16)
        const int kblocks = t.kblocks;
        const bool has_next = tile + tile_step < num_tiles;
        const int ncols = PAIR == 2 ? min(kBN, (g.N[t.p] - t.n0 + 31) & ~31) : min(kBN, (g.N[t.p] - t.n0 + 15) & ~15);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)((kBM * PAIR) >> 4) << 24);
        const uint32_t d_small = tmem_base + kTmemSmall;
        // One k-block's MMAs.  parts: bit 0 = hi*hi into d_big, bit 1 = the two lo products into d_small.  The A
        // sub-blocks are awaited when wait_a and handed back (with the B stage) when release.
        auto issue_kblock = [&](uint32_t itx, int kb, uint32_t d_big, bool big_first, int parts, bool wait_a, bool release) {
          const uint32_t s = itx % kStages;
          const uint64_t b_hi = desc0 + (uint64_t)((s * kStageBytes + kABytes) >> 4), b_lo = b_hi + (kBBytes >> 4);
#pragma unroll
          for (int half = 0; half < 2; ++half) {          // the two 32-k sub-blocks of the stage
            const uint32_t sub = 2 * itx + half, ss = sub & (kAStages - 1);
            if (wait_a) {
              mbar_wait(&a_ready[ss], (sub / kAStages) & 1);  // implies tma_full[s]: the converters waited on it
              GVQA_F16_TRACE(itx, 5 + half);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t a_hi = tmem_base + kTmemA + ss * kASubCols, a_lo = a_hi + kASubCols / 2;
            if (!(dbg & 8)) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {            // K = 16 per MMA: 8 TMEM columns of A, 32 bytes of B
                const int k = 2 * half + kk;
                if (parts & 2) {
                  umma_f16_ts_x<PAIR>(d_small, a_lo + 8 * kk, b_hi + 2 * k, idesc, (kb | k) != 0);
                  umma_f16_ts_x<PAIR>(d_small, a_hi + 8 * kk, b_lo + 2 * k, idesc, 1);
                }
                if (parts & 1) umma_f16_ts_x<PAIR>(d_big, a_hi + 8 * kk, b_hi + 2 * k, idesc, !(big_first && k == 0));
              }
            }
            if (release) umma_commit_to<PAIR>(&a_empty[ss]);
          }
          if (release) {
            umma_commit_to<PAIR>(&smem_empty[s]);
            if (kb == kblocks - 1) umma_commit_to<PAIR>(acc_full);
            GVQA_F16_TRACE(itx, 7);
          }
        };
        GVQA_F16_TRACE(1024 + tile_it, 0);
        if constexpr (DB == 0) {
          const int half_kb = (kblocks + 1) / 2;          // first k-block of the second K-half
          mbar_wait(acc_empty, (tile_it & 1) ^ 1);
          GVQA_F16_TRACE(1024 + tile_it, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            GVQA_F16_TRACE(it, 4);
            issue_kblock(it, kb, tmem_base + (kb >= half_kb ? (uint32_t)kBN : 0u), kb == 0 || kb == half_kb, 3, true, true);
            if (kb == 0 && has_next) tn = decode_tile<PAIR>(g, tile + tile_step, rank);
          }
        } else {
          const uint32_t buf = tile_it & 1, d_big = tmem_base + buf * kBN;
          // k-blocks whose hi*hi products are issued before the lo accumulator is known to be free.  One: its A
          // sub-blocks were converted while the previous tile's last MMAs ran (the ring slots of the second k-block
          // are still held by those MMAs; waiting for it here measured a 2400-cycle bubble per tile).
          const int pre = 1;
          mbar_wait(&big_empty[buf], ((tile_it >> 1) & 1) ^ 1);      // read by the epilogue two tiles ago
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < pre; ++kb) {
            GVQA_F16_TRACE(it + kb, 4);
            issue_kblock(it + kb, kb, d_big, kb == 0, 1, true, false);
          }
          mbar_wait(small_empty, (tile_it & 1) ^ 1);      // the previous tile's lo terms have been read
          GVQA_F16_TRACE(1024 + tile_it, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < pre; ++kb) issue_kblock(it + kb, kb, d_big, false, 2, false, true);
          if (has_next) tn = decode_tile<PAIR>(g, tile + tile_step, rank);
          for (int kb = pre; kb < kblocks; ++kb) {
            GVQA_F16_TRACE(it + kb, 4);
            issue_kblock(it + kb, kb, d_big, false, 3, true, true);
          }
          it += kblocks;
        }
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ===================== converters: thread = one row of the A tile (CH = 2: of one of its two boxes) =========
    const int quarter = warp & 3;                          // TMEM lane quarter a warp may touch = warp id % 4
    const int half0 = CH == 2 ? (warp - 2) >> 2 : 0;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0;
    float amax = 0.f;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int kblocks = tile_kblocks(g, tile);
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        if (threadIdx.x == 64) GVQA_F16_TRACE(it, 1);
        const uint32_t row_addr = smem_u32(smem + (size_t)s * kStageBytes) + (uint32_t)r * 128u;
#pragma unroll
        for (int half = half0; half < half0 + 2 / CH; ++half) {   // the [128 x 32] fp32 boxes of the stage this warp owns
          const uint32_t sub = 2 * it + half, ss = sub & (kAStages - 1);
          mbar_wait(&a_empty[ss], ((sub / kAStages) & 1) ^ 1);
          if (threadIdx.x == 64 && half == 0) GVQA_F16_TRACE(it, 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (!(dbg & 2)) {
            const uint32_t ta = tmem_base + lane_base + kTmemA + ss * kASubCols;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) {         // 128B swizzle: 16-byte chunk c lives at c ^ (row & 7)
              const float4 v = lds128(row_addr + half * kAHalfBytes + (uint32_t)((cidx ^ (r & 7)) * 16));
              amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
              const uint32_t h01 = pack_half2(v.x, v.y), h23 = pack_half2(v.z, v.w);
              const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
              hi[2 * cidx] = h01;
              hi[2 * cidx + 1] = h23;
              lo[2 * cidx] = pack_half2((v.x - f01.x) * kLoScale, (v.y - f01.y) * kLoScale);
              lo[2 * cidx + 1] = pack_half2((v.z - f23.x) * kLoScale, (v.w - f23.y) * kLoScale);
            }
            GVQA_TMEM_ST16(ta, hi, 0);
            GVQA_TMEM_ST16(ta + kASubCols / 2, lo, 0);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          }
          mbar_arrive_leader<PAIR>(&a_ready[ss]);
        }
        if (threadIdx.x == 64) GVQA_F16_TRACE(it, 3);
      }
    }
    if (overflow != nullptr && !(amax <= 65000.0f)) atomicOr(overflow, 1);   // also catches NaN / inf inputs
  } else {
    // ===================== epilogue (8 warps): TMEM -> registers, release TMEM, then TMA store =============
    const int quarter = warp & 3;
    const int chalf = warp >= kFirstEpiWarp + 4 ? 1 : 0;
    uint32_t tile_it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++tile_it) {
      const TileInfo t = decode_tile<PAIR>(g, tile, rank);
      const int z = t.z, m0 = t.m0, n0 = t.n0, N = g.N[t.p], kblocks = t.kblocks;
      const CUtensorMap* mc = GVQA_MAP(map_c, t.p);
      mbar_wait(acc_full, tile_it & 1);
      if (threadIdx.x == kFirstEpiWarp * 32) GVQA_F16_TRACE(1024 + tile_it, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int col0 = n0 + chalf * 64;
      float acc[64];
      const bool live = col0 < N && !(dbg & 4);
      const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64);
      if constexpr (DB == 0) {
        const bool two_chunks = kblocks >= 2;              // with a single k-block the second K-half is never written
        if (live) {
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const uint32_t taddr = tcol + (uint32_t)(pc * 16);
            uint32_t rs[16], r0[16], r1[16];
            GVQA_TMEM_LD16(rs, taddr + kTmemSmall);
            GVQA_TMEM_LD16(r0, taddr);
            if (two_chunks) GVQA_TMEM_LD16(r1, taddr + kBN);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float big = two_chunks ? __uint_as_float(r0[e]) + __uint_as_float(r1[e]) : __uint_as_float(r0[e]);
              acc[pc * 16 + e] = fmaf(__uint_as_float(rs[e]), kLoUnscale, big);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_leader<PAIR>(acc_empty);               // TMEM is free: the next tile's MMAs may start
        if (threadIdx.x == kFirstEpiWarp * 32) GVQA_F16_TRACE(1024 + tile_it, 3);
      } else {
        // the lo accumulator first: it is the only one the next tile (already running on the other hi*hi accumulator)
        // is waiting for
        if (live) {
          uint32_t lo_bits[64];
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) GVQA_TMEM_LD16((lo_bits + pc * 16), tcol + kTmemSmall + (uint32_t)(pc * 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 64; ++e) acc[e] = __uint_as_float(lo_bits[e]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_leader<PAIR>(small_empty);
        if (threadIdx.x == kFirstEpiWarp * 32) GVQA_F16_TRACE(1024 + tile_it, 3);
        const uint32_t buf = tile_it & 1;
        if (live) {
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t r0[16];
            GVQA_TMEM_LD16(r0, tcol + buf * kBN + (uint32_t)(pc * 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[pc * 16 + e] = fmaf(acc[pc * 16 + e], kLoUnscale, __uint_as_float(r0[e]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_leader<PAIR>(&big_empty[buf]);
      }
      if (live) {
        // optional Linear-layer epilogue (bias, ReLU): every thread of the warp reads the same bias words (broadcast)
        const float* bias = t.p == 0 ? g.bias[0] : (t.p == 1 ? g.bias[1] : g.bias[2]);
        const int relu = t.p == 0 ? g.relu[0] : (t.p == 1 ? g.relu[1] : g.relu[2]);
        if (bias != nullptr) {
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (col0 + e < N) acc[e] += __ldg(bias + col0 + e);
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 64; ++e) acc[e] = fmaxf(acc[e], 0.f);
        }
        const uint32_t stage = smem_u32(epi_stage + (size_t)(warp - kFirstEpiWarp) * kEpiStageBytes);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          if (col0 + pass * 32 < N) {
            if (pass == 1) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(stage + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) * 16), acc[pass * 32 + 4 * j],
                     acc[pass * 32 + 4 * j + 1], acc[pass * 32 + 4 * j + 2], acc[pass * 32 + 4 * j + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(mc, stage, col0 + pass * 32, m0 + quarter * 32, z);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      if (threadIdx.x == kFirstEpiWarp * 32) GVQA_F16_TRACE(1024 + tile_it, 4);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR == 2) {
    cluster_sync_all();       // neither CTA leaves (or frees tensor memory) while its peer may still signal or read it
    if (warp == 1)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  } else {
    __syncthreads();
    if (warp == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// hi = fp16(x), lo' = fp16((x - hi) * 2^11): the weight half of the split, done once at prepack time.
// Rows are re-pitched from ld_in floats to ld_out halves (ld_out a multiple of 8, zero padded).
__global__ void split_f16_kernel(const float* __restrict__ w, int64_t ld_in, __half* __restrict__ hi,
                                 __half* __restrict__ lo, int64_t ld_out, int64_t rows, int64_t cols) {
  const int64_t total = rows * ld_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_out, c = i - r * ld_out;
    const float x = c < cols ? w[r * ld_in + c] : 0.f;
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn((x - __half2float(h)) * kLoScale);
  }
}

}  // namespace f16gemm
}  // namespace gvqa

using namespace gvqa;

extern "C" GVQA_API int gvqa_split_f16(const float* w, int64_t ld_in, void* hi, void* lo, int64_t ld_out, int64_t rows,
                                       int64_t cols, void* stream_) {
  if (rows < 0 || cols <= 0 || ld_in < cols || ld_out < cols || (ld_out & 7)) return GVQA_ERR_BAD_SHAPE;
  if (rows == 0) return GVQA_OK;
  if (!w || !hi || !lo) return GVQA_ERR_NULL_POINTER;
  const int64_t blocks = (rows * ld_out + 255) / 256;
  f16gemm::split_f16_kernel<<<(unsigned)(blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs), 256, 0,
                              static_cast<cudaStream_t>(stream_)>>>(w, ld_in, static_cast<__half*>(hi),
                                                                    static_cast<__half*>(lo), ld_out, rows, cols);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_proj_gemm_3xf16_grouped(const gvqa_gemm_problem* problems, int32_t count, int32_t* overflow,
                                                     void* stream_) {
  using namespace f16gemm;
  if (!problems || count < 1 || count > kMaxProblems) return GVQA_ERR_BAD_SHAPE;
  // two-CTA pairs (default) or one CTA per tile: GVQA_GEMM_PAIR=1 / debug flag 32 select the single-CTA kernel
  static const int env_pair = [] {
    const char* e = getenv("GVQA_GEMM_PAIR");
    return e ? atoi(e) : 2;
  }();
  const int flags = gemm_debug_flags();
  const int pair = (env_pair == 1 || (flags & 32)) ? 1 : 2;
  const int conv_halves = 1;       // (eight converter warps, CH = 2, measured equal: not instantiated)
  GroupedParams g;
  memset(&g, 0, sizeof(g));
  int64_t tiles = 0;
  int live = 0;
  for (int i = 0; i < count; ++i) {
    const gvqa_gemm_problem& q = problems[i];
    if (q.m < 0 || q.n <= 0 || q.k <= 0 || q.batch <= 0 || q.lda < q.k || q.ldb < q.k || q.ldc < q.n || q.m >= (1ll << 31))
      return GVQA_ERR_BAD_SHAPE;
    if (q.m == 0) continue;
    if (!q.a || !q.b_hi || !q.b_lo || !q.c) return GVQA_ERR_NULL_POINTER;
    if ((q.k & 3) || (q.lda & 3) || (q.ldb & 7) || (q.ldc & 3) || (q.stride_a & 3) || (q.stride_b & 7) || (q.stride_c & 3))
      return GVQA_ERR_UNSUPPORTED;
    if (!aligned16(q.a) || !aligned16(q.b_hi) || !aligned16(q.b_lo) || !aligned16(q.c)) return GVQA_ERR_MISALIGNED;
    const int64_t sa = q.batch > 1 ? q.stride_a : q.m * q.lda, sb = q.batch > 1 ? q.stride_b : (int64_t)q.n * q.ldb,
                  sc = q.batch > 1 ? q.stride_c : q.m * q.ldc;
    if (!make_map_3d(&g.map_a[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, q.a, q.batch, q.m, q.k, q.lda, sa, kBM, 32) ||
        !make_map_3d(&g.map_bhi[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q.b_hi, q.batch, q.n, q.k, q.ldb, sb, kBN / pair, 64) ||
        !make_map_3d(&g.map_blo[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q.b_lo, q.batch, q.n, q.k, q.ldb, sb, kBN / pair, 64) ||
        !make_map_3d(&g.map_c[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, q.c, q.batch, q.m, q.n, q.ldc, sc, 32, 32))
      return GVQA_ERR_CUDA;
    g.M[live] = (int)q.m; g.N[live] = q.n; g.K[live] = q.k;
    g.bias[live] = q.bias; g.relu[live] = q.relu;
    g.n_tiles[live] = (q.n + kBN - 1) / kBN;
    g.tiles_per_batch[live] = (int)((q.m + kBM * pair - 1) / (kBM * pair)) * g.n_tiles[live];
    tiles += (int64_t)g.tiles_per_batch[live] * q.batch;
    if (tiles >= (1ll << 30)) return GVQA_ERR_BAD_SHAPE;
    g.tile_end[live] = (int)tiles;
    ++live;
  }
  if (live == 0) return GVQA_OK;
  for (int i = live; i < kMaxProblems; ++i) {     // unused slots: empty ranges, valid divisors
    g.tile_end[i] = (int)tiles;
    g.tiles_per_batch[i] = g.n_tiles[i] = 1;
    g.K[i] = kBK;
  }
  g.count = live;
  g.dbg = flags & 15;
  g.trace = gemm_debug_trace();
  // DB (double-buffered hi*hi accumulators, one chain per tile) is opt-in (GVQA_GEMM_DB=1 or debug flag 64), K <= 512
  // only: measured 47.4 vs 47.9 us at cfg2 for twice the truncation bias of the default (gemm_f16_merge_probe.txt)
  static const int env_db = [] {
    const char* e = getenv("GVQA_GEMM_DB");
    return e ? atoi(e) : 0;
  }();
  int max_k = 0;
  for (int i = 0; i < live; ++i) max_k = g.K[i] > max_k ? g.K[i] : max_k;
  const int db = ((env_db != 0 || (flags & 64)) && max_k <= 512) ? 1 : 0;
  auto kernel = pair == 2 ? (db ? proj_gemm_3xf16_kernel<1, 2, 1> : proj_gemm_3xf16_kernel<1, 2, 0>)
                          : (db ? proj_gemm_3xf16_kernel<1, 1, 1> : proj_gemm_3xf16_kernel<1, 1, 0>);
  static bool attr_done[2][2] = {{false, false}, {false, false}};
  if (!attr_done[pair - 1][db]) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) != cudaSuccess) {
      (void)cudaGetLastError();
      return GVQA_ERR_CUDA;
    }
    attr_done[pair - 1][db] = true;
  }
  // one persistent CTA per SM; with pairs, one cluster of two per TPC and tiles dealt to clusters
  const int64_t units = pair == 2 ? kNumSMs / 2 : kNumSMs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((tiles < units ? tiles : units) * pair));
  cfg.blockDim = dim3(64 + 128 * conv_halves + kEpiThreads);
  cfg.dynamicSmemBytes = kGemmSmem;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_mask() & 1) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (cudaLaunchKernelEx(&cfg, kernel, g, overflow) != cudaSuccess) {
    (void)cudaGetLastError();
    return GVQA_ERR_CUDA;
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_proj_gemm_3xf16_batched(const float* a, int64_t lda, int64_t stride_a, const void* b_hi,
                                                     const void* b_lo, int64_t ldb, int64_t stride_b, float* c,
                                                     int64_t ldc, int64_t stride_c, int64_t m, int32_t n, int32_t k,
                                                     int32_t batch, int32_t* overflow, void* stream_) {
  gvqa_gemm_problem q;
  q.a = a; q.lda = lda; q.stride_a = stride_a; q.b_hi = b_hi; q.b_lo = b_lo; q.ldb = ldb; q.stride_b = stride_b;
  q.c = c; q.ldc = ldc; q.stride_c = stride_c; q.m = m; q.n = n; q.k = k; q.batch = batch;
  q.bias = nullptr; q.relu = 0;
  return gvqa_proj_gemm_3xf16_grouped(&q, 1, overflow, stream_);
}

extern "C" GVQA_API int gvqa_linear_3xf16(const float* a, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                                          const float* bias, int32_t relu, float* c, int64_t ldc, int64_t m, int32_t n,
                                          int32_t k, int32_t* overflow, void* stream_) {
  gvqa_gemm_problem q;
  q.a = a; q.lda = lda; q.stride_a = 0; q.b_hi = b_hi; q.b_lo = b_lo; q.ldb = ldb; q.stride_b = 0;
  q.c = c; q.ldc = ldc; q.stride_c = 0; q.m = m; q.n = n; q.k = k; q.batch = 1;
  q.bias = bias; q.relu = relu;
  return gvqa_proj_gemm_3xf16_grouped(&q, 1, overflow, stream_);
}

extern "C" GVQA_API int gvqa_proj_gemm_3xf16(const float* a, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                                             float* c, int64_t ldc, int64_t m, int32_t n, int32_t k, int32_t* overflow,
                                             void* stream_) {
  return gvqa_proj_gemm_3xf16_batched(a, lda, 0, b_hi, b_lo, ldb, 0, c, ldc, 0, m, n, k, 1, overflow, stream_);
}
