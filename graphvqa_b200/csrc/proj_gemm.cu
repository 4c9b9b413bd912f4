// fp32-accurate node projection on the 5th-generation tensor cores (see include/gvqa_b200.h:
// gvqa_proj_gemm_3xtf32, gvqa_split_tf32).
//
//   C[M,N] = A[M,K] @ B[N,K]^T        (reference: x_l = lin_l(x_cat), gat_skip.py:133 -- a cuBLAS SGEMM there)
//
// The 1e-4 parity bar rules out plain TF32 (10-bit mantissa), so every fp32 operand is split as
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and three tcgen05.mma.kind::tf32 products are
// accumulated in fp32 in tensor memory:  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term
// is ~2^-22 relative) -- "3xTF32".  B (the weights) is split once at prepack time; A (the node
// state, new every hop) is split on the fly by the converter warps.
//
// Two measured facts shape the kernel (profiles/microbench/trace_gemm.py, ncu):
//  * the tensor core adds into its accumulator with truncation, which biases long chains (8x the
//    error of an fp32 SGEMM at K=512 with one accumulator).  The tile therefore keeps THREE
//    accumulators in TMEM: hi*hi of the two K-halves and one for all the small lo-terms (their
//    truncation is 2^-11 smaller); the epilogue adds them in fp32 RN.
//  * with both operands in shared memory a tf32 M128 x N128 x K8 MMA needs 8 KB of smem reads per
//    64 cycles = the whole 128 B/clk of the SM, so converter and TMA traffic stalled the tensor
//    pipe (40 % active).  A is therefore fed from TENSOR MEMORY (tcgen05.mma "TS" form): the
//    converters write A_hi / A_lo with tcgen05.st into a 2-stage TMEM ring and only B is read
//    from shared memory.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: per k-block one cp.async.bulk.tensor (UTMALDG) each for the raw A
//               tile and the B_hi / B_lo tiles, [128 x 32] floats, 128B-swizzled, 4-stage ring
//   warps 2-5   converters: thread = one row of the A tile: 8 swizzled 128-bit smem loads, split,
//               2 x tcgen05.st (hi, lo) into the TMEM A ring, arrive
//   warp 1      one elected lane issues 12 MMAs (M128 x N128 x K8, A from TMEM) per k-block and
//               commits the smem stage and the TMEM A stage back (tcgen05.commit -> mbarrier)
//   warps 6-9   epilogue: tcgen05.ld of their 32 TMEM lanes, fp32 sum of the three accumulators,
//               128-bit global stores
// TMEM map (512 columns): [0,128) hi*hi first K-half, [128,256) second K-half, [256,384) lo terms,
// [384,512) A ring: 2 stages x (32 columns hi | 32 columns lo).
#include "tcgen05_utils.cuh"

namespace gvqa {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // tile: rows of A, rows of B, k (floats; 128 bytes)
constexpr int kStages = 4;                          // shared-memory ring
constexpr int kAStages = 2;                         // tensor-memory ring of split A tiles
constexpr int kBigChunks = 2;                       // hi*hi accumulators over K-halves (+1 for the lo terms)
constexpr int kGemmThreads = 448;                   // warp 0 TMA, warp 1 MMA, warps 2-5 converters, warps 6-13 epilogue
constexpr int kConvThreads = 128;
constexpr int kEpiThreads = 256;                    // two warps per TMEM lane quarter, 64 accumulator columns each
constexpr uint32_t kABytes = kBM * kBK * 4;         // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 4;         // 16 KB
constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;       // A raw | B_hi | B_lo = 48 KB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemSmall = kBigChunks * kBN;             // 256
constexpr uint32_t kTmemA = (kBigChunks + 1) * kBN;           // 384
constexpr uint32_t kEpiStageBytes = 32 * 32 * 4;    // per epilogue warp: a [32 rows x 32 columns] block staged for the TMA store
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + (kEpiThreads / 32) * kEpiStageBytes +
                             1024 /*align slack*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(kGemmThreads, 1)
proj_gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                        const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_c,
                        const float* __restrict__ a_raw, int64_t lda, int M,
                        int N, int K, long long* __restrict__ trace, const int dbg) {
  // dbg (debug only, 0 in production): bit0 producer skips the TMA loads, bit1 converters skip their work,
  // bit2 epilogue skips TMEM loads + stores -- used by profiles/microbench/gemm_dbg.py to attribute time
  // trace (debug, may be null): CTA 0 records clock64() per k-block and role: [it][0..4] = producer issue,
  // converter start, converter done, mma start, mma committed; [1024+tile][0..1] = epilogue start / end
#define GVQA_TRACE(slot, col) do { if (trace && blockIdx.x == 0 && (slot) < 1100) trace[(slot) * 8 + (col)] = clock64(); } while (0)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi_stage = smem + (size_t)kStages * kStageBytes;   // [8 warps][4 KB], 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (kEpiThreads / 32) * kEpiStageBytes);
  uint64_t* tma_full = bars;                        // [kStages]   TMA bytes landed
  uint64_t* smem_empty = bars + kStages;            // [kStages]   MMAs that read the stage are done
  uint64_t* a_ready = bars + 2 * kStages;           // [kAStages]  converters filled the TMEM A stage
  uint64_t* a_empty = a_ready + kAStages;           // [kAStages]  MMAs that read the TMEM A stage are done
  uint64_t* acc_full = a_empty + kAStages;          // [1]
  uint64_t* acc_empty = acc_full + 1;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + kBN - 1) / kBN;
  const int num_tiles = m_tiles * n_tiles;
  // k-blocks rounded up to a multiple of the ring depth (TMA zero-fills the out-of-range ones): every tile then
  // starts at ring slot 0, so the MMA issuer's loop is unrolled over the slots with compile-time descriptors
  const int kblocks = ((K + kBK - 1) / kBK + kStages - 1) / kStages * kStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&smem_empty[s], 1);
    }
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_ready[s], kConvThreads);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiThreads);
    mbar_fence_init();
  }
  if (warp == 1) {  // one warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // PDL: barrier init and the TMEM allocation above overlap the tail of the previous kernel (the fused hop whose
  // output is this GEMM's A operand and which still reads the x_l buffer this GEMM overwrites)
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // A streamed exactly once (a single column tile, e.g. the edge-logit pre-pass over edge_attr): the tile loads
    // fetch 128-byte pieces of 128 different rows per k-block, which DRAM serves poorly.  Pull this CTA's first
    // row block into L2 with one contiguous bulk prefetch per row; the tile loads then hit L2.
    if (n_tiles == 1 && blockIdx.x < num_tiles && (K & 3) == 0) {
      const int m0 = blockIdx.x * kBM;
      for (int r = lane; r < kBM && m0 + r < M; r += 32)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a_raw + (int64_t)(m0 + r) * lda), "r"(K * 4) : "memory");
    }
    // ===================== TMA producer: one elected lane runs the whole loop =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&smem_empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* st = smem + (size_t)s * kStageBytes;
          GVQA_TRACE(it, 0);
          if (dbg & 1) {
            mbar_arrive(&tma_full[s]);
          } else {
            mbar_expect_tx(&tma_full[s], kABytes + 2 * kBBytes);
            tma_load_2d(st, &map_a, &tma_full[s], kb * kBK, m0);
            tma_load_2d(st + kABytes, &map_bhi, &tma_full[s], kb * kBK, n0);
            tma_load_2d(st + kABytes + kBBytes, &map_blo, &tma_full[s], kb * kBK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected lane runs the whole loop (leaving and re-entering the
    // elected region per k-block cost ~270 cycles of reconvergence after the tcgen05 instructions) ==========
    if (elect_one()) {
      uint32_t it = 0, tile_it = 0;
      static_assert(kAStages == 2 && kStages % kAStages == 0, "MMA loop unroll assumes 2 TMEM A stages");
      const uint64_t desc0 = umma_desc(smem_u32(smem));   // descriptor of (base + c) == desc0 + (c >> 4)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        // instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N = this tile's width rounded up to
        // 16 (a ragged last column tile, e.g. the 16 logit columns appended to W, costs N=16 MMAs, not 128)
        const int n0 = (tile % n_tiles) * kBN;
        const int ncols = min(kBN, (N - n0 + 15) & ~15);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncols >> 3) << 17) |
                               ((uint32_t)(kBM >> 4) << 24);
        const int half_kb = (kblocks + 1) / 2;             // first k-block of the second K-half (kBigChunks == 2)
        mbar_wait(acc_empty, (tile_it & 1) ^ 1);          // epilogue of the previous tile has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb0 = 0; kb0 < kblocks; kb0 += kStages) {
#pragma unroll
          for (int u = 0; u < kStages; ++u, ++it) {       // u == ring slot; slot and TMEM A stage are compile-time
            const int kb = kb0 + u;
            constexpr int kHalf = kAStages;               // kAStages == 2: TMEM A stage = u & 1, its parity = (u >> 1) & 1
            const int ts = u % kHalf;
            GVQA_TRACE(it, 5);
            mbar_wait(&a_ready[ts], (u / kHalf) & 1);     // implies tma_full[u]: the converters waited on it
            GVQA_TRACE(it, 6);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            GVQA_TRACE(it, 3);
            const uint64_t b_hi = desc0 + (uint64_t)((u * kStageBytes + kABytes) >> 4), b_lo = b_hi + (kBBytes >> 4);
            const uint32_t a_hi = tmem_base + kTmemA + (uint32_t)(ts * 2 * kBK), a_lo = a_hi + kBK;
            // hi*hi goes to the accumulator of this K-half, the lo terms to the last accumulator
            const bool second = kb >= half_kb;
            const bool chunk_first = kb == 0 || kb == half_kb;
            const uint32_t d_big = tmem_base + (second ? (uint32_t)kBN : 0u), d_small = tmem_base + kTmemSmall;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {            // K = 8 per MMA: 8 TMEM columns of A, 32 bytes of B
              umma_tf32_ts(d_small, a_lo + 8 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
              umma_tf32_ts(d_small, a_hi + 8 * k, b_lo + 2 * k, idesc, 1);
              umma_tf32_ts(d_big, a_hi + 8 * k, b_hi + 2 * k, idesc, !(chunk_first && k == 0));
            }
            umma_commit(&smem_empty[u]);                   // smem stage reusable once these MMAs are done
            umma_commit(&a_empty[ts]);                     // and so is the TMEM A stage
            if (kb == kblocks - 1) umma_commit(acc_full);  // accumulators complete
            GVQA_TRACE(it, 4);
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== converters (warps 2..5): thread = one row of the A tile ===============
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                     // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % kStages, ts = it % kAStages;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        mbar_wait(&a_empty[ts], ((it / kAStages) & 1) ^ 1);
        if (threadIdx.x == 64) GVQA_TRACE(it, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg & 2) { mbar_arrive(&a_ready[ts]); continue; }
        const uint32_t row_addr = smem_u32(smem + (size_t)s * kStageBytes) + (uint32_t)r * 128u;
        uint32_t hi[kBK], lo[kBK];
#pragma unroll
        for (int cidx = 0; cidx < kBK / 4; ++cidx) {       // 128B swizzle: 16-byte chunk c lives at c ^ (row & 7)
          const float4 v = lds128(row_addr + (uint32_t)((cidx ^ (r & 7)) * 16));
          hi[4 * cidx + 0] = to_tf32(v.x); hi[4 * cidx + 1] = to_tf32(v.y);
          hi[4 * cidx + 2] = to_tf32(v.z); hi[4 * cidx + 3] = to_tf32(v.w);
          lo[4 * cidx + 0] = to_tf32(v.x - __uint_as_float(hi[4 * cidx + 0]));
          lo[4 * cidx + 1] = to_tf32(v.y - __uint_as_float(hi[4 * cidx + 1]));
          lo[4 * cidx + 2] = to_tf32(v.z - __uint_as_float(hi[4 * cidx + 2]));
          lo[4 * cidx + 3] = to_tf32(v.w - __uint_as_float(hi[4 * cidx + 3]));
        }
        const uint32_t ta = tmem_base + lane_base + kTmemA + (uint32_t)(ts * 2 * kBK);
        GVQA_TMEM_ST16(ta, hi, 0);
        GVQA_TMEM_ST16(ta + 16, hi, 16);
        GVQA_TMEM_ST16(ta + kBK, lo, 0);
        GVQA_TMEM_ST16(ta + kBK + 16, lo, 16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&a_ready[ts]);
        if (threadIdx.x == 64) GVQA_TRACE(it, 2);
      }
    }
  } else {
    // ===================== epilogue (warps 6..13): TMEM -> registers, release TMEM, then global ====================
    // Two warps per TMEM lane quarter, 64 accumulator columns each.  The three accumulators are summed into
    // registers in 16-column pieces; TMEM is handed back to the MMA issuer as soon as the last tcgen05.ld has
    // landed, so the global stores of tile t overlap the main loop of tile t+1 (no second accumulator set needed).
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
    const int chalf = warp >= 10 ? 1 : 0;                  // which 64 columns of the tile
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
      mbar_wait(acc_full, tile_it & 1);
      if (threadIdx.x == 192) GVQA_TRACE(1024 + tile_it, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int col0 = n0 + chalf * 64;
      float acc[64];
      const bool live = col0 < N && !(dbg & 4);            // a ragged last column tile may leave this half empty
      if (live) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64 + pc * 16);
          uint32_t rs[16], r0[16], r1[16];
          GVQA_TMEM_LD16(rs, taddr + kTmemSmall);
          GVQA_TMEM_LD16(r0, taddr);
          GVQA_TMEM_LD16(r1, taddr + kBN);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 16; ++e)
            acc[pc * 16 + e] = (__uint_as_float(rs[e]) + __uint_as_float(r0[e])) + __uint_as_float(r1[e]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc_empty);                              // TMEM is free: the next tile's MMAs may start
      if (threadIdx.x == 192) GVQA_TRACE(1024 + tile_it, 1);
      if (live) {
        // registers -> 128B-swizzled shared block -> one TMA store per [32 x 32] block.  Thread-per-row
        // global stores would touch 32 different lines per instruction and saturate the LSU for ~4000
        // cycles per tile, starving the converters' shared-memory loads (measured); the TMA store costs the
        // LSU 8 conflict-light STS per block and clips rows >= M / columns >= N by itself.
        const uint32_t stage = smem_u32(epi_stage + (size_t)(warp - 6) * kEpiStageBytes);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          if (col0 + pass * 32 < N) {
            if (pass == 1) {                               // the block is reused: its previous store must have read it
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(stage + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) * 16),
                     acc[pass * 32 + 4 * j], acc[pass * 32 + 4 * j + 1], acc[pass * 32 + 4 * j + 2],
                     acc[pass * 32 + 4 * j + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              if (dbg & 8)
                tma_store_2d_hint(&map_c, stage, col0 + pass * 32, m0 + quarter * 32, l2_policy_evict_last());
              else
                tma_store_2d(&map_c, stage, col0 + pass * 32, m0 + quarter * 32);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // before the next tile restages
        __syncwarp();
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // all stores complete
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// hi = tf32(x), lo = tf32(x - hi): the weight half of the 3xTF32 split, done once at prepack time
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                  int64_t count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float h = __uint_as_float(to_tf32(x));
    hi[i] = h;
    lo[i] = __uint_as_float(to_tf32(x - h));
  }
}

}  // namespace gvqa

using namespace gvqa;

static long long* g_gemm_trace = nullptr;   // debug only: set through gvqa_debug_set_gemm_trace
static int g_gemm_dbg = 0;                  // debug only: set through gvqa_debug_set_gemm_flags
extern "C" GVQA_API void gvqa_debug_set_gemm_trace(long long* device_buffer) { g_gemm_trace = device_buffer; }
extern "C" GVQA_API void gvqa_debug_set_gemm_flags(int flags) { g_gemm_dbg = flags; }
namespace gvqa {
int gemm_debug_flags() { return g_gemm_dbg; }
long long* gemm_debug_trace() { return g_gemm_trace; }
}

extern "C" GVQA_API int gvqa_split_tf32(const float* w, float* hi, float* lo, int64_t count, void* stream_) {
  if (count < 0) return GVQA_ERR_BAD_SHAPE;
  if (count == 0) return GVQA_OK;
  if (!w || !hi || !lo) return GVQA_ERR_NULL_POINTER;
  const int64_t blocks = (count + 255) / 256;
  split_tf32_kernel<<<(unsigned)(blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs), 256, 0,
                      static_cast<cudaStream_t>(stream_)>>>(w, hi, lo, count);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}

extern "C" GVQA_API int gvqa_proj_gemm_3xtf32(const float* a, int64_t lda, const float* b_hi, const float* b_lo,
                                              int64_t ldb, float* c, int64_t ldc, int64_t m, int32_t n, int32_t k,
                                              void* stream_) {
  if (m < 0 || n <= 0 || k <= 0 || lda < k || ldb < k || ldc < n || m >= (1ll << 31)) return GVQA_ERR_BAD_SHAPE;
  if (m == 0) return GVQA_OK;
  if (!a || !b_hi || !b_lo || !c) return GVQA_ERR_NULL_POINTER;
  if ((k & 3) || (lda & 3) || (ldb & 3) || (ldc & 3)) return GVQA_ERR_UNSUPPORTED;
  if (!aligned16(a) || !aligned16(b_hi) || !aligned16(b_lo) || !aligned16(c)) return GVQA_ERR_MISALIGNED;
  CUtensorMap map_a, map_bhi, map_blo, map_c;
  if (!make_map(&map_a, a, m, k, lda, kBM) || !make_map(&map_bhi, b_hi, n, k, ldb, kBN) ||
      !make_map(&map_blo, b_lo, n, k, ldb, kBN) || !make_map(&map_c, c, m, n, ldc, 32, 32))
    return GVQA_ERR_CUDA;
  static const bool attr_ok =
      cudaFuncSetAttribute(proj_gemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) ==
      cudaSuccess;
  if (!attr_ok) return GVQA_ERR_CUDA;
  const int tiles = (int)((m + kBM - 1) / kBM) * ((n + kBN - 1) / kBN);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  if (launch_pdl(1, proj_gemm_3xtf32_kernel, dim3(grid), dim3(kGemmThreads), kGemmSmem, static_cast<cudaStream_t>(stream_),
                 map_a, map_bhi, map_blo, map_c, a, lda, (int)m, n, k, g_gemm_trace, g_gemm_dbg) != cudaSuccess) {
    (void)cudaGetLastError();
    return GVQA_ERR_CUDA;
  }
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
