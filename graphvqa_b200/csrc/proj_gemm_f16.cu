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
constexpr int kStages = 3;                          // shared-memory ring
constexpr int kAStages = 2;                         // tensor-memory ring of split A tiles
constexpr int kGemmThreads = 448;                   // warp 0 TMA, warp 1 MMA, warps 2-5 converters, warps 6-13 epilogue
constexpr int kConvThreads = 128;
constexpr int kEpiThreads = 256;
constexpr uint32_t kAHalfBytes = kBM * 32 * 4;      // one [128 x 32] fp32 box = 16 KB
constexpr uint32_t kABytes = 2 * kAHalfBytes;       // 32 KB
constexpr uint32_t kBBytes = kBN * kBK * 2;         // 16 KB
constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;       // 64 KB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemSmall = 2 * kBN;            // accumulators: [0,128) hi*hi first K-half, [128,256) second, [256,384) lo terms
constexpr uint32_t kTmemA = 3 * kBN;                // A ring: 2 stages x (32 columns hi | 32 columns lo'), 2 halves per column
constexpr uint32_t kACols = kBK / 2;                // 32 TMEM columns per 64 fp16
constexpr uint32_t kEpiStageBytes = 32 * 32 * 4;
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + (kEpiThreads / 32) * kEpiStageBytes + 1024 + 256;
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

// Up to kMaxProblems independent products share one persistent launch (tiles of problem 0 first, then 1, ...):
// gat_seq issues hop 0's projection together with the two small pre-pass products, whose ~170 short tiles then
// fill the last, partially empty wave instead of costing two more launches.
constexpr int kMaxProblems = 3;

struct alignas(64) GroupedParams {
  CUtensorMap map_a[kMaxProblems], map_bhi[kMaxProblems], map_blo[kMaxProblems], map_c[kMaxProblems];
  int M[kMaxProblems], N[kMaxProblems], K[kMaxProblems];
  int tiles_per_batch[kMaxProblems], n_tiles[kMaxProblems];
  int tile_end[kMaxProblems];      // exclusive prefix of tiles
  int count;
};

struct TileInfo {
  int p, z, m0, n0, kblocks;
};

__device__ __forceinline__ TileInfo decode_tile(const GroupedParams& g, int tile) {
  TileInfo t;
  t.p = tile < g.tile_end[0] ? 0 : (tile < g.tile_end[1] ? 1 : 2);
  const int local = tile - (t.p == 0 ? 0 : g.tile_end[t.p - 1]);
  t.z = local / g.tiles_per_batch[t.p];
  const int t2 = local - t.z * g.tiles_per_batch[t.p];
  t.m0 = (t2 / g.n_tiles[t.p]) * kBM;
  t.n0 = (t2 % g.n_tiles[t.p]) * kBN;
  t.kblocks = (g.K[t.p] + kBK - 1) / kBK;          // the ragged last k-block is zero-filled by TMA
  return t;
}

#define GVQA_MAP(field, p) ((p) == 0 ? &g.field[0] : ((p) == 1 ? &g.field[1] : &g.field[2]))

__global__ void __launch_bounds__(kGemmThreads, 1)
proj_gemm_3xf16_kernel(const __grid_constant__ GroupedParams g, int32_t* __restrict__ overflow) {
  // all tensor maps are 3-D [batch, rows, K]; a tile index decomposes into (problem, batch z, row tile, column tile)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi_stage = smem + (size_t)kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (kEpiThreads / 32) * kEpiStageBytes);
  uint64_t* tma_full = bars;                        // [kStages]
  uint64_t* smem_empty = bars + kStages;            // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;           // [kAStages]
  uint64_t* a_empty = a_ready + kAStages;           // [kAStages]
  uint64_t* acc_full = a_empty + kAStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = g.tile_end[g.count - 1];

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
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();                 // set-up above overlaps the previous kernel's tail (see common.cuh)
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer: one elected lane runs the whole loop =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileInfo t = decode_tile(g, tile);
        const CUtensorMap *ma = GVQA_MAP(map_a, t.p), *mh = GVQA_MAP(map_bhi, t.p), *ml = GVQA_MAP(map_blo, t.p);
        for (int kb = 0; kb < t.kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&smem_empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* st = smem + (size_t)s * kStageBytes;
          mbar_expect_tx(&tma_full[s], kStageBytes);
          tma_load_3d(st, ma, &tma_full[s], kb * kBK, t.m0, t.z);
          tma_load_3d(st + kAHalfBytes, ma, &tma_full[s], kb * kBK + 32, t.m0, t.z);
          tma_load_3d(st + kABytes, mh, &tma_full[s], kb * kBK, t.n0, t.z);
          tma_load_3d(st + kABytes + kBBytes, ml, &tma_full[s], kb * kBK, t.n0, t.z);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected lane runs the whole loop =====================
    if (elect_one()) {
      uint32_t it = 0, tile_it = 0;
      const uint64_t desc0 = umma_desc(smem_u32(smem));   // descriptor of (base + c) == desc0 + (c >> 4)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, M = 128, N = tile width (x16)
        const TileInfo t = decode_tile(g, tile);
        const int kblocks = t.kblocks;
        const int half_kb = (kblocks + 1) / 2;            // first k-block of the second K-half
        const int ncols = min(kBN, (g.N[t.p] - t.n0 + 15) & ~15);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        mbar_wait(acc_empty, (tile_it & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % kStages, ts = it & 1;
          mbar_wait(&a_ready[ts], (it >> 1) & 1);         // implies tma_full[s]: the converters waited on it
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t b_hi = desc0 + (uint64_t)((s * kStageBytes + kABytes) >> 4), b_lo = b_hi + (kBBytes >> 4);
          const uint32_t a_hi = tmem_base + kTmemA + ts * 2 * kACols, a_lo = a_hi + kACols;
          const bool second = kb >= half_kb;
          const bool chunk_first = kb == 0 || kb == half_kb;
          const uint32_t d_big = tmem_base + (second ? (uint32_t)kBN : 0u), d_small = tmem_base + kTmemSmall;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {            // K = 16 per MMA: 8 TMEM columns of A, 32 bytes of B
            umma_f16_ts(d_small, a_lo + 8 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
            umma_f16_ts(d_small, a_hi + 8 * k, b_lo + 2 * k, idesc, 1);
            umma_f16_ts(d_big, a_hi + 8 * k, b_hi + 2 * k, idesc, !(chunk_first && k == 0));
          }
          umma_commit(&smem_empty[s]);
          umma_commit(&a_empty[ts]);
          if (kb == kblocks - 1) umma_commit(acc_full);
        }
      }
    }
  } else if (warp < 6) {
    // ===================== converters (warps 2..5): thread = one row of the A tile ===============
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0;
    float amax = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int kblocks = decode_tile(g, tile).kblocks;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % kStages, ts = it & 1;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        mbar_wait(&a_empty[ts], ((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t row_addr = smem_u32(smem + (size_t)s * kStageBytes) + (uint32_t)r * 128u;
        const uint32_t ta = tmem_base + lane_base + kTmemA + (uint32_t)ts * 2 * kACols;
#pragma unroll
        for (int half = 0; half < 2; ++half) {             // the two [128 x 32] fp32 boxes of the stage
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int cidx = 0; cidx < 8; ++cidx) {           // 128B swizzle: 16-byte chunk c lives at c ^ (row & 7)
            const float4 v = lds128(row_addr + half * kAHalfBytes + (uint32_t)((cidx ^ (r & 7)) * 16));
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            const uint32_t h01 = pack_half2(v.x, v.y), h23 = pack_half2(v.z, v.w);
            const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
            hi[2 * cidx] = h01;
            hi[2 * cidx + 1] = h23;
            lo[2 * cidx] = pack_half2((v.x - f01.x) * kLoScale, (v.y - f01.y) * kLoScale);
            lo[2 * cidx + 1] = pack_half2((v.z - f23.x) * kLoScale, (v.w - f23.y) * kLoScale);
          }
          GVQA_TMEM_ST16(ta + half * 16, hi, 0);
          GVQA_TMEM_ST16(ta + kACols + half * 16, lo, 0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&a_ready[ts]);
      }
    }
    if (overflow != nullptr && !(amax <= 65000.0f)) atomicOr(overflow, 1);   // also catches NaN / inf inputs
  } else {
    // ===================== epilogue (warps 6..13): TMEM -> registers, release TMEM, then TMA store =============
    const int quarter = warp & 3;
    const int chalf = warp >= 10 ? 1 : 0;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const TileInfo t = decode_tile(g, tile);
      const int z = t.z, m0 = t.m0, n0 = t.n0, N = g.N[t.p], kblocks = t.kblocks;
      const CUtensorMap* mc = GVQA_MAP(map_c, t.p);
      mbar_wait(acc_full, tile_it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int col0 = n0 + chalf * 64;
      float acc[64];
      const bool live = col0 < N;
      const bool two_chunks = kblocks >= 2;                // with a single k-block the second K-half is never written
      if (live) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64 + pc * 16);
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
      mbar_arrive(acc_empty);                              // TMEM is free: the next tile's MMAs may start
      if (live) {
        const uint32_t stage = smem_u32(epi_stage + (size_t)(warp - 6) * kEpiStageBytes);
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
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
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
        !make_map_3d(&g.map_bhi[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q.b_hi, q.batch, q.n, q.k, q.ldb, sb, kBN, 64) ||
        !make_map_3d(&g.map_blo[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q.b_lo, q.batch, q.n, q.k, q.ldb, sb, kBN, 64) ||
        !make_map_3d(&g.map_c[live], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, q.c, q.batch, q.m, q.n, q.ldc, sc, 32, 32))
      return GVQA_ERR_CUDA;
    g.M[live] = (int)q.m; g.N[live] = q.n; g.K[live] = q.k;
    g.n_tiles[live] = (q.n + kBN - 1) / kBN;
    g.tiles_per_batch[live] = (int)((q.m + kBM - 1) / kBM) * g.n_tiles[live];
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
  static const bool attr_ok =
      cudaFuncSetAttribute(proj_gemm_3xf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) ==
      cudaSuccess;
  if (!attr_ok) return GVQA_ERR_CUDA;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  if (launch_pdl(1, proj_gemm_3xf16_kernel, dim3(grid), dim3(kGemmThreads), kGemmSmem, static_cast<cudaStream_t>(stream_),
                 g, overflow) != cudaSuccess) {
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
  return gvqa_proj_gemm_3xf16_grouped(&q, 1, overflow, stream_);
}

extern "C" GVQA_API int gvqa_proj_gemm_3xf16(const float* a, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                                             float* c, int64_t ldc, int64_t m, int32_t n, int32_t k, int32_t* overflow,
                                             void* stream_) {
  return gvqa_proj_gemm_3xf16_batched(a, lda, 0, b_hi, b_lo, ldb, 0, c, ldc, 0, m, n, k, 1, overflow, stream_);
}
