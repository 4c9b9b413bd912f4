// fp32-accurate node projection on the 5th-generation tensor cores (see include/gvqa_b200.h:
// gvqa_proj_gemm_3xtf32, gvqa_split_tf32).
//
//   C[M,N] = A[M,K] @ B[N,K]^T        (reference: x_l = lin_l(x_cat), gat_skip.py:133 -- a cuBLAS SGEMM there)
//
// The 1e-4 parity bar rules out plain TF32 (10-bit mantissa), so every fp32 operand is split as
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and three tcgen05.mma.kind::tf32 products are
// accumulated in fp32 in tensor memory:  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term
// is ~2^-22 relative) -- "3xTF32".  B (the weights) is split once at prepack time; A (the node
// state, new every hop) is split on the fly in shared memory by the converter warps.
//
// The tensor core adds into its accumulator with truncation, which biases long chains (measured:
// 8x the error of an fp32 SGEMM at K=512 with one accumulator).  So the tile keeps FOUR
// accumulators in TMEM: the hi*hi products of three K-thirds (chains 3x shorter) and one for all
// the small lo-terms (their truncation is 2^-11 smaller); the epilogue adds the four in fp32 RN.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: per k-block one cp.async.bulk.tensor (UTMALDG) each for the raw A
//               tile [128 x 32], B_hi and B_lo tiles [128 x 32], 128B-swizzled, 3-stage ring
//   warps 2-5   converters: split the A tile in place (hi) + side buffer (lo), fence.proxy.async,
//               arrive; after the last k-block they become the epilogue: tcgen05.ld of their 32
//               TMEM lanes -> registers -> 128-bit global stores
//   warp 1      one elected lane issues 12 MMAs (M128 x N128 x K8) per k-block and commits the
//               stage back to the producer (tcgen05.commit -> mbarrier)
// Accumulators: 128 lanes x 4 x 128 columns of TMEM (fp32).
#include <cuda.h>

#include "common.cuh"

namespace gvqa {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // tile: rows of A, rows of B, k (floats; 128 bytes)
constexpr int kStages = 3;
constexpr int kBigChunks = 3;                       // hi*hi accumulators over K-thirds (+1 for the lo terms)
constexpr int kGemmThreads = 192;
constexpr int kConvThreads = 128;
constexpr uint32_t kABytes = kBM * kBK * 4;         // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 4;         // 16 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // A(hi) | A_lo | B_hi | B_lo = 64 KB
constexpr uint32_t kTmemCols = (kBigChunks + 1) * kBN;        // 512: all of tensor memory
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3fff) | (1u << 16);          // start address, LBO = 1 (unused)
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);            // SBO = 1024 B, version 1, SWIZZLE_128B
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
proj_gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                        const __grid_constant__ CUtensorMap map_blo, float* __restrict__ c, int64_t ldc, int M,
                        int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* tma_full = bars;                    // [kStages]
  uint64_t* conv_done = bars + kStages;         // [kStages]
  uint64_t* empty = bars + 2 * kStages;         // [kStages]
  uint64_t* acc_full = bars + 3 * kStages;      // [1]
  uint64_t* acc_empty = acc_full + 1;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + kBN - 1) / kBN;
  const int num_tiles = m_tiles * n_tiles;
  const int kblocks = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&conv_done[s], kConvThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kConvThreads);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* st = smem + (size_t)s * kStageBytes;
          mbar_expect_tx(&tma_full[s], kABytes + 2 * kBBytes);
          tma_load_2d(st, &map_a, &tma_full[s], kb * kBK, m0);
          tma_load_2d(st + 2 * kABytes, &map_bhi, &tma_full[s], kb * kBK, n0);
          tma_load_2d(st + 2 * kABytes + kBBytes, &map_blo, &tma_full[s], kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N=128, M=128
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) |
                                 ((uint32_t)(kBM >> 4) << 24);
      uint32_t it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait(acc_empty, (tile_it & 1) ^ 1);          // epilogue of the previous tile has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&conv_done[s], (it / kStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes);
          const uint32_t a_lo = a_hi + kABytes, b_hi = a_hi + 2 * kABytes, b_lo = b_hi + kBBytes;
          // hi*hi goes to the accumulator of this K-third, the lo terms to the last accumulator
          const int chunk = (kb * kBigChunks) / kblocks;
          const bool chunk_first = kb == 0 || ((kb - 1) * kBigChunks) / kblocks != chunk;
          const uint32_t d_big = tmem_base + (uint32_t)(chunk * kBN), d_small = tmem_base + (uint32_t)(kBigChunks * kBN);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {            // K = 8 per MMA = 32 bytes inside the swizzle atom
            const uint32_t off = k * 32;
            umma_tf32(d_small, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, (kb | k) != 0);
            umma_tf32(d_small, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1);
            umma_tf32(d_big, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, !(chunk_first && k == 0));
          }
          umma_commit(&empty[s]);                          // stage reusable once these MMAs have read it
        }
        umma_commit(acc_full);                             // accumulator complete
      }
    }
  } else {
    // ===================== converters, then epilogue (warps 2..5) =====================
    const int t = threadIdx.x - 64;                        // 0..127
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
    uint32_t it = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&tma_full[s], (it / kStages) & 1);
        float4* a = reinterpret_cast<float4*>(smem + (size_t)s * kStageBytes);
        float4* alo = reinterpret_cast<float4*>(smem + (size_t)s * kStageBytes + kABytes);
#pragma unroll
        for (int j = 0; j < (int)(kABytes / 16) / kConvThreads; ++j) {
          const int idx = t + j * kConvThreads;
          const float4 v = a[idx];
          uint4 h, l;
          h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
          l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
          l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
          reinterpret_cast<uint4*>(a)[idx] = h;
          reinterpret_cast<uint4*>(alo)[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA
        mbar_arrive(&conv_done[s]);
      }
      // ---- epilogue: TMEM -> registers -> global ----
      mbar_wait(acc_full, tile_it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + quarter * 32 + lane;
      float* crow = c + (int64_t)row * ldc + n0;
      // accumulators that received nothing (fewer k-blocks than K-thirds) are skipped
      int used[kBigChunks];
#pragma unroll
      for (int q = 0; q < kBigChunks; ++q) {
        used[q] = 0;
        for (int kb = 0; kb < kblocks; ++kb) used[q] |= ((kb * kBigChunks) / kblocks) == q;
      }
#pragma unroll 1
      for (int cc = 0; cc < kBN / 32; ++cc) {
        float acc[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * 32);
#pragma unroll
        for (int q = kBigChunks; q >= 0; --q) {            // small terms first, then the K-thirds
          if (q < kBigChunks && !used[q]) continue;
          uint32_t r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr + (uint32_t)(q * kBN)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = (q == kBigChunks) ? __uint_as_float(r[e]) : acc[e] + __uint_as_float(r[e]);
        }
        if (row < M) {
          const int col0 = n0 + cc * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col0 + 4 * j + 3 < N) {
              *reinterpret_cast<float4*>(crow + cc * 32 + 4 * j) =
                  make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            } else {
              for (int e = 0; e < 4; ++e)
                if (col0 + 4 * j + e < N) crow[cc * 32 + 4 * j + e] = acc[4 * j + e];
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc_empty);
    }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [rows, K] fp32 row-major (row stride ld floats) -> 2-D map with a [box_rows x 32] 128B-swizzled box
static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gvqa

using namespace gvqa;

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
  CUtensorMap map_a, map_bhi, map_blo;
  if (!make_map(&map_a, a, m, k, lda, kBM) || !make_map(&map_bhi, b_hi, n, k, ldb, kBN) ||
      !make_map(&map_blo, b_lo, n, k, ldb, kBN))
    return GVQA_ERR_CUDA;
  static const bool attr_ok =
      cudaFuncSetAttribute(proj_gemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) ==
      cudaSuccess;
  if (!attr_ok) return GVQA_ERR_CUDA;
  const int tiles = (int)((m + kBM - 1) / kBM) * ((n + kBN - 1) / kBN);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  proj_gemm_3xtf32_kernel<<<grid, kGemmThreads, kGemmSmem, static_cast<cudaStream_t>(stream_)>>>(
      map_a, map_bhi, map_blo, c, ldc, (int)m, n, k);
  GVQA_LAUNCH_CHECK();
  return GVQA_OK;
}
