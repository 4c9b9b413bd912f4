// Shared device helpers for the gvqa_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/gvqa_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gvqa_b200 kernels are written for sm_100a (B200) only"
#endif

namespace gvqa {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define GVQA_LAUNCH_CHECK()                          \
  do {                                               \
    if (cudaPeekAtLastError() != cudaSuccess) {      \
      (void)cudaGetLastError();                      \
      return GVQA_ERR_CUDA;                          \
    }                                                \
  } while (0)

__host__ __device__ inline bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ float leaky_relu(float v, float slope) { return v > 0.f ? v : v * slope; }

// 128-bit read-only global load that does not allocate in L1 (streamed once per CTA).
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// 128-bit read-only global load through L1 (rows that are gathered more than once).
__device__ __forceinline__ float4 ldg_cached(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void stg_stream(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void fma4(float4& acc, float a, const float4& v) {
  acc.x = fmaf(a, v.x, acc.x);
  acc.y = fmaf(a, v.y, acc.y);
  acc.z = fmaf(a, v.z, acc.z);
  acc.w = fmaf(a, v.w, acc.w);
}

// ---- mbarrier + bulk-copy (TMA engine, UBLKCP in SASS) wrappers --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned),
// completion signalled on `bar` through complete_tx.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A kernel launched with launch_pdl() may begin while its predecessor in the stream is still draining: its CTAs
// are scheduled as the predecessor's CTAs exit, run whatever does not depend on the predecessor, and block in
// pdl_wait() until the predecessor grid has completed and its writes are visible.  Without the launch attribute
// (or with a non-kernel predecessor such as an event record) both calls are no-ops.
// Rule used by every PDL kernel here: independent set-up -> pdl_wait() -> pdl_launch_dependents() -> the rest.
// Triggering only after the own wait keeps the chain transitive: when kernel n+1 starts, kernel n has passed its
// wait, so everything before kernel n is complete.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// GVQA_PDL environment variable: bit0 enables the attribute for the projection GEMM, bit1 for the fused hop.
// Default 1: measured at cfg2 (B200, graph replay) the GEMM gains ~1.6 us per launch, while the hop kernel LOSES
// 5-7 us per launch (0.384 -> 0.410 ms/step with the index round trips before the wait, 0.419 with the wait at
// the top): a plainly stream-ordered launch already overlaps its ramp with the predecessor's drain, and
// griddepcontrol.wait additionally waits for the predecessor's memory flush.  The hop keeps its pdl_wait() calls
// (no-ops without the attribute).
inline int pdl_mask() {
  static const int mask = [] {
    const char* e = getenv("GVQA_PDL");
    return e ? atoi(e) : 1;
  }();
  return mask;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace gvqa
