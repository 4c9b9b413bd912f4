// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, M=128) on B200 by kind (tf32 / bf16), A source
// (shared-memory descriptor / tensor memory) and N.  One CTA per SM, one thread issues `reps` MMAs back to
// back, commits to an mbarrier and waits; cycles per MMA from clock64().  Operand contents are zeros.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../graphvqa_b200/csrc/common.cuh"
using namespace gvqa;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3fff) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

template <int KIND, int ASRC>  // KIND 0 tf32, 1 f16(bf16);  ASRC 0 smem, 1 tmem
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (KIND == 0 && ASRC == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 1 && ASRC == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 0 && ASRC == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 1 && ASRC == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xffffffffu));
  return pred;
}

template <int KIND, int ASRC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int nacc, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {     // whole warp; the MMAs are issued by one elected lane (no per-instruction elect loop)
    const uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 : bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16 * 1024;
    // accumulators: nacc of them, N columns each, from column 0; A (tmem) lives in the last 64 columns
    const uint32_t a_t = tmem + 448;
    for (int round = 0; round < 2; ++round) {   // round 0 warms up
      const long long t0 = clock64();
      if (elect_one()) {
        int acc = 0;
#pragma unroll 4
        for (int i = 0; i < reps; ++i) {
          const int k = i & 3;
          const uint32_t d = tmem + (uint32_t)(acc * N);
          acc = acc + 1 == nacc ? 0 : acc + 1;
          mma<KIND, ASRC>(d, a_t + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 32 * k), idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, round & 1);
      const long long t2 = clock64();
      if (round == 1 && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// Same MMAs, but issued the way the projection GEMM does it: groups of 12 with (mode bits)
// 1: two tcgen05.commit per group, 2: tcgen05.fence::after_thread_sync per group, 4: leave / re-enter the
// elected region (__syncwarp) per group, 8: a (satisfied) mbarrier wait per group, 16: 3 accumulators a la 3xTF32
template <int KIND, int ASRC>
__global__ void __launch_bounds__(320, 1) group_kernel(int N, int groups, int mode, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2, bar3, bar4;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_init(&bar4, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)) : "memory"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16 * 1024;
    const uint32_t a_t = tmem + 448;
    long long t0 = 0;
    for (int round = 0; round < 2; ++round) {
      t0 = clock64();
      if (mode & 4) {
        for (int g = 0; g < groups; ++g) {
          if (mode & 8) mbar_wait(&bar3, 0);
          if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma<KIND, ASRC>(tmem + 256, a_t + 32 + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 32 * k), idesc, 1);
              mma<KIND, ASRC>(tmem + 256, a_t + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 16384 + 32 * k), idesc, 1);
              mma<KIND, ASRC>(tmem + ((mode & 16) ? 128 * (g & 1) : 0), a_t + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 32 * k), idesc, 1);
            }
            if (mode & 1) {
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
          }
          __syncwarp();
        }
        if (elect_one())
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
      } else {
        if (elect_one()) {
          for (int g = 0; g < groups; ++g) {
            if (mode & 8) mbar_wait(&bar3, 0);
            if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma<KIND, ASRC>(tmem + 256, a_t + 32 + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 32 * k), idesc, 1);
              mma<KIND, ASRC>(tmem + 256, a_t + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 16384 + 32 * k), idesc, 1);
              mma<KIND, ASRC>(tmem + ((mode & 16) ? 128 * (g & 1) : 0), a_t + 8 * k, desc_sw128(a_s + 32 * k), desc_sw128(b_s + 32 * k), idesc, 1);
            }
            if (mode & 1) {
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
      }
      mbar_wait(&bar, round & 1);
      const long long t2 = clock64();
      if (round == 1 && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t2 - t0; out[1] = t2 - t0; }
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar4)) : "memory");
  } else if (mode & 32) {
    mbar_wait(&bar4, 0);     // 9 warps spinning on try_wait, like the idle roles of the GEMM
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int KIND, int ASRC>
int run(int N, int nacc, int grid, long long* d_out) {
  const int reps = 512;
  const size_t smem = 49 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<KIND, ASRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_kernel<KIND, ASRC><<<grid, 128, smem>>>(N, reps, nacc, d_out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  const int kelems = KIND == 0 ? 8 : 16;
  const double cyc = (double)h[1] / reps;
  printf("%-5s A=%-4s N=%3d nacc=%d grid=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA -> %6.0f FMA/clk/SM (K=%d per MMA)\n",
         KIND == 0 ? "tf32" : "bf16", ASRC ? "tmem" : "smem", N, nacc, grid, (double)h[0] / reps, cyc,
         128.0 * N * kelems / cyc, kelems);
  return 0;
}

// queue depth probe: one elected thread issues K MMAs (N=128 tf32, 64 cycles each) after the pipe is idle;
// out[0] = cycles until the issue loop is done, out[1] = until the warp has reconverged after the elected
// region, out[2] = until the commit barrier fires
__global__ void __launch_bounds__(128, 1) depth_kernel(int K, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16 * 1024;
    for (int round = 0; round < 2; ++round) {
      long long t1 = 0;
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < K; ++i)
          mma<0, 1>(tmem + 128 * (i & 1), tmem + 448 + 8 * (i & 3), 0, desc_sw128(b_s + 32 * (i & 3)), idesc, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
      }
      __syncwarp();
      const long long t2 = clock64();
      mbar_wait(&bar, round & 1);
      const long long t3 = clock64();
      t1 = __shfl_sync(0xffffffffu, t1, 0);   // elect.sync picks the lowest active lane
      if (round == 1 && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = t3 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int run_group(int mode, long long* d_out) {
  const int groups = 64;
  const size_t smem = 49 * 1024 + 1024;
  CK(cudaFuncSetAttribute(group_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  group_kernel<0, 1><<<148, 320, smem>>>(128, groups, mode, d_out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("group mode=%2d (1 commits, 2 fence, 4 re-elect, 8 mbar wait, 16 alt acc, 32 spinning warps): %7.1f cycles per group of 12 tf32 N=128 MMAs (768 ideal)\n",
         mode, (double)h[1] / groups);
  return 0;
}

int main() {
  long long* d_out0;
  CK(cudaMalloc(&d_out0, 32));
  CK(cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
  for (int K : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48}) {
    depth_kernel<<<148, 128, 50 * 1024>>>(K, d_out0);
    CK(cudaDeviceSynchronize());
    long long h[3];
    CK(cudaMemcpy(h, d_out0, sizeof(h), cudaMemcpyDeviceToHost));
    printf("depth K=%2d MMAs (%4d cycles of work): issue loop done %5lld, warp reconverged %5lld, commit fired %5lld\n", K, K * 64, h[0], h[1], h[2]);
  }
  for (int mode : {0, 15, 31, 32, 32 + 15, 32 + 31}) if (run_group(mode, d_out0)) return 1;
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  for (int grid : {148}) {
    for (int N : {64, 128, 256}) {
      const int nacc = N == 256 ? 1 : 2;
      if (run<0, 0>(N, nacc, grid, d_out)) return 1;
      if (run<0, 1>(N, nacc, grid, d_out)) return 1;
      if (run<1, 0>(N, nacc, grid, d_out)) return 1;
      if (run<1, 1>(N, nacc, grid, d_out)) return 1;
    }
  }
  run<0, 1>(128, 1, 148, d_out);
  run<0, 1>(128, 3, 148, d_out);
  run<1, 1>(128, 1, 148, d_out);
  return 0;
}
