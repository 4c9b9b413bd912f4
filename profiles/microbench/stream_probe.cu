// Micro-benchmark: what does the B200 memory system give a K1-sized problem (63 MB read + 16 MB
// read + 16 MB write, ~95 MB) under different access strategies?  Used to choose the design of
// the fused GAT hop kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../graphvqa_b200/csrc/common.cuh"
using namespace gvqa;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// A. plain streaming: out[i] = sum of 4 strided float4 (reads 4N, +N skip, writes N float4)
__global__ void copy_f4(const float4* __restrict__ x, const float4* __restrict__ skip, float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = ldg_stream((const float*)(x + i)), b = ldg_stream((const float*)(x + n4 + i));
    float4 c = ldg_stream((const float*)(x + 2 * n4 + i)), d = ldg_stream((const float*)(x + 3 * n4 + i));
    float4 s = ldg_stream((const float*)(skip + i));
    float4 o = make_float4(a.x + b.x + c.x + d.x + s.x, a.y + b.y + c.y + d.y + s.y, a.z + b.z + c.z + d.z + s.z, a.w + b.w + c.w + d.w + s.w);
    stg_stream((float*)(out + i), o);
  }
}

// B. TMA ring: persistent CTAs; each "tile" = R rows x SEG bytes gathered from rows of stride ROWB bytes
// (mimics node rows), streamed through a D-deep ring; consumers sum the tile rows into R/4... simple reduce
// and write tile_bytes/4 of output.
template <int D>
__global__ void __launch_bounds__(256) tma_ring(const float* __restrict__ x, const float* __restrict__ skip, float* __restrict__ out,
                                               int64_t total_tiles, int rows, int seg_floats, int64_t row_stride_floats, int tiles_per_rowblock) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tile_floats = rows * seg_floats;
  float* ring = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)D * tile_floats * 4);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { for (int b = 0; b < D; ++b) mbar_init(&full[b], 1); mbar_fence_init(); }
  __syncthreads();
  auto issue = [&](int64_t tile, int b) {
    // tile -> (rowblock rb, segment s): rows rb*rows .. +rows, columns s*seg
    const int64_t rb = tile / tiles_per_rowblock; const int s = (int)(tile - rb * tiles_per_rowblock);
    if (lane == 0) mbar_expect_tx(&full[b], (uint32_t)(tile_floats * 4));
    __syncwarp();
    for (int r = lane; r < rows; r += 32)
      bulk_g2s(ring + (size_t)b * tile_floats + (size_t)r * seg_floats, x + (rb * rows + r) * row_stride_floats + (int64_t)s * seg_floats, (uint32_t)(seg_floats * 4), &full[b]);
  };
  int64_t my = blockIdx.x; int it = 0;
  if (wid == 0) { int64_t t = my; for (int b = 0; b < D && t < total_tiles; ++b, t += gridDim.x) issue(t, b); }
  for (int64_t tile = my; tile < total_tiles; tile += gridDim.x, ++it) {
    const int b = it % D;
    mbar_wait(&full[b], (uint32_t)((it / D) & 1));
    const float* st = ring + (size_t)b * tile_floats;
    // consume: output quarter-size: out[tile][i] = sum over 4 consecutive row-groups + skip
    const int out_f4 = tile_floats / 16;
    for (int i = tid; i < out_f4; i += 256) {
      float4 a = *reinterpret_cast<const float4*>(st + 4 * i), bq = *reinterpret_cast<const float4*>(st + 4 * (i + out_f4));
      float4 c = *reinterpret_cast<const float4*>(st + 4 * (i + 2 * out_f4)), d = *reinterpret_cast<const float4*>(st + 4 * (i + 3 * out_f4));
      float4 s = ldg_stream(skip + (tile * out_f4 + i) * 4);
      float4 o = make_float4(a.x + bq.x + c.x + d.x + s.x, a.y + bq.y + c.y + d.y + s.y, a.z + bq.z + c.z + d.z + s.z, a.w + bq.w + c.w + d.w + s.w);
      stg_stream(out + (tile * out_f4 + i) * 4, o);
    }
    __syncthreads();
    const int64_t nxt = tile + (int64_t)D * gridDim.x;
    if (wid == 0 && nxt < total_tiles) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(nxt, b); }
  }
}

int main() {
  const int64_t N = 7680, HC = 2048, C = 512;
  const int64_t xl_floats = N * HC, h_floats = N * C;
  float *x, *skip, *out, *flush;
  CK(cudaMalloc(&x, xl_floats * 4)); CK(cudaMalloc(&skip, h_floats * 4)); CK(cudaMalloc(&out, h_floats * 4));
  const size_t flush_bytes = 512ull << 20; CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(x, 0, xl_floats * 4)); CK(cudaMemset(skip, 0, h_floats * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto&& launch) {
    float best = 1e9, sum = 0; const int reps = 10;
    for (int r = 0; r < reps + 2; ++r) {
      cudaMemsetAsync(flush, r, flush_bytes);   // evict L2
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    const double bytes = (xl_floats + 2 * h_floats) * 4.0;
    printf("%-44s best %7.2f us  mean %7.2f us  -> %6.0f GB/s (best)\n", name, best * 1e3, sum / reps * 1e3, bytes / (best * 1e-3) / 1e9);
  };
  for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    char nm[64]; snprintf(nm, 64, "copy_f4 grid=%d x256", g);
    timeit(nm, [&] { copy_f4<<<g, 256>>>((const float4*)x, (const float4*)skip, (float4*)out, h_floats / 4); });
  }
  // TMA ring: rows=120 (30 nodes x 4 heads as separate 'rows' of the [N*H, C] view), seg in floats
  struct Cfg { int rows, seg; int ctas_per_sm; int D; };
  const Cfg cfgs[] = {{120, 128, 3, 1}, {120, 128, 1, 3}, {120, 64, 2, 3}, {120, 64, 1, 6}, {120, 32, 2, 6}, {30, 512, 1, 3}, {30, 512, 2, 1}, {30, 256, 2, 3}, {30, 256, 1, 6}, {32, 2048, 1, 1}};
  for (const Cfg& c : cfgs) {
    // view x as [N*H rows][C floats] when rows=120 (row stride C), or [N rows][HC] when rows=30/32
    const bool headrows = c.rows == 120;
    const int64_t row_stride = headrows ? C : HC;
    const int64_t nrows = headrows ? N * 4 : N;
    const int tiles_per_rb = (int)((headrows ? C : HC) / c.seg);
    const int64_t total = (nrows / c.rows) * tiles_per_rb;
    const size_t smem = (size_t)c.D * c.rows * c.seg * 4 + 8 * c.D + 16;
    char nm[96]; snprintf(nm, 96, "tma_ring rows=%d seg=%dB ctas/sm=%d D=%d smem=%zuK", c.rows, c.seg * 4, c.ctas_per_sm, c.D, smem >> 10);
    const int grid = 148 * c.ctas_per_sm;
#define RUN(DD) { CK(cudaFuncSetAttribute(tma_ring<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      timeit(nm, [&] { tma_ring<DD><<<grid, 256, smem>>>(x, skip, out, total, c.rows, c.seg, row_stride, tiles_per_rb); }); }
    if (c.D == 1) RUN(1) else if (c.D == 3) RUN(3) else RUN(6)
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("  -> error %s\n", cudaGetErrorString(e)); return 1; }
  }
  return 0;
}
