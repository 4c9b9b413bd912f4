/*
 * gvqa_b200_debug.h -- profiling hooks of libgvqa_b200.so.  NOT part of the production C ABI (gvqa_b200.h):
 * these set process-global state read by the next launches and are meant for the micro-benchmarks under
 * profiles/microbench/ only.  Set them while no other host thread is launching; reset to NULL / 0 afterwards.
 */
#ifndef GVQA_B200_DEBUG_H_
#define GVQA_B200_DEBUG_H_

#include "gvqa_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* device buffer of 1100*8 int64 that CTA 0 of the projection GEMM fills with clock64() pipeline timestamps;
 * NULL = off (profiles/microbench/trace_gemm_f16.py) */
GVQA_API void gvqa_debug_set_gemm_trace(long long* device_buffer);
/* bit0 skip TMA loads, bit1 skip the A converters, bit2 skip the epilogue (results are then wrong; attributes
 * kernel time to stages, profiles/microbench/gemm_dbg.py).  0 = production behaviour. */
GVQA_API void gvqa_debug_set_gemm_flags(int flags);
/* per-CTA %globaltimer stamps of the hop kernels ([grid][8] uint64).  Block kernel: start, row pointers loaded,
 * logit terms loaded, softmax done, finish time of warps 0..3.  Warp-specialised kernel: start, first chunk ready,
 * chunks processed, SM id, finish time of consumer warps 0..3.  NULL disables (profiles/microbench/hop_trace.py). */
GVQA_API void gvqa_debug_set_hop_trace(unsigned long long* device_buffer);

/* device buffer of 1100*8 uint64 that CTA 0 of the fused aggregate-project hop fills with clock64() stamps
 * ([stage][0] producer, [1] converter past tma_full, [2] aggregated, [3] stored, [4]/[5] MMA past a_ready of the two
 * sub-blocks, [7] committed; [1024 + item][0..4] accumulator hand-over); NULL = off */
GVQA_API void gvqa_debug_set_fused_trace(unsigned long long* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* GVQA_B200_DEBUG_H_ */
