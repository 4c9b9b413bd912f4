// ABI version and error strings (see include/gvqa_b200.h).
#include "common.cuh"

extern "C" GVQA_API int gvqa_abi_version(void) { return GVQA_ABI_VERSION; }

extern "C" GVQA_API const char* gvqa_error_string(int status) {
  switch (status) {
    case GVQA_OK: return "ok";
    case GVQA_ERR_NULL_POINTER: return "a required pointer is NULL";
    case GVQA_ERR_BAD_SHAPE: return "negative or inconsistent sizes";
    case GVQA_ERR_UNSUPPORTED: return "unsupported configuration (channels % 4, channels > 1024, heads not in {1,2,4,8}, k > 32)";
    case GVQA_ERR_MISALIGNED: return "pointer or leading dimension is not 16-byte aligned";
    case GVQA_ERR_WORKSPACE: return "workspace too small";
    case GVQA_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown gvqa status";
  }
}

// L2 residency control for producer -> consumer tensors (x_l: written by the projection GEMM, read by
// the fused hop kernel of the same hop).  Sets / clears the stream's access-policy window; under
// CUDA-graph capture the window becomes an attribute of the captured kernel nodes.
extern "C" GVQA_API int gvqa_device_set_l2_persist_limit(size_t bytes) {
  int dev = 0, max_persist = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return GVQA_ERR_CUDA;
  if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) return GVQA_ERR_CUDA;
  if (bytes > (size_t)max_persist) bytes = (size_t)max_persist;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) return GVQA_ERR_CUDA;
  return (int)(bytes >> 20);   // MiB actually set aside
}

extern "C" GVQA_API int gvqa_stream_set_l2_window(const void* ptr, size_t bytes, float hit_ratio, void* stream_) {
  int dev = 0, max_window = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return GVQA_ERR_CUDA;
  if (cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess) return GVQA_ERR_CUDA;
  cudaStreamAttrValue v;
  v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  v.accessPolicyWindow.num_bytes = ptr ? (bytes > (size_t)max_window ? (size_t)max_window : bytes) : 0;
  v.accessPolicyWindow.hitRatio = hit_ratio;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(static_cast<cudaStream_t>(stream_), cudaStreamAttributeAccessPolicyWindow, &v) !=
      cudaSuccess)
    return GVQA_ERR_CUDA;
  return GVQA_OK;
}
