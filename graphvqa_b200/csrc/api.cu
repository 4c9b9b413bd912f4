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
