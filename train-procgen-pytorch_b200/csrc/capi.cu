// Library-level entry points of the C-ABI (include/tpp_b200.h).
#include "tpp_common.cuh"

extern "C" int tpp_version(void) { return 1; }

extern "C" int tpp_device_sm_count(int* out_sm_count) {
  TPP_CHECK_ARG(out_sm_count);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(out_sm_count, cudaDevAttrMultiProcessorCount, dev);
  return (int)e;
}

extern "C" const char* tpp_error_string(int code) {
  switch (code) {
    case TPP_OK: return "ok";
    case TPP_EINVAL: return "TPP_EINVAL: bad argument (null pointer, size or alignment)";
    case TPP_ENOTSUP: return "TPP_ENOTSUP: shape or configuration not supported by this kernel";
    default: return cudaGetErrorString((cudaError_t)code);
  }
}
