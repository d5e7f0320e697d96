// Test-only diagnostic (built into tests/native/libtpp_probe.so, NOT part of the product library or its ABI): one TMA im2col load (cp.async.bulk.tensor.4d ... .im2col) of an NHWC fp32 tensor into shared
// memory, copied out linearly.  tests/test_conv_ops.py uses it to pin the descriptor / coordinate conventions that
// the implicit-GEMM convolution path of csrc/gemm_tc.cu relies on (bounding-box corners, base pixel coordinates, tap
// offsets, zero fill of padding pixels, of channels beyond C and of pixels beyond the last image).
#include <cuda.h>

#include "../../train-procgen-pytorch_b200/csrc/tpp_common.cuh"

namespace tpp {
namespace probe {

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  return fn;
}

__global__ void __launch_bounds__(128) im2col_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int w, int h,
                                                           int n, int off_w, int off_h, int bytes,
                                                           float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -777.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2], {%7, %8};" ::"r"(dst),
        "l"(&tm), "r"(bar_a), "r"(c0), "r"(w), "r"(h), "r"(n), "h"((unsigned short)off_w), "h"((unsigned short)off_h)
        : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar_a), "r"(0)
        : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

}  // namespace probe
}  // namespace tpp

extern "C" int tpp_debug_tma_im2col(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t channels_per_pixel,
                                    int32_t pixels, int32_t w, int32_t h, int32_t n, int32_t off_w, int32_t off_h,
                                    int32_t swizzle, float* out, void* stream) {
  TPP_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0 && pixels > 0 && pixels <= 256);
  TPP_CHECK_ARG(channels_per_pixel > 0 && channels_per_pixel <= 64);
  auto enc = tpp::probe::encode_im2col();
  if (!enc) return TPP_ENOTSUP;
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lower[2] = {-1, -1}, upper[2] = {-1, -1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, lower, upper,
                   (cuuint32_t)channels_per_pixel, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1000 + (int)r;
  const int bytes = pixels * channels_per_pixel * 4;
  tpp::probe::im2col_probe_kernel<<<1, 128, bytes, tpp_stream(stream)>>>(tm, 0, w, h, n, off_w, off_h, bytes, out);
  TPP_LAUNCH_STATUS();
}
