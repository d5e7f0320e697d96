// Shared device helpers for the tpp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tpp_b200.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

#define TPP_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return TPP_EINVAL; \
  } while (0)

#define TPP_LAUNCH_STATUS()                          \
  do {                                               \
    cudaError_t e__ = cudaPeekAtLastError();         \
    if (e__ != cudaSuccess) {                        \
      cudaGetLastError();                            \
      return (int)e__;                               \
    }                                                \
    return TPP_OK;                                   \
  } while (0)

static inline cudaStream_t tpp_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int tpp_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

namespace tpp {

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter-based, no per-env RNG state in HBM; the four counter words
// are (env_id, tick_lo, tick_hi, draw) and the key is the 64-bit seed.
// ---------------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}

  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
  }

  __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t key0 = k0, key1 = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      key0 += 0x9E3779B9u;
      key1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// 24-bit uniform in [0, 1)
__host__ __device__ inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA; result valid in thread 0.  `smem` must hold >= 32 T's.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  T r = T(0);
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Streaming (read-once / write-once) accesses: keep L1 clean for data that is touched exactly once.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace tpp
