// Action draw shared by the rollout kernels (csrc/policy.cu, csrc/rollout_fused.cu): log-softmax, Philox inverse-CDF
// sample (or argmax), log-prob, value passthrough.  Replaces dist.sample() / log_prob in PPO.predict (agents/ppo.py:72-81).
#pragma once
#include <math_constants.h>

#include "tpp_common.cuh"

namespace tpp {

constexpr int MAX_A = 16;

// One env's action: softmax over the A logits of its head row, Philox inverse-CDF draw (or argmax), log-prob, value.
__device__ __forceinline__ void sample_row(const float* __restrict__ h, int A, int env, uint64_t seed,
                                           const uint64_t* tick, uint64_t t_offset, int greedy, int32_t* act,
                                           float* logp, float* value) {
  float z[MAX_A];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < MAX_A; ++j)
    if (j < A) { z[j] = h[j]; mx = fmaxf(mx, z[j]); }
  float se = 0.0f;
#pragma unroll
  for (int j = 0; j < MAX_A; ++j)
    if (j < A) se += expf(z[j] - mx);
  const float lse = mx + logf(se);
  int a = A - 1;
  if (greedy) {
    float best = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A && z[j] > best) { best = z[j]; a = j; }
  } else {
    const uint64_t tk = (tick ? *tick : 0ull) + t_offset;
    const uint4 r = Philox(seed)((uint32_t)env, (uint32_t)tk, (uint32_t)(tk >> 32), 0x5A17u);
    const float u = u01(r.x);
    float cdf = 0.0f;
    bool found = false;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A && !found) {
        cdf += expf(z[j] - lse);
        if (u < cdf) { a = j; found = true; }
      }
  }
  float la = 0.0f;
#pragma unroll
  for (int j = 0; j < MAX_A; ++j)
    if (j == a) la = z[j] - lse;
  *act = a;
  *logp = la;
  *value = h[A];
}

}  // namespace tpp
