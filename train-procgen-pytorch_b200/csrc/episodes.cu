// Device-side episode accounting for the logger.
//
// Replaces (reference, host Python): the O(T*N) double loop of Logger.feed (common/logger.py:119-147) and the
// [T, N] reward / done batches Storage.fetch_log_data hands it (common/storage.py:130-162).  Only what the logger
// keeps leaves the device: the number of episodes that finished during the rollout and the LAST `keep` of them
// (40 = the maxlen of the reference's deques, common/logger.py:33-35) as (return, length) records in the reference's
// ENV-MAJOR order (`for i in range(n_envs): for j in range(steps)`), i.e. a few hundred bytes per iteration instead of
// 5 bytes per env-step.  The running return / length of every env's open episode stays in HBM between rollouts.
#include "tpp_common.cuh"

namespace tpp {

// episodes finished per env during this rollout
__global__ void __launch_bounds__(128) episode_count_kernel(const uint8_t* __restrict__ done, int T, int N, int64_t ld,
                                                            int32_t* __restrict__ cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  int c = 0;
#pragma unroll 8
  for (int t = 0; t < T; ++t) c += done[(int64_t)t * ld + e] != 0;
  cnt[e] = c;
}

// in-place exclusive scan of cnt[0..N) by ONE CTA (cnt[N] = total): each thread owns a contiguous segment
__global__ void __launch_bounds__(1024) episode_offsets_kernel(int32_t* cnt, int N) {
  __shared__ int32_t part[1024];
  const int tid = threadIdx.x;
  const int seg = (N + 1023) / 1024;
  const int lo = tid * seg, hi = min(N, lo + seg);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {          // Hillis-Steele inclusive scan of the 1024 partial sums
    const int v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = tid ? part[tid - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    const int c = cnt[i];
    cnt[i] = run;
    run += c;
  }
  if (tid == 1023) cnt[N] = part[1023];
}

// walk each env's T steps, close episodes at done flags, keep the last `keep` records of the env-major order
__global__ void __launch_bounds__(128) episode_emit_kernel(const float* __restrict__ rew,
                                                           const uint8_t* __restrict__ done, int T, int N, int64_t ld,
                                                           double* __restrict__ run_ret, int32_t* __restrict__ run_len,
                                                           const int32_t* __restrict__ off, double* __restrict__ out,
                                                           int keep) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = off[N];
  const int first = total > keep ? total - keep : 0;
  if (e == 0) {
    out[0] = (double)total;
    out[1] = (double)(total - first);
  }
  if (e >= N) return;
  double acc = run_ret[e];
  int len = run_len[e];
  int g = off[e];
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const int64_t o = (int64_t)t * ld + e;
    acc += (double)rew[o];
    ++len;
    if (done[o]) {
      if (g >= first) {
        out[2 + 2 * (g - first)] = acc;
        out[3 + 2 * (g - first)] = (double)len;
      }
      ++g;
      acc = 0.0;
      len = 0;
    }
  }
  run_ret[e] = acc;
  run_len[e] = len;
}

}  // namespace tpp

extern "C" int tpp_episode_scan(const float* rew, const uint8_t* done, int32_t T, int32_t N, int64_t ld,
                                double* run_ret, int32_t* run_len, int32_t* scratch, double* out, int32_t keep,
                                void* stream) {
  TPP_CHECK_ARG(rew && done && run_ret && run_len && scratch && out && T > 0 && N > 0 && ld >= N && keep > 0);
  cudaStream_t s = tpp_stream(stream);
  const int grid = tpp_ceil_div(N, 128);
  tpp::episode_count_kernel<<<grid, 128, 0, s>>>(done, T, N, ld, scratch);
  tpp::episode_offsets_kernel<<<1, 1024, 0, s>>>(scratch, N);
  tpp::episode_emit_kernel<<<grid, 128, 0, s>>>(rew, done, T, N, ld, run_ret, run_len, scratch, out, keep);
  TPP_LAUNCH_STATUS();
}
