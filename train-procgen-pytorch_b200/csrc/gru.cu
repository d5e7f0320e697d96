// GRU cell of a recurrent policy: CategoricalPolicy(recurrent=True) passes the embedder's latent
// through nn.GRU(D, D) at PREDICTION time only -- PPO.optimize evaluates embedder + heads without it (its call through
// the policy is commented out upstream, agents/ppo.py:116-121), so there is no backward pass through time to build.
// Reference: common/model.py:219-226 (one cell step on hxs * masks), agents/ppo.py:72-81 (masks = 1 - done).
//
//   hm = h_prev * (1 - done)
//   gi = x  W_ih^T + b_ih        [N][3D]   (gate order r, z, n -- torch.nn.GRU)       \  two 3xTF32 tensor-core GEMMs
//   gh = hm W_hh^T + b_hh        [N][3D]                                                /  (tpp_gemm_tc, bias epilogue)
//   r = sigmoid(gi_r + gh_r); z = sigmoid(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1 - z) * n + z * hm
//
// Two elementwise kernels around the GEMMs: tpp_gru_mask_split forms the masked state as the GEMM's TF32 (hi, lo)
// operand; tpp_gru_gates finishes the cell and writes h' as plain fp32 (the next step's state, the rollout's
// hidden_states_batch slot) and as the (hi, lo) pair the head GEMM reads.
#include "tpp_common.cuh"

namespace tpp {

__device__ __forceinline__ float gru_tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(256) gru_mask_split_kernel(const float* __restrict__ h, int64_t ldh,
                                                             const uint8_t* __restrict__ done, int N, int D,
                                                             float* __restrict__ hi, float* __restrict__ lo, int64_t ld) {
  // one thread per (row, padded column): columns >= D of the pair are zero-filled (TMA reads whole 32-float k-blocks)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * ld) return;
  const int row = (int)(i / ld), col = (int)(i - (int64_t)row * ld);
  float v = 0.0f;
  if (col < D && !(done && done[row])) v = h[(int64_t)row * ldh + col];
  const float a = gru_tf32_round(v);
  hi[i] = a;
  lo[i] = gru_tf32_round(v - a);
}

__device__ __forceinline__ float gru_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) gru_gates_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                        int64_t ldg, const float* __restrict__ h_prev, int64_t ldh,
                                                        const uint8_t* __restrict__ done, int N, int D,
                                                        float* __restrict__ h_out, int64_t ldo, float* __restrict__ hi,
                                                        float* __restrict__ lo, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * ld) return;
  const int row = (int)(i / ld), col = (int)(i - (int64_t)row * ld);
  float v = 0.0f;
  if (col < D) {
    const float* a = gi + (int64_t)row * ldg;
    const float* b = gh + (int64_t)row * ldg;
    // h_prev may alias h_out (the bootstrap step updates the state in place): this thread reads its element first
    const float hm = (done && done[row]) ? 0.0f : h_prev[(int64_t)row * ldh + col];
    const float r = gru_sigmoid(a[col] + b[col]);
    const float z = gru_sigmoid(a[D + col] + b[D + col]);
    const float n = tanhf(a[2 * D + col] + r * b[2 * D + col]);
    v = (1.0f - z) * n + z * hm;
    h_out[(int64_t)row * ldo + col] = v;
  }
  if (hi) {
    const float t = gru_tf32_round(v);
    hi[i] = t;
    lo[i] = gru_tf32_round(v - t);
  }
}

}  // namespace tpp

extern "C" int tpp_gru_mask_split(const float* h, int64_t ldh, const uint8_t* done, int32_t N, int32_t D, float* hm_hi,
                                  float* hm_lo, int64_t ld, void* stream) {
  TPP_CHECK_ARG(h && hm_hi && hm_lo && N > 0 && D > 0 && ldh >= D && ld >= D);
  const int64_t total = (int64_t)N * ld;
  tpp::gru_mask_split_kernel<<<tpp_ceil_div(total, 256), 256, 0, tpp_stream(stream)>>>(h, ldh, done, N, D, hm_hi, hm_lo, ld);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_gru_gates(const float* gi, const float* gh, int64_t ldg, const float* h_prev, int64_t ldh,
                             const uint8_t* done, int32_t N, int32_t D, float* h_out, int64_t ldo, float* out_hi,
                             float* out_lo, int64_t ld, void* stream) {
  TPP_CHECK_ARG(gi && gh && h_prev && h_out && N > 0 && D > 0 && ldg >= 3 * D && ldh >= D && ldo >= D && ld >= D);
  TPP_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr));
  const int64_t total = (int64_t)N * ld;
  tpp::gru_gates_kernel<<<tpp_ceil_div(total, 256), 256, 0, tpp_stream(stream)>>>(gi, gh, ldg, h_prev, ldh, done, N, D,
                                                                                 h_out, ldo, out_hi, out_lo, ld);
  TPP_LAUNCH_STATUS();
}
