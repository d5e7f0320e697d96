// Policy-side kernels: exact-fp32 CUDA-core GEMM (generic strides), bias gradient, action sampling,
// the fused PPO loss forward+backward, gradient-norm reduction and clipped Adam.
//
// Replaces (reference, eager PyTorch): nn.Linear forward/backward of MLPModel (common/model.py:954-980) and the
// policy heads (common/policy.py:74-87); dist.sample()/log_prob in PPO.predict (agents/ppo.py:72-81); the loss
// chain agents/ppo.py:131-170 with cross_batch_entropy (common/misc_util.py:32-51) and its autograd;
// clip_grad_norm_ + optim.Adam(eps=1e-5) + zero_grad (agents/ppo.py:173-176).
#include <math_constants.h>

#include "policy_sample.cuh"

namespace tpp {

// ================================================================================================
// Exact fp32 GEMM on CUDA cores:  C(m,n) (+)= epi( sum_k A(m,k) B(n,k) )
// 64x64x16 tiles, 256 threads, 4x4 register tile per thread.  The tile loaders pick the thread->element
// map whose fastest index follows the operand's contiguous dimension, so forward (K-contiguous),
// data-gradient and weight-gradient (M-contiguous) calls are all coalesced.  This is the parity /
// small-shape path; the large policy GEMMs go through the tcgen05 kernel (gemm_tc.cu).
// ================================================================================================
constexpr int BM = 64, BN = 64, BK = 16;

template <bool ROW_FAST>   // ROW_FAST: consecutive threads walk the row (m or n) index; else the k index
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t s_row, int64_t s_k, int row0, int k0,
                                          int rows, int kend, float (*sm)[BM + 4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < (BM * BK) / 256; ++i) {
    const int lin = tid + i * 256;
    int r, k;
    if (ROW_FAST) { r = lin % BM; k = lin / BM; } else { k = lin % BK; r = lin / BK; }
    const int gr = row0 + r, gk = k0 + k;
    float v = 0.0f;
    if (gr < rows && gk < kend) v = P[(int64_t)gr * s_row + (int64_t)gk * s_k];
    sm[k][r] = v;
  }
}

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                       const float* __restrict__ B, int64_t sbn, int64_t sbk,
                                                       float* __restrict__ C, int64_t ldc,
                                                       const float* __restrict__ bias, const float* __restrict__ mask,
                                                       int M, int N, int K, int flags, int k_chunk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const bool a_row_fast = (sam == 1), b_row_fast = (sbn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    if (a_row_fast) load_tile<true>(A, sam, sak, m0, k0, M, kend, As); else load_tile<false>(A, sam, sak, m0, k0, M, kend, As);
    if (b_row_fast) load_tile<true>(B, sbn, sbk, n0, k0, N, kend, Bs); else load_tile<false>(B, sbn, sbk, n0, k0, N, kend, Bs);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if ((flags & TPP_EPI_BIAS) && blockIdx.z == 0) v += bias[n];
      if (flags & TPP_EPI_RELU) v = fmaxf(v, 0.0f);
      if (flags & TPP_EPI_MASK) v = mask[(int64_t)m * ldc + n] > 0.0f ? v : 0.0f;
      float* c = C + (int64_t)m * ldc + n;
      if (split) atomicAdd(c, v);
      else if (flags & TPP_EPI_ACCUM) *c += v;
      else *c = v;
    }
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dZ, int64_t ld, int M, int N,
                                                     float* __restrict__ out) {
  // block = 32 columns x 8 row-lanes; rows strided by 8*gridDim.y
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  float s = 0.0f;
  if (c < N)
    for (int m = blockIdx.y * 8 + rl; m < M; m += 8 * gridDim.y) s += dZ[(int64_t)m * ld + c];
  part[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
}

// ================================================================================================
// Action sampling (rollout): log-softmax, inverse-CDF draw from Philox, log-prob, value passthrough
// ================================================================================================
__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ head, int ld_head, int n_envs, int A,
                                                     int32_t* __restrict__ act, float* __restrict__ logp,
                                                     float* __restrict__ value, uint64_t seed, const uint64_t* tick,
                                                     uint64_t t_offset, int greedy, int env_offset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  sample_row(head + (int64_t)e * ld_head, A, env_offset + e, seed, tick, t_offset, greedy, act + e, logp + e, value + e);
}

// Rollout tail of an MLP policy in ONE launch: last embedder layer z = act(h W^T + b) (K <= 256 inputs, 64 outputs),
// both heads (A logits + value) and the action draw.  Exact fp32 FMAs on the CUDA cores: per step these are 3 dependent
// launches of 8-13 us each on the tensor-core path (4096 rows: launch- and prologue-bound), one of ~5 us here.
// CTA = 8 warps x 4 rows; a lane owns outputs (lane, lane + 32) of its warp's 4 rows: per 4 k it reads its two weight
// float4s once (conflict-free, row pitch K + 4) and each row's activation float4 as a warp-wide broadcast -- 85 FMAs
// per shared-memory wavefront (a row-per-thread mapping is shared-memory bound at 28); the loads of a k-step are issued
// together, two warps per scheduler cover their latency.
constexpr int TAIL_ROWS = 32, TAIL_L = 64, TAIL_RPW = 4, TAIL_WARPS = TAIL_ROWS / TAIL_RPW;
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__global__ void __launch_bounds__(TAIL_WARPS * 32) mlp_tail_sample_kernel(const float* __restrict__ h, int64_t ldh, int K,
                                                              const float* __restrict__ W, const float* __restrict__ b,
                                                              int relu, const float* __restrict__ Wh,
                                                              const float* __restrict__ bh, int A, int n_rows,
                                                              float* __restrict__ head_out, int ld_head,
                                                              int32_t* __restrict__ act, float* __restrict__ logp,
                                                              float* __restrict__ value, uint64_t seed,
                                                              const uint64_t* tick, uint64_t t_offset, int greedy,
                                                              int env_offset) {
  extern __shared__ __align__(16) float tail_sm[];
  constexpr int L = TAIL_L;
  const int P = K + 4, nh = A + 1;
  float* sW = tail_sm;                      // [L][P]
  float* sH = sW + L * P;                   // [TAIL_ROWS][P]
  float* sZ = sH + TAIL_ROWS * P;           // [TAIL_ROWS][L + 1]
  float* sWh = sZ + TAIL_ROWS * (L + 1);    // [nh][L + 1]
  float* sHd = sWh + nh * (L + 1);          // [TAIL_ROWS][nh + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = blockIdx.x * TAIL_ROWS;
  const int K4 = K >> 2;
  // cp.async fill: all 16-byte copies of a thread are in flight together (a load -> store loop pays one L2 round
  // trip per iteration: measured 10 us of a 20 us kernel)
  for (int l = warp; l < L; l += TAIL_WARPS)
    for (int k4 = lane; k4 < K4; k4 += 32)
      cp_async16(sW + l * P + 4 * k4, W + (int64_t)l * K + 4 * k4);
  for (int r = warp; r < TAIL_ROWS; r += TAIL_WARPS)
    for (int k4 = lane; k4 < K4; k4 += 32) {
      if (row0 + r < n_rows) cp_async16(sH + r * P + 4 * k4, h + (int64_t)(row0 + r) * ldh + 4 * k4);
      else *reinterpret_cast<float4*>(sH + r * P + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < nh * L; i += TAIL_WARPS * 32) sWh[(i / L) * (L + 1) + (i % L)] = __ldg(Wh + i);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float acc0[TAIL_RPW], acc1[TAIL_RPW];
#pragma unroll
  for (int r = 0; r < TAIL_RPW; ++r) acc0[r] = acc1[r] = 0.0f;
  const float* w0p = sW + lane * P;
  const float* w1p = sW + (lane + 32) * P;
  const float* hp = sH + warp * TAIL_RPW * P;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(w0p + k);
    const float4 w1 = *reinterpret_cast<const float4*>(w1p + k);
    float4 xs[TAIL_RPW];
#pragma unroll
    for (int r = 0; r < TAIL_RPW; ++r) xs[r] = *reinterpret_cast<const float4*>(hp + r * P + k);
#pragma unroll
    for (int r = 0; r < TAIL_RPW; ++r) {
      const float4 x = xs[r];
      acc0[r] = fmaf(x.x, w0.x, acc0[r]); acc1[r] = fmaf(x.x, w1.x, acc1[r]);
      acc0[r] = fmaf(x.y, w0.y, acc0[r]); acc1[r] = fmaf(x.y, w1.y, acc1[r]);
      acc0[r] = fmaf(x.z, w0.z, acc0[r]); acc1[r] = fmaf(x.z, w1.z, acc1[r]);
      acc0[r] = fmaf(x.w, w0.w, acc0[r]); acc1[r] = fmaf(x.w, w1.w, acc1[r]);
    }
  }
  const float b0 = __ldg(b + lane), b1 = __ldg(b + lane + 32);
#pragma unroll
  for (int r = 0; r < TAIL_RPW; ++r) {
    float v0 = acc0[r] + b0, v1 = acc1[r] + b1;
    if (relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
    float* z = sZ + (warp * TAIL_RPW + r) * (L + 1);
    z[lane] = v0;
    z[lane + 32] = v1;
  }
  __syncthreads();
  for (int o = tid; o < TAIL_ROWS * nh; o += TAIL_WARPS * 32) {     // heads: (row, output j) pairs
    const int rr = o / nh, j = o - rr * nh;
    float v = __ldg(bh + j);
    const float* z = sZ + rr * (L + 1);
    const float* w = sWh + j * (L + 1);
#pragma unroll 8
    for (int l = 0; l < L; ++l) v = fmaf(z[l], w[l], v);
    sHd[rr * (nh + 1) + j] = v;
  }
  __syncthreads();
  if (tid < TAIL_ROWS && row0 + tid < n_rows) {
    const int e = row0 + tid;
    const float* hd = sHd + tid * (nh + 1);
    if (head_out) {
      for (int j = 0; j < nh; ++j) head_out[(int64_t)e * ld_head + j] = hd[j];
    }
    sample_row(hd, A, env_offset + e, seed, tick, t_offset, greedy, act + e, logp + e, value + e);
  }
}

// ================================================================================================
// Fused PPO loss forward + backward.  One thread per sample; raw sums reduced per CTA
// and accumulated in double.  The min/max/clamp sub-gradients follow torch autograd exactly (ties split
// 1/2-1/2, clamp passes the gradient on the closed interval), and the clipped-value arithmetic is done
// with non-fused fp32 ops in the reference's order so that tie cases fall the same way.
// ================================================================================================
__global__ void __launch_bounds__(256) ppo_pbar_kernel(const float* __restrict__ head, int ld_head, int mb, int A,
                                                       float* pbar_sum) {
  __shared__ float red[32];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float p[MAX_A];
#pragma unroll
  for (int j = 0; j < MAX_A; ++j) p[j] = 0.0f;
  if (b < mb) {
    const float* h = head + (int64_t)b * ld_head;
    float mx = -CUDART_INF_F, se = 0.0f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) { p[j] = h[j]; mx = fmaxf(mx, p[j]); }
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) { p[j] = expf(p[j] - mx); se += p[j]; }
#pragma unroll
    for (int j = 0; j < MAX_A; ++j) p[j] = j < A ? p[j] / se : 0.0f;
  }
  for (int j = 0; j < A; ++j) {
    const float s = block_sum(p[j], red);
    if (threadIdx.x == 0) atomicAdd(pbar_sum + j, s);
  }
}

__global__ void __launch_bounds__(256) ppo_loss_kernel(tpp_loss_cfg c, const float* __restrict__ head, int ld_head,
                                                       const int32_t* __restrict__ act,
                                                       const float* __restrict__ old_logp,
                                                       const float* __restrict__ old_value,
                                                       const float* __restrict__ ret, const float* __restrict__ adv,
                                                       const float* __restrict__ pbar_sum, float* __restrict__ dhead,
                                                       double* stats, int rows, int stats_stride) {
  // rows = groups * c.mb samples: group g = rows [g*mb, (g+1)*mb) is one minibatch of the reference loop (its loss
  // terms are means over mb samples, its sums go to stats + g*stats_stride).  With groups > 1, mb % 256 == 0, so a
  // block never straddles two groups.
  __shared__ double red[32];
  if (c.coef_dev) {      // run-time coefficients (one captured graph serves every value)
    c.eps_clip = c.coef_dev[0]; c.value_coef = c.coef_dev[1]; c.entropy_coef = c.coef_dev[2];
    c.entropy_multiplier = c.coef_dev[3]; c.x_entropy_coef = c.coef_dev[4];
  }
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int A = c.n_actions;
  const float invB = 1.0f / (float)c.mb;
  const int block_row0 = blockIdx.x * blockDim.x;
  stats += (int64_t)(block_row0 / c.mb) * stats_stride;
  double s_pi = 0.0, s_v = 0.0, s_ent = 0.0;
  float p[MAX_A];
#pragma unroll
  for (int j = 0; j < MAX_A; ++j) p[j] = 0.0f;
  if (b < rows) {
    const float* h = head + (int64_t)b * ld_head;
    float l[MAX_A];
    float mx = -CUDART_INF_F, se = 0.0f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) { l[j] = h[j]; mx = fmaxf(mx, l[j]); }
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) se += expf(l[j] - mx);
    const float lse = mx + logf(se);
    float H = 0.0f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) { l[j] -= lse; p[j] = expf(l[j]); H -= p[j] * l[j]; }
    const int a = act[b];
    float la = 0.0f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j == a) la = l[j];

    // ---- clipped surrogate (agents/ppo.py:131-135) ----
    const float Ah = adv[b];
    const float r = expf(la - old_logp[b]);
    const float lo = 1.0f - c.eps_clip, hi = 1.0f + c.eps_clip;
    const float s1 = r * Ah, s2 = fminf(fmaxf(r, lo), hi) * Ah;
    s_pi = (double)fminf(s1, s2);
    const float u1 = s1 < s2 ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
    const float pass_r = (r >= lo && r <= hi) ? 1.0f : 0.0f;
    const float dla = -invB * Ah * r * (u1 + (1.0f - u1) * pass_r);   // d pi_loss / d logp_a

    // ---- clipped value error (agents/ppo.py:138-142) ----
    const float v = h[A], v0 = old_value[b], R = ret[b];
    const float dvv = __fsub_rn(v, v0);
    const float vc = __fadd_rn(v0, fminf(fmaxf(dvv, -c.eps_clip), c.eps_clip));
    const float e1 = __fsub_rn(v, R), e2 = __fsub_rn(vc, R);
    const float q1 = __fmul_rn(e1, e1), q2 = __fmul_rn(e2, e2);
    s_v = (double)fmaxf(q1, q2);
    const float w1 = q1 > q2 ? 1.0f : (q1 == q2 ? 0.5f : 0.0f);
    const float pass_v = (dvv >= -c.eps_clip && dvv <= c.eps_clip) ? 1.0f : 0.0f;
    const float dv = c.value_coef * invB * (w1 * e1 + (1.0f - w1) * e2 * pass_v);
    s_ent = (double)H;

    // ---- logits gradient ----
    const float ce = c.entropy_coef * c.entropy_multiplier;
    float xq[MAX_A];
    float xdot = 0.0f;
    const bool use_x = (c.x_entropy_coef != 0.0f) && pbar_sum;
    if (use_x) {
#pragma unroll
      for (int j = 0; j < MAX_A; ++j)
        if (j < A) { xq[j] = logf(pbar_sum[j] * invB) + 1.0f; xdot += p[j] * xq[j]; }
    }
    float* dh = dhead + (int64_t)b * ld_head;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
      if (j < A) {
        const float dent = -p[j] * (l[j] + H) * invB;                 // d entropy_mean / d z_j
        float g = dla * ((j == a ? 1.0f : 0.0f) - p[j]) - ce * dent;
        if (use_x) {
          const float dM = -invB * p[j] * (xq[j] - xdot);             // d marg_entropy / d z_j
          g -= c.x_entropy_coef * (dM - dent);
        }
        dh[j] = g;
      }
    dh[A] = dv;
    for (int j = A + 1; j < ld_head; ++j) dh[j] = 0.0f;
  }
  s_pi = block_sum(s_pi, red);
  s_v = block_sum(s_v, red);
  s_ent = block_sum(s_ent, red);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, s_pi);
    atomicAdd(stats + 1, s_v);
    atomicAdd(stats + 2, s_ent);
    if (block_row0 % c.mb == 0) atomicAdd(stats + 3, (double)c.mb);
  }
  for (int j = 0; j < A; ++j) {
    const double s = block_sum((double)p[j], red);
    if (threadIdx.x == 0) atomicAdd(stats + 4 + j, s);
  }
}

// ================================================================================================
// Global-norm clip + Adam over one flat fp32 buffer
// ================================================================================================
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(tpp_adam_state* st, const float* __restrict__ g, int64_t n) {
  __shared__ double red[32];
  const float gs = st->grad_scale;
  const int slot = st->step & 1;   // step is bumped by the Adam kernel, after every CTA here has retired
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = g[i] * gs;
    s += (double)x * (double)x;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&st->sqnorm[slot], s);
}

struct WeightViews {
  tpp_weight_view v[TPP_MAX_WEIGHT_VIEWS];
  int n;
};

__global__ void __launch_bounds__(256) adam_clip_kernel(tpp_adam_state* st, float* __restrict__ p,
                                                        float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, unsigned int* ticket,
                                                        const __grid_constant__ WeightViews views) {
  __shared__ float sc[6];
  const int step0 = st->step;
  const int slot = step0 & 1;
  const int t = step0 + 1;
  if (threadIdx.x == 0) {
    // scalars are formed in double exactly as torch's python code does, then rounded once to fp32
    const double b1 = st->beta1, b2 = st->beta2;
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    const float total_norm = (float)sqrt(st->sqnorm[slot]);
    sc[0] = fminf(st->max_grad_norm / (total_norm + 1e-6f), 1.0f) * st->grad_scale;   // clip_grad_norm_ semantics
    sc[1] = (float)(1.0 - b1);
    sc[2] = (float)b2;
    sc[3] = (float)(1.0 - b2);
    sc[4] = (float)(st->lr / bc1);          // step_size
    sc[5] = (float)sqrt(bc2);               // bias_correction2_sqrt
  }
  __syncthreads();
  const float coef = sc[0], omb1 = sc[1], b2 = sc[2], omb2 = sc[3], step_size = sc[4], bc2_sqrt = sc[5];
  const float eps = (float)st->eps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * omb1;                     // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * b2 + omb2 * gi * gi;                     // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    g[i] = 0.0f;
    // the tensor-core operand copies of this parameter (TF32 (hi, lo) pairs in GEMM layout, tpp_split_tf32's arithmetic):
    // written here instead of by ~9 re-split launches behind every optimizer step
    for (int k = 0; k < views.n; ++k) {
      const tpp_weight_view& w = views.v[k];
      const int64_t j = i - w.offset;
      if (j >= 0 && j < (int64_t)w.rows * w.cols) {
        const int r = (int)(j / w.cols), c = (int)(j - (int64_t)r * w.cols);
        const float x = w.scale != 0.0f ? pi * w.scale : pi;
        const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
        const int64_t o = (int64_t)r * w.ld + (w.col_of ? w.col_of[c] : c);
        w.hi[o] = h;
        w.lo[o] = x - h;
      }
    }
  }
  // the last CTA to finish publishes step+1 and clears the OTHER accumulator for the next reduction
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    st->sqnorm[slot] = 0.0;
    st->step = t;
    *ticket = 0u;
    __threadfence();
  }
}

}  // namespace tpp

extern "C" int tpp_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                            float* C, int64_t ldc, const float* bias, const float* mask, int32_t M, int32_t N,
                            int32_t K, int32_t flags, int32_t split_k, void* stream) {
  TPP_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N);
  TPP_CHECK_ARG(!(flags & TPP_EPI_BIAS) || bias);
  TPP_CHECK_ARG(!(flags & TPP_EPI_MASK) || mask);
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && (!(flags & TPP_EPI_ACCUM) || (flags & (TPP_EPI_RELU | TPP_EPI_MASK)))) return TPP_ENOTSUP;
  int k_chunk = tpp_ceil_div(tpp_ceil_div(K, split_k), tpp::BK) * tpp::BK;
  split_k = tpp_ceil_div(K, k_chunk);
  dim3 grid(tpp_ceil_div(N, tpp::BN), tpp_ceil_div(M, tpp::BM), split_k);
  tpp::gemm_f32_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(A, sam, sak, B, sbn, sbk, C, ldc, bias, mask, M, N, K,
                                                            flags, k_chunk);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_colsum_accum(const float* dZ, int64_t ld, int32_t M, int32_t N, float* out, void* stream) {
  TPP_CHECK_ARG(dZ && out && M > 0 && N > 0 && ld >= N);
  int gy = tpp_ceil_div(M, 8 * 16);
  if (gy > 64) gy = 64;
  dim3 grid(tpp_ceil_div(N, 32), gy);
  tpp::colsum_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(dZ, ld, M, N, out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_sample_actions(const float* head, int32_t ld_head, int32_t n_envs, int32_t n_actions, int32_t* act,
                                  float* logp, float* value, uint64_t seed, const uint64_t* tick, uint64_t t_offset,
                                  int32_t greedy, int32_t env_offset, void* stream) {
  TPP_CHECK_ARG(head && act && logp && value && n_envs > 0 && n_actions > 0 && n_actions <= tpp::MAX_A);
  TPP_CHECK_ARG(ld_head > n_actions);
  tpp::sample_kernel<<<tpp_ceil_div(n_envs, 256), 256, 0, tpp_stream(stream)>>>(head, ld_head, n_envs, n_actions, act,
                                                                                logp, value, seed, tick, t_offset,
                                                                                greedy, env_offset);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_mlp_tail_sample(const float* h, int64_t ldh, int32_t K, const float* W, const float* b, int32_t L,
                                   int32_t relu, const float* Wh, const float* bh, int32_t n_actions, int32_t n_rows,
                                   float* head_out, int32_t ld_head, int32_t* act, float* logp, float* value,
                                   uint64_t seed, const uint64_t* tick, uint64_t t_offset, int32_t greedy,
                                   int32_t env_offset, void* stream) {
  TPP_CHECK_ARG(h && W && b && Wh && bh && act && logp && value && n_rows > 0);
  TPP_CHECK_ARG(K > 0 && K <= 256 && (K & 3) == 0 && ldh >= K && (ldh & 3) == 0);
  if (L != tpp::TAIL_L) return TPP_ENOTSUP;
  TPP_CHECK_ARG((reinterpret_cast<uintptr_t>(h) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  TPP_CHECK_ARG(n_actions > 0 && n_actions < tpp::MAX_A && (!head_out || ld_head > n_actions));
  const int nh = n_actions + 1, P = K + 4;
  const size_t smem = sizeof(float) * ((size_t)L * P + (size_t)tpp::TAIL_ROWS * P + (size_t)tpp::TAIL_ROWS * (L + 1) +
                                       (size_t)nh * (L + 1) + (size_t)tpp::TAIL_ROWS * (nh + 1));
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tpp::mlp_tail_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         128 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  tpp::mlp_tail_sample_kernel<<<tpp_ceil_div(n_rows, tpp::TAIL_ROWS), tpp::TAIL_WARPS * 32, smem, tpp_stream(stream)>>>(
      h, ldh, K, W, b, relu, Wh, bh, n_actions, n_rows, head_out, ld_head, act, logp, value, seed, tick, t_offset, greedy,
      env_offset);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_ppo_pbar(const float* head, int32_t ld_head, int32_t mb, int32_t n_actions, float* pbar_sum,
                            void* stream) {
  TPP_CHECK_ARG(head && pbar_sum && mb > 0 && n_actions > 0 && n_actions <= tpp::MAX_A && ld_head > n_actions);
  tpp::ppo_pbar_kernel<<<tpp_ceil_div(mb, 256), 256, 0, tpp_stream(stream)>>>(head, ld_head, mb, n_actions, pbar_sum);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_ppo_loss_fwd_bwd(const tpp_loss_cfg* cfg, const float* head, int32_t ld_head, const int32_t* act,
                                    const float* old_logp, const float* old_value, const float* ret, const float* adv,
                                    const float* pbar, float* dhead, double* stats, void* stream) {
  TPP_CHECK_ARG(cfg && head && act && old_logp && old_value && ret && adv && dhead && stats);
  TPP_CHECK_ARG(cfg->mb > 0 && cfg->n_actions > 0 && cfg->n_actions <= tpp::MAX_A && ld_head > cfg->n_actions);
  TPP_CHECK_ARG(cfg->x_entropy_coef == 0.0f || pbar);
  tpp::ppo_loss_kernel<<<tpp_ceil_div(cfg->mb, 256), 256, 0, tpp_stream(stream)>>>(
      *cfg, head, ld_head, act, old_logp, old_value, ret, adv, pbar, dhead, stats, cfg->mb, 0);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_ppo_loss_fwd_bwd_grouped(const tpp_loss_cfg* cfg, int32_t groups, const float* head, int32_t ld_head,
                                            const int32_t* act, const float* old_logp, const float* old_value,
                                            const float* ret, const float* adv, float* dhead, double* stats,
                                            int32_t stats_stride, void* stream) {
  TPP_CHECK_ARG(cfg && head && act && old_logp && old_value && ret && adv && dhead && stats);
  TPP_CHECK_ARG(cfg->mb > 0 && cfg->n_actions > 0 && cfg->n_actions <= tpp::MAX_A && ld_head > cfg->n_actions);
  TPP_CHECK_ARG(groups >= 1 && (groups == 1 || cfg->mb % 256 == 0) && stats_stride >= 4 + cfg->n_actions);
  TPP_CHECK_ARG(cfg->x_entropy_coef == 0.0f);   // the cross-batch entropy needs per-group batch means: ungrouped path
  TPP_CHECK_ARG((int64_t)groups * cfg->mb < (int64_t)1 << 31);
  const int rows = groups * cfg->mb;
  tpp::ppo_loss_kernel<<<tpp_ceil_div(rows, 256), 256, 0, tpp_stream(stream)>>>(
      *cfg, head, ld_head, act, old_logp, old_value, ret, adv, nullptr, dhead, stats, rows, stats_stride);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_grad_sqnorm(tpp_adam_state* state, const float* g, int64_t n, void* stream) {
  TPP_CHECK_ARG(state && g && n > 0);
  int grid = tpp_ceil_div(n, 256 * 8);
  if (grid > 148 * 4) grid = 148 * 4;
  tpp::grad_sqnorm_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(state, g, n);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_adam_clip_step(tpp_adam_state* state, float* p, float* g, float* m, float* v, int64_t n,
                                  void* stream) {
  return tpp_adam_clip_step_views(state, p, g, m, v, n, nullptr, 0, stream);
}

extern "C" int tpp_adam_clip_step_views(tpp_adam_state* state, float* p, float* g, float* m, float* v, int64_t n,
                                        const tpp_weight_view* views, int32_t n_views, void* stream) {
  TPP_CHECK_ARG(state && p && g && m && v && n > 0);
  TPP_CHECK_ARG(n_views >= 0 && n_views <= TPP_MAX_WEIGHT_VIEWS && (n_views == 0 || views));
  tpp::WeightViews wv;
  wv.n = n_views;
  for (int k = 0; k < n_views; ++k) {
    wv.v[k] = views[k];
    TPP_CHECK_ARG(views[k].hi && views[k].lo && views[k].rows > 0 && views[k].cols > 0 && views[k].offset >= 0 &&
                  views[k].offset + (int64_t)views[k].rows * views[k].cols <= n && views[k].ld >= views[k].cols);
  }
  int grid = tpp_ceil_div(n, 256 * 4);
  if (grid > 148 * 4) grid = 148 * 4;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(&state->ticket);
  tpp::adam_clip_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(state, p, g, m, v, n, ticket, wv);
  TPP_LAUNCH_STATUS();
}
