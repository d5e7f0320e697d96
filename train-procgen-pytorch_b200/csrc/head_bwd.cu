// Backward pass of the policy/value heads in ONE kernel (the heads are [A+1 <= 16, H] — far too small for a GEMM
// launch each): given dLoss/d(logits, value) from the fused loss kernel it produces
//   dlatent = dhead @ Wh            (as a TF32 (hi, lo) pair: the A operand of the next tensor-core GEMMs)
//   gWh    += dhead^T @ latent       gbh += colsum(dhead)        gb_last += colsum(dlatent)
// Replaces four launches (two CUDA-core GEMMs, two column sums) plus a split kernel per minibatch.
// Reference: autograd through CategoricalPolicy.hidden_to_output (common/policy.py:74-87) in agents/ppo.py:170.
#include "tpp_common.cuh"

namespace tpp {

constexpr int HB_THREADS = 256;
constexpr int HB_ROWS = 64;       // samples per CTA: the latent tile [64][H] is staged in shared memory
constexpr int HB_MAX_NH = 16;

__device__ __forceinline__ float hb_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(HB_THREADS) head_backward_kernel(
    const float* __restrict__ dhead, int ld_head, const float* __restrict__ latent, const float* __restrict__ mask,
    int64_t ldl, const float* __restrict__ Wh, int nh, int H, float* __restrict__ dz_hi, float* __restrict__ dz_lo,
    float* __restrict__ dz_plain, int64_t ld_dz, float* __restrict__ gWh, float* __restrict__ gbh,
    float* __restrict__ gb_last, int mb) {
  extern __shared__ __align__(16) float sm[];
  float* sdh = sm;                                  // [HB_ROWS][HB_MAX_NH + 1]
  float* sW = sdh + HB_ROWS * (HB_MAX_NH + 1);      // [nh][H]
  float* slat = sW + HB_MAX_NH * H;                 // [HB_ROWS][H]  (16-byte aligned rows: filled with cp.async)
  float* sred = slat + HB_ROWS * H;                 // [HB_THREADS]
  const int tid = threadIdx.x;
  // A CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps its partial parameter-gradient sums in
  // registers: one flush of ~400 global atomics per CTA instead of one per 64-row tile (2048 tiles of a fused
  // accumulation window would queue 2048-deep on the same ~400 addresses: measured 85 us, mostly that).
  constexpr int GW_SLOTS = HB_MAX_NH * 256 / HB_THREADS;      // gWh entries a thread owns (o = tid + q * HB_THREADS)
  float gw_acc[GW_SLOTS];
#pragma unroll
  for (int q = 0; q < GW_SLOTS; ++q) gw_acc[q] = 0.0f;
  float gbh_acc = 0.0f, colsum = 0.0f;
  for (int i = tid; i < nh * H; i += HB_THREADS) sW[i] = Wh[i];
  const int ntiles = (mb + HB_ROWS - 1) / HB_ROWS;
  const int rows_per_pass = HB_THREADS / H;          // H in {16..256} -> 16..1 samples per pass
  const int k = tid % H, r = tid / H;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * HB_ROWS;
    const int nb = min(HB_ROWS, mb - b0);
    for (int i = tid; i < HB_ROWS * nh; i += HB_THREADS) {
      const int b = i / nh, a = i % nh;
      sdh[b * (HB_MAX_NH + 1) + a] = b < nb ? dhead[(int64_t)(b0 + b) * ld_head + a] : 0.0f;
    }
    // latent tile: 16-byte cp.async copies, all of a thread's in flight together (a load -> store loop pays one L2 round
    // trip per iteration, and a CTA walks several tiles)
    const int H4 = H >> 2;
    for (int i = tid; i < HB_ROWS * H4; i += HB_THREADS) {
      const int b = i / H4, k4 = i % H4;
      float* dst = slat + b * H + 4 * k4;
      if (b < nb) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                     "l"(latent + (int64_t)(b0 + b) * ldl + 4 * k4)
                     : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- gbh[a] += sum_b dhead[b][a] ----
    if (tid < nh) {
      for (int b = 0; b < nb; ++b) gbh_acc += sdh[b * (HB_MAX_NH + 1) + tid];
    }

    // ---- dlatent[b][k] = sum_a dhead[b][a] Wh[a][k]; thread -> (row lane r, column k), coalesced row writes ----
    for (int b = r; b < nb; b += rows_per_pass) {
      float acc = 0.0f;
      for (int a = 0; a < nh; ++a) acc += sdh[b * (HB_MAX_NH + 1) + a] * sW[a * H + k];
      const int64_t row = b0 + b;
      if (mask && !(mask[row * ldl + k] > 0.0f)) acc = 0.0f;
      colsum += acc;
      const float hi = hb_tf32(acc);
      dz_hi[row * ld_dz + k] = hi;
      dz_lo[row * ld_dz + k] = acc - hi;
      if (dz_plain) dz_plain[row * ld_dz + k] = acc;
    }

    // ---- gWh[a][k] += sum_b dhead[b][a] latent[b][k] from the staged tile ----
#pragma unroll
    for (int q = 0; q < GW_SLOTS; ++q) {
      const int o = tid + q * HB_THREADS;
      if (o < nh * H) {
        const int a = o / H, kk = o % H;
        float s = 0.0f;
#pragma unroll 8
        for (int b = 0; b < HB_ROWS; ++b) s += sdh[b * (HB_MAX_NH + 1) + a] * slat[b * H + kk];
        gw_acc[q] += s;
      }
    }
    __syncthreads();          // the staged tiles are rewritten by the next iteration
  }
  if (tid < nh) atomicAdd(gbh + tid, gbh_acc);
  sred[tid] = colsum;
  __syncthreads();
  if (tid < H) {
    float s = 0.0f;
    for (int q = 0; q < rows_per_pass; ++q) s += sred[q * H + tid];
    atomicAdd(gb_last + tid, s);
  }
#pragma unroll
  for (int q = 0; q < GW_SLOTS; ++q) {
    const int o = tid + q * HB_THREADS;
    if (o < nh * H) atomicAdd(gWh + o, gw_acc[q]);
  }
}

}  // namespace tpp

extern "C" int tpp_head_backward(const float* dhead, int32_t ld_head, const float* latent, const float* relu_mask,
                                 int64_t ldl, const float* Wh, int32_t nh, int32_t H, float* dz_hi, float* dz_lo,
                                 float* dz_plain, int64_t ld_dz, float* gWh, float* gbh, float* gb_last, int32_t mb,
                                 void* stream) {
  TPP_CHECK_ARG(dhead && latent && Wh && dz_hi && dz_lo && gWh && gbh && gb_last && mb > 0);
  TPP_CHECK_ARG(nh > 0 && nh <= tpp::HB_MAX_NH && ld_head >= nh && ldl >= H && ld_dz >= H);
  if (!(H == 64 || H == 128 || H == 256 || H == 32 || H == 16)) return TPP_ENOTSUP;
  // sized by the actual H: a 64-wide latent needs 25 KB, so 4-5 CTAs share an SM and overlap each other's tile loads
  const size_t smem = (tpp::HB_ROWS * (tpp::HB_MAX_NH + 1) + tpp::HB_MAX_NH * (size_t)H + tpp::HB_ROWS * (size_t)H +
                       tpp::HB_THREADS) * sizeof(float);
  const size_t smem_max = (tpp::HB_ROWS * (tpp::HB_MAX_NH + 1) + tpp::HB_MAX_NH * 256 + tpp::HB_ROWS * 256 +
                           tpp::HB_THREADS) * sizeof(float);
  TPP_CHECK_ARG((ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(latent) & 15) == 0);   // 16-byte row copies
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tpp::head_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_max);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = tpp_ceil_div(mb, tpp::HB_ROWS);
  const int per_sm = H <= 64 ? 4 : 2;
  if (grid > per_sm * 148) grid = per_sm * 148;   // tiles beyond the resident CTAs are walked by them
  tpp::head_backward_kernel<<<grid, tpp::HB_THREADS, smem, tpp_stream(stream)>>>(
      dhead, ld_head, latent, relu_mask, ldl, Wh, nh, H, dz_hi, dz_lo, dz_plain, ld_dz, gWh, gbh, gb_last, mb);
  TPP_LAUNCH_STATUS();
}
