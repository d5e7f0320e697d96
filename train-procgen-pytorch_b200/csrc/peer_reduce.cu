// Gradient all-reduce over NVLink peer memory, fused with the gradient-norm reduction (env-sharded data parallel).
//
// The reference has no distributed code; the sharded design needs ONE sum of the flat gradient (0.6 - 2.5 MB) per
// optimizer step, followed by clip_grad_norm_ + Adam (agents/ppo.py:173-176).  At this size a ring / tree all-reduce is
// pure latency (measured with NCCL inside the epoch graph: ~75 us per step at 8 GPUs, 1.8 ms per PPO iteration), so this
// kernel does a one-shot all-reduce instead: every rank publishes its gradient in a symmetric (peer-mapped) staging
// buffer, raises a flag in every peer's signal pad, and then reads all W staging buffers directly over NVLink / NVSwitch,
// summing them in rank order -- every rank forms bit-identical sums, so replicas stay bit-identical without a broadcast.
// The same pass accumulates sum (grad_scale * g)^2 for the clip, i.e. the collective and the norm kernel are one launch;
// tpp_adam_clip_step then runs on the reduced copy.
//
// Phases of the launch (grid = co-resident CTAs):
//   A. copy my chunk of the local gradient into staging[parity] and zero the local gradient (the next backward pass
//      can start accumulating at once); the last CTA to finish A (ticket) raises flag `epoch` in every rank's pad
//      (st.release.sys after a system-scope fence);
//   B. every CTA waits until all W flags of MY pad have reached `epoch` (ld.acquire.sys), then reduces its chunk from
//      the W staging buffers (volatile 16-byte loads, fixed rank order), writes the reduced gradient and its share of
//      the squared norm.
// staging is double-buffered by epoch parity: a rank rewrites staging[p] two launches later, which it can only reach
// after every peer has entered the launch in between, i.e. has finished reading staging[p].
#include "tpp_common.cuh"

namespace tpp {

constexpr int PEER_MAX = 8;
struct PeerCtx {
  const float* staging[PEER_MAX];    // rank r's staging buffer [2][n_pad] as mapped into THIS process
  uint32_t* pad[PEER_MAX];           // rank r's signal pad (uint32 words) as mapped into this process
  int rank, world, pad_offset;
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_v4(const float* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

__global__ void __launch_bounds__(256) peer_allreduce_sqnorm_kernel(PeerCtx c, float* __restrict__ g_local,
                                                                    float* __restrict__ g_reduced, tpp_adam_state* st,
                                                                    int64_t n, int64_t n_pad, uint32_t* epoch_counter,
                                                                    unsigned int* ticket, uint32_t* error_flag) {
  __shared__ double red[32];
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) epoch_s = *epoch_counter + 1u;
  __syncthreads();
  const uint32_t epoch = epoch_s;
  float* stage = const_cast<float*>(c.staging[c.rank]) + (int64_t)(epoch & 1u) * n_pad;
  const int64_t n4 = n >> 2;
  // ---- A: publish ----
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g_local)[i];
    reinterpret_cast<float4*>(stage)[i] = v;
    reinterpret_cast<float4*>(g_local)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    stage[i] = g_local[i];
    g_local[i] = 0.0f;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {                 // last CTA of this rank: the whole gradient is staged
      __threadfence_system();
      for (int r = 0; r < c.world; ++r) st_release_sys_u32(c.pad[r] + c.pad_offset + c.rank, epoch);
    }
  }
  // ---- B: wait for every rank's flag in my pad, reduce ----
  if (threadIdx.x < c.world) {
    const uint32_t* flag = c.pad[c.rank] + c.pad_offset + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys_u32(flag) - epoch) < 0) {
      if (++spins > (1ll << 28)) {                                 // a peer never arrived: flag the error instead of hanging
        *error_flag = 1u;
        break;
      }
    }
  }
  __syncthreads();
  const float gs = st->grad_scale;
  const int slot = st->step & 1;
  const int64_t off = (int64_t)(epoch & 1u) * n_pad;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < PEER_MAX; ++r) {
      if (r < c.world) {
        const float4 v = ld_volatile_v4(c.staging[r] + off + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    reinterpret_cast<float4*>(g_reduced)[i] = acc;
    const float x0 = acc.x * gs, x1 = acc.y * gs, x2 = acc.z * gs, x3 = acc.w * gs;
    s += (double)x0 * x0 + (double)x1 * x1 + (double)x2 * x2 + (double)x3 * x3;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < c.world; ++r) {
      float v;
      asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(c.staging[r] + off + i) : "memory");
      acc += v;
    }
    g_reduced[i] = acc;
    const float x = acc * gs;
    s += (double)x * x;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&st->sqnorm[slot], s);
  // the last CTA to finish advances the epoch (read by the next launch) and re-arms the ticket
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket + 1, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *epoch_counter = epoch;
    ticket[0] = 0u;
    ticket[1] = 0u;
    __threadfence();
  }
}

}  // namespace tpp

extern "C" int tpp_peer_allreduce_sqnorm(const uint64_t* staging_ptrs, const uint64_t* pad_ptrs, int32_t rank,
                                         int32_t world, int32_t pad_offset, float* g_local, float* g_reduced,
                                         tpp_adam_state* state, int64_t n, int64_t n_pad, uint32_t* epoch_counter,
                                         uint32_t* ticket2, uint32_t* error_flag, void* stream) {
  TPP_CHECK_ARG(staging_ptrs && pad_ptrs && g_local && g_reduced && state && epoch_counter && ticket2 && error_flag);
  TPP_CHECK_ARG(world >= 1 && world <= tpp::PEER_MAX && rank >= 0 && rank < world && n > 0 && n_pad >= n && (n_pad & 3) == 0);
  TPP_CHECK_ARG(((reinterpret_cast<uintptr_t>(g_local) | reinterpret_cast<uintptr_t>(g_reduced)) & 15) == 0);
  tpp::PeerCtx c;
  for (int r = 0; r < tpp::PEER_MAX; ++r) {
    c.staging[r] = r < world ? reinterpret_cast<const float*>(staging_ptrs[r]) : nullptr;
    c.pad[r] = r < world ? reinterpret_cast<uint32_t*>(pad_ptrs[r]) : nullptr;
    if (r < world) TPP_CHECK_ARG(c.staging[r] && c.pad[r] && (staging_ptrs[r] & 15) == 0);
  }
  c.rank = rank; c.world = world; c.pad_offset = pad_offset;
  // every CTA spins on flags that this rank's LAST CTA helps to raise: the grid must be co-resident
  int grid = tpp_ceil_div(n, 256 * 8);
  if (grid > 148) grid = 148;
  tpp::peer_allreduce_sqnorm_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(c, g_local, g_reduced, state, n, n_pad,
                                                                         epoch_counter, ticket2, error_flag);
  TPP_LAUNCH_STATUS();
}
