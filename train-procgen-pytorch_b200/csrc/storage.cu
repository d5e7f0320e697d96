// Rollout-storage kernels: GAE reverse scan (+ moments), advantage normalisation, minibatch gathers.
//
// Replaces (reference, torch CPU): Storage.compute_estimates (common/storage.py:56-79) and
// Storage.collate_data (common/storage.py:112-128), plus TransposeFrame/ScaledFloatFrame
// (common/env/procgen_wrappers.py:391-419) folded into the image gather.
#include <cstdlib>

#include "tpp_common.cuh"

namespace tpp {

// ------------------------------------------------------------------------------------------------
// GAE: one thread per env walks t = T-1 .. 0 (coalesced across the warp: consecutive envs are
// contiguous in [T][ld]).  The recurrence is two dependent fp32 ops per step; the loads do not depend on
// it and are issued U steps ahead.  Arithmetic uses explicit non-fused fp32 mul/add/sub in the
// reference's evaluation order, so the raw advantages and returns are bit-identical to torch CPU:
//   delta = (rew + (gamma * V[t+1]) * (1 - done)) - V[t]
//   A     = ((gamma*lambda) * A) * (1 - done) + delta
// ------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(128) gae_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                  const float* __restrict__ value, float* __restrict__ adv,
                                                  float* __restrict__ ret, double* moments, int T, int N, int64_t ld,
                                                  float gamma, float gl) {
  __shared__ double red[32];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (e < N) {
    float A = 0.0f;
    float v_next = value[(int64_t)T * ld + e];
    int t = T - 1;
    // software pipeline: the loads of batch i+1 are issued before the (serial) recurrence of batch i is evaluated
    float r[U], v[U], rn[U], vn[U];
    uint8_t d[U], dn[U];
    const bool have = t >= U - 1;
    if (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t o = (int64_t)(t - u) * ld + e;
        r[u] = __ldcs(rew + o); d[u] = __ldcs(done + o); v[u] = __ldcs(value + o);
      }
    }
    for (; t >= U - 1; t -= U) {
      const bool more = (t - U) >= U - 1;
      if (more) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t o = (int64_t)(t - U - u) * ld + e;
          rn[u] = __ldcs(rew + o); dn[u] = __ldcs(done + o); vn[u] = __ldcs(value + o);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float nd = 1.0f - (float)d[u];
        const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(__fmul_rn(gamma, v_next), nd)), v[u]);
        A = __fadd_rn(__fmul_rn(__fmul_rn(gl, A), nd), delta);
        const int64_t o = (int64_t)(t - u) * ld + e;
        __stcs(adv + o, A);
        __stcs(ret + o, __fadd_rn(A, v[u]));
        s1 += (double)A;
        s2 += (double)A * (double)A;
        v_next = v[u];
      }
      if (more) {
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = rn[u]; d[u] = dn[u]; v[u] = vn[u]; }
      }
    }
    for (; t >= 0; --t) {
      const int64_t o = (int64_t)t * ld + e;
      const float rr = rew[o], vv = value[o];
      const float nd = 1.0f - (float)done[o];
      const float delta = __fsub_rn(__fadd_rn(rr, __fmul_rn(__fmul_rn(gamma, v_next), nd)), vv);
      A = __fadd_rn(__fmul_rn(__fmul_rn(gl, A), nd), delta);
      adv[o] = A;
      ret[o] = __fadd_rn(A, vv);
      s1 += (double)A;
      s2 += (double)A * (double)A;
      v_next = vv;
    }
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(moments + 0, s1);
    atomicAdd(moments + 1, s2);
    if (blockIdx.x == 0) atomicAdd(moments + 2, (double)T * (double)N);
  }
}

// Small env counts (the PPO configs themselves: N = 256 .. 4096 per GPU).  With one thread per env there are too few
// threads to cover the load latency of 256 dependent batches (measured 20.7 us at T = 256, N = 4096, 0.19 of HBM peak).
// Here a CTA owns 32 envs and splits the work by what is parallel and what is not:
//   1. all 8 warps stage the env columns of rew / done / value for ALL T steps in shared memory (coalesced 128-byte rows,
//      every load in flight at once) and turn them into delta[t] = (rew + (gamma * V[t+1]) * (1 - done)) - V[t];
//   2. warp 0 walks the only sequential part, A = ((gamma*lambda) * A) * (1 - done) + delta -- three dependent fp32
//      operations per step, operands streamed from shared memory, nothing else on the chain;
//   3. all warps write adv / ret (= A + V) rows back and accumulate the moments.
// Every value is produced by the SAME non-fused fp32 operations in the reference's order, so raw advantages and returns
// stay bit-identical to torch CPU (an affine-map warp scan over time would re-associate the products; the sequential chain
// is ~2 us, it is not what made the small case slow).
__global__ void __launch_bounds__(256) gae_staged_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                         const float* __restrict__ value, float* __restrict__ adv,
                                                         float* __restrict__ ret, double* moments, int T, int N,
                                                         int64_t ld, float gamma, float gl) {
  extern __shared__ __align__(16) uint8_t gae_sm[];
  __shared__ double red[32];
  float* sdl = reinterpret_cast<float*>(gae_sm);           // [T][32]   rewards -> delta
  float* sv = sdl + (size_t)T * 32;                        // [T+1][32] values
  float* sa = sv + (size_t)(T + 1) * 32;                   // [T][32]   advantages
  float* snd = sa + (size_t)T * 32;                        // [T][32]   1 - done
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const bool on = e < N;
  // staging: cp.async moves rew / value rows global -> shared without passing through registers, so all 2 T row copies
  // of the CTA are in flight together (a load -> store loop keeps only a few rows per warp in flight: ~1 us per batch)
  // 16-byte copies: a lane moves four envs of a row, a warp instruction four rows (the rows are padded to ld, a multiple of
  // 32 floats, so the CTA's 32 columns are always readable and 16-byte aligned).  With 4-byte copies the kernel's time grew
  // by 32 ns per time step -- the LSU instruction count of the staging, not the recurrence, was what bounded it.
  const bool vec = ((ld & 3) == 0) && (int64_t)gridDim.x * 32 <= ld &&
                   (((reinterpret_cast<uintptr_t>(rew) | reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(adv) |
                      reinterpret_cast<uintptr_t>(ret)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(done) & 3) == 0);
  if (vec) {
    const int sub = lane >> 3, c4 = (lane & 7) * 4;          // row within the group of four, first env of this lane's chunk
    const int64_t col = (int64_t)blockIdx.x * 32 + c4;
    for (int t = warp * 4 + sub; t <= T; t += 32) {
      const int64_t o = (int64_t)t * ld + col;
      if (t < T)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdl + t * 32 + c4)),
                     "l"(rew + o) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sv + t * 32 + c4)),
                   "l"(value + o) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int t = warp * 4 + sub; t < T; t += 32) {
      const uchar4 d4 = *reinterpret_cast<const uchar4*>(done + (int64_t)t * ld + col);
      *reinterpret_cast<float4*>(snd + t * 32 + c4) =
          make_float4(1.0f - (float)d4.x, 1.0f - (float)d4.y, 1.0f - (float)d4.z, 1.0f - (float)d4.w);
    }
  } else {
  if (on) {
    for (int t = warp; t < T; t += 8) {
      const int64_t o = (int64_t)t * ld + e;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sdl + t * 32 + lane)),
                   "l"(rew + o) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sv + t * 32 + lane)),
                   "l"(value + o) : "memory");
    }
    if (warp == 0)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sv + T * 32 + lane)),
                   "l"(value + (int64_t)T * ld + e) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 8
  for (int t = warp; t < T; t += 8)
    snd[t * 32 + lane] = on ? 1.0f - (float)__ldcs(done + (int64_t)t * ld + e) : 0.0f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (on)
    for (int t = warp; t < T; t += 8) {
      const float nd = snd[t * 32 + lane];
      sdl[t * 32 + lane] = __fsub_rn(__fadd_rn(sdl[t * 32 + lane], __fmul_rn(__fmul_rn(gamma, sv[(t + 1) * 32 + lane]), nd)),
                                     sv[t * 32 + lane]);
    }
  __syncthreads();
  if (warp == 0) {
    // batches of 16 steps: operands into registers, the dependent chain on registers, results out (the compiler does
    // not move shared-memory loads across the stores of the previous step, which put a 30-cycle round trip on the chain)
    float A = 0.0f;
    for (int t0 = T - 1; t0 >= 0; t0 -= 16) {
      float nd[16], dl[16], a[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int t = t0 - u;
        nd[u] = t >= 0 ? snd[t * 32 + lane] : 0.0f;
        dl[u] = t >= 0 ? sdl[t * 32 + lane] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        A = __fadd_rn(__fmul_rn(__fmul_rn(gl, A), nd[u]), dl[u]);
        a[u] = A;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (t0 - u >= 0) sa[(t0 - u) * 32 + lane] = a[u];
    }
  }
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  if (vec) {
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const int64_t col = (int64_t)blockIdx.x * 32 + c4;
    const int nv = (int)min((int64_t)4, (int64_t)N - col);    // valid envs of this lane's chunk (<= 0: none)
    for (int t = warp * 4 + sub; t < T; t += 32) {
      const float4 a4 = *reinterpret_cast<const float4*>(sa + t * 32 + c4);
      const float4 v4 = *reinterpret_cast<const float4*>(sv + t * 32 + c4);
      const float A[4] = {a4.x, a4.y, a4.z, a4.w};
      const float R[4] = {__fadd_rn(a4.x, v4.x), __fadd_rn(a4.y, v4.y), __fadd_rn(a4.z, v4.z), __fadd_rn(a4.w, v4.w)};
      const int64_t o = (int64_t)t * ld + col;
      if (nv >= 4) {
        __stcs(reinterpret_cast<float4*>(adv + o), a4);
        __stcs(reinterpret_cast<float4*>(ret + o), make_float4(R[0], R[1], R[2], R[3]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nv) {
          if (nv < 4) { adv[o + k] = A[k]; ret[o + k] = R[k]; }
          s1 += (double)A[k];
          s2 += (double)A[k] * (double)A[k];
        }
    }
  } else if (on)
    for (int t = warp; t < T; t += 8) {
      const int64_t o = (int64_t)t * ld + e;
      const float A = sa[t * 32 + lane];
      __stcs(adv + o, A);
      __stcs(ret + o, __fadd_rn(A, sv[t * 32 + lane]));
      s1 += (double)A;
      s2 += (double)A * (double)A;
    }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(moments + 0, s1);
    atomicAdd(moments + 1, s2);
    if (blockIdx.x == 0) atomicAdd(moments + 2, (double)T * (double)N);
  }
}

// ------------------------------------------------------------------------------------------------
// GAE as a warp-level segmented scan over n_steps, fused with the advantage moments AND the normalisation
// (opt-in: Storage(gae_mode="warp_scan"); the kernels above stay the default because they are bit-exact).
//
// The recurrence A[t] = a[t] * A[t+1] + b[t], a = (gamma*lambda)(1 - done), b = delta, is a composition of affine maps;
// done = 1 makes a = 0 and cuts the chain (the segment boundary).  A CTA owns 32 envs and stages their columns in shared
// memory TRANSPOSED ([env][time], odd pitch: the coalesced global rows scatter over 32 banks); then the lanes of a warp
// split the time axis of one env: lane l composes its L = ceil(T/32) consecutive steps sequentially (2 L dependent fp32
// ops), the 32 per-lane maps are combined by a Kogge-Stone SUFFIX scan (5 shuffle rounds), and every lane replays its L
// steps from its incoming A.  Dependent chain per env: 3 L + 5 rounds instead of 3 T operations; the warp's 4 envs are
// independent chains.  The products are re-associated, so the result is NOT bit-identical to the sequential evaluation
// (measured max |dA| <= 4e-7 * max|A| at T = 256; stated tolerance 1e-5, tests/test_storage_parity.py).
// Fused normalisation: per-CTA double moments -> global atomics -> grid barrier (cooperative launch: all CTAs are
// co-resident) -> every CTA normalises ITS advantages straight from shared memory.  adv is written once, never re-read.
// ------------------------------------------------------------------------------------------------
// padded index of time step t: t + t / L when the lane stride L is even (magic = 2^32 / L + 1: t / L == umulhi(t, magic)
// for t < 2^32 / L; an integer division per staged element cost more than the scan itself)
__device__ __forceinline__ int gae_tp(int t, unsigned magic, int pad) { return t + (pad ? (int)__umulhi((unsigned)t, magic) : 0); }

__global__ void __launch_bounds__(256) gae_scan_fused_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                             const float* __restrict__ value, float* __restrict__ adv,
                                                             float* __restrict__ ret, double* moments, int T, int N,
                                                             int64_t ld, float gamma, float gl, int normalize, int PT) {
  extern __shared__ __align__(16) uint8_t gae_sm[];
  __shared__ double red[32];
  float* sb = reinterpret_cast<float*>(gae_sm);            // [32][PT]  rewards -> delta -> advantages
  float* sv = sb + (size_t)32 * PT;                        // [32][PT]  values 0 .. T
  float* sa = sv + (size_t)32 * PT;                        // [32][PT]  1 - done
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = (T + 31) >> 5;                             // time steps per lane
  const int pad = (L & 1) ? 0 : 1;                         // lane stride L (+1 if L is even): odd => conflict-free
  const unsigned magic = (unsigned)(0x100000000ull / (unsigned)L) + 1u;
  const int e0 = blockIdx.x * 32;
  const bool vec = ((ld & 3) == 0) && (int64_t)gridDim.x * 32 <= ld &&
                   (((reinterpret_cast<uintptr_t>(rew) | reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(adv) |
                      reinterpret_cast<uintptr_t>(ret)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(done) & 3) == 0);
  if (vec) {
    // ---- stage, 16 bytes per lane: four envs of a row per lane, four rows per warp instruction; the transposed stores of
    // a warp (4 consecutive time slots x 8 env groups 4 PT floats apart, PT odd) fall into 32 different banks ----
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const int64_t col = (int64_t)e0 + c4;
    for (int t = warp * 4 + sub; t <= T; t += 32) {
      const int64_t o = (int64_t)t * ld + col;
      const int q = gae_tp(t, magic, pad);
      const float4 v4 = __ldcs(reinterpret_cast<const float4*>(value + o));
      sv[(c4 + 0) * PT + q] = v4.x; sv[(c4 + 1) * PT + q] = v4.y; sv[(c4 + 2) * PT + q] = v4.z; sv[(c4 + 3) * PT + q] = v4.w;
      if (t < T) {
        const float4 r4 = __ldcs(reinterpret_cast<const float4*>(rew + o));
        const uchar4 d4 = *reinterpret_cast<const uchar4*>(done + o);
        sb[(c4 + 0) * PT + q] = r4.x; sb[(c4 + 1) * PT + q] = r4.y; sb[(c4 + 2) * PT + q] = r4.z; sb[(c4 + 3) * PT + q] = r4.w;
        sa[(c4 + 0) * PT + q] = 1.0f - (float)d4.x; sa[(c4 + 1) * PT + q] = 1.0f - (float)d4.y;
        sa[(c4 + 2) * PT + q] = 1.0f - (float)d4.z; sa[(c4 + 3) * PT + q] = 1.0f - (float)d4.w;
      }
    }
  } else {  // ---- stage (lanes over envs: coalesced rows; transposed shared-memory writes, pitch PT odd) ----
    const int e = e0 + lane;
    const bool on = e < N;
    if (on) {
      for (int t = warp; t < T; t += 8) {
        const int64_t o = (int64_t)t * ld + e;
        const int q = lane * PT + gae_tp(t, magic, pad);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sb + q)), "l"(rew + o) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sv + q)), "l"(value + o) : "memory");
      }
      if (warp == 0)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sv + lane * PT + gae_tp(T, magic, pad))),
                     "l"(value + (int64_t)T * ld + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 8
    for (int t = warp; t < T; t += 8)
      sa[lane * PT + gae_tp(t, magic, pad)] = on ? 1.0f - (float)__ldcs(done + (int64_t)t * ld + e) : 0.0f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  // ---- scan (lanes over time; warp w owns envs 4 w .. 4 w + 3, four independent chains) ----
  double s1 = 0.0, s2 = 0.0;
  {
    const int t_lo = lane * L, t_hi = min(T, t_lo + L);     // this lane's steps [t_lo, t_hi)
    float P[4], Q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* b = sb + (warp * 4 + j) * PT;
      const float* v = sv + (warp * 4 + j) * PT;
      const float* a = sa + (warp * 4 + j) * PT;
      float p = 1.0f, q = 0.0f;
      for (int t = t_hi - 1; t >= t_lo; --t) {
        const int i = t + pad * lane;                       // (t / L == lane inside this lane's chunk)
        const int i1 = t + 1 + pad * (t + 1 == t_lo + L ? lane + 1 : lane);
        const float nd = a[i], at = __fmul_rn(gl, nd);
        // delta in the reference's operation order (bit-identical to the sequential kernels' delta)
        const float dl = __fsub_rn(__fadd_rn(b[i], __fmul_rn(__fmul_rn(gamma, v[i1]), nd)), v[i]);
        b[i] = dl;
        q = fmaf(at, q, dl);                                // f_t o f_{t+1..}: x -> at (p x + q) + dl
        p = at * p;
      }
      P[j] = p; Q[j] = q;
    }
    // inclusive suffix scan of the lane maps: S_l = f_l o f_{l+1} o ... o f_31
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p2 = __shfl_down_sync(0xffffffffu, P[j], d), q2 = __shfl_down_sync(0xffffffffu, Q[j], d);
        if (lane + d < 32) { Q[j] = fmaf(P[j], q2, Q[j]); P[j] = P[j] * p2; }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float A = __shfl_down_sync(0xffffffffu, Q[j], 1);     // A entering this lane's chunk = S_{l+1}(0)
      if (lane == 31) A = 0.0f;
      float* b = sb + (warp * 4 + j) * PT;
      const float* a = sa + (warp * 4 + j) * PT;
      const bool on = e0 + warp * 4 + j < N;
      for (int t = t_hi - 1; t >= t_lo; --t) {
        const int i = t + pad * lane;
        A = fmaf(__fmul_rn(gl, a[i]), A, b[i]);
        b[i] = A;
        if (on) { s1 += (double)A; s2 += (double)A * (double)A; }
      }
    }
  }
  s1 = block_sum(s1, red);                                  // (block_sum's barriers also publish the advantages)
  s2 = block_sum(s2, red);
  float mean = 0.0f, denom = 1.0f;
  if (threadIdx.x == 0) {
    atomicAdd(moments + 0, s1);
    atomicAdd(moments + 1, s2);
    if (blockIdx.x == 0) atomicAdd(moments + 2, (double)T * (double)N);
  }
  if (normalize) {
    // grid barrier on the 4th slot of `moments` (an integer counter in its low word; the caller zeroes all four)
    unsigned int* ctr = reinterpret_cast<unsigned int*>(moments + 3);
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      while (*reinterpret_cast<volatile unsigned int*>(ctr) < gridDim.x) __nanosleep(20);
      __threadfence();
    }
    __syncthreads();
    const double n = __ldcg(moments + 2), m0 = __ldcg(moments + 0), m1 = __ldcg(moments + 1);
    const double mean_d = m0 / n;
    double var = (m1 - m0 * mean_d) / (n - 1.0);            // unbiased (torch.std default), as adv_normalize_kernel
    var = var > 0.0 ? var : 0.0;
    mean = (float)mean_d;
    denom = (float)sqrt(var) + 1e-8f;
  }
  // ---- write back (lanes over envs again) ----
  if (vec) {
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const int64_t col = (int64_t)e0 + c4;
    const int nv = (int)min((int64_t)4, (int64_t)N - col);
    for (int t = warp * 4 + sub; t < T; t += 32) {
      const int q = gae_tp(t, magic, pad);
      float A[4], R[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = sb[(c4 + k) * PT + q];
        R[k] = __fadd_rn(a, sv[(c4 + k) * PT + q]);
        A[k] = normalize ? (a - mean) / denom : a;
      }
      const int64_t o = (int64_t)t * ld + col;
      if (nv >= 4) {
        __stcs(reinterpret_cast<float4*>(adv + o), make_float4(A[0], A[1], A[2], A[3]));
        __stcs(reinterpret_cast<float4*>(ret + o), make_float4(R[0], R[1], R[2], R[3]));
      } else {
        for (int k = 0; k < nv; ++k) { adv[o + k] = A[k]; ret[o + k] = R[k]; }
      }
    }
    return;
  }
  const int e = e0 + lane;
  if (e < N)
    for (int t = warp; t < T; t += 8) {
      const int64_t o = (int64_t)t * ld + e;
      const int q = lane * PT + gae_tp(t, magic, pad);
      const float A = sb[q];
      __stcs(adv + o, normalize ? (A - mean) / denom : A);
      __stcs(ret + o, __fadd_rn(A, sv[q]));
    }
}

__global__ void __launch_bounds__(256) adv_normalize_kernel(float* adv, const double* moments, int T, int N,
                                                            int64_t ld) {
  // grid = (column chunks, T): no per-element div/mod; 128-bit accesses when the row allows it
  const double n = moments[2];
  const double mean_d = moments[0] / n;
  double var = (moments[1] - moments[0] * mean_d) / (n - 1.0);   // unbiased (torch.std default)
  var = var > 0.0 ? var : 0.0;
  const float mean = (float)mean_d;
  const float denom = (float)sqrt(var) + 1e-8f;
  float* row = adv + (int64_t)blockIdx.y * ld;
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= N) return;
  if (c4 + 4 <= N && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(adv) & 15) == 0) {
    float4 q = __ldcs(reinterpret_cast<const float4*>(row + c4));
    q.x = (q.x - mean) / denom; q.y = (q.y - mean) / denom; q.z = (q.z - mean) / denom; q.w = (q.w - mean) / denom;
    __stcs(reinterpret_cast<float4*>(row + c4), q);
  } else {
    for (int c = c4; c < N && c < c4 + 4; ++c) row[c] = (row[c] - mean) / denom;
  }
}

// ------------------------------------------------------------------------------------------------
// Minibatch gathers.  One thread per (sample, feature) with the feature index fastest, so the row-major
// output [mb][ld_out] is written fully coalesced; the 4-byte reads from the feature-major rollout are
// scattered by construction (random sample indices).  Thread (k, 0) also moves the per-sample scalars.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) {   // == cvt.rna.tf32.f32 (see csrc/gemm_tc.cu), full-rate form
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

struct GatherScalars {
  const int32_t* act; const float* logp; const float* value; const float* ret; const float* adv; const uint8_t* done;
  int32_t* o_act; float* o_logp; float* o_value; float* o_ret; float* o_adv; float* o_done;
};

__device__ __forceinline__ void gather_scalars(const GatherScalars& g, int64_t src, int k) {
  if (g.o_act) g.o_act[k] = g.act[src];
  if (g.o_logp) g.o_logp[k] = g.logp[src];
  if (g.o_value) g.o_value[k] = g.value[src];
  if (g.o_ret) g.o_ret[k] = g.ret[src];
  if (g.o_adv) g.o_adv[k] = g.adv[src];
  if (g.o_done) g.o_done[k] = (float)g.done[src];
}

__global__ void __launch_bounds__(256) gather_vec_kernel(const int64_t* __restrict__ idx, int mb, int N, int64_t ld,
                                                         int n_obs, const float* __restrict__ obs, GatherScalars g,
                                                         float* __restrict__ out_obs, float* __restrict__ out_lo,
                                                         int ld_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)mb * ld_out) return;
  const int k = (int)(i / ld_out), j = (int)(i % ld_out);
  const int64_t flat = idx[k];
  const int64_t t = flat / N, e = flat % N;
  const float v = j < n_obs ? obs[(t * n_obs + j) * ld + e] : 0.0f;
  if (out_lo) {             // TF32 pair for the tensor-core policy (hi in out_obs, residual in out_lo)
    const float h = tf32_hi(v);
    out_obs[i] = h;
    out_lo[i] = v - h;
  } else {
    out_obs[i] = v;
  }
  if (j == 0) gather_scalars(g, t * ld + e, k);
}

// frames uint8 NHWC -> float NCHW / 255.  One WARP per sample (8 samples per CTA): the H*W*C contiguous bytes of the
// frame are staged in the warp's shared-memory patch with 4-byte loads, then written channel-major with full
// 128-byte lines.  Warps never synchronise with each other, so 64 samples are in flight per SM.
constexpr int GI_WARPS = 8;
__global__ void __launch_bounds__(GI_WARPS * 32) gather_img_kernel(const int64_t* __restrict__ idx, int mb, int N,
                                                                   int64_t ld, int HW, int C,
                                                                   const uint8_t* __restrict__ frames, GatherScalars g,
                                                                   float* __restrict__ out_obs,
                                                                   float* __restrict__ out_lo, int ld_out, int raw) {
  extern __shared__ uint8_t px_all[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * (blockDim.x >> 5) + w;
  if (k >= mb) return;
  const int bytes = HW * C;
  const int patch = (bytes + 15) & ~15;
  uint8_t* px = px_all + w * patch;
  const int64_t flat = idx ? idx[k] : k;
  const uint8_t* src = frames + flat * bytes;      // flat = t*N + e indexes [T+1][N] frames directly
  if ((bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
    for (int i = lane; i < bytes / 4; i += 32)
      reinterpret_cast<uint32_t*>(px)[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
  } else {
    for (int b = lane; b < bytes; b += 32) px[b] = src[b];
  }
  __syncwarp();
  float* dst = out_obs + (int64_t)k * ld_out;
  float* dlo = out_lo ? out_lo + (int64_t)k * ld_out : nullptr;
  if (raw && !dlo && (HW & 3) == 0 && (ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    // raw pixel rows, 128-bit stores: a lane owns 4 consecutive outputs (same channel: HW % 4 == 0), 512 bytes per warp
    // instruction instead of 128 (the kernel is bound by the number of store instructions, not by their bytes)
    for (int o = 4 * lane; o < ld_out; o += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o < bytes) {
        const int cch = o / HW, pp = o - cch * HW;
        const uint8_t* q = px + pp * C + cch;
        v = make_float4((float)q[0], (float)q[C], (float)q[2 * C], (float)q[3 * C]);
      }
      *reinterpret_cast<float4*>(dst + o) = v;
    }
    if (lane == 0 && idx) gather_scalars(g, (flat / N) * ld + (flat % N), k);
    return;
  }
  int c = 0, p = lane;                              // o = c*HW + p, advanced without div/mod
  for (int o = lane; o < ld_out; o += 32) {
    float v = 0.0f;
    if (o < bytes) {
      while (p >= HW) { p -= HW; ++c; }
      // raw: the integer pixel value itself (exact in TF32: no lo half; the policy folds 1/255 into its first layer)
      v = raw ? (float)px[p * C + c] : (float)px[p * C + c] / 255.0f;
    }
    p += 32;
    if (dlo) {
      const float h = tf32_hi(v);
      dst[o] = h;
      dlo[o] = v - h;
    } else {
      dst[o] = v;
    }
  }
  if (lane == 0 && idx) gather_scalars(g, (flat / N) * ld + (flat % N), k);
}

// warps (= samples) per CTA such that the per-warp frame patches fit the default 48 KB of shared memory
static int gi_warps(int bytes) {
  const int patch = (bytes + 15) & ~15;
  int w = (48 * 1024) / patch;
  return w > GI_WARPS ? GI_WARPS : (w < 1 ? 1 : w);
}

}  // namespace tpp

extern "C" int tpp_gae(const float* rew, const uint8_t* done, const float* value, float* adv, float* ret,
                       double* moments, int32_t T, int32_t N, int64_t ld, float gamma, float lambda, void* stream) {
  TPP_CHECK_ARG(rew && done && value && adv && ret && moments && T > 0 && N > 0 && ld >= N);
  // (gamma*lambda) is formed in double and rounded once, as python does before it meets the fp32 tensor
  const float gl = (float)((double)gamma * (double)lambda);
  const size_t staged = ((size_t)T * 4 + 1) * 32 * 4;          // shared memory of the staged kernel (32 envs per CTA)
  if (N <= 8192 && staged <= 200 * 1024) {      // (beyond ~2 CTAs per SM the streaming kernel is faster: measured)
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(tpp::gae_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    tpp::gae_staged_kernel<<<tpp_ceil_div(N, 32), 256, staged, tpp_stream(stream)>>>(rew, done, value, adv, ret, moments, T,
                                                                                    N, ld, gamma, gl);
    TPP_LAUNCH_STATUS();
  }
  const int grid = tpp_ceil_div(N, 128);
  tpp::gae_kernel<8><<<grid, 128, 0, tpp_stream(stream)>>>(rew, done, value, adv, ret, moments, T, N, ld, gamma, gl);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_gae_scan(const float* rew, const uint8_t* done, const float* value, float* adv, float* ret,
                            double* moments4, int32_t T, int32_t N, int64_t ld, float gamma, float lambda,
                            int32_t normalize, void* stream) {
  TPP_CHECK_ARG(rew && done && value && adv && ret && moments4 && T > 0 && N > 0 && ld >= N);
  const float gl = (float)((double)gamma * (double)lambda);
  const int L = (T + 31) / 32;
  const int PT = (T + ((L & 1) ? 0 : T / L) + 1) | 1;       // > the padded index of slot T, odd
  const size_t smem = (size_t)3 * 32 * PT * 4;
  if (smem > 200 * 1024) return TPP_ENOTSUP;
  static int max_grid = 0;
  static size_t smem_seen = 0;
  if (smem != smem_seen) {         // (per shared-memory size: the occupancy query is not repeated on the hot path)
    cudaError_t e = cudaFuncSetAttribute(tpp::gae_scan_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tpp::gae_scan_fused_kernel, 256, smem) != cudaSuccess)
      return TPP_ENOTSUP;
    max_grid = per_sm * sms;
    smem_seen = smem;
  }
  const int grid = tpp_ceil_div(N, 32);
  if (normalize && grid > max_grid) return TPP_ENOTSUP;     // the grid barrier needs every CTA resident
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = tpp_stream(stream);
  cudaLaunchAttribute at[1];
  // Every CTA is resident by construction (grid <= max_grid, checked above), which is what the barrier needs; the
  // cooperative attribute makes the driver guarantee it whatever else shares the GPU.  It costs nothing measurable
  // (15.4 us either way, profiles/gae_scan_ab.py), so it is on unless TPP_GAE_COOP=0.
  static const bool coop = [] { const char* v = getenv("TPP_GAE_COOP"); return !(v && v[0] == '0'); }();
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = (normalize && coop) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tpp::gae_scan_fused_kernel, rew, done, value, adv, ret, moments4, (int)T, (int)N,
                                     (int64_t)ld, gamma, gl, (int)normalize, PT);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  return TPP_OK;
}

extern "C" int tpp_adv_normalize(float* adv, const double* moments, int32_t T, int32_t N, int64_t ld, void* stream) {
  TPP_CHECK_ARG(adv && moments && T > 0 && N > 0 && ld >= N);
  dim3 grid(tpp_ceil_div(N, 256 * 4), T);
  tpp::adv_normalize_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(adv, moments, T, N, ld);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_gather_vec(const int64_t* idx, int32_t mb, int32_t N, int64_t ld, int32_t n_obs, const float* obs,
                              const int32_t* act, const float* logp, const float* value, const float* ret,
                              const float* adv, const uint8_t* done, float* out_obs, float* out_obs_lo,
                              int32_t ld_out, int32_t* out_act, float* out_logp, float* out_value, float* out_ret,
                              float* out_adv, float* out_done, void* stream) {
  TPP_CHECK_ARG(idx && obs && out_obs && mb > 0 && N > 0 && ld >= N && n_obs > 0 && ld_out >= n_obs);
  tpp::GatherScalars g{act, logp, value, ret, adv, done, out_act, out_logp, out_value, out_ret, out_adv, out_done};
  const int64_t total = (int64_t)mb * ld_out;
  tpp::gather_vec_kernel<<<tpp_ceil_div(total, 256), 256, 0, tpp_stream(stream)>>>(idx, mb, N, ld, n_obs, obs, g,
                                                                                   out_obs, out_obs_lo, ld_out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_gather_img(const int64_t* idx, int32_t mb, int32_t N, int64_t ld, int32_t H, int32_t W, int32_t C,
                              const uint8_t* frames, const int32_t* act, const float* logp, const float* value,
                              const float* ret, const float* adv, const uint8_t* done, float* out_obs,
                              float* out_obs_lo, int32_t ld_out, int32_t* out_act, float* out_logp, float* out_value,
                              float* out_ret, float* out_adv, float* out_done, int32_t raw, void* stream) {
  TPP_CHECK_ARG(!(raw && out_obs_lo));
  TPP_CHECK_ARG(idx && frames && out_obs && mb > 0 && N > 0 && ld >= N && H > 0 && W > 0 && C > 0 &&
                ld_out >= H * W * C);
  const int bytes = H * W * C;
  TPP_CHECK_ARG(bytes <= 24 * 1024);  // one frame per warp in <= 48 KB of shared memory (588 B Box-World, 12 KB Procgen)
  tpp::GatherScalars g{act, logp, value, ret, adv, done, out_act, out_logp, out_value, out_ret, out_adv, out_done};
  const int warps = tpp::gi_warps(bytes);
  tpp::gather_img_kernel<<<tpp_ceil_div(mb, warps), warps * 32, warps * ((bytes + 15) & ~15), tpp_stream(stream)>>>(
      idx, mb, N, ld, H * W, C, frames, g, out_obs, out_obs_lo, ld_out, raw);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_frames_to_obs(const uint8_t* frames, int32_t N, int32_t H, int32_t W, int32_t C, float* out_obs,
                                 float* out_obs_lo, int32_t ld_out, int32_t raw, void* stream) {
  TPP_CHECK_ARG(!(raw && out_obs_lo));
  TPP_CHECK_ARG(frames && out_obs && N > 0 && ld_out >= H * W * C);
  const int bytes = H * W * C;
  TPP_CHECK_ARG(bytes <= 24 * 1024);
  tpp::GatherScalars g{};
  const int warps = tpp::gi_warps(bytes);
  tpp::gather_img_kernel<<<tpp_ceil_div(N, warps), warps * 32, warps * ((bytes + 15) & ~15), tpp_stream(stream)>>>(
      nullptr, N, N, N, H * W, C, frames, g, out_obs, out_obs_lo, ld_out, raw);
  TPP_LAUNCH_STATUS();
}
