// Rollout-step policy in ONE launch: the whole MLPModel forward (4 dense layers), the policy / value heads
// and the action draw for every env of a rollout step, activations never leaving the SM.
//
// Replaces (reference, per rollout step): PPO.predict = policy(obs) -> Categorical -> sample / log_prob
// (agents/ppo.py:72-81, common/policy.py:61-87, common/model.py:954-980).  Before this kernel the step ran three
// tensor-core GEMM launches + a tail kernel (4 dependent launches of 11-14 us each at 4096 envs, every one of them bound
// by launch / prologue latency and by re-reading activations and weights through L2).
//
// Decomposition.  A cluster of 4 CTAs owns a tile of 128 envs (rows).  Every layer's contraction is split over the 4
// CTAs (K-split): CTA j holds columns [j*ks, (j+1)*ks) of the layer's input (its "A slice", resident in shared memory
// as tcgen05 operand tiles) and streams the matching k-columns of the weights by TMA; tcgen05.mma (kind::tf32, 3xTF32:
// hi*hi + hi*lo + lo*hi, fp32 TMEM accumulator) produces a 128 x N partial product.  The partials are reduce-scattered
// through distributed shared memory: CTA j sends the fp32 columns that CTA p owns to p's inbox (st.shared::cluster),
// sums its own columns with the three it received (exact fp32 adds), adds the bias, applies ReLU and writes the result
// straight back as ITS A slice of the next layer (hi / lo tiles in the 128-byte-swizzled K-major layout the tensor
// core reads).  So per layer an SM ingests only 1/4 of the weights, no activation ever crosses L2, and the only
// exchange is 3/4 of a 128 x N fp32 tile per CTA over DSMEM.  The last layer's partials all go to CTA 0, which
// finishes the latent, evaluates both heads on CUDA cores and draws the action (same Philox stream as
// tpp_sample_actions).  Cross-CTA ordering is mbarrier-based (remote arrive.release.cluster / try_wait.acquire.cluster):
// "my slots are free" (peer_free) and "your slots are full" (inbox_early / inbox_late); no cluster-wide barrier in the loop.
//
// Layer 1 reads either row-major fp32 rows through TMA as an EXACT TF32 operand (integer pixel values 0..255 written
// by the Box-World step kernel: two passes, no lo half) or a feature-major rollout slot [n_obs][ld] (vector envs),
// which the epilogue threads split into hi / lo tiles themselves.
#include "policy_sample.cuh"
#include "tc_ptx.cuh"

namespace tpp {
namespace fused {
using namespace tpp::tc;

constexpr int CL = 4;                       // CTAs per cluster = K split
constexpr int NL = 4;                       // dense layers of the embedder
constexpr int TPR = 2;                      // epilogue threads per accumulator row (4: measured slower, 38.7 vs 36.5 us)
constexpr int CW = 64 / TPR, NCH = CW / 4;  // fp32 columns / 16-byte chunks of a 64-column slice per thread
constexpr int EPI_WARPS = 4 * TPR, EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;   // warp 0: TMA, warp 1: MMA + TMEM, then the epilogue warps (TPR threads per row)
constexpr int TILE = BLOCK_M * BLOCK_K * 4; // bytes of one [128][32] fp32 operand tile
// Region R (160 KB): the TMA ring.  Layer 1 uses it as 2 big stages [A tile 16 KB][W hi 32 KB][W lo 32 KB] (one k-block,
// all 256 output columns); layers 2..4 as 5 small stages [W hi 16 KB][W lo 16 KB] (one k-block of one 128-column half:
// N = 128 is the narrowest MMA shape that keeps the tensor pipe full -- N = 64 instructions take as long), so that most of
// the next layer's weights are already resident when its operand is ready.
// Half order: CTAs {0, 1} own the columns of half 0, CTAs {2, 3} those of half 1.  A CTA computes the OTHER pair's half
// first: those two partials travel (and the two it receives from the other pair are gathered) while its own half is
// computed; only the exchange with its pair mate follows the last MMA.
constexpr int BIG_STAGE = 5 * TILE, BIG_STAGES = 2, SMALL_STAGE = 2 * TILE, SMALL_STAGES = 5;
constexpr int R_BYTES = BIG_STAGE * BIG_STAGES;
static_assert(SMALL_STAGE * SMALL_STAGES == R_BYTES, "both stage layouts tile region R");
constexpr int Y_BYTES = 4 * TILE;           // A slice of layers 2..4: [2 k-blocks][hi, lo] tiles (layer 1, mode 1: hi, lo)
constexpr int OFF_R = 0, OFF_Y = R_BYTES, OFF_BAR = OFF_Y + Y_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 1024;
constexpr int SLOT_BYTES = BLOCK_M * 256;   // one exchange slot: 128 rows x 64 fp32
constexpr int SCRATCH_PER_CLUSTER = CL * (CL - 1) * SLOT_BYTES;

struct Params {
  int M, n_tiles;
  int ks[NL];                // K slice per CTA (multiple of 32)
  int nout[NL];              // output width per layer: 256 / 256 / 256 / 64
  const float* bias[NL];
  int relu[NL];
  int a1_mode;               // 0: TMA rows, exact operand; 1: feature-major slot, split in the kernel;
                             // 2: uint8 frames (K order = byte order), converted by the epilogue threads per k-block
  const float* x; long long ldx; int n_obs;      // mode 2: x = uint8 frames, ldx = bytes per frame, n_obs = bytes used
  const float* head_w; const float* head_b; int A;
  int32_t* act; float* logp; float* value; float* head_out; int ld_head;
  uint64_t seed; const uint64_t* tick; uint64_t t_offset; int greedy; int env_offset;
  uint8_t* scratch;          // exchange slots in global memory (L2-resident): [cluster][dst rank][3 slots][32 KB]
  long long* dbg;            // optional timeline probe: clock64 of cluster 0, [rank][layer][8 milestones] (+ [rank][32])
};

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// relaxed form: ordering comes from ONE fence.acq_rel.cluster issued by the same thread in front of a group of arrives
// (measured: every release-arrive costs a cluster-scope membar of ~1000 cycles; three per thread per layer were half of
// the exchange time)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void st_global_v4(void* ptr, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_global_cg_v4(const void* ptr) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr) : "memory");
  return r;
}
// byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a K-major, 128-byte-swizzled [rows][32] fp32 tile
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// byte offset of the 16-byte chunk `c` (0..15) of row `r` inside an exchange slot.  CHUNK-major ([16 chunks][128 rows]
// [16 B]): the thread-per-row stores / loads of one warp instruction touch 512 contiguous bytes
__device__ __forceinline__ uint32_t slot_off(int r, int c) { return (uint32_t)(c * 2048 + r * 16); }

__global__ void __launch_bounds__(THREADS, 1)
fused_policy_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW0h,
                    const __grid_constant__ CUtensorMap tmW0l, const __grid_constant__ CUtensorMap tmW1h,
                    const __grid_constant__ CUtensorMap tmW1l, const __grid_constant__ CUtensorMap tmW2h,
                    const __grid_constant__ CUtensorMap tmW2l, const __grid_constant__ CUtensorMap tmW3h,
                    const __grid_constant__ CUtensorMap tmW3l, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x / CL, n_clusters = (int)gridDim.x / CL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* big_full = bars;          // [2]
  uint64_t* big_empty = bars + 2;     // [2]
  uint64_t* sm_full = bars + 4;       // [5]
  uint64_t* sm_empty = bars + 9;      // [5]
  uint64_t* a1_full = bars + 14;      // mode 1: layer-1 operand tiles written by the epilogue threads
  uint64_t* a_ready = bars + 15;
  uint64_t* acc_full = bars + 16;     // [2]: the first / second half (in this CTA's order) of the accumulator is complete
  uint64_t* peer_free = bars + 18;    // [2], alternating: a fast peer's arrival for phase k+1 never lands in phase k
  uint64_t* inbox_early = bars + 20;  // the other pair's partials of my columns have landed
  uint64_t* inbox_late = bars + 21;   // my pair mate's
  uint64_t* inbox_all = bars + 22;    // last layer, CTA 0: all three peers'
  uint64_t* tile_done = bars + 23;
  uint64_t* a_tile_full = bars + 24;  // [2], mode 2: the epilogue threads have written big stage s's A tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  const CUtensorMap* tmw[NL][2] = {{&tmW0h, &tmW0l}, {&tmW1h, &tmW1l}, {&tmW2h, &tmW2l}, {&tmW3h, &tmW3l}};

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW0h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW0l) : "memory");
    for (int s = 0; s < BIG_STAGES; ++s) {
      mbar_init(big_full + s, 1);
      mbar_init(big_empty + s, 1);
    }
    for (int s = 0; s < SMALL_STAGES; ++s) {
      mbar_init(sm_full + s, 1);
      mbar_init(sm_empty + s, 1);
    }
    mbar_init(a1_full, EPI_THREADS);
    mbar_init(a_ready, EPI_THREADS);
    mbar_init(acc_full, 1);
    mbar_init(acc_full + 1, 1);
    mbar_init(peer_free, 3 * EPI_WARPS);        // 3 peers x their epilogue warps
    mbar_init(peer_free + 1, 3 * EPI_WARPS);
    mbar_init(inbox_early, 2 * EPI_WARPS);      // (lane 0 of every peer epilogue warp arrives after the warp's stores)
    mbar_init(inbox_late, EPI_WARPS);
    mbar_init(inbox_all, 3 * EPI_WARPS);
    mbar_init(tile_done, EPI_THREADS);
    mbar_init(a_tile_full, EPI_THREADS);
    mbar_init(a_tile_full + 1, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // every peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb0 = p.ks[0] / BLOCK_K;

  if (warp == 0) {
    // ===== TMA producer: per tile the big stages of layer 1, then the small weight chunks of layers 2..4 (MMA order) =====
    if (lane == 0) {
      int big_it = 0, sm_it = 0, titer = 0;
      for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++titer) {
        const int row0 = tile * BLOCK_M;
        if (titer > 0) mbar_wait(tile_done, (uint32_t)((titer - 1) & 1));     // every small stage has been consumed
        if (titer == 0 && p.a1_mode == 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // (A rows come from the env kernel)
        for (int kb = 0; kb < nkb0; ++kb, ++big_it) {
          const int s = big_it % BIG_STAGES;
          mbar_wait(big_empty + s, (uint32_t)(((big_it / BIG_STAGES) & 1) ^ 1));
          mbar_expect_tx(big_full + s, (uint32_t)((p.a1_mode == 0 ? 5 : 4) * TILE));
          uint8_t* dst = smem + OFF_R + s * BIG_STAGE;
          const int kc = (int)rank * p.ks[0] + kb * BLOCK_K;
          if (p.a1_mode == 0) tma_load_2d(&tmA1, big_full + s, dst, kc, row0);
          tma_load_2d(tmw[0][0], big_full + s, dst + TILE, kc, 0);
          tma_load_2d(tmw[0][1], big_full + s, dst + 3 * TILE, kc, 0);
        }
        // the small-stage layout aliases both big stages: wait until their last uses have been consumed
        for (int u = big_it - 1; u >= 0 && u >= big_it - BIG_STAGES; --u)
          mbar_wait(big_empty + (u % BIG_STAGES), (uint32_t)((u / BIG_STAGES) & 1));
        for (int l = 1; l < NL; ++l) {
          const int nkb = p.ks[l] / BLOCK_K, halves = p.nout[l] > 128 ? 2 : 1, rows = p.nout[l] > 128 ? 128 : p.nout[l];
          for (int i = 0; i < halves; ++i) {
            const int hh = halves == 2 ? (i ^ 1 ^ (int)(rank >> 1)) : 0;       // the other pair's half first
            for (int kb = 0; kb < nkb; ++kb, ++sm_it) {
              const int s = sm_it % SMALL_STAGES;
              mbar_wait(sm_empty + s, (uint32_t)(((sm_it / SMALL_STAGES) & 1) ^ 1));
              mbar_expect_tx(sm_full + s, (uint32_t)(2 * rows * 128));
              uint8_t* dst = smem + OFF_R + s * SMALL_STAGE;
              const int kc = (int)rank * p.ks[l] + kb * BLOCK_K;
              tma_load_2d(tmw[l][0], sm_full + s, dst, kc, hh * 128);
              tma_load_2d(tmw[l][1], sm_full + s, dst + TILE, kc, hh * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int big_it = 0, sm_it = 0, titer = 0;
    uint32_t ar_ph = 0;
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++titer) {
      {   // layer 1: N = 256 in one instruction shape; the A tile travels with its weight k-block (mode 0) or sits in Y
        if (p.a1_mode == 1) {
          mbar_wait(a1_full, (uint32_t)(titer & 1));
          tc_fence_after();
        }
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
        for (int kb = 0; kb < nkb0; ++kb, ++big_it) {
          const int s = big_it % BIG_STAGES;
          mbar_wait(big_full + s, (uint32_t)((big_it / BIG_STAGES) & 1));
          if (p.a1_mode == 2) mbar_wait(a_tile_full + s, (uint32_t)((big_it / BIG_STAGES) & 1));
          tc_fence_after();
          if (lane == 0) {
            const uint32_t st = smem_u32(smem + OFF_R + s * BIG_STAGE);
            const uint32_t a_hi = p.a1_mode != 1 ? st : smem_u32(smem + OFF_Y), a_lo = a_hi + TILE;
            const uint32_t b_hi = st + TILE, b_lo = st + 3 * TILE;
            bool first = kb == 0;
            for (int pass = (p.a1_mode == 1 ? 2 : 1); pass >= 0; --pass) {
              const uint32_t a = pass == 2 ? a_lo : a_hi, b = pass == 1 ? b_lo : b_hi;
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                umma_tf32(tmem_base, make_desc(a + k * UMMA_K * 4, BLOCK_K), make_desc(b + k * UMMA_K * 4, BLOCK_K), idesc,
                          first ? 0u : 1u);
                first = false;
              }
            }
            umma_commit(big_empty + s);
            if (kb == nkb0 - 1) umma_commit(acc_full);
          }
          __syncwarp();
        }
      }
      for (int l = 1; l < NL; ++l) {
        mbar_wait(a_ready, ar_ph);
        ar_ph ^= 1u;
        tc_fence_after();
        const int nkb = p.ks[l] / BLOCK_K, halves = p.nout[l] > 128 ? 2 : 1, rows = p.nout[l] > 128 ? 128 : p.nout[l];
        const uint32_t a_base = smem_u32(smem + OFF_Y);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(rows >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
        for (int i = 0; i < halves; ++i) {
          const int hh = halves == 2 ? (i ^ 1 ^ (int)(rank >> 1)) : 0;
          for (int kb = 0; kb < nkb; ++kb, ++sm_it) {
            const int s = sm_it % SMALL_STAGES;
            mbar_wait(sm_full + s, (uint32_t)((sm_it / SMALL_STAGES) & 1));
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_hi = a_base + (uint32_t)(kb * 2 * TILE), a_lo = a_hi + TILE;
              const uint32_t b_hi = smem_u32(smem + OFF_R + s * SMALL_STAGE), b_lo = b_hi + TILE;
              const uint32_t d = tmem_base + (uint32_t)(hh * 128);
              bool first = kb == 0;
              // small terms first: a_lo * b_hi, a_hi * b_lo, a_hi * b_hi
              for (int pass = 2; pass >= 0; --pass) {
                const uint32_t a = pass == 2 ? a_lo : a_hi, b = pass == 1 ? b_lo : b_hi;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  umma_tf32(d, make_desc(a + k * UMMA_K * 4, BLOCK_K), make_desc(b + k * UMMA_K * 4, BLOCK_K), idesc,
                            first ? 0u : 1u);
                  first = false;
                }
              }
              umma_commit(sm_empty + s);
              if (kb == nkb - 1) umma_commit(acc_full + i);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===== epilogue / exchange: two threads per accumulator row (TMEM lane), 32 columns of every 64-column slice each =====
    const int q = warp & 3, ch = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* my_slots = p.scratch + ((size_t)cluster_id * CL + rank) * (CL - 1) * SLOT_BYTES;   // where peers write for me
    uint32_t acc_ph[2] = {0u, 0u}, ie_ph = 0, il_ph = 0, ia_ph = 0;
    int pf_sig = 0, pf_wait = 0;             // peer_free signals sent / waits done (barrier = count & 1, parity = count >> 1)
    int titer = 0, big_it_e = 0;
    auto signal_peer_free = [&]() {          // this warp's part of MY exchange slots may be overwritten
      __syncwarp();                          // (its loads have returned: their values were consumed before this point)
      if (lane == 0) {
#pragma unroll
        for (int dp = 1; dp < CL; ++dp)
          mbar_arrive_cluster_relaxed(mapa(smem_u32(peer_free + (pf_sig & 1)), (rank + dp) & 3));
      }
      ++pf_sig;
    };
    auto wait_peer_free = [&]() {
      mbar_wait_cluster(peer_free + (pf_wait & 1), (uint32_t)((pf_wait >> 1) & 1));
      ++pf_wait;
    };
    auto wait_acc = [&](int h) {
      mbar_wait(acc_full + h, acc_ph[h]);
      acc_ph[h] ^= 1u;
      tc_fence_after();
    };
    auto tmem_ld_cw = [&](uint32_t taddr, float* v) {
      if (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
    };
    // my CW columns of the 64-column slice starting at accumulator column col0 -> slot `slot` of CTA `dst_rank`
    // (st.global, L2-resident; a warp instruction writes 512 contiguous bytes), then the release-arrive on its barrier
    auto send_slice = [&](uint32_t col0, uint32_t dst_rank, int slot) {
      float v[CW];
      tmem_ld_cw(t_row + col0 + (uint32_t)(CW * ch), v);
      uint8_t* dst = p.scratch + (((size_t)cluster_id * CL + dst_rank) * (CL - 1) + slot) * SLOT_BYTES;
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        st_global_v4(dst + slot_off(row, NCH * ch + j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    };
    // publish this warp's stores to the CTAs in `mask`: warp barrier, ONE cluster-scope fence by lane 0, relaxed arrives
    auto publish = [&](uint32_t mask, uint64_t* bar) {
      __syncwarp();
      if (lane == 0) {
        fence_cluster();
#pragma unroll
        for (uint32_t r = 0; r < CL; ++r)
          if ((mask >> r) & 1u) mbar_arrive_cluster_relaxed(mapa(smem_u32(bar), r));
      }
    };
    // acc[32] += the partials of my 32 columns in slots [s0, s0 + n) (16-byte L2 loads, all in flight together)
    auto gather_slots = [&](float* acc, int s0, int n) {
      float4 t[3][NCH];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < n)
#pragma unroll
          for (int j = 0; j < NCH; ++j)
            t[k][j] = ld_global_cg_v4(my_slots + ((s0 + k) % 3) * SLOT_BYTES + slot_off(row, NCH * ch + j));
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < n)
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            acc[4 * j] += t[k][j].x; acc[4 * j + 1] += t[k][j].y;
            acc[4 * j + 2] += t[k][j].z; acc[4 * j + 3] += t[k][j].w;
          }
    };
    // slot of source s at destination d: ((s - d) mod 4) - 1.  My pair mate's slot, and the first of the other two
    const uint32_t mate = rank ^ 1u;
    const int late_slot = (int)((mate - rank + CL) & 3) - 1, early_slot0 = (late_slot + 1) % 3;
    const bool probe = p.dbg && cluster_id == 0 && warp == 2 && lane == 0;
#define FPROBE(l, i) do { if (probe) p.dbg[rank * 64 + (l) * 8 + (i)] = clock64(); } while (0)
    if (probe) p.dbg[rank * 64 + 32] = clock64();
    // programmatic dependent launch: everything above (barriers, TMEM, the weight stream of layer 1) may run while the
    // previous kernel of the stream (the env's step) is still finishing; its outputs are first touched below
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++titer) {
      const int grow = tile * BLOCK_M + row;
      if (p.a1_mode == 1) {
        // layer-1 operand from a feature-major slot x[k][ldx]: this CTA's 32 k-columns, split into hi / lo tiles (in Y)
        uint8_t* hi_t = smem + OFF_Y;
        uint8_t* lo_t = hi_t + TILE;
#pragma unroll
        for (int cc = 0; cc < 8 / TPR; ++cc) {
          const int c = (8 / TPR) * ch + cc;
          float v[4], h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = (int)rank * BLOCK_K + c * 4 + j;
            v[j] = (k < p.n_obs && grow < p.M) ? __ldg(p.x + (long long)k * p.ldx + grow) : 0.0f;
            h[j] = tf32_round(v[j]);
          }
          const uint32_t o = tile_off(row, c);
          *reinterpret_cast<float4*>(hi_t + o) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(lo_t + o) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(a1_full);
      }
      if (p.a1_mode == 2) {
        // layer-1 operand straight from the uint8 frame: this CTA's k-slice is a contiguous byte range of the frame (the
        // first layer's weight columns are stored in frame byte order), 32 / TPR bytes per thread and k-block; converted to the
        // exact fp32 pixel values and written into the A slot of the big stage its weight k-block streams into
        const uint8_t* fr = reinterpret_cast<const uint8_t*>(p.x) + (long long)grow * p.ldx;
        constexpr int WPT = 8 / TPR;            // 32-bit words (= 16-byte chunks of the tile) per thread and k-block
        uint32_t px[6][WPT];
#pragma unroll
        for (int kb = 0; kb < 6; ++kb)
#pragma unroll
          for (int w = 0; w < WPT; ++w) {
            const int b0 = (int)rank * p.ks[0] + kb * BLOCK_K + 4 * WPT * ch + 4 * w;
            px[kb][w] = (kb < nkb0 && grow < p.M && b0 < p.n_obs) ? __ldg(reinterpret_cast<const uint32_t*>(fr + b0)) : 0u;
          }
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) {
          if (kb < nkb0) {
            const int it = big_it_e + kb, s = it % BIG_STAGES;
            mbar_wait(big_empty + s, (uint32_t)(((it / BIG_STAGES) & 1) ^ 1));
            uint8_t* a_t = smem + OFF_R + s * BIG_STAGE;
#pragma unroll
            for (int w = 0; w < WPT; ++w) {
              const uint32_t v = px[kb][w];
              *reinterpret_cast<float4*>(a_t + tile_off(row, WPT * ch + w)) =
                  make_float4((float)(v & 255u), (float)((v >> 8) & 255u), (float)((v >> 16) & 255u), (float)(v >> 24));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(a_tile_full + s);
          }
        }
        big_it_e += nkb0;
      }
      signal_peer_free();      // (layer 1 of this tile) my slots were consumed in the previous tile
      for (int l = 0; l < NL; ++l) {
        FPROBE(l, 0);
        if (l < NL - 1) {
          // ---- reduce-scatter of the 128 x 256 partial: 64 columns per owner (owners 0, 1 in column half 0) ----
          float own[CW], bv[CW];
          const float* bias = p.bias[l] + rank * 64 + CW * ch;
#pragma unroll
          for (int j = 0; j < CW; ++j) bv[j] = __ldg(bias + j);
          wait_peer_free();                    // the peers' slots are writable
          FPROBE(l, 2);
          wait_acc(0);                         // the other pair's half (layer 1: all 256 columns, one commit)
          {
            const uint32_t other = ((rank >> 1) ^ 1u) * 2u;          // owners other, other + 1
            send_slice(other * 64, other, (int)((rank - other + CL) & 3) - 1);
            send_slice((other + 1) * 64, other + 1, (int)((rank - other - 1 + CL) & 3) - 1);
            publish(3u << other, inbox_early);
          }
#pragma unroll
          for (int j = 0; j < CW; ++j) own[j] = 0.0f;
          mbar_wait_cluster(inbox_early, ie_ph);     // the other pair computed MY half first: gather it under my MMAs
          ie_ph ^= 1u;
          gather_slots(own, early_slot0, 2);
          if (l > 0) wait_acc(1);
          FPROBE(l, 1);
          send_slice(mate * 64, mate, (int)((rank - mate + CL) & 3) - 1);
          publish(1u << mate, inbox_late);
          {
            float mine[CW];
            tmem_ld_cw(t_row + rank * 64 + (uint32_t)(CW * ch), mine);
#pragma unroll
            for (int j = 0; j < CW; ++j) own[j] += mine[j];
          }
          tc_fence_before();
          FPROBE(l, 3);
          mbar_wait_cluster(inbox_late, il_ph);
          il_ph ^= 1u;
          FPROBE(l, 4);
          gather_slots(own, late_slot, 1);
          const float floor_v = p.relu[l] ? 0.0f : -3.402823466e38f;
          // my CW columns of my A slice of the next layer (k-block = column / 32): hi / lo swizzled tiles
          uint8_t* hi_t = smem + OFF_Y + ((CW * ch) >> 5) * 2 * TILE;
          const int c_in = ((CW * ch) & 31) >> 2;          // first 16-byte chunk inside the k-block's row
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            float x[4], h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[i] = fmaxf(own[4 * j + i] + bv[4 * j + i], floor_v);
              h[i] = tf32_round(x[i]);
            }
            const uint32_t o = tile_off(row, c_in + j);
            *reinterpret_cast<float4*>(hi_t + o) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(hi_t + TILE + o) = make_float4(x[0] - h[0], x[1] - h[1], x[2] - h[2], x[3] - h[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          signal_peer_free();                  // my slots have been consumed: writable for the next layer
          mbar_arrive(a_ready);
          FPROBE(l, 5);
        } else {
          // ---- last embedder layer (N = 64): every partial goes to CTA 0, which finishes the step ----
          const int nh = p.A + 1;
          float wreg[(MAX_A + 1) * 65 / EPI_THREADS + 1];       // CTA 0: head weights + biases, in flight during the waits
          float bz[CW];
          uint64_t tick_v = 0;
          if (rank == 0) {
            tick_v = p.tick ? *p.tick : 0ull;          // (the draw's inputs: loaded now, used ~10 us of latency later)
#pragma unroll
            for (int i = 0; i < (MAX_A + 1) * 65 / EPI_THREADS + 1; ++i) {
              const int idx = i * EPI_THREADS + (warp - 2) * 32 + lane;
              wreg[i] = idx < nh * 64 ? __ldg(p.head_w + idx) : (idx < nh * 65 ? __ldg(p.head_b + idx - nh * 64) : 0.0f);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) bz[j] = __ldg(p.bias[l] + CW * ch + j);
          }
          wait_peer_free();
          wait_acc(0);
          FPROBE(l, 1);
          if (rank != 0) {
            send_slice(0, 0, (int)rank - 1);
            publish(1u, inbox_all);
            tc_fence_before();
          } else {
            float z[CW];
            tmem_ld_cw(t_row + (uint32_t)(CW * ch), z);
            tc_fence_before();
            // head weights -> shared memory (region Y is free: this layer's MMAs have retired)
            float* sWh = reinterpret_cast<float*>(smem + OFF_Y);
            float* sHp = sWh + (MAX_A + 1) * 65;                    // [TPR][128][MAX_A + 1] partial head sums
#pragma unroll
            for (int i = 0; i < (MAX_A + 1) * 65 / EPI_THREADS + 1; ++i) {
              const int idx = i * EPI_THREADS + (warp - 2) * 32 + lane;
              if (idx < nh * 65) sWh[idx] = wreg[i];                  // [nh][64] weights, then the nh biases
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            mbar_wait_cluster(inbox_all, ia_ph);
            ia_ph ^= 1u;
            FPROBE(l, 4);
            gather_slots(z, 0, 3);
            const float floor_v = p.relu[l] ? 0.0f : -3.402823466e38f;
#pragma unroll
            for (int j = 0; j < CW; ++j) z[j] = fmaxf(z[j] + bz[j], floor_v);
            FPROBE(l, 6);
            for (int j = 0; j < nh; ++j) {
              float a = 0.0f;
              const float* w = sWh + j * 64 + CW * ch;
#pragma unroll
              for (int k = 0; k < CW; ++k) a = fmaf(z[k], w[k], a);
              sHp[(ch * BLOCK_M + row) * (MAX_A + 1) + j] = a;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            FPROBE(l, 7);
            if (ch == 0 && grow < p.M) {
              float hd[MAX_A + 1];
#pragma unroll
              for (int j = 0; j < MAX_A + 1; ++j)
                hd[j] = 0.0f;
              for (int j = 0; j < nh; ++j) {
                float a = 0.0f;
#pragma unroll
                for (int c = 0; c < TPR; ++c) a += sHp[(c * BLOCK_M + row) * (MAX_A + 1) + j];
                hd[j] = sWh[nh * 64 + j] + a;
              }
              if (p.head_out)
                for (int j = 0; j < nh; ++j) p.head_out[(long long)grow * p.ld_head + j] = hd[j];
              sample_row(hd, p.A, p.env_offset + grow, p.seed, &tick_v, p.t_offset, p.greedy, p.act + grow, p.logp + grow,
                         p.value + grow);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // nobody reads Y when the next tile rewrites it
          }
          mbar_arrive(tile_done);
          FPROBE(l, 5);
        }
      }
    }
#undef FPROBE
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace fused
}  // namespace tpp

extern "C" int tpp_policy_rollout_fused(const tpp_fused_policy* f, void* stream) {
  using namespace tpp::fused;
  TPP_CHECK_ARG(f && f->x && f->n_rows > 0 && f->act && f->logp && f->value && f->head_w && f->head_b);
  TPP_CHECK_ARG(f->n_actions > 0 && f->n_actions + 1 <= tpp::MAX_A + 1 && f->n_actions <= tpp::MAX_A);
  TPP_CHECK_ARG(f->a1_mode >= 0 && f->a1_mode <= 2);
  for (int l = 0; l < NL; ++l) TPP_CHECK_ARG(f->w_hi[l] && f->w_lo[l] && f->bias[l] && f->k[l] > 0 && f->ldw[l] >= f->k[l]);
  // shapes this kernel is built for: depth-4 MLPModel, hidden width 256, latent 64 (the reference's mlpmodel sets)
  if (f->n[0] != 256 || f->n[1] != 256 || f->n[2] != 256 || f->n[3] != 64) return TPP_ENOTSUP;
  if (f->k[1] != 256 || f->k[2] != 256 || f->k[3] != 256) return TPP_ENOTSUP;
  Params p;
  p.M = f->n_rows;
  p.n_tiles = (f->n_rows + tpp::tc::BLOCK_M - 1) / tpp::tc::BLOCK_M;
  for (int l = 0; l < NL; ++l) {
    p.ks[l] = ((f->k[l] + CL - 1) / CL + 31) / 32 * 32;
    p.nout[l] = f->n[l];
    p.bias[l] = f->bias[l];
    p.relu[l] = f->relu[l];
  }
  if (f->a1_mode == 1 && p.ks[0] != 32) return TPP_ENOTSUP;      // feature-major slots: n_obs <= 128
  if (f->a1_mode == 2 && (p.ks[0] > 6 * 32 || (f->k[0] & 3) || (f->ldx & 3) || (reinterpret_cast<uintptr_t>(f->x) & 3)))
    return TPP_ENOTSUP;                                            // uint8 frames: <= 768 bytes, word-aligned
  TPP_CHECK_ARG(f->scratch);
  p.a1_mode = f->a1_mode; p.x = f->x; p.ldx = f->ldx; p.n_obs = f->k[0];
  p.head_w = f->head_w; p.head_b = f->head_b; p.A = f->n_actions;
  p.act = f->act; p.logp = f->logp; p.value = f->value; p.head_out = f->head_out; p.ld_head = f->ld_head;
  p.dbg = reinterpret_cast<long long*>(f->dbg);
  p.seed = f->seed; p.tick = f->tick; p.t_offset = f->t_offset; p.greedy = f->greedy; p.env_offset = f->env_offset;
  CUtensorMap tmA1, tmW[NL][2];
  int rc, bytes;
  for (int l = 0; l < NL; ++l) {
    const int rows = l == 0 ? 256 : (f->n[l] > 128 ? 128 : f->n[l]);      // layer 1: all 256 output rows per box
    if ((rc = tpp::tc::make_map(&tmW[l][0], f->w_hi[l], f->ldw[l], f->n[l], (int)f->ldw[l], rows, 0, &bytes))) return rc;
    if ((rc = tpp::tc::make_map(&tmW[l][1], f->w_lo[l], f->ldw[l], f->n[l], (int)f->ldw[l], rows, 0, &bytes))) return rc;
  }
  if (f->a1_mode == 0) {
    if ((rc = tpp::tc::make_map(&tmA1, f->x, f->ldx, f->n_rows, (int)f->ldx, tpp::tc::BLOCK_M, 0, &bytes))) return rc;
  } else {
    tmA1 = tmW[0][0];
  }
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fused_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(CL, 1, 1);
    q.blockDim = dim3(THREADS, 1, 1);
    q.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, fused_policy_kernel, &q) != cudaSuccess || max_clusters < 1) {
      cudaGetLastError();
      max_clusters = 32;
    }
    attr_set = true;
  }
  int clusters = p.n_tiles < max_clusters ? p.n_tiles : max_clusters;
  const long long fit = f->scratch_bytes / SCRATCH_PER_CLUSTER;
  if (fit < 1) return TPP_EINVAL;
  if (clusters > fit) clusters = (int)fit;
  p.scratch = reinterpret_cast<uint8_t*>(f->scratch);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CL * clusters), 1, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = tpp_stream(stream);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // prologue + weight prefetch overlap the env kernel
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = f->no_pdl ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fused_policy_kernel, tmA1, tmW[0][0], tmW[0][1], tmW[1][0], tmW[1][1],
                                     tmW[2][0], tmW[2][1], tmW[3][0], tmW[3][1], p);
  if (e != cudaSuccess) return (int)e;
  TPP_LAUNCH_STATUS();
}
