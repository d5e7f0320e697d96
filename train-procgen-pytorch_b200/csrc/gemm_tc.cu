// Tensor-core path of the policy's dense layers AND 3x3 convolutions: tcgen05.mma (kind::tf32) with TMEM
// accumulators, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through an mbarrier pipeline,
// warp-specialised (1 TMA warp, 1 MMA warp, 4 or 16 epilogue warps).  sm_100a only.
//
// Convolutions (IMPALA-CNN, reference common/model.py:134-208) are implicit GEMMs on the narrow tile instances
// (BLOCK_N = 16 / 32 = output channels): the A tiles are gathered straight from the NHWC activation pair by TMA
// *im2col* loads (cp.async.bulk.tensor.4d...im2col: 128 consecutive output pixels x one filter tap per k-block, the
// TMA unit zero-fills padding pixels, unused channel slots and the tail behind the last image), so no col matrix
// exists.  Forward and data gradient are K-major (k-block = tap; 16-channel tensors use 64-byte rows / 64-byte
// swizzle, 32-channel ones 128-byte rows); the weight gradient gathers the same tensor as an MN-major operand
// (k-block = 32 pixels, one im2col box per tap).  The narrow instances are PERSISTENT (two TMEM accumulators,
// producer / MMA warps run ahead into the next tile while the epilogue drains the previous one) and carry the
// fused residual add, ReLU-pair output and running bias-gradient column sums.  Measured on B200 (block-1 shape,
// 2.1 M pixels, 16 -> 16 channels, 3xTF32): forward 322 us, data gradient 338 us, weight gradient 820 us; the bound is
// the L2 -> shared-memory gather of 9 taps x (hi, lo) (7.5 TB/s achieved), see profiles/README.md.
//
// Precision.  The parity bar of this path is fp32 (<= 1e-5 relative against the reference's torch-fp32 result),
// which a single TF32 product (10-bit mantissa) cannot meet.  Every operand therefore exists as a pair
//   x = hi + lo,  hi = tf32_round(x),  lo = x - hi   (both stored as fp32 words),
// and `precision == 3` accumulates  hi*hi + hi*lo + lo*hi  in the same fp32 TMEM accumulator ("3xTF32",
// relative error ~2^-21).  `precision == 1` issues only hi*hi (fast mode, ~1e-3; documented tolerance).
// The pairs are produced where the tensors are produced: by this kernel's own epilogue, by the minibatch gather,
// and by tpp_split_tf32 for tensors that come from elsewhere (weights after an optimizer step, loss gradients).
//
// C[m][n] = sum_k A(m,k) * B(n,k).  Each operand is either K-major (rows of the matrix are m / n, the contraction
// index is contiguous) or MN-major (rows are the contraction index, m / n is contiguous); tcgen05 reads both
// straight from 128B-swizzled shared memory (instruction-descriptor a_major / b_major bits), so no tensor is ever
// transposed in memory:
//   forward        : A = activations [mb][in]   K-major,   B = W  [out][in]   K-major
//   rollout, vector envs: A = feature-major obs slot [in][N]  MN-major
//   data gradient  : A = dZ [mb][out]           K-major,   B = W  [out][in]   MN-major  (contraction over out)
//   weight gradient: A = dZ [mb][out]           MN-major,  B = X  [mb][in]    MN-major  (contraction over mb),
//                    split over the mb contraction across CTAs, fp32 atomics into the flat gradient.
//
// Replaces: nn.Linear forward/backward of MLPModel and the policy heads on cuBLAS (reference
// common/model.py:954-980, common/policy.py:74-87, autograd of agents/ppo.py:170).
#include <cuda.h>

#include <algorithm>

#include "tc_ptx.cuh"

namespace tpp {
namespace tc {

constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
// Epilogue warps: 4 per TMEM lane quarter for wide tiles (the epilogue is instruction-latency bound, not bandwidth
// bound); narrow tiles (one 32-column group: convolution outputs) have work for 4 warps only, and the small CTA
// (192 threads x 80 registers) lets 3-4 tiles share an SM so that prologue / epilogue overlap other tiles' loads.
// (SPLIT kernels of the 256-wide tiles: 8 epilogue warps, i.e. 36 KB of transpose patches instead of 72 -- the room
// goes to the stage rings)
// LEAN kernels (256 x 256 persistent pairs): also 8 epilogue warps, with 4 KB XOR-swizzled patches, so that THREE pair
// stages fit the 227 KB -- for launches whose epilogue is light enough to hide behind the main loop with 8 warps
constexpr int epi_warps(int block_n, bool split = false, bool lean = false) {
  return block_n <= 32 ? 4 : (((split && block_n == 256) || lean) ? 8 : 16);
}
// wide tiles carry 4 more warps behind the epilogue warps: the on-chip operand splitters (see Params::split_a)
// two groups, group g splits the k-blocks it = g (mod 2); 256-wide tiles (8 epilogue warps) have room for 2 x 8 warps
constexpr int conv_warps(int block_n) { return block_n == 256 ? 16 : 8; }
constexpr int num_threads(int block_n, bool split = false, bool lean = false) {   // warp 0: TMA, warp 1: MMA + TMEM, epilogue warps, [splitters]
  return 64 + epi_warps(block_n, split, lean) * 32 + (split ? conv_warps(block_n) * 32 : 0);
}
constexpr int STG_PITCH = 36;            // floats per staged row (16-byte aligned, conflict-free 128-bit LDS/STS)
constexpr int stg_pitch(bool lean) { return lean ? 32 : STG_PITCH; }
constexpr int stg_bytes(int block_n, bool split = false, bool lean = false) {     // a 32x32 patch per warp
  return epi_warps(block_n, split, lean) * 32 * stg_pitch(lean) * 4;
}
// float index of (row, col) in a patch; LEAN: pitch 32 with the 16-byte chunk XOR-swizzled by the row (conflict-free
// 128-bit accesses from both sides of the transpose: a quarter-warp touches 8 distinct chunks)
template <bool LEAN>
__device__ __forceinline__ int stg_at(int row, int col) {
  return LEAN ? row * 32 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3)) : row * STG_PITCH + col;
}

struct Params {
  int M, N, K;                 // logical problem (rows of A, rows of B, contraction)
  int npass;                   // 1 (tf32) or 3 (3xTF32)
  int nops_a, nops_b;          // operand copies staged per k-block: 2 = (hi, lo), 1 = hi only (single pass, or the
                               // operand is exact in TF32 -- e.g. integer pixel values -- and has no lo half)
  int split_a, split_b;        // wide tiles: the operand arrives as ONE plain fp32 copy; the tensor core reads it as its hi
                               // half (it truncates fp32 words to TF32) and the splitter warps form lo = x - trunc(x) in
                               // shared memory between the TMA landing and the MMA -- half the HBM / L2 bytes of a pair
  int tma_a, tma_b;            // operand copies the TMA producer loads (1 when split on chip or exact)
  int lo_stages, lo_offset;    // split on chip: the lo halves live in their own ring of lo_stages slots at smem +
                               // lo_offset ([A lo][B lo], split operands only), so that the TMA ring holds plain tiles
                               // only and is deeper for the same shared memory
  int skip;                    // bit `pass` set: that pass is not issued (2: a_lo*b_hi, 1: a_hi*b_lo)
  float alpha;                 // scale of the atomically accumulated product (TPP_EPI_ACCUM)
  int stages;
  int kb_per_split;            // k-blocks handled by one blockIdx.z
  int bar_offset;              // byte offset of the mbarriers behind the stage / staging region
  int flags;
  const float* bias;
  const float* mask; long long ld_mask;
  uint32_t* bits_out;          // wide tiles: (result > 0) of every 32 x 32 output block as 32 words (128 B instead of 4 KB)
  const uint32_t* bits_in;     // ... and the same words read back as the ReLU mask of a data gradient
  int bits_ncb;                // 32-column blocks per row of blocks (ceil(N / 32))
  float* out; long long ldc;   // plain fp32 (or atomic accumulation target)
  float* out_hi; float* out_lo;
  float* colsum;               // optional: colsum[n] += sum over this tile's rows of the (masked) result (bias grad)
  int a_mn, b_mn;              // operand majors (0 = K-major, 1 = MN-major)
  long long* dbg;              // optional timeline probe (one CTA): clock64 at 8 milestones
  int dbg_y;                   // blockIdx.y of the probed CTA (wide tiles)
  int ntn;                     // number of N tiles
  int ntiles;                  // ntn * number of M tiles
  int total_work;              // ntiles * number of k-splits: work items (tile, split), split-major
  int stg_offset;              // byte offset of the epilogue's transpose patches (0: they alias the pipeline stages)
  int tx_bytes;                // bytes one stage receives (narrow MN-major operands use smaller TMA boxes)
  int a_slot, b_slot;          // bytes of one A / B operand copy inside a stage
  int b_tx;                    // bytes one B box delivers
  int conv_W, conv_HW;         // implicit 3x3 convolution: A tiles come from TMA im2col loads of an NHWC tensor (0 = off)
  int bk;                      // fp32 words per k-block row: 32 (128-byte swizzle) or 16 (64-byte swizzle: 16-channel
                               // convolutions, whose taps would otherwise be half zero fill)
  int conv_wgrad;              // 1: weight-gradient form, A = im2col(X) MN-major (rows = (tap, channel slot), k = pixels)
  const float* addend; long long ld_add;   // optional residual: result += addend[m][n] (after bias / relu / mask)
};

enum { F_BIAS = 1, F_RELU = 2, F_MASK = 4, F_ATOMIC = 8, F_ADD = 16, F_RELU_OUT = 32, F_PAIR_RELU = 64 };
// (1-bit ReLU masks: Params::bits_out / bits_in, wide tiles, see tpp_tc_gemm.mask_bits)

// ---------------------------------------------------------------------------------------------------
// The kernel: one 128 x BLOCK_N output tile (x one k-split) per CTA
// ---------------------------------------------------------------------------------------------------
// PAIR (BLOCK_N = 256 only): the CTA is one half of a cluster of two (cta_group::2).  The pair computes a 256 x 256
// output tile: each CTA stages ITS 128 rows of A and ITS 128 of the 256 B rows (half the L2 -> shared-memory bytes per
// FLOP of a single-CTA 128 x 256 tile, a quarter of two 128 x 128 tiles -- the bound of these kernels), the leader
// (rank 0) issues the 256 x 256 x 8 MMAs for both, each CTA's TMEM receives its own 128 accumulator rows and each CTA
// runs the epilogue of those rows.  Both CTAs' TMA loads signal the LEADER's full barrier; tcgen05.commit multicasts
// the stage-free / accumulator-complete arrivals to both CTAs.
// PERSIST (with PAIR): the pair loops over 256 x 256 work items like the narrow tiles do -- two TMEM accumulators (all 512
// columns), the producer / MMA warps run into the next item while both CTAs' epilogue warps drain the previous one
// (their transpose patches get their own shared memory; the pipeline keeps two 64 KB stages); the peer's epilogue warps
// return an accumulator by arriving on the LEADER's barrier.
// lo half of an operand whose hi half is the fp32 word itself: the tensor core drops the 13 low mantissa bits of a
// kind::tf32 operand (measured, profiles/tf32_operand_probe.py), so hi = trunc(x) and x - trunc(x) is exact in fp32 (13
// significant bits); rounding it to TF32's 11 keeps the residual error unbiased at 2^-21 |x|.
__device__ __forceinline__ float split_lo(float x) {
  return tf32_round(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

template <int BLOCK_N, bool PAIR = false, bool PERSIST = false, bool SPLIT = false, bool LEAN = false>
__global__ void __launch_bounds__(num_threads(BLOCK_N, SPLIT, LEAN), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, Params p) {
  // Narrow tiles (BLOCK_N <= 32) are the convolution tiles.  Only they carry the residual / ReLU-pair epilogue extras,
  // the 64-byte-row k-blocks and the im2col producer, and they run PERSISTENT: a CTA loops over work items with two
  // TMEM accumulators, so that the producer / MMA warps start the next tile while the epilogue warps drain the
  // previous one (per-tile prologue, epilogue and CTA launch were ~2/3 of a 9-k-block tile's time).  Wide tiles keep
  // compile-time constants, the lean epilogue and one work item per CTA.
  constexpr bool NARROW = BLOCK_N <= 32;
  constexpr int CONV_GROUP = conv_warps(BLOCK_N) / 2;                 // splitter warps per group
  static_assert(!PAIR || BLOCK_N == 256 || BLOCK_N == 64, "CTA-pair forms: 256 x 256 and 256 x 64 tiles");
  constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;                // B rows this CTA stages
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  constexpr uint32_t ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;          // TMEM columns of one accumulator
  static_assert(!PERSIST || PAIR, "persistent wide tiles exist in the CTA-pair form only");
  constexpr bool PERS = NARROW || PERSIST;                            // the CTA loops over work items
  constexpr uint32_t TMEM_COLS = PERS ? 2 * ACC_COLS : ACC_COLS;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), a_major bit 15,
  // b_major bit 16 (1 = MN-major), N>>3 at bit 17, M>>4 at bit 24
  const uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                         ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                         ((uint32_t)((PAIR ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);

  // 128B-swizzled tiles need 1024-byte alignment.  The alignment is requested on the declaration (and checked) instead
  // of rounding the address through an integer: that cast loses the shared address space and turns every staging
  // access into a generic LD/ST (measured: ~430 cycles per dependent access instead of ~30).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  // (exact-operand modes exist on the wide tiles only: the narrow persistent kernels are sensitive to every extra
  // instruction of their single-thread producer / MMA roles -- measured +9 % on the convolution tiles)
  const int nops_a = NARROW ? (p.npass == 3 ? 2 : 1) : p.nops_a, nops_b = NARROW ? nops_a : p.nops_b;
  const int bk = NARROW ? p.bk : BLOCK_K;
  const int A_BYTES = NARROW ? p.a_slot : BLOCK_M * BLOCK_K * 4, B_BYTES = NARROW ? p.b_slot : B_ROWS * BLOCK_K * 4;
  const int ta = NARROW ? nops_a : p.tma_a, tb = NARROW ? nops_b : p.tma_b;   // copies per operand in a TMA stage
  const int stage_bytes = A_BYTES * ta + B_BYTES * tb;
  const int lo_bytes = NARROW ? 0 : A_BYTES * p.split_a + B_BYTES * p.split_b;
  const int LS = NARROW ? 0 : p.lo_stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.bar_offset);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;      // [2]: accumulator i complete (MMA -> epilogue)
  uint64_t* tmem_empty = tmem_full + 2;            // [2]: accumulator i drained (epilogue -> MMA)
  uint64_t* conv_bar = tmem_empty + 2;             // [lo_stages]: the splitter warps have filled lo slot l
  uint64_t* lo_empty = conv_bar + LS;              // [lo_stages]: the MMAs reading lo slot l have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_empty + LS);
  // on-chip split: every CTA's loads complete on its OWN full barrier (its splitters wait there) and the MMA issuer waits
  // for the splitters of the CTA (of both CTAs of a pair) instead of for the loads
  static_assert(!SPLIT || BLOCK_N > 32, "operands are split on chip by the wide tiles only");
  constexpr bool split_mode = SPLIT;     // (a separate kernel: the pair-operand kernels keep their round-1 shape)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items (tile, k-split) are flattened on grid.x (n tile fastest, then m tile, then split), so M is not limited
  // by the 65535 bound of grid.y.  A persistent CTA takes items blockIdx.x, blockIdx.x + gridDim.x, ...
  // Wide tiles: one work item per CTA on a (n tile, m tile, split) grid -- no index arithmetic on the critical path.
  const int wbegin = PERS ? (int)(PAIR ? blockIdx.x >> 1 : blockIdx.x) : 0, wend = PERS ? p.total_work : 1;
  const int wstride = PERS ? (int)(PAIR ? gridDim.x >> 1 : gridDim.x) : 1;
  const bool probe = p.dbg && blockIdx.x == 0 && (int)blockIdx.y == p.dbg_y && blockIdx.z == 0 && lane == 0;
#define TPP_PROBE(i) do { if (probe) p.dbg[i] = clock64(); } while (0)
  if (warp == 0) TPP_PROBE(0);
  const int total_kb = (p.K + bk - 1) / bk;
#define TPP_DECODE_WORK(w)                                                           \
  int m0, n0, kb0;                                                                   \
  if (PERS) {                                                                        \
    const int tile_ = (w) % p.ntiles, split_ = (w) / p.ntiles;                       \
    m0 = PAIR ? ((tile_ / p.ntn) * 2 + (int)cta_rank) * BLOCK_M : (tile_ / p.ntn) * BLOCK_M; \
    n0 = (tile_ % p.ntn) * BLOCK_N;                                                  \
    kb0 = split_ * p.kb_per_split;                                                   \
  } else if (PAIR) {   /* grid.x = 2 x n tiles (the pair), grid.y = 256-row tiles: this CTA owns 128 of the rows */ \
    m0 = ((int)blockIdx.y * 2 + (int)cta_rank) * BLOCK_M;                            \
    n0 = ((int)blockIdx.x >> 1) * BLOCK_N;                                           \
    kb0 = (int)blockIdx.z * p.kb_per_split;                                          \
  } else {                                                                           \
    m0 = (int)blockIdx.y * BLOCK_M;                                                  \
    n0 = (int)blockIdx.x * BLOCK_N;                                                  \
    kb0 = (int)blockIdx.z * p.kb_per_split;                                          \
  }                                                                                  \
  const int nkb = min(p.kb_per_split, total_kb - kb0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int l = 0; l < LS; ++l) {
      mbar_init(conv_bar + l, CONV_GROUP + (PAIR && leader ? 1 : 0));   // own splitters (+ the peer's relay)
      mbar_init(lo_empty + l, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, epi_warps(BLOCK_N, SPLIT, LEAN) * (PAIR ? 2 : 1));   // PAIR: both CTAs' epilogue warps, on the leader
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // wide tiles: per-CTA column sums are first combined in shared memory (behind the barriers), then flushed with one
  // global atomic per column -- the 4 quarter-warps of 64 M tiles hitting the same 8 cache lines directly cost ~8 us
  float* cs_sh = reinterpret_cast<float*>(smem + p.bar_offset + 256);
  if (p.colsum)
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) cs_sh[i] = 0.0f;
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anything of this CTA signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) TPP_PROBE(1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
     int it = 0;                                   // k-blocks issued so far (stage ring position across work items)
     for (int w = wbegin; w < wend; w += wstride) {
      TPP_DECODE_WORK(w)
      int cw = 0, ch = 0, cn = 0;
      if (NARROW && p.conv_W && !p.conv_wgrad) {   // base pixel of this tile in bounding-box coordinates
        cn = m0 / p.conv_HW;
        const int r = m0 - cn * p.conv_HW;
        ch = r / p.conv_W - 1;
        cw = r % p.conv_W - 1;
      }
      // weight-gradient form: this M tile covers filter taps tap0 .. tap0 + ntaps - 1 (32 channel slots each)
      const int tap0 = (m0 / BLOCK_M) * (BLOCK_M / 32);
      const int ntaps = min(BLOCK_M / 32, 9 - tap0);
      const uint32_t tx = (NARROW && p.conv_wgrad) ? (uint32_t)(ntaps * 4096 * nops_a + p.b_tx * nops_b) : (uint32_t)p.tx_bytes;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        if (it == 0) TPP_PROBE(2);
        if (split_mode) mbar_expect_tx(full_bar + s, PAIR ? tx / 2 : tx);      // this CTA's own boxes, own barrier
        else if (!PAIR || leader) mbar_expect_tx(full_bar + s, tx);   // PAIR: tx counts both CTAs' boxes, on the leader
        uint8_t* st = smem + s * stage_bytes;
        const int kc = (kb0 + kb) * bk;
        // A.  K-major: 2-D box {32 k, rows}; MN-major: 3-D box {32 m/n, 32 k-rows, blocks of 32 m/n};
        // implicit convolution: k-block = filter tap, rows = 128 consecutive output pixels gathered by TMA im2col;
        // its weight-gradient form: k-block = 32 consecutive pixels, one im2col box {32 slots, 32 pixels} per tap
        for (int o = 0; o < 2; ++o) {
          if (o >= ta && o >= tb) break;
          if (NARROW || o < ta) {
          const CUtensorMap* tmA = o ? &tmA_lo : &tmA_hi;
          uint8_t* dst = st + o * A_BYTES;
          if (NARROW && p.conv_wgrad) {
            const int n = kc / p.conv_HW, r = kc - n * p.conv_HW;
            const int h = r / p.conv_W - 1, w = r % p.conv_W - 1;
            for (int j = 0; j < ntaps; ++j)
              tma_load_im2col(tmA, full_bar + s, dst + j * 4096, w, h, n, (tap0 + j) % 3, (tap0 + j) / 3);
          } else if (NARROW && p.conv_W) {
            const int tap = kb0 + kb;
            tma_load_im2col(tmA, full_bar + s, dst, cw, ch, cn, tap % 3, tap / 3);
          } else if (PAIR && !split_mode) {
            if (p.a_mn) tma_load_3d_pair(tmA, full_bar + s, dst, 0, kc, m0 >> 5);
            else tma_load_2d_pair(tmA, full_bar + s, dst, kc, m0);
          } else if (p.a_mn) {
            tma_load_3d(tmA, full_bar + s, dst, 0, kc, m0 >> 5);
          } else {
            tma_load_2d(tmA, full_bar + s, dst, kc, m0);
          }
          }
          if (!NARROW && o >= tb) continue;
          const CUtensorMap* tmB = o ? &tmB_lo : &tmB_hi;
          uint8_t* dstb = st + A_BYTES * ta + o * B_BYTES;
          if (PAIR) {                                   // this CTA's half of the tile's 256 B rows
            const int nb = n0 + (int)cta_rank * B_ROWS;
            if (split_mode) {
              if (p.b_mn) tma_load_3d(tmB, full_bar + s, dstb, 0, kc, nb >> 5);
              else tma_load_2d(tmB, full_bar + s, dstb, kc, nb);
            } else if (p.b_mn) tma_load_3d_pair(tmB, full_bar + s, dstb, 0, kc, nb >> 5);
            else tma_load_2d_pair(tmB, full_bar + s, dstb, kc, nb);
          } else if (p.b_mn) tma_load_3d(tmB, full_bar + s, dstb, 0, kc, n0 >> 5);
          else tma_load_2d(tmB, full_bar + s, dstb, kc, n0);
        }
      }
     }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread; the leader CTA of a pair issues for both) =====
    if (PAIR && !leader && split_mode) {
      // the peer CTA's warp 1 has no MMAs to issue: it forwards "my lo halves of stage s are written" to the leader
      int it = 0;
      for (int w = wbegin; w < wend; w += wstride) {
        TPP_DECODE_WORK(w)
        (void)m0; (void)n0; (void)kb0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int l = it % LS;
          mbar_wait(conv_bar + l, (it / LS) & 1);
          if (lane == 0)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                             smem_u32(conv_bar + l) & PEER_BIT_MASK)
                         : "memory");
          __syncwarp();
        }
      }
    }
    int it = 0, acc_i = 0;
    for (int w = wbegin; w < (PAIR && !leader ? wbegin : wend); w += wstride, ++acc_i) {
     TPP_DECODE_WORK(w)
     (void)m0; (void)n0; (void)kb0;
     const uint32_t tmem_acc = tmem_base + (uint32_t)(acc_i & 1) * ACC_COLS;
     if (PERS) {                                     // wait until the epilogue has drained this accumulator
       mbar_wait(tmem_empty + (acc_i & 1), ((acc_i >> 1) & 1) ^ 1);
       tc_fence_after();
     }
     for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      const int l = split_mode ? it % LS : 0;
      if (split_mode) mbar_wait(conv_bar + l, (it / LS) & 1);     // (implies the stage has landed, in both CTAs)
      else mbar_wait(full_bar + s, ph);
      tc_fence_after();
      if (it == 0) TPP_PROBE(3);
      if (split_mode && probe && it < 12) p.dbg[64 + it] = clock64();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
        const uint32_t b_hi = a_hi + A_BYTES * ta;
        uint32_t a_lo = a_hi + A_BYTES, b_lo = b_hi + B_BYTES;
        if (split_mode) {
          const uint32_t lo = smem_u32(smem + p.lo_offset + l * lo_bytes);
          if (p.split_a) a_lo = lo;
          if (p.split_b) b_lo = lo + (p.split_a ? A_BYTES : 0);
        }
        // small terms first, then hi*hi
        bool first = kb == 0;                       // the first MMA of a work item overwrites the accumulator
        for (int pass = p.npass - 1; pass >= 0; --pass) {
          if (!NARROW && ((p.skip >> pass) & 1)) continue;
          const uint32_t a = (pass == 2) ? a_lo : a_hi;
          const uint32_t b = (pass == 1) ? b_lo : b_hi;
          const int ksteps = bk / UMMA_K;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (NARROW && k >= ksteps) break;
            const uint32_t acc = NARROW ? ((kb > 0 || pass != p.npass - 1 || k > 0) ? 1u : 0u) : (first ? 0u : 1u);
            first = false;
            // k-step: K-major advances 32 bytes inside the swizzled row, MN-major one 8-row group (1024 bytes)
            const uint64_t da = p.a_mn ? make_desc_mn(a + k * 1024) : make_desc(a + k * UMMA_K * 4, bk);
            const uint64_t db = p.b_mn ? make_desc_mn(b + k * 1024) : make_desc(b + k * UMMA_K * 4, bk);
            if (PAIR) umma_tf32_pair(tmem_acc, da, db, IDESC, acc); else umma_tf32(tmem_acc, da, db, IDESC, acc);
          }
        }
        // frees the smem stage when these MMAs retire (PAIR: in both CTAs)
        if (PAIR) umma_commit_pair(empty_bar + s); else umma_commit(empty_bar + s);
        if (split_mode) { if (PAIR) umma_commit_pair(lo_empty + l); else umma_commit(lo_empty + l); }
        if (kb == nkb - 1) {                        // accumulator complete
          if (PAIR) umma_commit_pair(tmem_full + (acc_i & 1)); else umma_commit(tmem_full + (acc_i & 1));
          TPP_PROBE(4);
        }
      }
      __syncwarp();
     }
    }
  } else if (SPLIT && warp >= 2 + epi_warps(BLOCK_N, SPLIT, LEAN)) {
    // ===== operand splitters (wide tiles): lo = x - trunc_tf32(x), elementwise in place of the landed tile =====
    // The conversion is elementwise, so the swizzled layout of the hi tile is the layout of the lo tile: a thread walks
    // 16-byte chunks linearly (conflict-free).  One arrive per warp on the MMA issuer's barrier (the leader's, in a pair).
    if (split_mode) {
      const int cwarp = warp - 2 - epi_warps(BLOCK_N, SPLIT, LEAN);
      const int ctid = (cwarp % CONV_GROUP) * 32 + lane, cgroup = cwarp / CONV_GROUP;
      int it = 0;
      for (int w = wbegin; w < wend; w += wstride) {
        TPP_DECODE_WORK(w)
        (void)m0; (void)n0; (void)kb0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          if ((it & 1) != cgroup) continue;
          const int s = it % p.stages, l = it % LS;
          mbar_wait(lo_empty + l, ((it / LS) & 1) ^ 1);
          mbar_wait(full_bar + s, (it / p.stages) & 1);
          const bool cprobe = probe && ctid == 0 && it < 12;      // (dbg is int64[96] when a split flag is set)
          if (cprobe) p.dbg[16 + 4 * it] = clock64();
          uint8_t* st = smem + s * stage_bytes;
#pragma unroll
          for (int op = 0; op < 2; ++op) {
            if (!(op == 0 ? p.split_a : p.split_b)) continue;
            const int bytes = op == 0 ? A_BYTES : B_BYTES;
            const float4* hi = reinterpret_cast<const float4*>(st + (op == 0 ? 0 : A_BYTES * ta));
            float4* lo = reinterpret_cast<float4*>(smem + p.lo_offset + l * lo_bytes + (op == 1 && p.split_a ? A_BYTES : 0));
            for (int base = 0; base < bytes / 16; base += CONV_GROUP * 32 * 4) {
              float4 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = base + ctid + CONV_GROUP * 32 * j;
                if (i < bytes / 16) v[j] = hi[i];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = base + ctid + CONV_GROUP * 32 * j;
                if (i < bytes / 16) {
                  const float4 x = v[j];
                  lo[i] = make_float4(split_lo(x.x), split_lo(x.y), split_lo(x.z), split_lo(x.w));
                }
              }
            }
          }
          if (cprobe) p.dbg[16 + 4 * it + 1] = clock64();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the tensor core's reads
          __syncwarp();
          if (cprobe) p.dbg[16 + 4 * it + 2] = clock64();
          if (lane == 0) mbar_arrive(conv_bar + l);            // this CTA's barrier: a cluster-scope release costs ~1000
          if (cprobe) p.dbg[16 + 4 * it + 3] = clock64();       // cycles, the peer's idle MMA warp relays it instead
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> smem transpose -> (bias / relu / mask / tf32 split) -> coalesced global ====
    // A thread owns one accumulator ROW (TMEM lane).  Writing rows directly would touch 32 different 128-byte lines
    // per store instruction, so each warp first transposes 32x32 blocks through a private shared-memory patch; after
    // that a lane owns 4 fixed COLUMNS of 8 rows: bias is 4 registers, mask / output accesses are full 128-byte
    // lines (8 lanes per row, 4 rows per instruction), column sums are register accumulations.  EPI_WARPS/4 warps serve each
    // TMEM lane quarter (column groups round-robin).  The pipeline stages are free by now (tmem_full => all MMAs
    // retired), the patches alias them.
    constexpr int EPI_WARPS = epi_warps(BLOCK_N, SPLIT, LEAN);
    const int ew = warp - 2;                            // 0 .. EPI_WARPS-1
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = ew >> 2;                           // which column groups this warp takes (round-robin)
    float* stg = reinterpret_cast<float*>(smem + (PERS ? p.stg_offset : 0)) + ew * (32 * stg_pitch(LEAN));
    constexpr int GW = BLOCK_N < 32 ? BLOCK_N : 32;     // columns per staged group
    const int rr = lane >> 3, cc = (lane & 7) * 4;      // post-transpose mapping: 4 rows x (8 lanes x 4 columns)
    const bool atomic = p.flags & F_ATOMIC;
    int acc_i = 0;
    // running column sums of a persistent CTA (see the colsum flush below): kept in the 4 spare floats of this lane's
    // patch row (STG_PITCH = 36) rather than in registers -- at 192 threads the third resident CTA is lost above 96
    float* cs_run = stg + lane * STG_PITCH + 32;
    if (NARROW) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cs_run[j] = 0.0f;
    }
    int cs_n = -1;
    for (int w = wbegin; w < wend; w += wstride, ++acc_i) {
    TPP_DECODE_WORK(w)
    (void)kb0;
    const int mrow0 = m0 + quarter * 32;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc_i & 1) * ACC_COLS;
    if (NARROW && (p.flags & (F_MASK | F_ADD)) && cc < GW) {
      // the ReLU mask / residual rows of this tile are known before its accumulator is complete: pull them into L1
      // now, the epilogue's dependent loads then hit (they were ~1/3 of the data-gradient tiles' time)
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long gm = mrow0 + it * 4 + rr;
        if (gm < p.M) {
          if (p.flags & F_MASK) prefetch_l1(p.mask + gm * p.ld_mask + n0 + cc);
          if (p.flags & F_ADD) prefetch_l1(p.addend + gm * p.ld_add + n0 + cc);
        }
      }
    }
    if (PERSIST && (p.flags & F_MASK) && (lane & 7) == 0) {
      // persistent wide tiles: the same for the ReLU-mask lines of this warp's column groups (one lane per 128-byte
      // line).  The epilogue of a data-gradient item otherwise waits on 16 dependent HBM round trips per warp.
      for (int c = half * GW; c < BLOCK_N; c += (EPI_WARPS / 4) * GW) {
        if (n0 + c >= p.N) break;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const long long gm = mrow0 + it * 4 + rr;
          if (gm < p.M) prefetch_l1(p.mask + gm * p.ld_mask + n0 + c);
        }
      }
    }
    mbar_wait(tmem_full + (acc_i & 1), (acc_i >> 1) & 1);
    tc_fence_after();
    if (warp == 2) TPP_PROBE(5);
#pragma unroll 1
    for (int c = half * GW; c < BLOCK_N; c += (EPI_WARPS / 4) * GW) {
      const int n = n0 + c;
      if (n >= p.N) break;
      float v[GW];
      if (GW == 32) tmem_ld32(t_row + (uint32_t)c, v); else tmem_ld16(t_row + (uint32_t)c, v);
      if (warp == 2 && c == 0) TPP_PROBE(8);
      __syncwarp();                                     // previous group's readers are done with the patch
#pragma unroll
      for (int j = 0; j < GW; j += 4)
        *reinterpret_cast<float4*>(stg + stg_at<LEAN>(lane, j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      if (warp == 2 && c == 0) TPP_PROBE(9);
      const bool lane_on = cc < GW;                     // (BLOCK_N == 16: lanes owning columns >= 16 idle)
      const int nc = n + cc;                            // first of this lane's 4 columns
      const bool cols4 = lane_on && nc + 4 <= p.N;
      float b4[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.flags & F_BIAS) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (lane_on && nc + j < p.N) b4[j] = __ldg(p.bias + nc + j);
      }
      float cs4[4] = {0.f, 0.f, 0.f, 0.f};
      const bool vec = cols4 && ((p.ldc & 3) == 0);
      // Fast path: interior 32x32 block, everything 16-byte aligned -> branch-free body with running pointers.
      const bool interior = lane_on && (mrow0 + 32 <= p.M) && (n + GW <= p.N) && ((p.ldc & 3) == 0) &&
                            (((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.out_hi) |
                               reinterpret_cast<uintptr_t>(p.out_lo)) & 15) == 0) &&
                            (!(p.flags & F_MASK) || (((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0))) &&
                            (!(p.flags & F_ADD) || (((p.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0)));
      if (__all_sync(0xffffffffu, interior || !lane_on) && !atomic) {
        if (lane_on) {
          const long long o0 = (long long)(mrow0 + rr) * p.ldc + nc, ostep = 4 * p.ldc;
          float* po = p.out ? p.out + o0 : nullptr;
          float* ph = p.out_hi ? p.out_hi + o0 : nullptr;
          float* pl = p.out_lo ? p.out_lo + o0 : nullptr;
          const float* pm = (p.flags & F_MASK) ? p.mask + (long long)(mrow0 + rr) * p.ld_mask + nc : nullptr;
          const long long mstep = 4 * p.ld_mask;
          const float* pa = (NARROW && (p.flags & F_ADD)) ? p.addend + (long long)(mrow0 + rr) * p.ld_add + nc : nullptr;
          const long long astep = 4 * p.ld_add;
          const float floor_v = (p.flags & F_RELU) ? 0.0f : -3.402823466e38f;
          const float floor_out = (p.flags & F_RELU_OUT) ? 0.0f : -3.402823466e38f;
          const float floor_pair = (p.flags & F_PAIR_RELU) ? 0.0f : -3.402823466e38f;   // pair = relu(plain)
          const float* sp = stg + rr * STG_PITCH + cc;    // (LEAN: per-row swizzle, see the load below)
          // wide tiles: the 8 ReLU-mask loads of this lane are issued together, in front of the loop (one dependent
          // global load per iteration left the epilogue of a data-gradient item latency-bound)
          constexpr int UNR = NARROW ? 2 : 8;
          float4 mq8[NARROW ? 1 : 8];
          if (!NARROW && pm) {
#pragma unroll
            for (int it = 0; it < (NARROW ? 1 : 8); ++it) mq8[it] = *reinterpret_cast<const float4*>(pm + it * mstep);
          }
          // 1-bit masks: the 32 x 32 block at (mrow0, n) owns 32 words; word it*4 + j holds, at bit `lane`, whether the
          // element this lane handles in iteration it (row it*4 + rr, column cc + j) is positive.  The producing and the
          // consuming epilogue share this lane mapping, so the layout is private to the kernel.
          const long long wbase = !NARROW && (p.bits_out || p.bits_in)
                                      ? ((long long)(mrow0 >> 5) * p.bits_ncb + (n >> 5)) * 32 : 0;
          uint32_t wmine = 0u;
          if (!NARROW && p.bits_in) wmine = p.bits_in[wbase + lane];
#pragma unroll UNR
          for (int it = 0; it < 8; ++it) {
            const float4 q = LEAN ? *reinterpret_cast<const float4*>(stg + stg_at<true>(it * 4 + rr, cc))
                                  : *reinterpret_cast<const float4*>(sp + it * 4 * STG_PITCH);
            float x[4] = {fmaxf(q.x + b4[0], floor_v), fmaxf(q.y + b4[1], floor_v), fmaxf(q.z + b4[2], floor_v),
                          fmaxf(q.w + b4[3], floor_v)};
            if (pm) {
              float4 mq;
              if (NARROW) { mq = *reinterpret_cast<const float4*>(pm); pm += mstep; }
              else mq = mq8[NARROW ? 0 : it];
              x[0] = mq.x > 0.0f ? x[0] : 0.0f;
              x[1] = mq.y > 0.0f ? x[1] : 0.0f;
              x[2] = mq.z > 0.0f ? x[2] : 0.0f;
              x[3] = mq.w > 0.0f ? x[3] : 0.0f;
            }
            if (!NARROW && p.bits_in) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t wd = __shfl_sync(0xffffffffu, wmine, it * 4 + j);
                x[j] = ((wd >> lane) & 1u) ? x[j] : 0.0f;
              }
            }
            if (!NARROW && p.bits_out) {
              uint32_t wd = 0u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t b = __ballot_sync(0xffffffffu, x[j] > 0.0f);
                if (lane == j) wd = b;
              }
              if (lane < 4) p.bits_out[wbase + it * 4 + lane] = wd;
            }
            if (NARROW) {
              if (pa) {
                const float4 aq = *reinterpret_cast<const float4*>(pa);
                pa += astep;
                x[0] += aq.x; x[1] += aq.y; x[2] += aq.z; x[3] += aq.w;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], floor_out);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) cs4[j] += x[j];
            if (po) { *reinterpret_cast<float4*>(po) = make_float4(x[0], x[1], x[2], x[3]); po += ostep; }
            if (ph) {
              if (NARROW) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], floor_pair);
              }
              const float h0 = tf32_round(x[0]), h1 = tf32_round(x[1]), h2 = tf32_round(x[2]), h3 = tf32_round(x[3]);
              *reinterpret_cast<float4*>(ph) = make_float4(h0, h1, h2, h3);
              *reinterpret_cast<float4*>(pl) = make_float4(x[0] - h0, x[1] - h1, x[2] - h2, x[3] - h3);
              ph += ostep;
              pl += ostep;
            }
          }
        }
      } else
      // Generic path (edges, unaligned, atomic accumulation).
      // NOT unrolled: the body is ~200 instructions; eight unrolled copies (x2 groups, x4 template instances) do not
      // fit the instruction caches and every iteration then stalls ~400 cycles on instruction fetch (measured).
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rr, gm = mrow0 + r;
        const float4 q = *reinterpret_cast<const float4*>(stg + stg_at<LEAN>(r, cc));
        float x[4] = {q.x, q.y, q.z, q.w};
        if (gm >= p.M || !lane_on) continue;
        const long long off = (long long)gm * p.ldc + nc;
        if (atomic) {
          if (nkb > 0) {
            float* dst = p.out + off;
            if (!NARROW) {
#pragma unroll
              for (int j = 0; j < 4; ++j) x[j] *= p.alpha;
            }
            if (vec && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x[0]), "f"(x[1]), "f"(x[2]),
                           "f"(x[3])
                           : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (nc + j < p.N) atomicAdd(dst + j, x[j]);
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          x[j] += b4[j];
          if (p.flags & F_RELU) x[j] = fmaxf(x[j], 0.0f);
        }
        if (p.flags & F_MASK) {
          const float* mk = p.mask + (long long)gm * p.ld_mask + nc;
          if (cols4 && ((p.ld_mask & 3) == 0) && (reinterpret_cast<uintptr_t>(mk) & 15) == 0) {
            const float4 mq = *reinterpret_cast<const float4*>(mk);
            if (!(mq.x > 0.0f)) x[0] = 0.0f;
            if (!(mq.y > 0.0f)) x[1] = 0.0f;
            if (!(mq.z > 0.0f)) x[2] = 0.0f;
            if (!(mq.w > 0.0f)) x[3] = 0.0f;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (nc + j < p.N && !(mk[j] > 0.0f)) x[j] = 0.0f;
          }
        }
        if (NARROW && (p.flags & F_ADD)) {
          const float* ad = p.addend + (long long)gm * p.ld_add + nc;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nc + j < p.N) x[j] += ad[j];
        }
        if (NARROW && (p.flags & F_RELU_OUT)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cs4[j] += x[j];
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float y = (NARROW && (p.flags & F_PAIR_RELU)) ? fmaxf(x[j], 0.0f) : x[j];
          h[j] = tf32_round(y);
          l[j] = y - h[j];
        }
        float* outs[3] = {p.out, p.out_hi, p.out_lo};
        const float* src[3] = {x, h, l};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (!outs[a]) continue;
          float* dst = outs[a] + off;
          if (vec && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(src[a][0], src[a][1], src[a][2], src[a][3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (nc + j < p.N) dst[j] = src[a][j];
          }
        }
      }
      if (warp == 2 && c == 0) TPP_PROBE(10);
      if (p.colsum && !atomic) {   // bias gradient: this warp's 32 rows of columns nc..nc+3 (lanes l, l+8, l+16, l+24)
        if (NARROW) {
          // persistent CTA, one column group per tile: keep the sums in registers across tiles and flush them when the
          // column group changes / at the end (16 addresses would otherwise take ~130 000 same-address atomics each)
          if (cs_n != n && cs_n >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v = cs_run[j];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (rr == 0 && cc < GW && cs_n + cc + j < p.N) atomicAdd(p.colsum + cs_n + cc + j, v);
              cs_run[j] = 0.0f;
            }
          }
          cs_n = n;
#pragma unroll
          for (int j = 0; j < 4; ++j) cs_run[j] += cs4[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cs4[j] += __shfl_xor_sync(0xffffffffu, cs4[j], 8);
            cs4[j] += __shfl_xor_sync(0xffffffffu, cs4[j], 16);
          }
          if (rr == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (lane_on && nc + j < p.N) atomicAdd(cs_sh + c + cc + j, cs4[j]);
          }
        }
      }
    }
    if (!NARROW && p.colsum && !atomic) {      // all epilogue warps of the CTA: combine, then one atomic per column
      asm volatile("bar.sync 1, %0;" ::"n"(epi_warps(BLOCK_N, SPLIT, LEAN) * 32) : "memory");
      for (int i = ew * 32 + lane; i < BLOCK_N; i += epi_warps(BLOCK_N, SPLIT, LEAN) * 32) {
        if (n0 + i < p.N) atomicAdd(p.colsum + n0 + i, cs_sh[i]);
        if (PERSIST) cs_sh[i] = 0.0f;          // the next work item starts from zero ...
      }
      if (PERSIST) asm volatile("bar.sync 1, %0;" ::"n"(epi_warps(BLOCK_N, SPLIT, LEAN) * 32) : "memory");   // ... in every warp's view
    }
    if (PERS) {                         // accumulator drained: hand it back to the MMA warp (PAIR: the leader's)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) {
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                           smem_u32(tmem_empty + (acc_i & 1)) & PEER_BIT_MASK)
                       : "memory");
        } else {
          mbar_arrive(tmem_empty + (acc_i & 1));
        }
      }
    }
    }   // work items
    if (NARROW && p.colsum && !atomic) {
      // final flush of the running sums: the CTA's 4 epilogue warps combine in shared memory, one warp adds to global
      if (cs_n >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = cs_run[j];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (rr == 0 && cc < GW) atomicAdd(cs_sh + cc + j, v);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(epi_warps(BLOCK_N, SPLIT, LEAN) * 32) : "memory");
      if (ew == 0 && cs_n >= 0 && lane < GW && cs_n + lane < p.N) atomicAdd(p.colsum + cs_n + lane, cs_sh[lane]);
    }
  }
  if (warp == 2) TPP_PROBE(6);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // neither CTA's shared memory / TMEM goes away while the pair's MMAs may use it
  if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
  if (warp == 1) TPP_PROBE(7);
#undef TPP_PROBE
}

// x -> (hi, lo) [+ transposed (hi, lo)]; rows x cols, source row stride ld_in, destination row stride ld_out
// (columns in [cols, ld_out) are zero-filled so that padded operands stay exact).
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, long long ld_in, int rows, int cols,
                                                         float* __restrict__ hi, float* __restrict__ lo,
                                                         long long ld_out, float* __restrict__ t_hi,
                                                         float* __restrict__ t_lo, long long ld_t) {
  __shared__ float th[32][33], tl[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.0f;
    if (r < rows && c < cols) v = x[(long long)r * ld_in + c];
    const float h = tf32_round(v), l = v - h;
    if (r < rows && c < ld_out) {
      if (hi) hi[(long long)r * ld_out + c] = h;
      if (lo) lo[(long long)r * ld_out + c] = l;
    }
    th[ty + i * 8][tx] = h;
    tl[ty + i * 8][tx] = l;
  }
  if (!t_hi) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;       // transposed: row index of the output is the source column
    if (c < cols && r < ld_t) {
      const bool in = r < rows;
      t_hi[(long long)c * ld_t + r] = in ? th[tx][ty + i * 8] : 0.0f;
      t_lo[(long long)c * ld_t + r] = in ? tl[tx][ty + i * 8] : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
// K-major operand: matrix [rows][ld] with the contraction index contiguous; box = {32 k, box_rows}.
// MN-major operand: matrix [K rows][ld] with the m/n index contiguous; viewed as a 3-D tensor
// (32 m/n, K rows, blocks of 32 m/n) so that one TMA box {32, 32 k-rows, box_rows/32} lands in shared memory as the
// canonical MN-major SW128_32B layout (atoms 4 k-rows x 128 B; m/n blocks BLOCK_K*128 B apart).
// cuTensorMapEncodeTiled is a driver-API symbol: resolve it through the runtime at first use so that the library has
// no link-time dependency on libcuda.so (it must load, and export its symbols, on machines without a driver).
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

// NHWC fp32 tensor [B][H][W][C] as the A operand of a 3x3 / pad-1 convolution: box = BLOCK_M pixels x 32 channel slots
// (one 128-byte swizzled row per pixel, the same shared-memory tile as the K-major tiled map), conventions pinned by
// tests/test_conv_ops.py::test_tma_im2col_conventions.
static int make_map_im2col(CUtensorMap* tm, const float* base, int B, int H, int W, int C, int wgrad, int bk) {
  if (!base || (reinterpret_cast<uintptr_t>(base) & 15) || (C & 3) || C > 32) return TPP_EINVAL;
  EncodeIm2colFn enc = encode_im2col();
  if (!enc) return TPP_ENOTSUP;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lower[2] = {-1, -1}, upper[2] = {-1, -1};   // -pad ; pad - (3 - 1)
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // forward / data gradient: K-major A tile, 128 pixels x 128-byte rows; weight gradient: MN-major A blocks of
  // 32 pixels (k rows) x 32 channel slots in the 32-byte-atom swizzle (the only legal MN-major fp32 layout)
  // (bk = 16: 16 channel slots per pixel, 64-byte rows in the 64-byte swizzle)
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, lower, upper,
                   (cuuint32_t)bk, wgrad ? BLOCK_K : BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   wgrad ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                         : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TPP_OK : TPP_EINVAL;
}

template <int BLOCK_N, bool PAIR = false, bool PERSIST = false, bool SPLIT = false, bool LEAN = false>
static int launch(const tpp_tc_gemm* g, int split_k, cudaStream_t s) {
  constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;     // B rows one CTA stages
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  const int npass = (g->precision & 15) == 3 ? 3 : 1;
  if (BLOCK_N <= 32 && ((g->precision & 48) || (g->alpha != 0.0f && g->alpha != 1.0f))) return TPP_ENOTSUP;
  const bool a_exact = npass == 3 && (g->precision & 16), b_exact = npass == 3 && (g->precision & 32);
  const bool a_split = npass == 3 && (g->precision & 64) && !a_exact, b_split = npass == 3 && (g->precision & 128) && !b_exact;
  if ((a_split || b_split) != SPLIT) return TPP_EINVAL;     // (dispatched by tpp_gemm_tc)
  const int nops_a = (npass == 3 && !a_exact) ? 2 : 1, nops_b = (npass == 3 && !b_exact) ? 2 : 1;
  if (BLOCK_N > 32 && ((g->flags & (F_ADD | F_RELU_OUT | F_PAIR_RELU)) || g->conv_C > 0))
    return TPP_ENOTSUP;   // residual / ReLU-pair epilogues and convolution mode are compiled into the narrow tiles only
  int rc, a_bytes = A_BYTES, b_bytes = B_ROWS * BLOCK_K * 4;
  const bool conv = g->conv_C > 0;
  const int wgrad = conv && g->conv_wgrad;
  int bk = BLOCK_K;
  if (conv) {
    const long long pixels = (long long)g->conv_B * g->conv_H * g->conv_W;
    if (wgrad) {
      if (!g->a_mn || !g->b_mn || g->M != 9 * 32 || pixels != g->K || !(g->flags & F_ATOMIC)) return TPP_EINVAL;
    } else {
      if (g->K == 9 * 16 && g->conv_C <= 16 && !g->b_mn) bk = 16;    // 16 channel slots per tap
      if (g->a_mn || g->K != 9 * bk || pixels != g->M) return TPP_EINVAL;
      a_bytes = BLOCK_M * bk * 4;
    }
    if ((rc = make_map_im2col(&tmA_hi, g->a_hi, g->conv_B, g->conv_H, g->conv_W, g->conv_C, wgrad, bk))) return rc;
    if (nops_a == 2 && (rc = make_map_im2col(&tmA_lo, g->a_lo, g->conv_B, g->conv_H, g->conv_W, g->conv_C, wgrad, bk)))
      return rc;
  } else if ((rc = make_map(&tmA_hi, g->a_hi, g->lda, g->M, g->K, BLOCK_M, g->a_mn, &a_bytes))) return rc;
  if ((rc = make_map(&tmB_hi, g->b_hi, g->ldb, g->N, g->K, B_ROWS, g->b_mn, &b_bytes, bk))) return rc;
  if (nops_a == 2 && !a_split) {
    if (!conv && (rc = make_map(&tmA_lo, g->a_lo, g->lda, g->M, g->K, BLOCK_M, g->a_mn, &a_bytes))) return rc;
  } else {
    tmA_lo = tmA_hi;
  }
  if (nops_b == 2 && !b_split) {
    if ((rc = make_map(&tmB_lo, g->b_lo, g->ldb, g->N, g->K, B_ROWS, g->b_mn, &b_bytes, bk))) return rc;
  } else {
    tmB_lo = tmB_hi;
  }
  Params p;
  p.M = g->M; p.N = g->N; p.K = g->K; p.npass = npass; p.flags = g->flags;
  p.bias = g->bias; p.mask = g->mask; p.ld_mask = g->ld_mask;
  p.bits_out = g->mask_bits_out; p.bits_in = g->mask_bits; p.bits_ncb = (g->N + 31) / 32;
  if (g->mask_bits_out || g->mask_bits) {
    // the bit words are written / read by the interior fast path of the wide tiles' epilogue only
    if (BLOCK_N <= 32 || (g->M & 31) || (g->N & 31) || (g->ldc & 3) || (g->flags & F_ATOMIC) ||
        ((reinterpret_cast<uintptr_t>(g->out) | reinterpret_cast<uintptr_t>(g->out_hi) |
          reinterpret_cast<uintptr_t>(g->out_lo)) & 15))
      return TPP_ENOTSUP;
  }
  p.addend = g->addend; p.ld_add = g->ld_add;
  p.nops_a = nops_a; p.nops_b = nops_b;
  p.split_a = a_split ? 1 : 0; p.split_b = b_split ? 1 : 0;
  p.tma_a = a_split ? 1 : nops_a; p.tma_b = b_split ? 1 : nops_b;
  p.skip = (a_exact ? 4 : 0) | (b_exact ? 2 : 0);
  p.alpha = g->alpha == 0.0f ? 1.0f : g->alpha;
  p.tx_bytes = (a_bytes * p.tma_a + b_bytes * p.tma_b) * (PAIR ? 2 : 1);   // PAIR: both CTAs' boxes land on one barrier
  p.conv_W = conv ? g->conv_W : 0; p.conv_HW = conv ? g->conv_H * g->conv_W : 0;
  p.conv_wgrad = wgrad; p.b_tx = b_bytes;
  p.out = g->out; p.ldc = g->ldc; p.out_hi = g->out_hi; p.out_lo = g->out_lo;
  p.colsum = g->colsum; p.a_mn = g->a_mn ? 1 : 0; p.b_mn = g->b_mn ? 1 : 0;
  p.dbg = reinterpret_cast<long long*>(g->dbg);
  p.dbg_y = BLOCK_N <= 32 ? 0 : g->_reserved;
  p.bk = bk;
  const int total_kb = (g->K + bk - 1) / bk;
  if (split_k < 1) split_k = 1;
  if (split_k > total_kb) split_k = total_kb;
  p.kb_per_split = (total_kb + split_k - 1) / split_k;
  split_k = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  // slots keep the full tile size: the MMA reads all BLOCK_M x BLOCK_N operand rows, also the ones a smaller TMA box
  // left unfilled (their products land in accumulator rows / columns that are never stored)
  p.a_slot = BLOCK_M * bk * 4;
  p.b_slot = B_ROWS * bk * 4;
  const int stage_bytes = p.a_slot * p.tma_a + p.b_slot * p.tma_b;
  const int lo_bytes = p.a_slot * p.split_a + p.b_slot * p.split_b;
  p.lo_stages = lo_bytes ? 2 : 0;
  // (split on chip: every byte counts -- the exact 227 KB limit; otherwise the round-1 budget the tuning was done with)
  const int budget = ((lo_bytes || LEAN) ? 227 * 1024 - 1024 - 256 : 224 * 1024 - 1024 - 256) - p.lo_stages * lo_bytes;
  int stages = (budget - ((BLOCK_N <= 32 || PERSIST) ? stg_bytes(BLOCK_N, SPLIT, LEAN) : 0)) / stage_bytes;
  if (stages < 1) return TPP_ENOTSUP;
  if (stages > (BLOCK_N <= 32 ? 8 : 4)) stages = BLOCK_N <= 32 ? 8 : 4;
  // many more tiles than SMs and a short contraction (convolution rows): trade pipeline depth for 2-3 resident CTAs
  // per SM so that one tile's epilogue / prologue overlaps another tile's loads (measured on the IMPALA shapes:
  // throughput follows the number of resident CTAs, not the depth: 3 CTAs x 2 stages beat 1 CTA x 6 stages by 1.4-2x)
  const long long n_tiles = (long long)((g->N + BLOCK_N - 1) / BLOCK_N) * ((g->M + BLOCK_M - 1) / BLOCK_M) * split_k;
  // (narrow tiles only: the wide CTAs -- 576 threads x 67 registers -- are alone on their SM whatever their stage count)
  if (n_tiles >= 4 * 148 && BLOCK_N <= 32) {
    int few = (74 * 1024 - (BLOCK_N <= 32 ? stg_bytes(BLOCK_N, SPLIT, LEAN) : 0)) / stage_bytes;   // three resident CTAs
    if (few < 2) few = 2;
    if (stages > few) stages = few;
  }
  if (stages > p.kb_per_split && BLOCK_N > 32 && !PERSIST) stages = p.kb_per_split < 1 ? 1 : p.kb_per_split;
  p.stages = stages;
  constexpr bool NARROW = BLOCK_N <= 32;
  size_t region = (size_t)stages * stage_bytes;
  p.lo_offset = (int)region;
  region += (size_t)p.lo_stages * lo_bytes;
  if (NARROW || PERSIST) {
    // persistent CTAs: the epilogue's transpose patches get their own region (the next tile's loads are in flight)
    p.stg_offset = (int)region;
    region += stg_bytes(BLOCK_N, SPLIT, LEAN);
  } else {
    // the patches alias the pipeline stages (free once the accumulator is complete): the region must hold them
    p.stg_offset = 0;
    if (region < (size_t)stg_bytes(BLOCK_N, SPLIT, LEAN)) region = stg_bytes(BLOCK_N, SPLIT, LEAN);
  }
  p.bar_offset = (int)region;
  const size_t smem = region + 1024 + 256;
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, PAIR, PERSIST, SPLIT, LEAN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.ntn = (g->N + BLOCK_N - 1) / BLOCK_N;
  p.ntiles = p.ntn * ((g->M + (PAIR ? 2 : 1) * BLOCK_M - 1) / ((PAIR ? 2 : 1) * BLOCK_M));   // PAIR: 256-row tiles
  p.total_work = p.ntiles * split_k;
  unsigned grid_x = (unsigned)p.total_work;
  if (NARROW) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // resident CTAs per SM from the kernel's own resources (registers, shared memory, TMEM columns)
    static int regs_per_cta = 0;
    if (!regs_per_cta) {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, gemm_tc_kernel<BLOCK_N, PAIR, PERSIST, SPLIT, LEAN>) != cudaSuccess) return TPP_ENOTSUP;
      regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * num_threads(BLOCK_N, SPLIT, LEAN);
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    // registers are allocated per warp inside each of the 4 SM sub-partitions (16384 registers each)
    const int regs_per_warp = regs_per_cta / (num_threads(BLOCK_N, SPLIT, LEAN) / 32);
    const int warps_per_sm = 4 * (16384 / regs_per_warp);
    if (per_sm > warps_per_sm / (num_threads(BLOCK_N, SPLIT, LEAN) / 32)) per_sm = warps_per_sm / (num_threads(BLOCK_N, SPLIT, LEAN) / 32);
    if (per_sm > 512 / (2 * (BLOCK_N < 32 ? 32 : BLOCK_N))) per_sm = 512 / (2 * (BLOCK_N < 32 ? 32 : BLOCK_N));
    if (per_sm < 1) per_sm = 1;
    const unsigned resident = (unsigned)(per_sm * sms);
    if (grid_x > resident) grid_x = resident;
  }
  dim3 grid(grid_x, 1, 1);
  if (!NARROW) {
    const int ntm = (g->M + BLOCK_M - 1) / BLOCK_M;
    if (ntm > 65535) return TPP_ENOTSUP;
    grid = dim3((unsigned)p.ntn, (unsigned)ntm, (unsigned)split_k);
  }
  if (PAIR) {
    // a cluster of two CTAs along grid.x per 256 x 256 tile; grid.y counts 256-row tiles
    const int ntm2 = (g->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    if (ntm2 > 65535) return TPP_ENOTSUP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)p.ntn, (unsigned)ntm2, (unsigned)split_k);
    if (PERSIST) {                 // one pair per two SMs, each looping over the (tile, split) work items
      int sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const unsigned pairs = (unsigned)std::min(p.total_work, sms / 2);
      cfg.gridDim = dim3(2u * pairs, 1, 1);
    }
    cfg.blockDim = dim3(num_threads(BLOCK_N, SPLIT, LEAN), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BLOCK_N, PAIR, PERSIST, SPLIT, LEAN>, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    if (e != cudaSuccess) return (int)e;
    TPP_LAUNCH_STATUS();
  }
  gemm_tc_kernel<BLOCK_N, PAIR, false, SPLIT, LEAN><<<grid, num_threads(BLOCK_N, SPLIT, LEAN), smem, s>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
  TPP_LAUNCH_STATUS();
}

}  // namespace tc
}  // namespace tpp

extern "C" int tpp_gemm_tc(const tpp_tc_gemm* g, void* stream) {
  TPP_CHECK_ARG(g && g->a_hi && g->b_hi && g->M > 0 && g->N > 0 && g->K > 0);
  TPP_CHECK_ARG((g->precision & ~240) == 1 || (g->precision & ~240) == 3);
  TPP_CHECK_ARG((g->precision & 15) == 1 || ((g->a_lo || (g->precision & (16 | 64))) && (g->b_lo || (g->precision & (32 | 128)))));
  const bool atomic = g->flags & tpp::tc::F_ATOMIC;
  TPP_CHECK_ARG(!atomic || g->out);
  TPP_CHECK_ARG(g->split_k <= 1 || atomic);
  TPP_CHECK_ARG(!(g->flags & tpp::tc::F_BIAS) || g->bias);
  TPP_CHECK_ARG(!(g->flags & tpp::tc::F_ADD) || (g->addend && !(g->flags & tpp::tc::F_ATOMIC)));
  TPP_CHECK_ARG(!(g->flags & tpp::tc::F_MASK) || g->mask);
  TPP_CHECK_ARG((g->out_hi == nullptr) == (g->out_lo == nullptr));
  cudaStream_t s = tpp_stream(stream);
  int bn = g->block_n;
  if (bn == 0) bn = g->N <= 16 ? 16 : (g->N <= 32 ? 32 : (g->N <= 64 ? 64 : 128));
  if ((g->precision & 15) == 3 && (g->precision & (TPP_TC_A_SPLIT | TPP_TC_B_SPLIT)) &&
      !((g->precision & TPP_TC_A_EXACT) && (g->precision & TPP_TC_B_EXACT))) {
    // lo halves formed on chip: the SPLIT kernels (extra splitter warps, separate lo ring)
    const bool a_s = (g->precision & TPP_TC_A_SPLIT) && !(g->precision & TPP_TC_A_EXACT);
    const bool b_s = (g->precision & TPP_TC_B_SPLIT) && !(g->precision & TPP_TC_B_EXACT);
    if (a_s || b_s) {
      switch (bn) {
        case 64: return tpp::tc::launch<64, false, false, true>(g, g->split_k, s);
        case 128: return tpp::tc::launch<128, false, false, true>(g, g->split_k, s);
        case 256: return tpp::tc::launch<256, false, false, true>(g, g->split_k, s);
        case TPP_TC_TILE_PAIR: return tpp::tc::launch<256, true, false, true>(g, g->split_k, s);
        case TPP_TC_TILE_PAIR_PERSISTENT: return tpp::tc::launch<256, true, true, true>(g, g->split_k, s);
        case TPP_TC_TILE_PAIR64_PERSISTENT: return tpp::tc::launch<64, true, true, true>(g, g->split_k, s);
        default: return TPP_ENOTSUP;      // the narrow (convolution) tiles take pairs only
      }
    }
  }
  switch (bn) {
    case 16: return tpp::tc::launch<16>(g, g->split_k, s);
    case 32: return tpp::tc::launch<32>(g, g->split_k, s);
    case 64: return tpp::tc::launch<64>(g, g->split_k, s);
    case 128: return tpp::tc::launch<128>(g, g->split_k, s);
    case 256: return tpp::tc::launch<256>(g, g->split_k, s);
    case TPP_TC_TILE_PAIR: return tpp::tc::launch<256, true>(g, g->split_k, s);   // 256 x 256 tile, one per CTA pair
    case TPP_TC_TILE_PAIR_PERSISTENT: return tpp::tc::launch<256, true, true>(g, g->split_k, s);
    case TPP_TC_TILE_PAIR_PERSISTENT_LEAN: return tpp::tc::launch<256, true, true, false, true>(g, g->split_k, s);
    case TPP_TC_TILE_PAIR64_PERSISTENT: return tpp::tc::launch<64, true, true>(g, g->split_k, s);   // 256 x 64 tiles
    default: return TPP_ENOTSUP;
  }
}

extern "C" int tpp_split_tf32(const float* x, int64_t ld_in, int32_t rows, int32_t cols, float* hi, float* lo,
                              int64_t ld_out, float* t_hi, float* t_lo, int64_t ld_t, void* stream) {
  TPP_CHECK_ARG(x && rows > 0 && cols > 0 && ld_in >= cols);
  TPP_CHECK_ARG((hi == nullptr) == (lo == nullptr) && (t_hi == nullptr) == (t_lo == nullptr) && (hi || t_hi));
  TPP_CHECK_ARG(!hi || ld_out >= cols);
  TPP_CHECK_ARG(!t_hi || ld_t >= rows);
  const int64_t ext_c = hi && ld_out > cols ? ld_out : cols, ext_r = t_hi && ld_t > rows ? ld_t : rows;
  dim3 grid(tpp_ceil_div(ext_c, 32), tpp_ceil_div(ext_r, 32));
  tpp::tc::split_tf32_kernel<<<grid, 256, 0, tpp_stream(stream)>>>(x, ld_in, rows, cols, hi, lo, ld_out, t_hi, t_lo,
                                                                   ld_t);
  TPP_LAUNCH_STATUS();
}
