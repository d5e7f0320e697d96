// 3x3 convolution weight gradient for the IMPALA-CNN's narrow layers (16 / 32 channels) on the fp32 FMA pipe.
//
//   gw[(ky*3 + kx)*32 + ci][co] += sum over (b, y, x) of  X[b][y + ky - 1][x + kx - 1][ci] * dY[b][y][x][co]
//
// (reference: autograd of nn.Conv2d(k=3, padding=1) in common/model.py:134-208, ResidualBlock / ImpalaBlock).
//
// Why not the tensor core here.  The output is tiny (9 * Cin * Cout <= 9216 numbers) and the contraction runs over
// millions of pixels.  On tcgen05 (csrc/gemm_tc.cu, conv_wgrad form) that is an M = 288, N = 16..32 problem: fp32 parity
// needs three TF32 passes, the im2col operand is re-gathered once per tap from L2 (6.4 GB for a 268 MB tensor at the
// block-1 shape, half of it zero-fill because MN-major fp32 rows must be 128 bytes) and N = 16 uses an eighth of an MMA's
// columns: measured 825 us = 11.7 TFLOP/s at the block-1 shape (2 M pixels, 16 -> 16 channels), ten times the HBM time of
// its operands.  The same contraction is 4.8 G fp32 FMAs -- 134 us of the B200's FMA pipe -- when every operand byte
// is read from HBM once, staged in shared memory once and reused from registers:
//   * a CTA walks tiles of R = 8 image rows; the (R + 2) x (W + 2) input halo and the R x W dY tile are staged by
//     cp.async (zero fill outside the image) into a double buffer, so the next tile loads under this tile's FMAs;
//   * a thread owns a 4-ci x 4-co block of ALL nine taps (144 accumulators) for one pixel group; it slides along a row
//     segment keeping the 3 x 3 window of float4 input vectors in registers: per pixel 3 new LDS.128 of X, one of dY,
//     144 FFMA (36 FMAs per shared-memory load);
//   * the CTA's pixel groups (16 for 16 x 16 channels) take different rows of the tile; groups sharing a warp read rows an
//     ODD number of pixels apart (odd row pitches), i.e. disjoint bank halves -- every LDS.128 is one wavefront;
//   * accumulation chains stay short (a thread sums ~1000 pixels), partial sums meet in shared memory, then one
//     red.global.add.v4.f32 per four outputs and CTA.
// Exact fp32 products (no TF32 split), input read as relu(plain) -- the activation the forward pass convolved.
// Measured on B200 (profiles/ncu_conv_fma_r02.md): 267 us at the block-1 shape, DRAM read 268.7 MB = X + dY once each, FMA pipe
// 53 % active (what is left: shared-memory load latency at 8 warps per SM and the 2.5 K-instruction unrolled segment).
#include "tpp_common.cuh"

namespace tpp {

template <int CIN, int COUT, int W>
struct WgCfg {
  static constexpr int R = 8;                              // image rows per tile
  static constexpr int CI_B = CIN / 4, CO_B = COUT / 4;
  static constexpr int COMBOS = CI_B * CO_B;               // threads of one pixel group
  static constexpr int THREADS = 256;
  static constexpr int G = THREADS / COMBOS;               // pixel groups per CTA
  static constexpr int SEG = W >= 16 ? 16 : W;             // pixels a group walks in one unit
  static constexpr int SEGS = W / SEG;
  static constexpr int UNITS = R * SEGS;                   // (row, segment) units of a tile
  static constexpr int XP = (W + 2) | 1;                   // halo row pitch in pixels (odd)
  static constexpr int DP = W | 1;                         // dY row pitch in pixels (odd)
  static constexpr int X_FLOATS = (R + 2) * XP * CIN;
  static constexpr int D_FLOATS = R * DP * COUT;
  static constexpr int BUF_FLOATS = X_FLOATS + D_FLOATS;
  static constexpr int OUT_FLOATS = 9 * CIN * COUT;
  // (the final reduction gives every pixel group its own copy of the outputs: G * OUT_FLOATS = 36864 floats for every shape)
  static constexpr size_t SMEM = sizeof(float) * (2 * BUF_FLOATS > G * OUT_FLOATS ? 2 * BUF_FLOATS : G * OUT_FLOATS);
  static_assert(THREADS % COMBOS == 0 && UNITS % G == 0 && W % SEG == 0, "tile does not divide");
};

__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n)
               : "memory");
}

template <int CIN, int COUT, int W>
__global__ void __launch_bounds__(256, 1) conv3x3_wgrad_cc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ gw, int B, int relu) {
  using C = WgCfg<CIN, COUT, W>;
  constexpr int H = W, R = C::R;
  extern __shared__ __align__(16) float wg_sm[];
  const int tid = threadIdx.x;
  const int combo = tid % C::COMBOS, grp = tid / C::COMBOS;
  const int ci4 = combo % C::CI_B, co4 = combo / C::CI_B;
  const int tiles_per_img = H / R;
  const int ntiles = B * tiles_per_img;

  auto stage = [&](int tile, float* buf) {
    const int b = tile / tiles_per_img, y0 = (tile % tiles_per_img) * R;
    const float* xb = x + (size_t)b * H * W * CIN;
    const float* db = dy + ((size_t)b * H + y0) * W * COUT;
    constexpr int XC = (R + 2) * (W + 2) * C::CI_B;          // 16-byte chunks of the halo
    for (int i = tid; i < XC; i += C::THREADS) {
      const int c = i % C::CI_B, hx = (i / C::CI_B) % (W + 2), hr = i / (C::CI_B * (W + 2));
      const int gy = y0 - 1 + hr, gx = hx - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async16_zfill(buf + (hr * C::XP + hx) * CIN + c * 4, ok ? xb + ((size_t)gy * W + gx) * CIN + c * 4 : x, ok);
    }
    float* dbuf = buf + C::X_FLOATS;
    constexpr int DC = R * W * C::CO_B;
    for (int i = tid; i < DC; i += C::THREADS) {
      const int c = i % C::CO_B, px = (i / C::CO_B) % W, r = i / (C::CO_B * W);
      cp_async16_zfill(dbuf + (r * C::DP + px) * COUT + c * 4, db + ((size_t)r * W + px) * COUT + c * 4, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[9][4][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.0f;

  int it = 0;
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, wg_sm);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    float* buf = wg_sm + (it & 1) * C::BUF_FLOATS;
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      stage(next, wg_sm + ((it + 1) & 1) * C::BUF_FLOATS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* dbuf = buf + C::X_FLOATS;
#pragma unroll 1
    for (int u = grp; u < C::UNITS; u += C::G) {
      const int r = u % R, x0 = (u / R) * C::SEG;            // groups g, g + 1 (same warp): adjacent rows, odd pitch apart
      const float* xp = buf + (r * C::XP + x0) * CIN + ci4 * 4;      // halo (r, x0): tap (0, 0) of output pixel (r, x0)
      const float* dp = dbuf + (r * C::DP + x0) * COUT + co4 * 4;
      float4 col[3][3];                                      // [halo column x0 + k][ky]
      auto ldx = [&](int hx, int ky) {
        float4 v = *reinterpret_cast<const float4*>(xp + (ky * C::XP + hx) * CIN);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        return v;
      };
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) { col[0][ky] = ldx(0, ky); col[1][ky] = ldx(1, ky); }
      // one pixel: the new halo column replaces the oldest of the three register columns (k = px mod 3 is a compile-time
      // constant: the segment is fully unrolled.  Unrolling by the rotation period only -- to keep the 2.5 K-instruction
      // body inside the instruction cache, ncu shows 21 % "no instruction" stalls -- measured slower: 285 vs 269 us)
      auto pixel = [&](int px, int k) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) col[(k + 2) % 3][ky] = ldx(px + 2, ky);
        const float4 d4 = *reinterpret_cast<const float4*>(dp + px * COUT);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 xv = col[(k + kx) % 3][ky];
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[ky * 3 + kx][i][j] = fmaf(xs[i], d[j], acc[ky * 3 + kx][i][j]);
          }
      };
#pragma unroll
      for (int px = 0; px < C::SEG; ++px) pixel(px, px % 3);
    }
    __syncthreads();                                         // this buffer is refilled two tiles from now
  }
  // ---- the CTA's pixel groups meet in shared memory (one copy of the outputs per group: plain stores, no atomics), then
  // one vector reduction per four outputs into the global gradient ----
  float* red = wg_sm + grp * C::OUT_FLOATS;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(red + (t * CIN + ci4 * 4 + i) * COUT + co4 * 4) =
          make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
  __syncthreads();
  for (int i = tid; i < C::OUT_FLOATS / 4; i += C::THREADS) {
    const int co = (i % C::CO_B) * 4, ci = (i / C::CO_B) % CIN, t = i / (C::CO_B * CIN);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < C::G; ++g) {
      const float4 q = *reinterpret_cast<const float4*>(wg_sm + g * C::OUT_FLOATS + i * 4);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    float* dst = gw + ((size_t)t * 32 + ci) * COUT + co;     // GEMM layout of the engine: 32 channel slots per tap
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

template <int CIN, int COUT, int W>
static int launch_wgrad_cc(const float* x, const float* dy, float* gw, int B, int relu, cudaStream_t s) {
  using C = WgCfg<CIN, COUT, W>;
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_cc_kernel<CIN, COUT, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return TPP_ENOTSUP;
    attr_set = true;
  }
  const int ntiles = B * (W / C::R);
  const int grid = ntiles < sms ? ntiles : sms;
  conv3x3_wgrad_cc_kernel<CIN, COUT, W><<<grid, C::THREADS, C::SMEM, s>>>(x, dy, gw, B, relu);
  TPP_LAUNCH_STATUS();
}

// ------------------------------------------------------------------------------------------------------------------
// First convolution of the network (3 -> 16 channels on the 64 x 64 observation): the input is the channel-PLANAR float
// observation [B][3][H][W] (element strides sb / sc / sh, unit stride along x), the output has 27 x 16 numbers.  Same
// scheme as above with the roles re-balanced: a thread owns all nine taps x 3 input channels x 4 output channels (108
// accumulators), a pixel group is 4 threads, the CTA's 64 groups take the 64 (row, 8-pixel segment) units of an
// 8 x 64 tile.  gw layout = the engine's explicit form: gw[(ky*3 + kx)*3 + ci][co].  (The tensor-core path needed a
// materialised col matrix and ran this contraction -- M = 27 -- at 911 us.)
// ------------------------------------------------------------------------------------------------------------------
template <int COUT, int W>
struct WgFirstCfg {
  static constexpr int CIN = 3, R = 8, SEG = 8;
  static constexpr int CO_B = COUT / 4;
  static constexpr int THREADS = 256;
  static constexpr int G = THREADS / CO_B;
  static constexpr int UNITS = R * (W / SEG);
  static constexpr int XP = (W + 2) | 1;                   // halo row pitch in floats (odd)
  static constexpr int DP = W | 1;                         // dY row pitch in pixels (odd)
  static constexpr int X_FLOATS = ((CIN * (R + 2) * XP + 3) / 4) * 4;
  static constexpr int D_FLOATS = R * DP * COUT;
  static constexpr int BUF_FLOATS = X_FLOATS + D_FLOATS;
  static constexpr int OUT_FLOATS = 9 * CIN * COUT;
  static constexpr size_t SMEM = sizeof(float) * (2 * BUF_FLOATS > G * OUT_FLOATS ? 2 * BUF_FLOATS : G * OUT_FLOATS);
  static_assert(UNITS % G == 0, "tile does not divide");
};

template <int COUT, int W>
__global__ void __launch_bounds__(256, 1) conv3x3_wgrad_first_kernel(const float* __restrict__ x, long long sb, long long sc,
                                                                     long long sh, const float* __restrict__ dy,
                                                                     float* __restrict__ gw, int B) {
  using C = WgFirstCfg<COUT, W>;
  constexpr int H = W, R = C::R, CIN = 3;
  extern __shared__ __align__(16) float wg_sm[];
  const int tid = threadIdx.x;
  const int co4 = tid % C::CO_B, grp = tid / C::CO_B;
  const int tiles_per_img = H / R;
  const int ntiles = B * tiles_per_img;

  auto stage = [&](int tile, float* buf) {
    const int b = tile / tiles_per_img, y0 = (tile % tiles_per_img) * R;
    const float* xb = x + (size_t)b * sb;
    constexpr int XE = CIN * (R + 2) * (W + 2);              // halo elements (4-byte copies: planar rows are not 16-byte
    for (int i = tid; i < XE; i += C::THREADS) {             // aligned behind the left halo column)
      const int hx = i % (W + 2), hr = (i / (W + 2)) % (R + 2), c = i / ((W + 2) * (R + 2));
      const int gy = y0 - 1 + hr, gx = hx - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* src = ok ? xb + c * sc + gy * sh + gx : x;
      const int n = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(buf + (c * (R + 2) + hr) * C::XP + hx)), "l"(src), "r"(n) : "memory");
    }
    float* dbuf = buf + C::X_FLOATS;
    const float* db = dy + ((size_t)b * H + y0) * W * COUT;
    constexpr int DC = R * W * C::CO_B;
    for (int i = tid; i < DC; i += C::THREADS) {
      const int c = i % C::CO_B, px = (i / C::CO_B) % W, r = i / (C::CO_B * W);
      cp_async16_zfill(dbuf + (r * C::DP + px) * COUT + c * 4, db + ((size_t)r * W + px) * COUT + c * 4, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[9][CIN][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < CIN; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.0f;

  int it = 0;
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, wg_sm);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    float* buf = wg_sm + (it & 1) * C::BUF_FLOATS;
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      stage(next, wg_sm + ((it + 1) & 1) * C::BUF_FLOATS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* dbuf = buf + C::X_FLOATS;
#pragma unroll 1
    for (int u = grp; u < C::UNITS; u += C::G) {
      const int r = u % R, x0 = (u / R) * C::SEG;            // a warp's 8 groups: the 8 rows of one segment (odd pitch apart)
      const float* xp = buf + r * C::XP + x0;                // plane 0, halo (r, x0)
      const float* dp = dbuf + (r * C::DP + x0) * COUT + co4 * 4;
      float col[3][3][CIN];                                  // [halo column x0 + k][ky][ci]
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int c = 0; c < CIN; ++c) col[k][ky][c] = xp[(c * (R + 2) + ky) * C::XP + k];
      auto pixel = [&](int px, int k) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int c = 0; c < CIN; ++c) col[(k + 2) % 3][ky][c] = xp[(c * (R + 2) + ky) * C::XP + px + 2];
        const float4 d4 = *reinterpret_cast<const float4*>(dp + px * COUT);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int c = 0; c < CIN; ++c)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                acc[ky * 3 + kx][c][j] = fmaf(col[(k + kx) % 3][ky][c], d[j], acc[ky * 3 + kx][c][j]);
      };
#pragma unroll
      for (int px = 0; px < C::SEG; ++px) pixel(px, px % 3);
    }
    __syncthreads();
  }
  float* red = wg_sm + grp * C::OUT_FLOATS;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CIN; ++c)
      *reinterpret_cast<float4*>(red + (t * CIN + c) * COUT + co4 * 4) =
          make_float4(acc[t][c][0], acc[t][c][1], acc[t][c][2], acc[t][c][3]);
  __syncthreads();
  for (int i = tid; i < C::OUT_FLOATS / 4; i += C::THREADS) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int g = 0; g < C::G; ++g) {
      const float4 q = *reinterpret_cast<const float4*>(wg_sm + g * C::OUT_FLOATS + i * 4);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    float* dst = gw + (size_t)i * 4;                         // [(tap*3 + ci)][co], rows contiguous (ldc == COUT)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// First convolution, forward: out[b][y][x][co] = bias[co] + sum_{ky,kx,ci} X[b][ci][y+ky-1][x+kx-1] * W[co][ci][ky][kx]
// from the channel-planar observation to the NHWC activation (plain fp32: its consumer is the max-pool).  K = 27 is far
// too short for a tensor-core tile (the GEMM form needed a materialised 32-wide col matrix: im2col 453 us + GEMM 465 us
// for 2048 frames); here a CTA of 128 threads stages the 10 x 66 x 3 halo of an 8-row tile and the 432 weights, and a
// thread computes 8 pixels x 8 channels (64 accumulators, 90 scalar + 54 vector shared-memory loads per 1728 FFMA).
// Small CTAs (10 KB, ~100 registers): several tiles per SM overlap their staging with each other's arithmetic.
// ------------------------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(128) conv3x3_fwd_first_kernel(const float* __restrict__ x, long long sb, long long sc,
                                                                long long sh, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ out, int B) {
  constexpr int H = W, R = 8, CIN = 3, COUT = 16, XP = (W + 2) | 1, SEG = 8;
  static_assert(R * (W / SEG) * 2 == 128, "one (row, segment, channel half) per thread");
  __shared__ __align__(16) float xs[CIN * (R + 2) * XP];
  __shared__ __align__(16) float ws[27 * COUT];
  const int tid = threadIdx.x;
  const int tiles_per_img = H / R;
  const int b = blockIdx.x / tiles_per_img, y0 = (blockIdx.x % tiles_per_img) * R;
  const float* xb = x + (size_t)b * sb;
  for (int i = tid; i < CIN * (R + 2) * (W + 2); i += 128) {
    const int hx = i % (W + 2), hr = (i / (W + 2)) % (R + 2), c = i / ((W + 2) * (R + 2));
    const int gy = y0 - 1 + hr, gx = hx - 1;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const float* src = ok ? xb + c * sc + gy * sh + gx : x;
    const int n = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(xs + (c * (R + 2) + hr) * XP + hx)), "l"(src), "r"(n) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < 27 * COUT; i += 128) {              // ws[(ky*3 + kx)*3 + ci][co] <- W[co][ci][ky][kx]
    const int co = i % COUT, k = i / COUT, ci = k % 3, tap = k / 3;
    ws[i] = __ldg(w + (co * 3 + ci) * 9 + tap);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int half = tid & 1, u = tid >> 1;                    // a warp: 2 segments x 8 rows (odd pitch: distinct banks)
  const int r = u % R, x0 = (u / R) * SEG;
  float acc[SEG][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + half * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + half * 8 + 4));
#pragma unroll
    for (int p = 0; p < SEG; ++p) {
      acc[p][0] = b0.x; acc[p][1] = b0.y; acc[p][2] = b0.z; acc[p][3] = b0.w;
      acc[p][4] = b1.x; acc[p][5] = b1.y; acc[p][6] = b1.z; acc[p][7] = b1.w;
    }
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float xr[SEG + 2];
      const float* xp = xs + (ci * (R + 2) + r + ky) * XP + x0;
#pragma unroll
      for (int k = 0; k < SEG + 2; ++k) xr[k] = xp[k];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wp = ws + ((ky * 3 + kx) * 3 + ci) * COUT + half * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int p = 0; p < SEG; ++p)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(xr[p + kx], wv[j], acc[p][j]);
      }
    }
  float* op = out + (((size_t)b * H + y0 + r) * W + x0) * COUT + half * 8;
#pragma unroll
  for (int p = 0; p < SEG; ++p) {
    __stcs(reinterpret_cast<float4*>(op + p * COUT), make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
    __stcs(reinterpret_cast<float4*>(op + p * COUT + 4), make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Forward / data gradient of the 3x3 convolutions that touch a 16-channel tensor at 32 x 32 (IMPALA block 1 and the
// first convolution of block 2), on the fp32 FMA pipe:
//   y[b][y][x][co] = sum_{ky,kx,ci} X[b][y+ky-1][x+kx-1][ci] * Wg[co][(ky*3 + kx)*S + ci]      (S = channel slots per tap)
// with the epilogue of the tensor-core tiles (bias -> ReLU mask -> residual add -> column sums -> plain / TF32 pair).
// The data gradient is the same contraction on dY with the flipped weights the engine already keeps (Wd).
// Why: with 16 output (or input) channels a tcgen05 tile uses an eighth of an MMA's columns, needs three TF32 passes for
// fp32 parity and re-gathers its operand nine times from L2 (measured 335 - 350 us for 4.8 GFMA); the FMA pipe does the
// same arithmetic exactly, once, from a halo staged once.
//   * CTA = R rows x 32 pixels of one image; the halo is staged CHANNEL-CHUNK-PLANAR ([ci / 4][row][pixel] float4s, odd
//     row pitch): lanes of a warp sit on different rows of the same pixel column, i.e. 8 distinct 16-byte bank groups;
//   * a thread computes 8 pixels x 8 output channels (64 accumulators): per (ky, 4 input channels) 10 LDS.128 of X and
//     24 broadcast LDS.128 of weights for 768 FFMA.
// ------------------------------------------------------------------------------------------------------------------
template <int CI, int CO, int W, int R>
struct FmaCfg {
  static constexpr int SEG = 8, CO_Q = CO / 8, UNITS = R * (W / SEG), THREADS = UNITS * CO_Q;
  static constexpr int XP = (W + 2) | 1;                   // halo row pitch in pixels (odd)
  static constexpr int X_FLOATS = (CI / 4) * (R + 2) * XP * 4;
  static constexpr int W_FLOATS = 9 * CI * CO;
  static constexpr size_t SMEM = sizeof(float) * (X_FLOATS + W_FLOATS);
  static_assert(THREADS % 32 == 0 && THREADS <= 256, "CTA shape");
};

struct FmaEpi {
  const float* bias; const float* mask; const float* addend;
  float* out; float* out_hi; float* out_lo; float* colsum;
  int relu_in, pair_relu;
};

template <int CI, int CO, int W, int R>
__global__ void __launch_bounds__(FmaCfg<CI, CO, W, R>::THREADS) conv3x3_fma_kernel(const float* __restrict__ x,
                                                                                 const float* __restrict__ wg, int slots,
                                                                                 FmaEpi e, int B) {
  using C = FmaCfg<CI, CO, W, R>;
  constexpr int H = W, SEG = C::SEG, CI4 = CI / 4;
  extern __shared__ __align__(16) float fm_sm[];
  __shared__ float cs_sh[C::THREADS / 32][CO];
  float* xs = fm_sm;                                         // [CI4][R + 2][XP] float4
  float* ws = fm_sm + C::X_FLOATS;                           // [(tap * CI + ci)][CO]
  const int tid = threadIdx.x;
  const int tiles_per_img = H / R;
  const int b = blockIdx.x / tiles_per_img, y0 = (blockIdx.x % tiles_per_img) * R;
  const float* xb = x + (size_t)b * H * W * CI;
  for (int i = tid; i < CI4 * (R + 2) * (W + 2); i += C::THREADS) {
    const int c = i % CI4, hx = (i / CI4) % (W + 2), hr = i / (CI4 * (W + 2));
    const int gy = y0 - 1 + hr, gx = hx - 1;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    cp_async16_zfill(xs + ((c * (R + 2) + hr) * C::XP + hx) * 4, ok ? xb + ((size_t)gy * W + gx) * CI + c * 4 : x, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < C::W_FLOATS; i += C::THREADS) {      // ws[tap][ci][co] <- Wg[co][tap * slots + ci]
    const int ci = i % CI, tap = (i / CI) % 9, co = i / (CI * 9);
    ws[(tap * CI + ci) * CO + co] = __ldg(wg + (size_t)co * 9 * slots + tap * slots + ci);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int q = tid % C::CO_Q, u = tid / C::CO_Q;
  const int r = u % R, x0 = (u / R) * SEG;                   // a warp's units: rows of one pixel column (odd pitch apart)
  float acc[SEG][8];
#pragma unroll
  for (int p = 0; p < SEG; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.0f;
#pragma unroll 1                                             // (one (ky, 4 channels) body = ~800 instructions: stays in the
  for (int ky = 0; ky < 3; ++ky)                             // instruction cache)
#pragma unroll 1
    for (int c = 0; c < CI4; ++c) {
      float4 xr[SEG + 2];
      const float* xp = xs + ((c * (R + 2) + r + ky) * C::XP + x0) * 4;
#pragma unroll
      for (int k = 0; k < SEG + 2; ++k) {
        float4 v = *reinterpret_cast<const float4*>(xp + k * 4);
        if (e.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        xr[k] = v;
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* wp = ws + ((ky * 3 + kx) * CI + c * 4 + i) * CO + q * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < SEG; ++p) {
            const float4 xq = xr[p + kx];
            const float xv = i == 0 ? xq.x : (i == 1 ? xq.y : (i == 2 ? xq.z : xq.w));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(xv, wv[j], acc[p][j]);
          }
        }
    }
  // ---- epilogue: bias -> mask -> residual -> column sums -> plain / pair (the order of csrc/gemm_tc.cu) ----
  float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = __ldg(e.bias + q * 8 + j);
  }
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t row0 = ((size_t)b * H + y0 + r) * W + x0;
#pragma unroll
  for (int p = 0; p < SEG; ++p) {
    const size_t off = (row0 + p) * CO + q * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[p][j] + bv[j];
    if (e.mask) {
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(e.mask + off)), m1 = __ldg(reinterpret_cast<const float4*>(e.mask + off + 4));
      const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = mv[j] > 0.0f ? v[j] : 0.0f;
    }
    if (e.addend) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(e.addend + off)), a1 = __ldg(reinterpret_cast<const float4*>(e.addend + off + 4));
      v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w; v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] += v[j];
    if (e.out) {
      *reinterpret_cast<float4*>(e.out + off) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(e.out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (e.out_hi) {
      float h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = e.pair_relu ? fmaxf(v[j], 0.0f) : v[j];
        h[j] = __uint_as_float((__float_as_uint(y) + 0x1000u) & 0xFFFFE000u);      // tf32_round of csrc/gemm_tc.cu
        l[j] = y - h[j];
      }
      *reinterpret_cast<float4*>(e.out_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(e.out_hi + off + 4) = make_float4(h[4], h[5], h[6], h[7]);
      *reinterpret_cast<float4*>(e.out_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
      *reinterpret_cast<float4*>(e.out_lo + off + 4) = make_float4(l[4], l[5], l[6], l[7]);
    }
  }
  if (e.colsum) {                                            // lanes with the same channel octet: tid = q (mod CO_Q)
#pragma unroll
    for (int o = C::CO_Q; o < 32; o <<= 1)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane < C::CO_Q) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cs_sh[warp][lane * 8 + j] = cs[j];
    }
    __syncthreads();
    if (tid < CO) {
      float t = 0.0f;
#pragma unroll
      for (int w2 = 0; w2 < C::THREADS / 32; ++w2) t += cs_sh[w2][tid];
      atomicAdd(e.colsum + tid, t);
    }
  }
}

template <int CI, int CO, int W, int R>
static int launch_fma(const float* x, const float* wg, int slots, const FmaEpi& e, int B, cudaStream_t s) {
  using C = FmaCfg<CI, CO, W, R>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(conv3x3_fma_kernel<CI, CO, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C::SMEM);
    if (err != cudaSuccess) return (int)err;
    attr_set = true;
  }
  conv3x3_fma_kernel<CI, CO, W, R><<<B * (W / R), C::THREADS, C::SMEM, s>>>(x, wg, slots, e, B);
  TPP_LAUNCH_STATUS();
}

}  // namespace tpp

extern "C" int tpp_conv3x3_wgrad(const float* x, int32_t relu, const float* dy, float* gw, int32_t B, int32_t H, int32_t W,
                                 int32_t cin, int32_t cout, void* stream) {
  TPP_CHECK_ARG(x && dy && gw && B > 0);
  if (H != W) return TPP_ENOTSUP;
  cudaStream_t s = tpp_stream(stream);
  if (cin == 16 && cout == 16 && W == 32) return tpp::launch_wgrad_cc<16, 16, 32>(x, dy, gw, B, relu, s);
  if (cin == 16 && cout == 32 && W == 32) return tpp::launch_wgrad_cc<16, 32, 32>(x, dy, gw, B, relu, s);
  if (cin == 32 && cout == 32 && W == 16) return tpp::launch_wgrad_cc<32, 32, 16>(x, dy, gw, B, relu, s);
  if (cin == 32 && cout == 32 && W == 8) return tpp::launch_wgrad_cc<32, 32, 8>(x, dy, gw, B, relu, s);
  return TPP_ENOTSUP;
}

extern "C" int tpp_conv3x3_wgrad_first(const float* x, int64_t sb, int64_t sc, int64_t sh, const float* dy, float* gw,
                                       int32_t B, int32_t H, int32_t W, int32_t cout, void* stream) {
  TPP_CHECK_ARG(x && dy && gw && B > 0);
  if (H != 64 || W != 64 || cout != 16) return TPP_ENOTSUP;
  using C = tpp::WgFirstCfg<16, 64>;
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tpp::conv3x3_wgrad_first_kernel<16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return TPP_ENOTSUP;
    attr_set = true;
  }
  const int ntiles = B * (H / C::R);
  const int grid = ntiles < sms ? ntiles : sms;
  tpp::conv3x3_wgrad_first_kernel<16, 64><<<grid, C::THREADS, C::SMEM, tpp_stream(stream)>>>(x, sb, sc, sh, dy, gw, B);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_conv3x3_fwd_first(const float* x, int64_t sb, int64_t sc, int64_t sh, const float* w, const float* bias,
                                     float* out, int32_t B, int32_t H, int32_t W, int32_t cout, void* stream) {
  TPP_CHECK_ARG(x && w && bias && out && B > 0);
  if (H != 64 || W != 64 || cout != 16) return TPP_ENOTSUP;
  tpp::conv3x3_fwd_first_kernel<64><<<B * (64 / 8), 128, 0, tpp_stream(stream)>>>(x, sb, sc, sh, w, bias, out, B);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_conv3x3_fma(const float* x, int32_t relu_in, const float* wg, int32_t slots, const float* bias,
                               const float* mask, const float* addend, int32_t pair_relu, float* out, float* out_hi,
                               float* out_lo, float* colsum, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout,
                               void* stream) {
  TPP_CHECK_ARG(x && wg && B > 0 && (out || out_hi) && (!out_hi == !out_lo) && slots >= cin);
  if (H != W) return TPP_ENOTSUP;
  tpp::FmaEpi e{bias, mask, addend, out, out_hi, out_lo, colsum, relu_in, pair_relu};
  cudaStream_t s = tpp_stream(stream);
  if (cin == 16 && cout == 16 && W == 32) return tpp::launch_fma<16, 16, 32, 16>(x, wg, slots, e, B, s);
  if (cin == 16 && cout == 32 && W == 32) return tpp::launch_fma<16, 32, 32, 8>(x, wg, slots, e, B, s);
  if (cin == 32 && cout == 16 && W == 32) return tpp::launch_fma<32, 16, 32, 16>(x, wg, slots, e, B, s);
  return TPP_ENOTSUP;
}
