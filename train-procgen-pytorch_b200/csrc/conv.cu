// IMPALA-CNN building blocks in NHWC (reference: common/model.py:134-208): the 3x3 / pad-1 convolutions are run as
// GEMMs on the tcgen05 kernel (csrc/gemm_tc.cu) over an explicitly materialised im2col matrix,
//     col[p][tap*C + c] = act(x[b, y + ky - 1, x + kx - 1, c]),   p = (b*H + y)*W + x,  tap = ky*3 + kx,
// written directly as the TF32 (hi, lo) operand pair; ReLU-on-load and the uint8 -> [0,1] frame scaling are fused
// into this gather.  forward: Y = col(X) Wf^T; data gradient: dX = col(dY) Wd^T with flipped taps; weight gradient:
// dWf = dY^T col(X) (both operands MN-major).  Max-pooling 3x3 / stride 2 / pad 1 keeps the argmax tap for the
// backward pass, which is written in gather form (no atomics, deterministic).
// Round-1 note: materialising col costs 9x the activation bytes; the TMA-im2col (implicit GEMM) variant is next.
#include "tpp_common.cuh"

namespace tpp {

__device__ __forceinline__ float cv_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// One thread per 16-byte chunk of a col row (4 consecutive k): a warp writes 512 contiguous bytes of hi and of lo.
// C % 4 == 0 and unit channel stride: the 4 values are one float4 of the source pixel.
__global__ void __launch_bounds__(256) im2col3x3_vec_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                            long long sb, long long sy_, long long sx_, int relu,
                                                            float scale, float* __restrict__ hi,
                                                            float* __restrict__ lo, int Kp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int kq = Kp >> 2;
  const long long M = (long long)B * H * W;
  if (t >= M * kq) return;
  const long long p = t / kq;
  const int k = (int)(t - p * kq) * 4;
  const int tap = k / C, c = k - tap * C;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tap < 9) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W)
      v = __ldg(reinterpret_cast<const float4*>(x + b * sb + sy * sy_ + sx * sx_ + c));
  }
  float a[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  float h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (relu) a[j] = fmaxf(a[j], 0.0f);
    h[j] = cv_tf32(a[j]);
  }
  *reinterpret_cast<float4*>(hi + p * Kp + k) = make_float4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<float4*>(lo + p * Kp + k) = make_float4(a[0] - h[0], a[1] - h[1], a[2] - h[2], a[3] - h[3]);
}

// Few input channels (9C <= 32 = Kp: the first convolution reading NCHW observation rows or uint8 frames): a warp
// owns 32 consecutive pixels; lane = pixel while reading (consecutive x: coalesced for NCHW planes), lane = k while
// writing (one 128-byte col row per instruction), transposed through a per-warp shared-memory tile.
template <typename TIn>
__global__ void __launch_bounds__(256) im2col3x3_small_kernel(const TIn* __restrict__ x, int B, int H, int W, int C,
                                                              long long sb, long long sy_, long long sx_, long long sc,
                                                              int relu, float scale, float* __restrict__ hi,
                                                              float* __restrict__ lo) {
  __shared__ float tile[8][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long M = (long long)B * H * W;
  const long long p0 = ((long long)blockIdx.x * 8 + w) * 32;
  if (p0 >= M) return;
  const long long p = p0 + lane;
  const bool live = p < M;
  const int xx = (int)(p % W), yy = (int)((p / W) % H);
  const long long b = p / ((long long)W * H);
  const TIn* base = x + b * sb;
  int k = 0;
  for (int tap = 0; tap < 9; ++tap) {
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    const bool inside = live && sy >= 0 && sy < H && sx >= 0 && sx < W;
    const TIn* src = base + sy * sy_ + sx * sx_;
    for (int c = 0; c < C; ++c, ++k) {
      float q = inside ? (float)src[c * sc] * scale : 0.0f;
      if (relu) q = fmaxf(q, 0.0f);
      tile[w][lane][k] = q;
    }
  }
  for (; k < 32; ++k) tile[w][lane][k] = 0.0f;
  __syncwarp();
  const int n = (int)(M - p0 < 32 ? M - p0 : 32);
  for (int i = 0; i < n; ++i) {
    const float q = tile[w][i][lane];
    const float h = cv_tf32(q);
    hi[(p0 + i) * 32 + lane] = h;
    if (lo) lo[(p0 + i) * 32 + lane] = q - h;
  }
}

// Generic layout: one thread per col element, k fastest -> coalesced stores.
template <typename TIn>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const TIn* __restrict__ x, int B, int H, int W, int C,
                                                        long long sb, long long sy_, long long sx_, long long sc,
                                                        int relu, float scale, float* __restrict__ hi,
                                                        float* __restrict__ lo, int Kp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)B * H * W;
  if (t >= M * Kp) return;
  const long long p = t / Kp;
  const int k = (int)(t - p * Kp);
  const int tap = k / C, c = k - tap * C;
  float q = 0.0f;
  if (tap < 9) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) q = (float)x[b * sb + sy * sy_ + sx * sx_ + c * sc] * scale;
  }
  if (relu) q = fmaxf(q, 0.0f);
  const float h = cv_tf32(q);
  hi[t] = h;
  if (lo) lo[t] = q - h;
}

// NHWC max-pool 3x3, stride 2, pad 1: y [B, Ho, Wo, C], arg = winning tap 0..8 (first maximum, like torch).
// One thread per output pixel and 4 channels (C % 4 == 0): float4 loads, uchar4 / float4 stores.
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                          float* __restrict__ y, uint8_t* __restrict__ arg,
                                                          float* __restrict__ r_hi, float* __restrict__ r_lo, int Ho,
                                                          int Wo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = C >> 2;
  const long long total = (long long)B * Ho * Wo * cq;
  if (t >= total) return;
  const int c = (int)(t % cq) * 4;
  const int ox = (int)((t / cq) % Wo), oy = (int)((t / ((long long)cq * Wo)) % Ho);
  const long long b = t / ((long long)cq * Wo * Ho);
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  int bi[4] = {0, 0, 0, 0};
  bool any = false;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int sy = oy * 2 + tap / 3 - 1, sx = ox * 2 + tap % 3 - 1;
    if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + ((b * H + sy) * W + sx) * C + c));
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (!any || v[j] > best[j]) { best[j] = v[j]; bi[j] = tap; }
    any = true;
  }
  const long long o = t * 4;
  *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1], best[2], best[3]);
  *reinterpret_cast<uchar4*>(arg + o) = make_uchar4((uint8_t)bi[0], (uint8_t)bi[1], (uint8_t)bi[2], (uint8_t)bi[3]);
  if (r_hi) {   // TF32 pair of relu(y): the operand of the residual block's first convolution
    float q[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = fmaxf(best[j], 0.0f);
      h[j] = cv_tf32(q[j]);
    }
    *reinterpret_cast<float4*>(r_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(r_lo + o) = make_float4(q[0] - h[0], q[1] - h[1], q[2] - h[2], q[3] - h[3]);
  }
}

// dx[b, y, x, c] = sum over the (<= 4) pooling windows that contain (y, x) and whose argmax is this pixel.
// One thread per pixel and 4 channels (C % 4 == 0): float4 / uchar4 accesses.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg,
                                                          int B, int H, int W, int C, int Ho, int Wo,
                                                          float* __restrict__ dx, float* __restrict__ dx_hi,
                                                          float* __restrict__ dx_lo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = C >> 2;
  const long long total = (long long)B * H * W * cq;
  if (t >= total) return;
  const int c = (int)(t % cq) * 4;
  const int xx = (int)((t / cq) % W), yy = (int)((t / ((long long)cq * W)) % H);
  const long long b = t / ((long long)cq * W * H);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  // windows (oy, ox) with oy*2 - 1 <= yy <= oy*2 + 1
  for (int oy = yy / 2; oy <= (yy + 1) / 2; ++oy) {
    if (oy >= Ho) continue;
    const int ky = yy - (oy * 2 - 1);
    for (int ox = xx / 2; ox <= (xx + 1) / 2; ++ox) {
      if (ox >= Wo) continue;
      const int kx = xx - (ox * 2 - 1);
      const long long o = ((b * Ho + oy) * Wo + ox) * C + c;
      const uchar4 a = *reinterpret_cast<const uchar4*>(arg + o);
      const int want = ky * 3 + kx;
      if (a.x == want || a.y == want || a.z == want || a.w == want) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + o));
        if (a.x == want) s[0] += g.x;
        if (a.y == want) s[1] += g.y;
        if (a.z == want) s[2] += g.z;
        if (a.w == want) s[3] += g.w;
      }
    }
  }
  const long long o = t * 4;
  *reinterpret_cast<float4*>(dx + o) = make_float4(s[0], s[1], s[2], s[3]);
  if (dx_hi) {
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = cv_tf32(s[j]);
    *reinterpret_cast<float4*>(dx_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(dx_lo + o) = make_float4(s[0] - h[0], s[1] - h[1], s[2] - h[2], s[3] - h[3]);
  }
}

// The same for even H, W: one thread per 2 x 2 block of input pixels and 4 channels.  The block (by, bx) is touched by the
// four windows (by .. by + 1) x (bx .. bx + 1) only, so their (arg, dy) vectors are loaded once -- 4 loads for 4 output
// pixels instead of 1 + 2 + 2 + 4 -- and four times fewer threads walk the index arithmetic.  Sums run in the per-pixel
// kernel's window order (bit-identical results).
__global__ void __launch_bounds__(256) maxpool_bwd_quad_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg,
                                                               int B, int H, int W, int C, int Ho, int Wo,
                                                               float* __restrict__ dx, float* __restrict__ dx_hi,
                                                               float* __restrict__ dx_lo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = C >> 2, Wb = W >> 1, Hb = H >> 1;
  const long long total = (long long)B * Hb * Wb * cq;
  if (t >= total) return;
  const int c = (int)(t % cq) * 4;
  const int bx = (int)((t / cq) % Wb), by = (int)((t / ((long long)cq * Wb)) % Hb);
  const long long b = t / ((long long)cq * Wb * Hb);
  uchar4 a[2][2];
  float4 g[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int oy = by + i, ox = bx + j;
      a[i][j] = make_uchar4(255, 255, 255, 255);
      g[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oy < Ho && ox < Wo) {
        const long long o = ((b * Ho + oy) * Wo + ox) * C + c;
        a[i][j] = *reinterpret_cast<const uchar4*>(arg + o);
        g[i][j] = __ldg(reinterpret_cast<const float4*>(dy + o));
      }
    }
  // input pixel (2 by + py, 2 bx + px) lies in window (by + i, bx + j) at tap ((py + 1 - 2 i) * 3 + (px + 1 - 2 j)), for the
  // (i, j) with that tap inside 0 .. 2
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i <= py; ++i)
#pragma unroll
        for (int j = 0; j <= px; ++j) {
          const int want = (py + 1 - 2 * i) * 3 + (px + 1 - 2 * j);
          if (a[i][j].x == want) s[0] += g[i][j].x;
          if (a[i][j].y == want) s[1] += g[i][j].y;
          if (a[i][j].z == want) s[2] += g[i][j].z;
          if (a[i][j].w == want) s[3] += g[i][j].w;
        }
      const long long o = ((b * H + 2 * by + py) * W + 2 * bx + px) * C + c;
      *reinterpret_cast<float4*>(dx + o) = make_float4(s[0], s[1], s[2], s[3]);
      if (dx_hi) {
        float h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = cv_tf32(s[k]);
        *reinterpret_cast<float4*>(dx_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dx_lo + o) = make_float4(s[0] - h[0], s[1] - h[1], s[2] - h[2], s[3] - h[3]);
      }
    }
}

// Column sums of a narrow row-major matrix [M][C], C in {4, 8, 16, 32, 64} (bias gradient of a convolution from its
// NHWC output gradient): the matrix is read as one flat float4 stream, a thread's 4 columns never change.
__global__ void __launch_bounds__(256) colsum_narrow_kernel(const float* __restrict__ x, long long total4, int C,
                                                            float* __restrict__ out) {
  __shared__ float part[256][4];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) part[threadIdx.x][j] = a[j];
  __syncthreads();
  // threads t, t + C/4, t + 2C/4, ... own the same columns (stride and blockDim are multiples of C/4)
  const int groups = C / 4;
  if ((int)threadIdx.x < groups) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = threadIdx.x; t < 256; t += groups)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += part[t][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(out + threadIdx.x * 4 + j, s[j]);
  }
}

// y = act(x + bias) as plain fp32 and/or TF32 pair: finishes a split-K (atomically accumulated) dense layer.
__global__ void __launch_bounds__(256) bias_act_split_kernel(const float* __restrict__ x, long long ld_in, int M, int N,
                                                             const float* __restrict__ bias, int relu,
                                                             float* __restrict__ out, float* __restrict__ hi,
                                                             float* __restrict__ lo, long long ld_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * N) return;
  const int n = (int)(t % N);
  const long long m = t / N;
  float v = x[m * ld_in + n] + (bias ? bias[n] : 0.0f);
  if (relu) v = fmaxf(v, 0.0f);
  if (out) out[m * ld_out + n] = v;
  if (hi) {
    const float h = cv_tf32(v);
    hi[m * ld_out + n] = h;
    lo[m * ld_out + n] = v - h;
  }
}

// ------------------------------------------------------------------------------------------------
// IMPALA feature sparsity (common/model.py:203-208): fs = mean_j max_b tanh(|100 h_bj|) over the flattened, ReLU'd
// block-3 features h [M][E] (h >= 0, so |.| is the identity and fp32 bit patterns order like the values).
// Forward: per column the 64-bit key (bits(h) << 32 | ~row) is max-reduced with atomicMax: the largest value and, on
// ties, the smallest row -- torch.max(dim=0)'s choice.  Backward: d(coef*fs)/dh[b*, j] = coef/E * 100 * (1 - tanh^2)
// for the argmax row of every column with h > 0 (ReLU and abs both have zero derivative at 0), added to the
// gradient of the features and re-split into its TF32 pair.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fs_colmax_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                        int M, int E, int rows_per_cta,
                                                        unsigned long long* __restrict__ key) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  unsigned long long best = 0ull;
  for (int r = r0; r < r1; ++r) {
    const int64_t o = (int64_t)r * E + j;
    const float v = hi[o] + (lo ? lo[o] : 0.0f);
    const unsigned long long k = ((unsigned long long)__float_as_uint(fmaxf(v, 0.0f)) << 32) |
                                 (unsigned long long)(0xFFFFFFFFu - (unsigned)r);
    best = k > best ? k : best;
  }
  atomicMax(key + j, best);
}

__global__ void __launch_bounds__(1024) fs_value_kernel(const unsigned long long* __restrict__ key, int E,
                                                        float* __restrict__ fs_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int j = threadIdx.x; j < E; j += blockDim.x)
    s += (double)tanhf(100.0f * __uint_as_float((unsigned)(key[j] >> 32)));
  s = block_sum(s, red);
  if (threadIdx.x == 0) fs_out[0] = (float)(s / (double)E);
}

__global__ void __launch_bounds__(256) fs_grad_kernel(const unsigned long long* __restrict__ key, int E, float coef,
                                                      float* __restrict__ dx, float* __restrict__ dx_hi,
                                                      float* __restrict__ dx_lo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E) return;
  const unsigned long long k = key[j];
  const float h = __uint_as_float((unsigned)(k >> 32));
  if (!(h > 0.0f)) return;
  const unsigned row = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
  const float th = tanhf(100.0f * h);
  const float g = coef / (float)E * 100.0f * (1.0f - th * th);
  const int64_t o = (int64_t)row * E + j;
  const float v = dx[o] + g;
  dx[o] = v;
  if (dx_hi) {
    const float vh = cv_tf32(v);
    dx_hi[o] = vh;
    dx_lo[o] = v - vh;
  }
}

}  // namespace tpp

extern "C" int tpp_bias_act_split(const float* x, int64_t ld_in, int32_t M, int32_t N, const float* bias, int32_t relu,
                                  float* out, float* out_hi, float* out_lo, int64_t ld_out, void* stream) {
  TPP_CHECK_ARG(x && M > 0 && N > 0 && ld_in >= N && ld_out >= N && (out || out_hi));
  TPP_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr));
  tpp::bias_act_split_kernel<<<tpp_ceil_div((long long)M * N, 256), 256, 0, tpp_stream(stream)>>>(
      x, ld_in, M, N, bias, relu, out, out_hi, out_lo, ld_out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_im2col3x3(const void* x, int32_t x_is_u8, int32_t B, int32_t H, int32_t W, int32_t C, int64_t sb,
                             int64_t sy, int64_t sx, int64_t sc, int32_t relu, float scale, float* col_hi,
                             float* col_lo, int32_t Kp, void* stream) {
  TPP_CHECK_ARG(x && col_hi && B > 0 && H > 0 && W > 0 && C > 0 && Kp >= 9 * C && (Kp & 3) == 0);
  TPP_CHECK_ARG((reinterpret_cast<uintptr_t>(col_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(col_lo) & 15) == 0);
  const long long rows = (long long)B * H * W;
  cudaStream_t s = tpp_stream(stream);
  if (Kp == 32 && 9 * C <= 32) {
    const int grid = tpp_ceil_div(rows, 256);
    if (x_is_u8)
      tpp::im2col3x3_small_kernel<uint8_t><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(x), B, H, W, C, sb, sy,
                                                               sx, sc, relu, scale, col_hi, col_lo);
    else
      tpp::im2col3x3_small_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), B, H, W, C, sb, sy, sx,
                                                             sc, relu, scale, col_hi, col_lo);
  } else if (x_is_u8) {
    tpp::im2col3x3_kernel<uint8_t><<<tpp_ceil_div(rows * Kp, 256), 256, 0, s>>>(
        reinterpret_cast<const uint8_t*>(x), B, H, W, C, sb, sy, sx, sc, relu, scale, col_hi, col_lo, Kp);
  } else if ((C & 3) == 0 && sc == 1 && ((sb | sy | sx) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    tpp::im2col3x3_vec_kernel<<<tpp_ceil_div(rows * (Kp / 4), 256), 256, 0, s>>>(
        reinterpret_cast<const float*>(x), B, H, W, C, sb, sy, sx, relu, scale, col_hi, col_lo, Kp);
  } else {
    tpp::im2col3x3_kernel<float><<<tpp_ceil_div(rows * Kp, 256), 256, 0, s>>>(
        reinterpret_cast<const float*>(x), B, H, W, C, sb, sy, sx, sc, relu, scale, col_hi, col_lo, Kp);
  }
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_maxpool3x3s2_fwd(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, float* y, uint8_t* arg,
                                    float* relu_hi, float* relu_lo, void* stream) {
  TPP_CHECK_ARG(x && y && arg && B > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0);
  TPP_CHECK_ARG((relu_hi == nullptr) == (relu_lo == nullptr));
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  tpp::maxpool_fwd_kernel<<<tpp_ceil_div(total, 256), 256, 0, tpp_stream(stream)>>>(x, B, H, W, C, y, arg, relu_hi,
                                                                                    relu_lo, Ho, Wo);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_maxpool3x3s2_bwd(const float* dy, const uint8_t* arg, int32_t B, int32_t H, int32_t W, int32_t C,
                                    float* dx, float* dx_hi, float* dx_lo, void* stream) {
  TPP_CHECK_ARG(dy && arg && dx && B > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0);
  TPP_CHECK_ARG((dx_hi == nullptr) == (dx_lo == nullptr));
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  if (((H | W) & 1) == 0) {       // even sizes (every IMPALA stage): one thread per 2 x 2 input block
    const long long quads = (long long)B * (H / 2) * (W / 2) * (C / 4);
    tpp::maxpool_bwd_quad_kernel<<<tpp_ceil_div(quads, 256), 256, 0, tpp_stream(stream)>>>(dy, arg, B, H, W, C, Ho, Wo, dx,
                                                                                           dx_hi, dx_lo);
    TPP_LAUNCH_STATUS();
  }
  const long long total = (long long)B * H * W * (C / 4);
  tpp::maxpool_bwd_kernel<<<tpp_ceil_div(total, 256), 256, 0, tpp_stream(stream)>>>(dy, arg, B, H, W, C, Ho, Wo, dx,
                                                                                    dx_hi, dx_lo);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_colsum_narrow(const float* x, int64_t M, int32_t C, float* out, void* stream) {
  TPP_CHECK_ARG(x && out && M > 0 && (C == 4 || C == 8 || C == 16 || C == 32 || C == 64));
  TPP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const long long total4 = M * C / 4;
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  tpp::colsum_narrow_kernel<<<(int)blocks, 256, 0, tpp_stream(stream)>>>(x, total4, C, out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_feature_sparsity(const float* h_hi, const float* h_lo, int32_t M, int32_t E, uint64_t* scratch,
                                    float* fs_out, void* stream) {
  TPP_CHECK_ARG(h_hi && scratch && fs_out && M > 0 && E > 0);
  cudaStream_t s = tpp_stream(stream);
  cudaMemsetAsync(scratch, 0, sizeof(uint64_t) * (size_t)E, s);
  const int rows_per_cta = 64;
  dim3 grid(tpp_ceil_div(E, 256), tpp_ceil_div(M, rows_per_cta));
  tpp::fs_colmax_kernel<<<grid, 256, 0, s>>>(h_hi, h_lo, M, E, rows_per_cta,
                                             reinterpret_cast<unsigned long long*>(scratch));
  tpp::fs_value_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const unsigned long long*>(scratch), E, fs_out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_feature_sparsity_grad(const uint64_t* scratch, int32_t E, float coef, float* dx, float* dx_hi,
                                         float* dx_lo, void* stream) {
  TPP_CHECK_ARG(scratch && dx && E > 0 && (!dx_hi == !dx_lo));
  tpp::fs_grad_kernel<<<tpp_ceil_div(E, 256), 256, 0, tpp_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(scratch), E, coef, dx, dx_hi, dx_lo);
  TPP_LAUNCH_STATUS();
}
