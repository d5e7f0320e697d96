// Fused step + truncation + auto-reset + obs/reward/done emit for the pre-vectorised env families.
//
// Replaces (reference, pure numpy): PreVecEnv.step/set/reset (discrete_env/pre_vec_env.py:78-119) and the
// transition_model of CartPoleVecEnv (cartpole_pre_vec.py:210-256), CartPoleSwingVecEnv
// (cartpole_swing_pre_vec.py:195-239), MountainCarVecEnv (mountain_car_pre_vec.py:194-209) and AcrobotVecEnv
// (acrobot_pre_vec.py:279-394,450-541).
//
// HBM-bound streaming kernel: feature-major fp32 columns, each thread owns VEC=4
// consecutive envs and moves every column with one 128-bit load and one 128-bit store; the Philox reset
// stream is counter-based (no RNG state in memory); obs/reward/done go straight into the rollout slot.
#include <math_constants.h>

#include "tpp_common.cuh"

namespace tpp {

struct EnvParams {          // by-value kernel argument (device copy of tpp_env_cfg + launch data)
  int32_t n_envs, max_steps;
  uint64_t seed;
  float lo[16], hi[16];
  float p[8];
};

template <int F> struct Fam;
template <> struct Fam<TPP_CARTPOLE>       { static constexpr int NS = 9,  NO = 9,  DYN = 0; };
template <> struct Fam<TPP_CARTPOLE_SWING> { static constexpr int NS = 9,  NO = 9,  DYN = 0; };
template <> struct Fam<TPP_MOUNTAIN_CAR>   { static constexpr int NS = 5,  NO = 5,  DYN = 0; };
template <> struct Fam<TPP_ACROBOT>        { static constexpr int NS = 12, NO = 14, DYN = 4; };
template <> struct Fam<TPP_LUNAR_LANDER>   { static constexpr int NS = 8,  NO = 8,  DYN = 0; };

// ------------------------------------------------------------------------------------------------
// Dynamics (one env, registers only).  fp32 with IEEE division and accurate sin/cos: the one-step
// error against the float64 reference stays below 2e-6 absolute (tests/test_env_parity.py).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cartpole_euler(float* s, int a, float tau) {
  const float x = s[0], xd = s[1], th = s[2], thd = s[3], g = s[4], len = s[5], mc = s[6], mp = s[7], fm = s[8];
  const float force = (a == 0) ? -fm : fm;
  float sn, cs;
  sincosf(th, &sn, &cs);
  const float pml = mp * len, tot = mp + mc;
  const float temp = (force + pml * thd * thd * sn) / tot;
  const float thacc = (g * sn - cs * temp) / (len * (4.0f / 3.0f - mp * cs * cs / tot));
  const float xacc = temp - pml * thacc * cs / tot;
  s[0] = x + tau * xd;
  s[1] = xd + tau * xacc;
  s[2] = th + tau * thd;
  s[3] = thd + tau * thacc;
}

template <int F>
__device__ __forceinline__ void transition(float* s, int a, const EnvParams& c, float& rew, bool& term);

template <>
__device__ __forceinline__ void transition<TPP_CARTPOLE>(float* s, int a, const EnvParams& c, float& rew, bool& term) {
  cartpole_euler(s, a, c.p[2]);
  term = (s[0] < -c.p[0]) | (s[0] > c.p[0]) | (s[2] < -c.p[1]) | (s[2] > c.p[1]);
  rew = 1.0f;
}

template <>
__device__ __forceinline__ void transition<TPP_CARTPOLE_SWING>(float* s, int a, const EnvParams& c, float& rew,
                                                                bool& term) {
  cartpole_euler(s, a, c.p[2]);
  term = (s[0] < -c.p[0]) | (s[0] > c.p[0]);
  const float r_theta = fmaxf(cosf(s[2]), 0.0f);
  const float r_x = cosf((s[0] / c.p[0]) * 1.57079632679489662f);
  rew = r_theta * r_x;
}

template <>
__device__ __forceinline__ void transition<TPP_MOUNTAIN_CAR>(float* s, int a, const EnvParams& c, float& rew,
                                                              bool& term) {
  const float force = c.p[0], max_speed = c.p[1], left = c.p[2], goal_vel = c.p[3];
  float pos = s[0], vel = s[1];
  const float g = s[2], right = s[3], goal = s[4];
  vel = vel + ((float)(a - 1) * force + cosf(3.0f * pos) * (-g));
  vel = fminf(fmaxf(vel, -max_speed), max_speed);
  pos = pos + vel;
  pos = fminf(fmaxf(pos, left), right);
  if (pos == left && vel < 0.0f) vel = 0.0f;
  term = (pos >= goal) & (vel >= goal_vel);
  s[0] = pos;
  s[1] = vel;
  rew = (c.p[4] != 0.0f) ? -1.0f : (sinf(3.0f * pos) * 0.45f + 0.55f - 1.0f);
}

// "book" equations, acrobot_pre_vec.py:359-394.  y = (th1, th2, dth1, dth2); pr = 8 physical parameters.
// tr = (cos th1, sin th1, cos th2, sin th2) of y: supplied by the caller (stage 1 reuses the previous observation,
// whose first four columns are exactly these values) so that a stage costs two sincosf instead of sincosf + 2 sinf;
// sin(th1 + th2) follows from the addition theorem.  The kernel is ALU-bound (ncu: 22 % DRAM), trig dominates.
__device__ __forceinline__ void acrobot_dsdt(const float* y, const float* tr, float torque, const float* pr, float* dy) {
  const float g = pr[0], l1 = pr[1], m1 = pr[3], m2 = pr[4], lc1 = pr[5], lc2 = pr[6], moi = pr[7];
  const float d1v = y[2], d2v = y[3];
  const float c1 = tr[0], s1 = tr[1], c2 = tr[2], s2 = tr[3];
  const float s12 = s1 * c2 + c1 * s2;
  const float dd1 = m1 * lc1 * lc1 + m2 * (l1 * l1 + lc2 * lc2 + 2.0f * l1 * lc2 * c2) + moi + moi;
  const float dd2 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + moi;
  // cos(x - pi/2) == sin(x): evaluated as sin to avoid the fp32 rounding of the shifted argument
  const float phi2 = m2 * lc2 * g * s12;
  const float phi1 = -m2 * l1 * lc2 * d2v * d2v * s2 - 2.0f * m2 * l1 * lc2 * d2v * d1v * s2 +
                     (m1 * lc1 + m2 * l1) * g * s1 + phi2;
  // two IEEE divisions per stage instead of four (1 / dd1 is shared: each division is ~10 instructions and the kernel is
  // instruction-bound; <= 1-2 ulp per term, inside the 2e-6 + 1e-5 |b| tolerance of tests/test_env_parity.py)
  const float inv_dd1 = 1.0f / dd1, r21 = dd2 * inv_dd1;
  const float ddth2 = (torque + r21 * phi1 - m2 * l1 * lc2 * d1v * d1v * s2 - phi2) /
                      (m2 * lc2 * lc2 + moi - dd2 * r21);
  const float ddth1 = -(dd2 * ddth2 + phi1) * inv_dd1;
  dy[0] = d1v;
  dy[1] = d2v;
  dy[2] = ddth1;
  dy[3] = ddth2;
}

__device__ __forceinline__ float wrap_pi(float x) {   // per-env wrap to [-pi, pi] (rare path: in double)
  double v = (double)x;
  while (v > CUDART_PI) v -= 2.0 * CUDART_PI;
  while (v < -CUDART_PI) v += 2.0 * CUDART_PI;
  return (float)v;
}

template <>
__device__ __forceinline__ void transition<TPP_ACROBOT>(float* s, int a, const EnvParams& c, float& rew, bool& term) {
  const float torque = (float)(a - 1);
  const float dt = c.p[2], dt2 = 0.5f * c.p[2];
  const float* pr = s + 4;
  float k1[4], k2[4], k3[4], k4[4], y[4], tr[4];
  acrobot_dsdt(s, s + 12, torque, pr, k1);               // s[12..15] = (cos, sin) of th1, th2 from the previous obs
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = s[i] + dt2 * k1[i];
  sincosf(y[0], &tr[1], &tr[0]);
  sincosf(y[1], &tr[3], &tr[2]);
  acrobot_dsdt(y, tr, torque, pr, k2);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = s[i] + dt2 * k2[i];
  sincosf(y[0], &tr[1], &tr[0]);
  sincosf(y[1], &tr[3], &tr[2]);
  acrobot_dsdt(y, tr, torque, pr, k3);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = s[i] + dt * k3[i];
  sincosf(y[0], &tr[1], &tr[0]);
  sincosf(y[1], &tr[3], &tr[2]);
  acrobot_dsdt(y, tr, torque, pr, k4);
  const float dt6 = dt * (1.0f / 6.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = s[i] + dt6 * (k1[i] + 2.0f * k2[i] + 2.0f * k3[i] + k4[i]);
  if (fabsf(y[0]) > 3.1415925f) y[0] = wrap_pi(y[0]);   // cheap pre-filter just below pi; exact test in double
  if (fabsf(y[1]) > 3.1415925f) y[1] = wrap_pi(y[1]);
  s[0] = y[0];
  s[1] = y[1];
  s[2] = fminf(fmaxf(y[2], -c.p[0]), c.p[0]);
  s[3] = fminf(fmaxf(y[3], -c.p[1]), c.p[1]);
  sincosf(s[0], &s[13], &s[12]);                          // reused by the terminal test and by the observation
  sincosf(s[1], &s[15], &s[14]);
  term = (-s[12] - (s[12] * s[14] - s[13] * s[15])) > 1.0f;   // -cos th1 - cos(th1 + th2)
  rew = term ? 0.0f : -1.0f;
}

// lunar_lander_pre_vec: the reference has no implementation (discrete_env/lunar_lander_pre_vec.py:16 raises; the
// rest of that file is Gymnasium's scalar Box2D lander).  Own vectorised semantics, restated in oracle/lunar.py
// ("parity unpinned"): one rigid hull, two massless legs as spring-damper foot contacts on flat ground at the
// helipad height, no engine dispersion, dt = 1/50.  Taken from the reference file: observation layout and
// normalisation (:606-615), engine impulse geometry (:520-601), reward shaping (:617-633), terminal rewards.
// State == the 8-wide normalised observation [x, y, vx, vy, angle, 20*omega/FPS, leg1, leg2].
__device__ __forceinline__ float lander_shaping(const float* s) {
  return -100.0f * sqrtf(s[0] * s[0] + s[1] * s[1]) - 100.0f * sqrtf(s[2] * s[2] + s[3] * s[3]) -
         100.0f * fabsf(s[4]) + 10.0f * s[6] + 10.0f * s[7];
}

template <>
__device__ __forceinline__ void transition<TPP_LUNAR_LANDER>(float* s, int a, const EnvParams& c, float& rew,
                                                              bool& term) {
  constexpr float FPS = 50.0f, SC = 30.0f, W2 = 10.0f, H2 = 400.0f / 30.0f / 2.0f;
  constexpr float PAD = (400.0f / 30.0f) / 4.0f, LEG_DOWN = 18.0f / SC, DT = 1.0f / FPS;
  constexpr float MASS = 4.82f, INERTIA = 0.84f;
  // divisions by constants as multiplications by the (compile-time) reciprocal: an IEEE division is ~10 instructions, and
  // this transition is instruction-bound (<= 1 ulp per operation, far inside the stated 1e-5 tolerance)
  constexpr float I_MASS = 1.0f / MASS, I_INERTIA = 1.0f / INERTIA, I_W2 = 1.0f / W2, I_H2 = 1.0f / H2, I_FPS = 1.0f / FPS;
  constexpr float K_N = 1500.0f, C_N = 60.0f, C_T = 30.0f, MU = 1.0f;
  const float gravity = c.p[0], main_power = c.p[1], side_power = c.p[2];
  const float prev = lander_shaping(s);
  float x = s[0] * W2 + W2, y = s[1] * H2 + (PAD + LEG_DOWN);
  float vx = s[2] * (FPS * I_W2), vy = s[3] * (FPS * I_H2), ang = s[4], om = s[5] * (FPS / 20.0f);
  float sn, cs;
  sincosf(ang, &sn, &cs);
  const float tipx = sn, tipy = cs, sidex = -cs, sidey = sn;
  const float mainf = (a == 2) ? 1.0f : 0.0f;
  {
    const float ox = tipx * (4.0f / SC), oy = -tipy * (4.0f / SC);
    const float jx = -ox * main_power * mainf, jy = -oy * main_power * mainf;
    vx += jx * I_MASS;
    vy += jy * I_MASS;
    om += (ox * jy - oy * jx) * I_INERTIA;
  }
  const float sidef = (a == 1 || a == 3) ? 1.0f : 0.0f;
  {
    const float d = (float)(a - 2) * sidef;
    const float ox = sidex * (d * 12.0f / SC), oy = -sidey * (d * 12.0f / SC);
    const float jx = -ox * side_power, jy = -oy * side_power;
    const float rx = ox - tipx * 17.0f / SC, ry = oy + tipy * 14.0f / SC;
    vx += jx * I_MASS;
    vy += jy * I_MASS;
    om += (rx * jy - ry * jx) * I_INERTIA * sidef;
  }
  const float footx[2] = {-20.0f / SC, 20.0f / SC}, footy = -26.0f / SC;
  float fx = 0.0f, fy = 0.0f, tq = 0.0f;
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const float rx = cs * footx[f] - sn * footy, ry = sn * footx[f] + cs * footy;
    const float pen = PAD - (y + ry);
    if (pen > 0.0f) {
      const float vfx = vx - om * ry, vfy = vy + om * rx;
      const float fn = fmaxf(K_N * pen - C_N * vfy, 0.0f);
      const float ft = fminf(fmaxf(-C_T * vfx, -MU * fn), MU * fn);
      fx += ft;
      fy += fn;
      tq += rx * fn - ry * ft;
    }
  }
  vx += fx * (DT * I_MASS);
  vy += (fy * I_MASS - gravity) * DT;
  om += tq * (DT * I_INERTIA);
  x += vx * DT;
  y += vy * DT;
  ang += om * DT;
  sincosf(ang, &sn, &cs);
  const float c0 = (y + sn * footx[0] + cs * footy <= PAD) ? 1.0f : 0.0f;
  const float c1 = (y + sn * footx[1] + cs * footy <= PAD) ? 1.0f : 0.0f;
  const float hull = fminf(y + sn * (-17.0f / SC) + cs * (-10.0f / SC), y + sn * (17.0f / SC) + cs * (-10.0f / SC));
  s[0] = (x - W2) * I_W2;
  s[1] = (y - (PAD + LEG_DOWN)) * I_H2;
  s[2] = vx * (W2 * I_FPS);
  s[3] = vy * (H2 * I_FPS);
  s[4] = ang;
  s[5] = om * (20.0f * I_FPS);
  s[6] = c0;
  s[7] = c1;
  rew = lander_shaping(s) - prev - 0.30f * mainf - 0.03f * sidef;
  const bool crashed = (hull <= PAD) | (fabsf(s[0]) >= 1.0f);
  const bool landed = (c0 > 0.0f) & (c1 > 0.0f) & (fabsf(vx) < 0.05f) & (fabsf(vy) < 0.05f) & (fabsf(om) < 0.05f) & !crashed;
  if (crashed) rew = -100.0f;
  if (landed) rew = 100.0f;
  term = crashed | landed;
}

// ------------------------------------------------------------------------------------------------
// Start-space sampling: Philox(seed; env, tick_lo, tick_hi, block)
// ------------------------------------------------------------------------------------------------
template <int F>
__device__ __forceinline__ void draw_start(float* s, const EnvParams& c, uint32_t env, uint64_t tick) {
  constexpr int NS = Fam<F>::NS;
  Philox rng(c.seed);
  const uint32_t t0 = (uint32_t)tick, t1 = (uint32_t)(tick >> 32);
  if (F == TPP_MOUNTAIN_CAR) {
    // rejection of goal_position > right_boundary (mountain_car_pre_vec.py:158-161, helper_pre_vec.py:32-38)
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
      const uint4 r = rng(env, t0, t1, attempt);
      s[0] = c.lo[0] + (c.hi[0] - c.lo[0]) * u01(r.x);
      s[1] = c.lo[1];
      s[2] = c.lo[2] + (c.hi[2] - c.lo[2]) * u01(r.y);
      s[3] = c.lo[3] + (c.hi[3] - c.lo[3]) * u01(r.z);
      s[4] = c.lo[4] + (c.hi[4] - c.lo[4]) * u01(r.w);
      if (!(s[4] > s[3])) return;
    }
    s[4] = s[3];
    return;
  }
#pragma unroll
  for (int b = 0; b < (NS + 3) / 4; ++b) {
    const uint4 r = rng(env, t0, t1, (uint32_t)b);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = b * 4 + j;
      if (col < NS) s[col] = c.lo[col] + (c.hi[col] - c.lo[col]) * u01(w[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Vector helpers: VEC consecutive envs of one column in one access
// ------------------------------------------------------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<4> {
  static __device__ __forceinline__ void ldf(const float* p, float* o) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p)); o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
  static __device__ __forceinline__ void stf(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  static __device__ __forceinline__ void ldi(const int32_t* p, int32_t* o) {
    const int4 q = __ldcs(reinterpret_cast<const int4*>(p)); o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
  static __device__ __forceinline__ void sti(int32_t* p, const int32_t* v) {
    __stcs(reinterpret_cast<int4*>(p), make_int4(v[0], v[1], v[2], v[3]));
  }
  static __device__ __forceinline__ void stb(uint8_t* p, const uint8_t* v) {
    __stcs(reinterpret_cast<uchar4*>(p), make_uchar4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct Vec<2> {
  static __device__ __forceinline__ void ldf(const float* p, float* o) {
    const float2 q = __ldcs(reinterpret_cast<const float2*>(p)); o[0] = q.x; o[1] = q.y;
  }
  static __device__ __forceinline__ void stf(float* p, const float* v) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  }
  static __device__ __forceinline__ void ldi(const int32_t* p, int32_t* o) {
    const int2 q = __ldcs(reinterpret_cast<const int2*>(p)); o[0] = q.x; o[1] = q.y;
  }
  static __device__ __forceinline__ void sti(int32_t* p, const int32_t* v) {
    __stcs(reinterpret_cast<int2*>(p), make_int2(v[0], v[1]));
  }
  static __device__ __forceinline__ void stb(uint8_t* p, const uint8_t* v) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2(v[0], v[1]);
  }
};
template <> struct Vec<1> {
  static __device__ __forceinline__ void ldf(const float* p, float* o) { o[0] = __ldcs(p); }
  static __device__ __forceinline__ void stf(float* p, const float* v) { __stcs(p, v[0]); }
  static __device__ __forceinline__ void ldi(const int32_t* p, int32_t* o) { o[0] = __ldcs(p); }
  static __device__ __forceinline__ void sti(int32_t* p, const int32_t* v) { __stcs(p, v[0]); }
  static __device__ __forceinline__ void stb(uint8_t* p, const uint8_t* v) { *p = v[0]; }
};

template <int F>
__device__ __forceinline__ void emit_obs(const float* s, float* o) {
  if (F == TPP_ACROBOT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = s[12 + j];        // (cos, sin) pairs cached behind the state
#pragma unroll
    for (int j = 2; j < 12; ++j) o[j + 2] = s[j];
  } else {
#pragma unroll
    for (int j = 0; j < Fam<F>::NO; ++j) o[j] = s[j];
  }
}

// ------------------------------------------------------------------------------------------------
// The step kernel
// ------------------------------------------------------------------------------------------------
template <int F, int VEC>
__global__ void __launch_bounds__(256) env_step_kernel(EnvParams c, const float* obs_in, float* obs_out,
                                                       float* dyn, const int32_t* action, int32_t* step_ctr,
                                                       float* rew_out, uint8_t* done_out, const float* reset_rows,
                                                       const uint64_t* tick, uint64_t t_offset, int64_t ld) {
  constexpr int NS = Fam<F>::NS, NO = Fam<F>::NO, DYN = Fam<F>::DYN;
  const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e0 >= c.n_envs) return;

  float s[NS + 4][VEC];        // + 4: acrobot's cached (cos, sin) pairs; unused (and eliminated) elsewhere
  int32_t act[VEC], ctr[VEC];
  // ---- all loads first (independent 128-bit requests in flight) ----
  if (DYN) {
#pragma unroll
    for (int j = 0; j < DYN; ++j) Vec<VEC>::ldf(dyn + j * ld + e0, s[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) Vec<VEC>::ldf(obs_in + (int64_t)j * ld + e0, s[NS + j]);
#pragma unroll
    for (int j = DYN; j < NS; ++j) Vec<VEC>::ldf(obs_in + (int64_t)(j + NO - NS) * ld + e0, s[j]);
  } else {
#pragma unroll
    for (int j = 0; j < NS; ++j) Vec<VEC>::ldf(obs_in + (int64_t)j * ld + e0, s[j]);
  }
  Vec<VEC>::ldi(action + e0, act);
  Vec<VEC>::ldi(step_ctr + e0, ctr);
  const uint64_t tk = (tick ? *tick : 0ull) + t_offset;

  float o[NO][VEC], rew[VEC];
  uint8_t dn[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float st[NS + 4], ob[NO];
#pragma unroll
    for (int j = 0; j < NS + (DYN ? 4 : 0); ++j) st[j] = s[j][i];
    bool term;
    transition<F>(st, act[i], c, rew[i], term);
    const int32_t n = ctr[i] + 1;
    const bool done = term | (n >= c.max_steps);
    if (done) {
      if (reset_rows) {
#pragma unroll
        for (int j = 0; j < NS; ++j) st[j] = reset_rows[(int64_t)j * ld + e0 + i];
      } else {
        draw_start<F>(st, c, (uint32_t)(e0 + i), tk);
      }
      if (F == TPP_ACROBOT) {
        sincosf(st[0], &st[13], &st[12]);
        sincosf(st[1], &st[15], &st[14]);
      }
    }
    ctr[i] = done ? 0 : n;
    dn[i] = done ? 1 : 0;
    emit_obs<F>(st, ob);
#pragma unroll
    for (int j = 0; j < NO; ++j) o[j][i] = ob[j];
    if (DYN) {
#pragma unroll
      for (int j = 0; j < DYN; ++j) s[j][i] = st[j];
    }
  }
  // ---- stores ----
#pragma unroll
  for (int j = 0; j < NO; ++j) Vec<VEC>::stf(obs_out + (int64_t)j * ld + e0, o[j]);
  if (DYN) {
#pragma unroll
    for (int j = 0; j < DYN; ++j) Vec<VEC>::stf(dyn + j * ld + e0, s[j]);
  }
  Vec<VEC>::sti(step_ctr + e0, ctr);
  Vec<VEC>::stf(rew_out + e0, rew);
  Vec<VEC>::stb(done_out + e0, dn);
}

template <int F>
__global__ void __launch_bounds__(256) env_reset_kernel(EnvParams c, float* obs_out, float* dyn, int32_t* step_ctr,
                                                        const float* reset_rows, const uint64_t* tick,
                                                        uint64_t t_offset, int64_t ld) {
  constexpr int NS = Fam<F>::NS, NO = Fam<F>::NO, DYN = Fam<F>::DYN;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= c.n_envs) return;
  float st[NS + 4], ob[NO];
  if (reset_rows) {
#pragma unroll
    for (int j = 0; j < NS; ++j) st[j] = reset_rows[(int64_t)j * ld + e];
  } else {
    draw_start<F>(st, c, (uint32_t)e, (tick ? *tick : 0ull) + t_offset);
  }
  if (F == TPP_ACROBOT) {
    sincosf(st[0], &st[13], &st[12]);
    sincosf(st[1], &st[15], &st[14]);
  }
  emit_obs<F>(st, ob);
#pragma unroll
  for (int j = 0; j < NO; ++j) obs_out[(int64_t)j * ld + e] = ob[j];
#pragma unroll
  for (int j = 0; j < DYN; ++j) dyn[(int64_t)j * ld + e] = st[j];
  step_ctr[e] = 0;
}

__global__ void tick_advance_kernel(uint64_t* tick, uint64_t delta) { *tick += delta; }

static EnvParams to_params(const tpp_env_cfg* cfg) {
  EnvParams p;
  p.n_envs = cfg->n_envs;
  p.max_steps = cfg->max_steps;
  p.seed = cfg->seed;
  for (int i = 0; i < 16; ++i) { p.lo[i] = cfg->start_low[i]; p.hi[i] = cfg->start_high[i]; }
  for (int i = 0; i < 8; ++i) p.p[i] = cfg->p[i];
  return p;
}

template <int F>
static int launch_step(const tpp_env_cfg* cfg, const float* obs_in, float* obs_out, float* dyn, const int32_t* action,
                       int32_t* step_ctr, float* rew_out, uint8_t* done_out, const float* reset_rows,
                       const uint64_t* tick, uint64_t t_offset, int64_t ld, cudaStream_t s) {
  const EnvParams p = to_params(cfg);
  const int64_t N = cfg->n_envs;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec4 = (N % 4 == 0) && (ld % 4 == 0) && al16(obs_in) && al16(obs_out) && al16(action) && al16(step_ctr) &&
                    al16(rew_out) && ((reinterpret_cast<uintptr_t>(done_out) & 3) == 0) && (!dyn || al16(dyn));
  // cfg->p[7] = envs per thread hint (0 = default 4): the ALU-bound families trade vector width for occupancy
  const int hint = (int)cfg->p[7];
  if (vec4 && hint == 2) {
    const int grid = tpp_ceil_div(N / 2, 256);
    env_step_kernel<F, 2><<<grid, 256, 0, s>>>(p, obs_in, obs_out, dyn, action, step_ctr, rew_out, done_out,
                                               reset_rows, tick, t_offset, ld);
  } else if (vec4 && hint != 1) {
    const int grid = tpp_ceil_div(N / 4, 256);
    env_step_kernel<F, 4><<<grid, 256, 0, s>>>(p, obs_in, obs_out, dyn, action, step_ctr, rew_out, done_out,
                                               reset_rows, tick, t_offset, ld);
  } else {
    const int grid = tpp_ceil_div(N, 256);
    env_step_kernel<F, 1><<<grid, 256, 0, s>>>(p, obs_in, obs_out, dyn, action, step_ctr, rew_out, done_out,
                                               reset_rows, tick, t_offset, ld);
  }
  TPP_LAUNCH_STATUS();
}

}  // namespace tpp

extern "C" int tpp_env_step(const tpp_env_cfg* cfg, const float* obs_in, float* obs_out, float* dyn_state,
                            const int32_t* action, int32_t* step_ctr, float* rew_out, uint8_t* done_out,
                            const float* reset_rows, const uint64_t* tick, uint64_t t_offset, int64_t ld,
                            void* stream) {
  TPP_CHECK_ARG(cfg && obs_in && obs_out && action && step_ctr && rew_out && done_out);
  TPP_CHECK_ARG(cfg->n_envs > 0 && ld >= cfg->n_envs);
  cudaStream_t s = tpp_stream(stream);
  switch (cfg->family) {
    case TPP_CARTPOLE:
      return tpp::launch_step<TPP_CARTPOLE>(cfg, obs_in, obs_out, nullptr, action, step_ctr, rew_out, done_out,
                                            reset_rows, tick, t_offset, ld, s);
    case TPP_CARTPOLE_SWING:
      return tpp::launch_step<TPP_CARTPOLE_SWING>(cfg, obs_in, obs_out, nullptr, action, step_ctr, rew_out, done_out,
                                                  reset_rows, tick, t_offset, ld, s);
    case TPP_MOUNTAIN_CAR:
      return tpp::launch_step<TPP_MOUNTAIN_CAR>(cfg, obs_in, obs_out, nullptr, action, step_ctr, rew_out, done_out,
                                                reset_rows, tick, t_offset, ld, s);
    case TPP_ACROBOT:
      TPP_CHECK_ARG(dyn_state);
      return tpp::launch_step<TPP_ACROBOT>(cfg, obs_in, obs_out, dyn_state, action, step_ctr, rew_out, done_out,
                                           reset_rows, tick, t_offset, ld, s);
    case TPP_LUNAR_LANDER:
      return tpp::launch_step<TPP_LUNAR_LANDER>(cfg, obs_in, obs_out, nullptr, action, step_ctr, rew_out, done_out,
                                                reset_rows, tick, t_offset, ld, s);
    default:
      return TPP_ENOTSUP;
  }
}

extern "C" int tpp_env_reset(const tpp_env_cfg* cfg, float* obs_out, float* dyn_state, int32_t* step_ctr,
                             const float* reset_rows, const uint64_t* tick, uint64_t t_offset, int64_t ld,
                             void* stream) {
  TPP_CHECK_ARG(cfg && obs_out && step_ctr);
  TPP_CHECK_ARG(cfg->n_envs > 0 && ld >= cfg->n_envs);
  cudaStream_t s = tpp_stream(stream);
  const tpp::EnvParams p = tpp::to_params(cfg);
  const int grid = tpp_ceil_div(cfg->n_envs, 256);
  switch (cfg->family) {
    case TPP_CARTPOLE:
      tpp::env_reset_kernel<TPP_CARTPOLE><<<grid, 256, 0, s>>>(p, obs_out, nullptr, step_ctr, reset_rows, tick, t_offset, ld);
      break;
    case TPP_CARTPOLE_SWING:
      tpp::env_reset_kernel<TPP_CARTPOLE_SWING><<<grid, 256, 0, s>>>(p, obs_out, nullptr, step_ctr, reset_rows, tick, t_offset, ld);
      break;
    case TPP_MOUNTAIN_CAR:
      tpp::env_reset_kernel<TPP_MOUNTAIN_CAR><<<grid, 256, 0, s>>>(p, obs_out, nullptr, step_ctr, reset_rows, tick, t_offset, ld);
      break;
    case TPP_ACROBOT:
      TPP_CHECK_ARG(dyn_state);
      tpp::env_reset_kernel<TPP_ACROBOT><<<grid, 256, 0, s>>>(p, obs_out, dyn_state, step_ctr, reset_rows, tick, t_offset, ld);
      break;
    case TPP_LUNAR_LANDER:
      tpp::env_reset_kernel<TPP_LUNAR_LANDER><<<grid, 256, 0, s>>>(p, obs_out, nullptr, step_ctr, reset_rows, tick, t_offset, ld);
      break;
    default:
      return TPP_ENOTSUP;
  }
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_tick_advance(uint64_t* tick, uint64_t delta, void* stream) {
  TPP_CHECK_ARG(tick);
  tpp::tick_advance_kernel<<<1, 1, 0, tpp_stream(stream)>>>(tick, delta);
  TPP_LAUNCH_STATUS();
}
