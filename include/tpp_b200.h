/* tpp_b200.h — C-ABI of the B200-native PPO rollout-and-update hot path.
 *
 * One shared library (train-procgen-pytorch_b200/csrc/libtpp_b200.so, built by nvcc for sm_100a only).
 * Every entry point takes plain device pointers + sizes + a cudaStream_t passed as void*, launches
 * asynchronously on that stream, never allocates or frees, keeps no hidden global state, and returns 0 or a
 * cudaError_t / TPP_E* code (never throws).  The caller (PyTorch on the Python side, ctypes binding in
 * train-procgen-pytorch_b200/_lib.py) owns every buffer.
 *
 * The reference (tbuckworth/train-procgen-pytorch) is pure Python; it has no FFI.  Each entry point therefore
 * cites the reference *function* it replaces (path:line under the reference root).
 *
 * Layouts:
 *   vector observations  : feature-major ("SoA") float32  obs[T+1][n_obs][ld]   (ld >= N envs, ld % 32 == 0)
 *   image observations   : uint8 NHWC                     frame[T+1][N][H][W][3]
 *   per-step scalars     : [T][N] row-major (act int32, logp/rew/value/adv/ret float32, done uint8)
 */
#ifndef TPP_B200_H
#define TPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPP_OK 0
#define TPP_EINVAL 10001   /* bad argument (null pointer, size, alignment)            */
#define TPP_ENOTSUP 10002  /* shape/config outside what the kernel was built for      */

/* ---- library / device -------------------------------------------------------------------------------- */
int tpp_version(void);                       /* ABI version, bumped on any signature change             */
int tpp_device_sm_count(int* out_sm_count);  /* multiProcessorCount of the current device               */
const char* tpp_error_string(int code);

/* ---- pre-vectorised env families --------------------------------------------------------------------- */
enum { TPP_CARTPOLE = 0, TPP_CARTPOLE_SWING = 1, TPP_MOUNTAIN_CAR = 2, TPP_ACROBOT = 3, TPP_LUNAR_LANDER = 4 };

/* Static description of one env family instance (replaces the ctor kwargs of the reference classes:
 * discrete_env/cartpole_pre_vec.py:92-109, cartpole_swing_pre_vec.py:91-106, mountain_car_pre_vec.py:107-126,
 * acrobot_pre_vec.py:164-180).  `p[]` meaning per family:
 *   cartpole        p0 x_threshold, p1 theta_threshold_radians, p2 tau
 *   cartpole_swing  p0 x_threshold, p2 tau
 *   mountain_car    p0 force, p1 max_speed, p2 left_boundary, p3 goal_velocity, p4 sparse_rewards(0/1)
 *   acrobot         p0 max_vel_1, p1 max_vel_2, p2 dt
 *   lunar_lander    see DESIGN.md (own semantics; the reference has no implementation)                  */
typedef struct {
  int32_t family;
  int32_t n_envs;          /* N                                                                        */
  int32_t max_steps;       /* truncation horizon (pre_vec_env.py:84-86)                                */
  int32_t n_state;         /* columns of the start space: 9 / 9 / 5 / 12 / 8                           */
  uint64_t seed;           /* Philox4x32-10 key                                                        */
  float start_low[16];     /* StartSpace low/high per state column (helper_pre_vec.py:4-12)            */
  float start_high[16];
  float p[8];
} tpp_env_cfg;

/* Fused transition + step counter + truncation + auto-reset + obs/reward/done emit for one step of N envs.
 * Replaces PreVecEnv.step + set (discrete_env/pre_vec_env.py:78-93,108-119) and the family's
 * transition_model.  obs_in / obs_out are rollout slots t and t+1 (feature-major, column stride ld).
 * For families whose observation IS the state (cartpole, swing, mountain car) the state lives only in the
 * slots; acrobot additionally updates dyn_state[4][ld] (theta1, theta2, dtheta1, dtheta2) in place.
 * done_out holds the PRE-reset flag, obs_out the POST-reset observation (reference semantics).
 * reset_rows: NULL -> start states drawn in-kernel from Philox(seed, env, tick); otherwise feature-major
 * [n_state][ld] rows used for the finished envs (teacher forcing for parity tests).
 * tick: device counter mixed into the Philox counter (may be NULL -> 0); t_offset is added to it.            */
int tpp_env_step(const tpp_env_cfg* cfg, const float* obs_in, float* obs_out, float* dyn_state,
                 const int32_t* action, int32_t* step_ctr, float* rew_out, uint8_t* done_out,
                 const float* reset_rows, const uint64_t* tick, uint64_t t_offset, int64_t ld, void* stream);

/* Full reset of all N envs into obs_out (+ dyn_state) and step_ctr = 0.
 * Replaces PreVecEnv.reset (discrete_env/pre_vec_env.py:98-106).                                         */
int tpp_env_reset(const tpp_env_cfg* cfg, float* obs_out, float* dyn_state, int32_t* step_ctr,
                  const float* reset_rows, const uint64_t* tick, uint64_t t_offset, int64_t ld, void* stream);

/* tick[0] += delta (one thread).  Lets a captured CUDA graph advance the RNG stream between replays.     */
int tpp_tick_advance(uint64_t* tick, uint64_t delta, void* stream);

/* torch.randperm(n) on torch's default CPU generator, restated on the host (MT19937 + ATen's Fisher-Yates loop): the
 * permutation `out[n]` and the advanced generator state are bit-identical to torch's, several times faster.
 * state624 / left / next: the MT19937 fields of torch.get_rng_state() (624 words, one per uint64; countdown; index).
 * Replaces torch.randperm in Storage.fetch_train_generator (common/storage.py:87).  n < 2^32 / 20.            */
int tpp_randperm_mt19937(uint64_t* state624, int32_t* left, uint64_t* next, int64_t n, int64_t* out);
/* The same permutation written as int32 (n < 2^31 always holds here): half the pinned bytes / H2D traffic of an epoch's
 * minibatch indices; the device widens them when it copies the staging buffer into the gather's index buffer.      */
int tpp_randperm_mt19937_i32(uint64_t* state624, int32_t* left, uint64_t* next, int64_t n, int32_t* out);

/* ---- Box-World ----------------------------------------------------------------------------------------- */
/* Device-resident state of N Box-World envs (replaces the numpy members of BoxWorldVec,
 * boxworld/box_world_env_vec.py:24-60).  cells = (n+2)*(n+2).                                            */
typedef struct {
  int32_t n_envs, n;             /* grid side n (frames are (n+2) x (n+2) x 3)                            */
  int32_t max_steps;
  int32_t n_levels;              /* >0: levels cycle through a bank of n_levels; 0: unbounded (streamed)  */
  int64_t start_seed;
  int32_t goal_length, num_distractor, distractor_length, _pad;
  uint8_t* world;                /* [N][n+2][n+2][3]                                                      */
  int8_t* world_dic;             /* [N][n+2][n+2]  -1 / 0 / 1   (lock status)                            */
  int32_t* player_pos;           /* [N][2]                                                                */
  uint8_t* owned_key;            /* [N][4]  (rgb + pad)                                                   */
  int32_t* num_env_steps;        /* [N]                                                                   */
  int32_t* episode_reward;       /* [N]                                                                   */
  int64_t* seed_counter;         /* [1] next level seed (box_world_env_vec.py:294-297)                    */
  const uint8_t* bank_world;     /* [n_levels][cells][3]   (NULL when n_levels == 0)                      */
  const int8_t* bank_dic;        /* [n_levels][cells]                                                     */
  const int32_t* bank_pos;       /* [n_levels][2]                                                         */
  int32_t* scratch;              /* [2 N + 4096], zero-initialised: per-CTA done counts, ticket, scan workspace         */
} tpp_boxworld_state;

/* One step of all envs: grid transition, rewards, done, level replacement in env-index order with the
 * sequential seed counter, and emit of the post-reset frame into frame_out (rollout slot t+1, uint8 NHWC).
 * Replaces BoxWorldVec.step (boxworld/box_world_env_vec.py:70-209).
 * reward_out int32 [N] (raw env reward), done_out uint8 [N]; fin_ret/fin_len/fin_solved nullable [N].
 * obs_out (nullable) fp32 [N][ld_obs]: the same post-reset frames as the policy's next input rows, channel-major
 * integer pixel values 0..255 (= tpp_frames_to_obs(raw) of frame_out, without a launch of its own in the rollout's
 * dependent kernel chain).                                                                                  */
int tpp_boxworld_step(const tpp_boxworld_state* st, const int32_t* action, int32_t* reward_out,
                      uint8_t* done_out, uint8_t* frame_out, int32_t* fin_ret, int32_t* fin_len,
                      uint8_t* fin_solved, float* obs_out, int32_t ld_obs, void* stream);

/* Host-side level generator (MT19937 + CPython random.sample/choices semantics), `count` consecutive seeds
 * starting at seed0.  Replaces world_gen / sampling_pairs (boxworld/boxworld_gen_vec.py:4-97).
 * Outputs are HOST buffers: world [count][cells][3], dic [count][cells], pos [count][2].                  */
int tpp_boxworld_gen_levels_host(int32_t n, int32_t goal_length, int32_t num_distractor,
                                 int32_t distractor_length, int64_t seed0, int32_t count, uint8_t* world,
                                 int8_t* dic, int32_t* pos);

/* Same generator on the device, one level per thread, writing straight into env slots `env_ids[i]`
 * (or slots 0..count-1 when env_ids is NULL) of a tpp_boxworld_state; seeds[i] per level.                */
int tpp_boxworld_gen_levels_device(const tpp_boxworld_state* st, const int32_t* env_ids, const int64_t* seeds,
                                   int32_t count, void* stream);

/* Copy frames of all envs into a rollout slot (used after reset()).                                      */
int tpp_boxworld_emit_frames(const tpp_boxworld_state* st, uint8_t* frame_out, void* stream);

/* Return-based reward normalisation for N envs, one step.  Replaces VecNormalize.step_wait +
 * RunningMeanStd.update (common/env/procgen_wrappers.py:282-342, ob=False).
 * ret [N] f64 running discounted return; rms [3] f64 = (mean, var, count); raw_rew int32 or float32
 * (raw_is_int); out_rew f32 [N] normalised+clipped; done u8 [N].  n_envs <= 65536 (single CTA pass).      */
int tpp_vecnormalize_step(double* ret, double* rms, const void* raw_rew, int raw_is_int, const uint8_t* done,
                          float* out_rew, int32_t n_envs, double gamma, double cliprew, double epsilon,
                          void* stream);
/* The same for the T steps of a finished rollout in one launch (identical arithmetic, step after step): raw_rew int32
 * [T][ld], done u8 [T][ld] -> out_rew f32 [T][ld] normalised+clipped and (nullable) out_raw f32 [T][ld] = the raw
 * rewards as floats for the logger.  Rewards never feed back into the rollout, so the per-step cross-env reduction
 * can leave the step's critical path (and env ranges can step independently).  With scratch (>= T*ld + 3T doubles)
 * the work is four short parallel launches -- returns per env, batch moments per step, the sequential Chan merge of T
 * moment pairs, normalisation -- instead of one CTA walking the T steps (n_envs <= 65536 there).                */
int tpp_vecnormalize_rollout(double* ret, double* rms, const int32_t* raw_rew, const uint8_t* done, float* out_rew,
                             float* out_raw, int32_t T, int32_t n_envs, int64_t ld, double gamma, double cliprew,
                             double epsilon, double* scratch, int64_t scratch_doubles, void* stream);

/* ---- rollout storage ----------------------------------------------------------------------------------- */
/* GAE(gamma, lambda) reverse scan with done-masking + returns + global first/second moments.
 * Replaces the loop of Storage.compute_estimates (common/storage.py:56-77).
 * rew,done: [T][ld]; value: [T+1][ld]; adv,ret: [T][ld]; moments: double[3] = (sum, sum of squares, count),
 * ACCUMULATED into (caller zeroes).  Arithmetic order equals the reference's fp32 ops (bit-exact raw adv). */
int tpp_gae(const float* rew, const uint8_t* done, const float* value, float* adv, float* ret,
            double* moments, int32_t T, int32_t N, int64_t ld, float gamma, float lambda, void* stream);

/* The same estimates as a warp-level segmented scan over n_steps (affine-map composition, done = segment boundary),
 * fused with the advantage moments and -- normalize != 0 -- with the normalisation itself (one cooperative launch with
 * a grid barrier; adv is written once, already normalised).  Replaces common/storage.py:56-79 in ONE launch.
 * moments4: double[4], zeroed by the caller: (sum, sum of squares, count) of the RAW advantages + the barrier's counter.
 * Re-associates the products: raw advantages / returns agree with tpp_gae to ~4e-7 of their scale, not bit for bit.
 * TPP_ENOTSUP when T needs more than 200 KB of shared memory or (normalize) the grid (N / 32 CTAs) cannot be
 * co-resident -- callers then use tpp_gae + tpp_adv_normalize.  normalize = 0: raw adv (sharded runs all-reduce the
 * moments first).                                                                                              */
int tpp_gae_scan(const float* rew, const uint8_t* done, const float* value, float* adv, float* ret,
                 double* moments4, int32_t T, int32_t N, int64_t ld, float gamma, float lambda, int32_t normalize,
                 void* stream);

/* adv <- (adv - mean) / (std_unbiased + 1e-8) with the moments from tpp_gae (after an optional cross-rank
 * all-reduce of the three doubles).  Replaces common/storage.py:78-79.                                   */
int tpp_adv_normalize(float* adv, const double* moments, int32_t T, int32_t N, int64_t ld, void* stream);

/* Device-side episode accounting for the logger.  Replaces the O(T*N) host loop of Logger.feed
 * (common/logger.py:119-147) over the batches of Storage.fetch_log_data (common/storage.py:130-162).
 * rew f32 / done u8: [T][ld] (the raw env rewards when the env normalises them).  run_ret f64 [N] / run_len i32 [N]:
 * return and length of every env's open episode, carried between rollouts (updated in place).  scratch: i32 [N+1].
 * out: f64 [2 + 2*keep] = (episodes finished during this rollout, records written = min(that, keep), then
 * (return, length) pairs of the LAST `keep` finished episodes in the reference's env-major walk order).     */
int tpp_episode_scan(const float* rew, const uint8_t* done, int32_t T, int32_t N, int64_t ld, double* run_ret,
                     int32_t* run_len, int32_t* scratch, double* out, int32_t keep, void* stream);

/* Minibatch gather from the resident rollout by flat sample index k = t*N + e (common/storage.py:112-128).
 * Vector observations: obs feature-major [T+1][n_obs][ld] -> out_obs row-major [mb][ld_out] (cols >= n_obs
 * zero-filled).  out_* arrays are [mb].  idx: int64 [mb] (from torch.randperm on the host: bit-exact).    */
int tpp_gather_vec(const int64_t* idx, int32_t mb, int32_t N, int64_t ld, int32_t n_obs, const float* obs,
                   const int32_t* act, const float* logp, const float* value, const float* ret,
                   const float* adv, const uint8_t* done, float* out_obs, float* out_obs_lo, int32_t ld_out,
                   int32_t* out_act, float* out_logp, float* out_value, float* out_ret, float* out_adv,
                   float* out_done, void* stream);
/* out_obs_lo (both gathers, nullable): when given, out_obs receives tf32_round(x) and out_obs_lo the residual, i.e.
 * the (hi, lo) operand pair of tpp_gemm_tc is produced by the gather itself.                                  */

/* Image observations: frames uint8 NHWC [T+1][N][H][W][C] -> out_obs float32 NCHW [mb][C][H][W] / 255
 * (TransposeFrame + ScaledFloatFrame, common/env/procgen_wrappers.py:391-419, applied at gather time).    */
int tpp_gather_img(const int64_t* idx, int32_t mb, int32_t N, int64_t ld, int32_t H, int32_t W, int32_t C,
                   const uint8_t* frames, const int32_t* act, const float* logp, const float* value,
                   const float* ret, const float* adv, const uint8_t* done, float* out_obs, float* out_obs_lo,
                   int32_t ld_out, int32_t* out_act, float* out_logp, float* out_value, float* out_ret,
                   float* out_adv, float* out_done, int32_t raw, void* stream);
/* raw != 0 (out_obs_lo must be NULL): out_obs receives the integer pixel values 0..255 as float32 (exact in TF32, so
 * the tensor-core policy needs no lo half: TPP_TC_A_EXACT) and the consumer folds ScaledFloatFrame's 1/255 into its
 * first layer's weights.                                                                                       */

/* uint8 NHWC frames of one rollout slot -> float32 NCHW/255 rows [N][ld_out] (policy input at rollout).   */
int tpp_frames_to_obs(const uint8_t* frames, int32_t N, int32_t H, int32_t W, int32_t C, float* out_obs,
                      float* out_obs_lo, int32_t ld_out, int32_t raw, void* stream);

/* ---- policy: dense layers ------------------------------------------------------------------------------ */
/* C[m][n] (+)= epilogue( sum_k A(m,k) * B(n,k) ), fp32 in / fp32 accumulate, generic strides:
 *   A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk],  C(m,n) = C[m*ldc + n].
 * flags: TPP_EPI_BIAS (add bias[n]), TPP_EPI_RELU, TPP_EPI_MASK (multiply by mask(m,n) > 0, mask ld = ldc),
 *        TPP_EPI_ACCUM (C += result; with split_k > 1 partial sums are combined with atomics).
 * This one entry serves nn.Linear forward (common/model.py:962-967), its data gradient and its weight
 * gradient (autograd of agents/ppo.py:170).  CUDA-core exact-fp32 path; the tensor-core path is below.    */
enum { TPP_EPI_BIAS = 1, TPP_EPI_RELU = 2, TPP_EPI_MASK = 4, TPP_EPI_ACCUM = 8, TPP_EPI_ADD = 16,
       TPP_EPI_RELU_OUT = 32, /* tpp_gemm_tc only: max(.,0) once more AFTER the addend (the ReLU behind the last
                                 residual block, common/model.py:182) */
       TPP_EPI_PAIR_RELU = 64 /* tpp_gemm_tc only: out_hi/out_lo receive the TF32 pair of max(result, 0) while `out`
                                 keeps the result itself: the next convolution's ReLU'd input, common/model.py:146-150 */ };
int tpp_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                 int64_t ldc, const float* bias, const float* mask, int32_t M, int32_t N, int32_t K,
                 int32_t flags, int32_t split_k, void* stream);

/* out[n] += sum_m dZ[m*ld + n]   (bias gradient).                                                        */
int tpp_colsum_accum(const float* dZ, int64_t ld, int32_t M, int32_t N, float* out, void* stream);

/* Tensor-core dense layer (tcgen05.mma kind::tf32, TMEM accumulators, TMA-fed, csrc/gemm_tc.cu):
 *   C[m][n] = epilogue( sum_k A(m,k) * B(n,k) ).
 * Operand layouts: K-major (a_mn/b_mn = 0): matrix [M or N rows][ld], contraction index contiguous;
 *                  MN-major (= 1): matrix [K rows][ld], m / n index contiguous, ld >= ceil32(M or N).
 * So forward (X[mb][in], W[out][in]: K,K), data gradient (dZ[mb][out] K-major, W[out][in] MN-major) and weight
 * gradient (dZ[mb][out], X[mb][in]: MN,MN, contraction over mb) all read the SAME row-major arrays.
 * Every operand is a (hi, lo) pair of fp32 arrays, hi = tf32_round(x), lo = x - hi; precision 3 accumulates
 * hi*hi + hi*lo + lo*hi (3xTF32, fp32-grade: the parity path), precision 1 only hi*hi (fast mode; *_lo unused).
 * Addresses must be 16-byte aligned and lda/ldb multiples of 4 (TMA); out-of-range k is zero-filled by TMA.
 * flags: TPP_EPI_BIAS, TPP_EPI_RELU, TPP_EPI_MASK (zero where mask[m*ld_mask+n] <= 0), TPP_EPI_ACCUM (fp32 atomic
 *        accumulation of the raw product into `out`; the only mode that allows split_k > 1: weight gradients).
 * Outputs (each nullable): out (plain fp32 [M][ldc]), out_hi/out_lo (TF32 pair, [M][ldc]), colsum ([N], += column
 * sums of the result: the bias gradient of the layer below).  block_n: 0 = auto, 16/32/64/128/256 = 128 x block_n tile
 * on one CTA, or a TPP_TC_TILE_* code: 256 x 256 (256 x 64) tiles on CTA pairs (clusters of two CTAs, tcgen05
 * cta_group::2: each CTA stages 128 rows of A and half of the tile's B rows), optionally persistent (74 pairs loop over
 * the work items with two TMEM accumulators, the epilogue of one item overlapping the next item's loads and MMAs).
 * Replaces nn.Linear forward / backward (common/model.py:954-980, common/policy.py:74-87).                  */
typedef struct {
  const float* a_hi; const float* a_lo; int64_t lda;
  const float* b_hi; const float* b_lo; int64_t ldb;
  int32_t M, N, K;
  int32_t precision, split_k, flags, block_n;
  int32_t a_mn, b_mn, conv_wgrad;   /* conv_wgrad: see conv_C */
  const float* bias;
  const float* mask; int64_t ld_mask;
  float* out; float* out_hi; float* out_lo; int64_t ldc;
  float* colsum;
  void* dbg;          /* optional int64[16] (int64[96] with a TPP_TC_*_SPLIT flag: + the splitter warps' per-k-block
                         times): clock64 timeline of one CTA (x = 0, y = _reserved, z = 0) -- profiling aid, normally
                         NULL */
  const float* addend; int64_t ld_add;   /* TPP_EPI_ADD (tpp_gemm_tc only): result += addend[m*ld_add + n], applied
                                            after bias / relu / mask: the residual connection (forward) and the
                                            skip-path gradient (backward) of ResidualBlock, common/model.py:134-153 */
  int32_t conv_B, conv_H, conv_W, conv_C;  /* conv_C > 0: implicit 3x3 / pad-1 convolution.  a_hi/a_lo are NHWC fp32
                                            tensors [conv_B][conv_H][conv_W][conv_C] (conv_C in 4..32, multiple of 4);
                                            the A tiles are gathered by TMA im2col loads (no col matrix): M must be
                                            conv_B*conv_H*conv_W, K = 288 = 9 taps x 32 channel slots, and the K-major
                                            B operand holds W[n][tap*32 + c] (zero for c >= conv_C).  lda / a_mn unused.
                                            conv_C <= 16 may instead pass K = 144 = 9 taps x 16 slots with
                                            W[n][tap*16 + c]: 64-byte rows, half the shared-memory fill.
                                            nn.Conv2d(k=3, pad=1) forward and data gradient, common/model.py:137-163.
                                            With conv_wgrad = 1 the same tensor is the MN-major A operand of the
                                            weight gradient: out[tap*32 + c][n] += sum_p X[p + tap][c] * dY[p][n] with
                                            M = 288, K = conv_B*conv_H*conv_W pixels, B = dY [K][ldb] MN-major
                                            (a_mn = b_mn = 1, TPP_EPI_ACCUM, any split_k).                          */
  float alpha; int32_t _reserved;          /* TPP_EPI_ACCUM: out += alpha * product (0 means 1)                      */
  uint32_t* mask_bits_out;                 /* 1-bit ReLU masks (tiles wider than 32 columns; M, N multiples of 32):     */
  const uint32_t* mask_bits;               /* mask_bits_out receives (result > 0) of every element, 32 words per 32 x 32
                                              block ([M/32][N/32][32] uint32, 1/32 of the fp32 activation it stands
                                              for); mask_bits zeroes the result where the bit is clear -- relu'(H) of
                                              the data gradient without re-reading H (TPP_EPI_MASK's 4-byte form).
                                              The word / bit order inside a block is the kernel's own: only a buffer
                                              written through mask_bits_out may be passed as mask_bits.             */
} tpp_tc_gemm;
/* precision: 1 = single-pass TF32, 3 = 3xTF32; with 3, | TPP_TC_A_EXACT / TPP_TC_B_EXACT declares that operand exactly
 * representable in TF32 (e.g. integer pixel values 0..255): it has no lo half (a_lo / b_lo unused, not loaded) and the
 * pass that would multiply it is skipped -- two passes, 3/4 of the operand traffic.                                */
enum { TPP_TC_A_EXACT = 16, TPP_TC_B_EXACT = 32 };
/* with 3, | TPP_TC_A_SPLIT / TPP_TC_B_SPLIT (tiles wider than 32 columns): a_hi / b_hi is the PLAIN fp32 operand and the
 * kernel forms the lo half in shared memory (the tensor core reads an fp32 word as its truncated TF32 value, so the plain
 * tile IS the hi half; four extra warps write lo = tf32_round(x - trunc(x)) next to it before the MMAs of the stage are
 * issued).  Same three passes, half the operand bytes from HBM / L2, no separate split pass.  a_lo / b_lo unused.
 * Measured (profiles/README.md, round 2): the HBM bytes halve but the launch is NOT faster -- the plain tiles and the
 * lo ring share the 227 KB of shared memory and the splitters sit between the TMA landing and the MMAs, so the k-block
 * period is bound by (TMA latency + split + cross-CTA signal + MMA) / stages.  The engine keeps pairs by default.   */
enum { TPP_TC_A_SPLIT = 64, TPP_TC_B_SPLIT = 128 };
/* block_n codes of the CTA-pair tiles */
enum { TPP_TC_TILE_PAIR = 512, TPP_TC_TILE_PAIR_PERSISTENT = 513, TPP_TC_TILE_PAIR64_PERSISTENT = 65,
       /* the persistent pair tile with 8 instead of 16 epilogue warps and 4 KB swizzled transpose patches: THREE
        * instead of two 64 KB operand stages fit (two stages do not cover the TMA latency: the k-block period is
        * (latency + MMAs) / 2 = 2380 cycles against 1840 of MMAs).  For launches with a light epilogue (forward). */
       TPP_TC_TILE_PAIR_PERSISTENT_LEAN = 514 };
int tpp_gemm_tc(const tpp_tc_gemm* g, void* stream);

/* Backward of the policy/value heads ([nh = A+1 <= 16][H] weights) in one kernel: from dhead [mb][ld_head]
 * (tpp_ppo_loss_fwd_bwd) and the latent activations [mb][ldl] it produces dlatent = dhead @ Wh as a TF32 (hi, lo)
 * pair [mb][ld_dz] (optionally masked by relu_mask > 0 and/or also as plain fp32), and accumulates
 * gWh += dhead^T latent, gbh += colsum(dhead), gb_last += colsum(dlatent) (bias gradient of the last embedder layer).
 * H in {16, 32, 64, 128, 256}.  Replaces autograd through CategoricalPolicy.hidden_to_output
 * (common/policy.py:74-87).                                                                                  */
int tpp_head_backward(const float* dhead, int32_t ld_head, const float* latent, const float* relu_mask, int64_t ldl,
                      const float* Wh, int32_t nh, int32_t H, float* dz_hi, float* dz_lo, float* dz_plain,
                      int64_t ld_dz, float* gWh, float* gbh, float* gb_last, int32_t mb, void* stream);

/* x [rows][ld_in] -> (hi, lo) [rows][ld_out] (columns >= cols zero-filled) and/or transposed (t_hi, t_lo)
 * [cols][ld_t] (columns >= rows zero-filled): makes a foreign fp32 tensor (gathered observations, weights after an
 * optimizer step, loss gradients) an operand of tpp_gemm_tc.                                                */
int tpp_split_tf32(const float* x, int64_t ld_in, int32_t rows, int32_t cols, float* hi, float* lo, int64_t ld_out,
                   float* t_hi, float* t_lo, int64_t ld_t, void* stream);

/* ---- policy: IMPALA-CNN pieces (NHWC activations; common/model.py:134-208) ------------------------------ */
/* 3x3 / pad-1 window gather: col[p][tap*C + c] = act(scale * x[b, y+ky-1, x+kx-1, c]) for p = (b*H + y)*W + x,
 * tap = ky*3 + kx (zero outside the image; columns [9C, Kp) zero), written as the TF32 (hi, lo) operand pair
 * [B*H*W][Kp] of tpp_gemm_tc.  x is float32 NHWC, or uint8 NHWC frames when x_is_u8 (scale = 1/255 then fuses
 * ScaledFloatFrame, common/env/procgen_wrappers.py:407-419); relu != 0 applies max(.,0) on load (the nn.ReLU in
 * front of the residual convolutions, common/model.py:146-150).  nn.Conv2d(k=3, pad=1) forward is then
 * Y[p][co] = col . Wf[co] with Wf[co][tap*Cin + ci] = W[co][ci][ky][kx]; the data gradient is the same operation on
 * dY with Wd[ci][tap*Cout + co] = W[co][ci][2-ky][2-kx]; the weight gradient is dY^T col (both MN-major).
 * col_lo may be NULL (single-pass TF32 mode needs only the hi half).                                           */
int tpp_im2col3x3(const void* x, int32_t x_is_u8, int32_t B, int32_t H, int32_t W, int32_t C, int64_t sb, int64_t sy,
                  int64_t sx, int64_t sc, int32_t relu, float scale, float* col_hi, float* col_lo, int32_t Kp,
                  void* stream);
/* (sb, sy, sx, sc): element strides of x over (sample, row, column, channel): NHWC = (HWC, WC, C, 1); the gathered
 * NCHW observation rows of tpp_gather_img = (ld_out, W, 1, HW).                                                 */

/* Weight gradient of a 3x3 / padding-1 convolution on the fp32 FMA pipe (csrc/conv_cc.cu), for the IMPALA-CNN's
 * narrow layers: gw[(ky*3 + kx)*32 + ci][co] += sum_{b,y,x} X[b][y+ky-1][x+kx-1][ci] * dY[b][y][x][co], X = relu(x)
 * when relu != 0.  x: NHWC [B][H][W][cin] plain fp32 (the tensor the forward convolution's input pair was split
 * from), dy: NHWC [B][H][W][cout] plain fp32, gw: the engine's GEMM layout [9*32][cout] (32 channel slots per tap),
 * ACCUMULATED into.  Replaces autograd's conv2d weight gradient (reference common/model.py:134-208).
 * Shapes: (cin, cout, H = W) in {(16,16,32), (16,32,32), (32,32,16), (32,32,8)}; anything else returns TPP_ENOTSUP
 * and the caller uses the tensor-core form (tpp_gemm_tc with conv_wgrad).                                        */
int tpp_conv3x3_wgrad(const float* x, int32_t relu, const float* dy, float* gw, int32_t B, int32_t H, int32_t W,
                      int32_t cin, int32_t cout, void* stream);

/* The same for the network's first convolution (3 -> 16 channels, 64 x 64): x is the channel-planar float observation
 * [B][3][H][W] with element strides sb (sample), sc (channel), sh (row), unit stride along x; gw is the engine's
 * explicit layout [(ky*3 + kx)*3 + ci][cout] (row stride cout), accumulated into.  Other shapes: TPP_ENOTSUP.    */
int tpp_conv3x3_wgrad_first(const float* x, int64_t sb, int64_t sc, int64_t sh, const float* dy, float* gw, int32_t B,
                            int32_t H, int32_t W, int32_t cout, void* stream);

/* Forward of that first convolution: out[b][y][x][co] = bias[co] + sum X[b][ci][y+ky-1][x+kx-1] * w[co][ci][ky][kx],
 * x as above, w = the reference's nn.Conv2d weight [cout][3][3][3] and bias [cout] as they lie in the flat parameter
 * buffer, out = NHWC [B][H][W][cout] plain fp32 (exact fp32 FMAs).  Replaces ImpalaBlock.conv on the observation
 * (reference common/model.py:134-150).  H = W = 64, cout = 16; other shapes: TPP_ENOTSUP.                       */
int tpp_conv3x3_fwd_first(const float* x, int64_t sb, int64_t sc, int64_t sh, const float* w, const float* bias,
                          float* out, int32_t B, int32_t H, int32_t W, int32_t cout, void* stream);

/* Forward / data gradient of a 3x3 / padding-1 convolution with a 16-channel tensor on one side at 32 x 32, on the
 * fp32 FMA pipe (exact fp32): y[p][co] = sum_{ky,kx,ci} X[p + (ky-1, kx-1)][ci] * wg[co][(ky*3 + kx)*slots + ci],
 * X = relu(x) when relu_in.  x: NHWC [B][H][W][cin] plain; wg: the engine's GEMM-layout plain weights (forward: Wf,
 * data gradient: the flipped Wd, with cin / cout swapped by the caller).  Epilogue in the order of tpp_gemm_tc's:
 * + bias (or NULL), zero where mask <= 0 (or NULL; [rows][cout]), + addend (or NULL), colsum[co] += column sums (or
 * NULL), then out (plain, or NULL) and the TF32 pair out_hi / out_lo (or NULL) of relu(y) when pair_relu else of y.
 * Replaces nn.Conv2d forward / autograd data gradient of ResidualBlock / ImpalaBlock (reference
 * common/model.py:134-208).  (cin, cout, H = W) in {(16,16,32), (16,32,32), (32,16,32)}; else TPP_ENOTSUP.      */
int tpp_conv3x3_fma(const float* x, int32_t relu_in, const float* wg, int32_t slots, const float* bias,
                    const float* mask, const float* addend, int32_t pair_relu, float* out, float* out_hi, float* out_lo,
                    float* colsum, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout, void* stream);

/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC (common/model.py:163,171): y [B][(H+1)/2][(W+1)/2][C],
 * arg = winning tap (first maximum); backward routes dy to the winning input pixel (gather form, no atomics).   */
int tpp_maxpool3x3s2_fwd(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, float* y, uint8_t* arg,
                         float* relu_hi, float* relu_lo, void* stream);  /* relu_hi/lo (nullable): TF32 pair of relu(y) */
int tpp_maxpool3x3s2_bwd(const float* dy, const uint8_t* arg, int32_t B, int32_t H, int32_t W, int32_t C, float* dx,
                         float* dx_hi, float* dx_lo, void* stream);   /* dx_hi/dx_lo (nullable): TF32 pair of dx */
/* out[c] += sum_m x[m*C + c] for a narrow row-major [M][C] matrix (C in 4/8/16/32/64): a convolution's bias gradient
 * from its NHWC output gradient (autograd of nn.Conv2d bias, common/model.py:137-138,158).                     */
int tpp_colsum_narrow(const float* x, int64_t M, int32_t C, float* out, void* stream);
/* IMPALA feature sparsity fs = mean_j max_b tanh(|100 h_bj|) over the flattened ReLU'd block-3 features h [M][E]
 * (ImpalaModel.forward_with_attn_indices, common/model.py:203-208; h = h_hi + h_lo, h_lo nullable).  fs_out[0] is
 * written; scratch (uint64 [E]) keeps every column's (max value, first argmax row) for the backward call, which adds
 * d(coef * fs)/dh to the feature gradient dx [M][E] (and refreshes the touched entries of its TF32 pair):
 * the `fs_coef * feature_sparsity` term of the PPO loss (agents/ppo.py:164-169).                                */
int tpp_feature_sparsity(const float* h_hi, const float* h_lo, int32_t M, int32_t E, uint64_t* scratch, float* fs_out,
                         void* stream);
int tpp_feature_sparsity_grad(const uint64_t* scratch, int32_t E, float coef, float* dx, float* dx_hi, float* dx_lo,
                              void* stream);
/* y = act(x + bias) -> plain fp32 (out) and/or TF32 pair (out_hi, out_lo), rows [M][ld]: finishes a dense layer whose
 * contraction was split across CTAs (TPP_EPI_ACCUM) -- the tensor core adds into its fp32 accumulator with
 * truncation, so contractions longer than ~512 terms are chunked and the chunks summed with IEEE adds to stay
 * fp32-grade (IMPALA's 2048-wide fc layer, common/model.py:175).                                               */
int tpp_bias_act_split(const float* x, int64_t ld_in, int32_t M, int32_t N, const float* bias, int32_t relu, float* out,
                       float* out_hi, float* out_lo, int64_t ld_out, void* stream);

/* ---- recurrent policy: GRU cell at prediction time ------------------------------------------------------ */
/* CategoricalPolicy(recurrent=True) (common/policy.py:49-69) runs the embedder's latent through nn.GRU(D, D) in
 * PPO.predict (agents/ppo.py:72-81, common/model.py:219-226: one cell step on hxs * (1 - done)); PPO.optimize does
 * not call it (agents/ppo.py:116-121).  The cell is two tpp_gemm_tc launches (gi = x W_ih^T + b_ih and
 * gh = hm W_hh^T + b_hh, [N][3D], gate order r, z, n) between these two kernels:
 * tpp_gru_mask_split: hm = h * (1 - done) as the TF32 (hi, lo) operand [N][ld] (columns >= D zero); done nullable.
 * tpp_gru_gates: r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r gh_n), h' = (1 - z) n + z hm -> h_out
 *   [N][ldo] plain fp32 (may alias h_prev) and, when non-NULL, the (hi, lo) pair [N][ld] the head GEMM reads.      */
int tpp_gru_mask_split(const float* h, int64_t ldh, const uint8_t* done, int32_t N, int32_t D, float* hm_hi,
                       float* hm_lo, int64_t ld, void* stream);
int tpp_gru_gates(const float* gi, const float* gh, int64_t ldg, const float* h_prev, int64_t ldh, const uint8_t* done,
                  int32_t N, int32_t D, float* h_out, int64_t ldo, float* out_hi, float* out_lo, int64_t ld,
                  void* stream);

/* ---- policy: action sampling at rollout --------------------------------------------------------------- */
/* head: [N][ld_head] rows of (A logits, 1 value).  Writes act int32, logp, value for slot t.
 * Replaces dist.sample()/log_prob in PPO.predict (agents/ppo.py:72-81, common/policy.py:74-87).
 * Philox(seed, env_offset + env, tick + t_offset); greedy != 0 -> argmax.  env_offset: index of row 0 among the
 * rank's envs when the rollout runs in several env ranges (each range keeps its envs' random streams).      */
int tpp_sample_actions(const float* head, int32_t ld_head, int32_t n_envs, int32_t n_actions, int32_t* act,
                       float* logp, float* value, uint64_t seed, const uint64_t* tick, uint64_t t_offset,
                       int32_t greedy, int32_t env_offset, void* stream);

/* Rollout tail of an MLP policy in one launch: z = act(h W^T + b) for the LAST embedder layer (h [n_rows][ldh], K <= 256
 * inputs, K % 4 == 0; W [L][K], L == 64 outputs (else TPP_ENOTSUP); relu != 0 applies max(.,0)), the policy / value heads
 * (Wh [A+1][L], bh [A+1]) and tpp_sample_actions' draw, exact fp32 on the CUDA cores.  head_out (nullable)
 * [n_rows][ld_head] receives the A logits + value.  Replaces MLPModel's last Linear (common/model.py:954-980),
 * CategoricalPolicy.hidden_to_output (common/policy.py:74-87) and dist.sample()/log_prob (agents/ppo.py:77-79).  */
int tpp_mlp_tail_sample(const float* h, int64_t ldh, int32_t K, const float* W, const float* b, int32_t L,
                        int32_t relu, const float* Wh, const float* bh, int32_t n_actions, int32_t n_rows,
                        float* head_out, int32_t ld_head, int32_t* act, float* logp, float* value, uint64_t seed,
                        const uint64_t* tick, uint64_t t_offset, int32_t greedy, int32_t env_offset, void* stream);

/* The whole rollout-step policy in ONE launch (csrc/rollout_fused.cu): a depth-4 MLPModel forward (hidden width 256,
 * latent 64), both heads and tpp_sample_actions' draw for n_rows envs; a cluster of 4 CTAs per 128 envs splits every
 * layer's contraction four ways and exchanges partial sums over distributed shared memory, so activations never leave
 * the SMs.  Replaces PPO.predict's policy(obs) -> Categorical -> sample / log_prob per rollout step
 * (agents/ppo.py:72-81, common/policy.py:61-87, common/model.py:954-980).
 * a1_mode 0: x = row-major fp32 rows [n_rows][ldx] that are EXACT in TF32 (integer pixel values; ldx % 4 == 0, zero
 *            beyond k[0]); a1_mode 1: x = feature-major slot [k[0]][ldx] (k[0] <= 128), split into hi / lo in-kernel.
 * w_hi / w_lo[l]: the layer's TF32 weight pair [n[l]][ldw[l]] (zero beyond k[l]); bias[l] fp32; relu[l] != 0 applies
 * max(., 0).  head_w [A+1][64] / head_b [A+1] plain fp32.  head_out (nullable) [n_rows][ld_head] receives logits + value.
 * Returns TPP_ENOTSUP for other shapes (callers fall back to the per-layer GEMM path).                          */
typedef struct {
  int32_t n_rows, a1_mode;
  const float* x; int64_t ldx;
  const float* w_hi[4]; const float* w_lo[4]; int64_t ldw[4];
  int32_t k[4], n[4];
  const float* bias[4]; int32_t relu[4];
  const float* head_w; const float* head_b; int32_t n_actions, ld_head;
  int32_t* act; float* logp; float* value; float* head_out;
  uint64_t seed; const uint64_t* tick; uint64_t t_offset; int32_t greedy, env_offset;
  void* dbg;      /* nullable: int64 [4][64] clock64 timeline of cluster 0 (profiles/fused_rollout_timeline.py)     */
  int32_t no_pdl, _pad;                   /* != 0: plain stream order (default: programmatic dependent launch -- the
                                             kernel's prologue and weight prefetch overlap the tail of the previous one) */
  void* scratch; int64_t scratch_bytes;   /* exchange slots of the clusters (device memory, stays in L2): 384 KB per
                                             128 envs processed concurrently; 33 x 384 KB covers a full B200            */
} tpp_fused_policy;
int tpp_policy_rollout_fused(const tpp_fused_policy* f, void* stream);

/* ---- PPO loss, fused forward + backward ---------------------------------------------------------------- */
typedef struct {
  float eps_clip, value_coef, entropy_coef, entropy_multiplier, x_entropy_coef;
  int32_t n_actions, mb;
  int32_t _pad;
  /* nullable DEVICE float[5] = (eps_clip, value_coef, entropy_coef, entropy_multiplier, x_entropy_coef): when given,
   * the kernel reads the five coefficients from it at run time instead of from the by-value fields above, so a
   * captured CUDA graph serves every value (agents/ppo.py:97-101 changes entropy_multiplier every optimize() under
   * entropy_scaling).  The host fields still say whether the cross-batch entropy term is in use.               */
  const float* coef_dev;
} tpp_loss_cfg;

/* head [mb][ld_head] = (A logits, value) -> dhead [mb][ld_head] = dLoss/d(logits, value) and raw sums
 * stats[0..3+A] (double, ACCUMULATED): [0] sum min(s1,s2), [1] sum max(v1,v2), [2] sum entropy_b,
 * [3] sample count, [4..4+A) sum_b p_b.  pbar (nullable float[A]): batch-mean probabilities, required when
 * x_entropy_coef != 0 (compute with tpp_ppo_pbar first).
 * Replaces agents/ppo.py:131-170 + common/misc_util.py:32-51 and their autograd.       */
int tpp_ppo_loss_fwd_bwd(const tpp_loss_cfg* cfg, const float* head, int32_t ld_head, const int32_t* act,
                         const float* old_logp, const float* old_value, const float* ret, const float* adv,
                         const float* pbar, float* dhead, double* stats, void* stream);
int tpp_ppo_pbar(const float* head, int32_t ld_head, int32_t mb, int32_t n_actions, float* pbar_sum,
                 void* stream);
/* The same over `groups` consecutive minibatches of cfg->mb samples in one launch (head / per-sample arrays hold
 * groups*mb rows): group g's loss terms are means over ITS mb samples and its sums go to stats + g*stats_stride.
 * Used when the reference accumulates gradients over several minibatches before one optimizer step
 * (agents/ppo.py:111,173-177: the weights do not change in between, so the minibatches can share one forward /
 * backward pass).  groups > 1 needs mb % 256 == 0; x_entropy_coef must be 0.                                   */
int tpp_ppo_loss_fwd_bwd_grouped(const tpp_loss_cfg* cfg, int32_t groups, const float* head, int32_t ld_head,
                                 const int32_t* act, const float* old_logp, const float* old_value, const float* ret,
                                 const float* adv, float* dhead, double* stats, int32_t stats_stride, void* stream);

/* ---- optimizer ------------------------------------------------------------------------------------------ */
typedef struct {
  double lr, beta1, beta2, eps;      /* python floats of torch.optim.Adam (kept in double: torch forms 1-beta and
                                        the bias corrections in double before they meet the fp32 tensors)       */
  float max_grad_norm, grad_scale;   /* grad_scale: 1/world_size after the gradient all-reduce                   */
  int32_t step;       /* optimizer steps taken so far; bumped by tpp_adam_clip_step                               */
  int32_t ticket;     /* completion counter of the Adam kernel (must start at 0)                                   */
  double sqnorm[2];   /* ping-pong accumulators of sum (grad_scale*g)^2                                            */
} tpp_adam_state;     /* lives in DEVICE memory; host updates lr with a small memcpy                      */

/* sqnorm[step & 1] += sum (grad_scale*g)^2.                                                               */
int tpp_grad_sqnorm(tpp_adam_state* state, const float* g, int64_t n, void* stream);
/* clip_grad_norm_(max_grad_norm) + Adam(lr, betas, eps) + zero grad over flat buffers, one pass.
 * Replaces agents/ppo.py:173-176 (torch.nn.utils.clip_grad_norm_, optim.Adam(eps=1e-5)).                 */
int tpp_adam_clip_step(tpp_adam_state* state, float* p, float* g, float* m, float* v, int64_t n, void* stream);
/* The same step that also refreshes the tensor-core operand copies of the parameters it moves: view k says that the
 * flat range [offset, offset + rows*cols) is a row-major [rows][cols] weight whose TF32 (hi, lo) pair lives at
 * hi / lo [rows][ld] (x = p * scale when scale != 0; column c written at col_of[c] when col_of != NULL -- the
 * frame-byte-order copy of a first layer).  Arithmetic of tpp_split_tf32.  Replaces the re-split launches behind every
 * optimizer step (nine for the depth-4 MLP).                                                                    */
#define TPP_MAX_WEIGHT_VIEWS 8
typedef struct {
  int64_t offset; int32_t rows, cols;
  float* hi; float* lo; int64_t ld;
  float scale; int32_t _pad;
  const int32_t* col_of;
} tpp_weight_view;
int tpp_adam_clip_step_views(tpp_adam_state* state, float* p, float* g, float* m, float* v, int64_t n,
                             const tpp_weight_view* views, int32_t n_views, void* stream);

/* ---- multi-GPU: gradient all-reduce over peer memory, fused with the norm reduction ----------------------- */
/* One-shot all-reduce of the flat gradient across `world` (<= 8) GPUs of one node + tpp_grad_sqnorm of the SUM, in one
 * launch (csrc/peer_reduce.cu).  The sharded design's only collective: the sum of the per-rank gradients in front of
 * clip_grad_norm_ + Adam (agents/ppo.py:173-176 under env sharding).
 * staging_ptrs / pad_ptrs: HOST arrays [world] of device addresses as mapped into this process -- rank r's symmetric
 * staging buffer (float [2][n_pad], double-buffered by launch parity) and rank r's signal pad (uint32 words; flags live at
 * pad_offset + rank).  g_local [n]: this rank's gradient, copied out and ZEROED; g_reduced [n]: the sum in rank order
 * (bit-identical on every rank); state->sqnorm[step & 1] += sum (grad_scale * g_reduced)^2.  epoch_counter / ticket2[2] /
 * error_flag: zero-initialised device words owned by the caller (error_flag != 0: a peer never arrived).
 * Every rank must issue the call in the same order; all CTAs of the launch are co-resident (<= 148).              */
int tpp_peer_allreduce_sqnorm(const uint64_t* staging_ptrs, const uint64_t* pad_ptrs, int32_t rank, int32_t world,
                              int32_t pad_offset, float* g_local, float* g_reduced, tpp_adam_state* state, int64_t n,
                              int64_t n_pad, uint32_t* epoch_counter, uint32_t* ticket2, uint32_t* error_flag,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPP_B200_H */
