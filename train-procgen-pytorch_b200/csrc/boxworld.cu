// Box-World on the device: integer grid transition, level replacement with the sequential seed counter,
// frame emit into the rollout, the MT19937/CPython-compatible level generator (host + device), and the
// VecNormalize reward wrapper.
//
// Replaces (reference, numpy + python): BoxWorldVec.step / replace_world_i / increment_seed
// (boxworld/box_world_env_vec.py:70-209, 233-240, 294-297), world_gen / sampling_pairs
// (boxworld/boxworld_gen_vec.py:4-97) and VecNormalize.step_wait (common/env/procgen_wrappers.py:282-342).
// All grid state is integer: results are bit-exact against the oracle (tests/test_boxworld_parity.py).
#include "tpp_common.cuh"

namespace tpp {

constexpr int BW_WARPS = 8;                  // envs per CTA (one warp each)
constexpr uint8_t C_AGENT = 128, C_GOAL = 255, C_GRID = 220;

// ================================================================================================
// Level generator: MT19937 with CPython's seeding and sampling semantics
// ================================================================================================
struct MT {
  uint32_t* mt;   // 624 words provided by the caller (stack on host, shared memory on device)
  int idx;

  __host__ __device__ void init_genrand(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  // random.seed(int): key = little-endian 32-bit words of |seed| (at least one word)
  __host__ __device__ void seed(int64_t sd) {
    uint64_t a = sd < 0 ? (uint64_t)(-sd) : (uint64_t)sd;
    uint32_t key[2] = {(uint32_t)a, (uint32_t)(a >> 32)};
    const int len = key[1] ? 2 : 1;
    init_genrand(19650218u);
    int i = 1, j = 0;
    for (int k = 624; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      ++i; ++j;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
      if (j >= len) j = 0;
    }
    for (int k = 623; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      ++i;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    idx = 624;
  }
  __host__ __device__ uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  __host__ __device__ uint32_t randbelow(uint32_t n) {      // Random._randbelow_with_getrandbits
    int k = 0;
    for (uint32_t t = n; t; t >>= 1) ++k;                   // n.bit_length()
    uint32_t r = next() >> (32 - k);
    while (r >= n) r = next() >> (32 - k);
    return r;
  }
  __host__ __device__ double random() {                     // Random.random: 53-bit
    const uint32_t a = next() >> 5, b = next() >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
  // random.sample(population of n items, k) for n <= 21 ("pool" branch); out receives positions
  __host__ __device__ void sample_pool(int* pool, int n, int k, int* out) {
    for (int i = 0; i < k; ++i) {
      const int j = (int)randbelow((uint32_t)(n - i));
      out[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
  }
};

__constant__ uint8_t d_bw_colors[19][3] = {
    {0, 0, 117},    {230, 190, 255}, {170, 255, 195}, {255, 250, 200}, {255, 216, 177}, {250, 190, 190}, {240, 50, 230},
    {145, 30, 180}, {67, 99, 216},   {66, 212, 244},  {60, 180, 75},   {191, 239, 69},  {255, 255, 25},  {245, 130, 49},
    {230, 25, 75},  {128, 0, 0},     {154, 99, 36},   {128, 128, 0},   {70, 153, 144}};
static const uint8_t h_bw_colors[19][3] = {
    {0, 0, 117},    {230, 190, 255}, {170, 255, 195}, {255, 250, 200}, {255, 216, 177}, {250, 190, 190}, {240, 50, 230},
    {145, 30, 180}, {67, 99, 216},   {66, 212, 244},  {60, 180, 75},   {191, 239, 69},  {255, 255, 25},  {245, 130, 49},
    {230, 25, 75},  {128, 0, 0},     {154, 99, 36},   {128, 128, 0},   {70, 153, 144}};

__host__ __device__ inline const uint8_t* bw_color(int c) {
#ifdef __CUDA_ARCH__
  return d_bw_colors[c];
#else
  return h_bw_colors[c];
#endif
}

constexpr int BW_MAX_N = 20, BW_MAX_PAIRS = 64, BW_MAX_GOAL = 19;

struct LevelSpec { int n, goal_length, num_distractor, distractor_length; };

__host__ __device__ inline bool level_spec_ok(const LevelSpec& s) {
  const int pairs = s.goal_length - 1 + s.num_distractor * s.distractor_length;
  return s.n >= 3 && s.n <= BW_MAX_N && s.goal_length >= 2 && s.goal_length - 1 <= BW_MAX_GOAL && pairs <= BW_MAX_PAIRS &&
         s.num_distractor >= 0 && s.distractor_length >= 0 && s.distractor_length <= 19 - (s.goal_length - 1) &&
         s.num_distractor <= 16;
}

// world: [(n+2)^2][3], dic: [(n+2)^2], pos: [2].  `mtbuf` = 624 words of scratch.
__host__ __device__ inline void gen_level(const LevelSpec& sp, int64_t seed, uint32_t* mtbuf, uint8_t* world,
                                          int8_t* dic, int32_t* pos) {
  const int n = sp.n, S = n + 2, w = n - 1;
  MT rng{mtbuf, 624};
  rng.seed(seed);
  for (int r = 0; r < S; ++r)
    for (int c = 0; c < S; ++c) {
      const bool wall = (r == 0) | (c == 0) | (r == S - 1) | (c == S - 1);
      uint8_t* px = world + (r * S + c) * 3;
      px[0] = px[1] = px[2] = wall ? 0 : C_GRID;
      dic[r * S + c] = -1;
    }
  auto paint = [&](int r, int c, const uint8_t* col) {   // inner-grid coordinates
    uint8_t* px = world + ((r + 1) * S + (c + 1)) * 3;
    px[0] = col[0]; px[1] = col[1]; px[2] = col[2];
  };
  const uint8_t white[3] = {C_GOAL, C_GOAL, C_GOAL}, grey[3] = {C_AGENT, C_AGENT, C_AGENT};

  // goal_colors = sample(range(19), goal_length-1)
  int pool[19], goal_cols[BW_MAX_GOAL];
  for (int i = 0; i < 19; ++i) pool[i] = i;
  const int ng = sp.goal_length - 1;
  rng.sample_pool(pool, 19, ng, goal_cols);
  // distractor colours: sample(colours not on the goal path, distractor_length) per branch
  int free_cols[19], nfree = 0;
  for (int c = 0; c < 19; ++c) {
    bool used = false;
    for (int i = 0; i < ng; ++i) used |= (goal_cols[i] == c);
    if (!used) free_cols[nfree++] = c;
  }
  int dis_cols[16][19], dis_root[16];
  for (int d = 0; d < sp.num_distractor; ++d) {
    for (int i = 0; i < nfree; ++i) pool[i] = free_cols[i];
    rng.sample_pool(pool, nfree, sp.distractor_length, dis_cols[d]);
  }
  // distractor_roots = choices(range(goal_length-1), k=num_distractor)
  for (int d = 0; d < sp.num_distractor; ++d) dis_root[d] = (int)floor(rng.random() * (double)ng);

  // sampling_pairs: population = ascending tuple of the remaining cell codes 1 .. n*(n-1)-1
  uint32_t bits[16];
  for (int i = 0; i < 16; ++i) bits[i] = 0;
  const int ncode = n * w;
  int remaining = 0;
  for (int v = 1; v < ncode; ++v) { bits[v >> 5] |= 1u << (v & 31); ++remaining; }
  auto drop = [&](int v) {
    if (v >= 0 && v < 512 && (bits[v >> 5] >> (v & 31) & 1u)) { bits[v >> 5] &= ~(1u << (v & 31)); --remaining; }
  };
  auto pick = [&]() {   // sample(population, 1)[0] == population[randbelow(len)]
    int j = (int)rng.randbelow((uint32_t)remaining);
    for (int v = 1; v < ncode; ++v)
      if (bits[v >> 5] >> (v & 31) & 1u) {
        if (j == 0) return v;
        --j;
      }
    return 1;
  };
  const int npair = ng + sp.num_distractor * sp.distractor_length;
  int kx[BW_MAX_PAIRS], ky[BW_MAX_PAIRS];
  for (int k = 0; k < npair; ++k) {
    const int key = pick();
    kx[k] = key / w; ky[k] = key % w;
    drop(kx[k] * w + ky[k]);
    const int fwd = (2 < n - 2 - ky[k]) ? 2 : (n - 2 - ky[k]);
    for (int i = 1; i <= fwd; ++i) drop(kx[k] * w + i + ky[k]);
    const int back = (2 < ky[k]) ? 2 : ky[k];
    for (int i = 1; i <= back; ++i) drop(kx[k] * w - i + ky[k]);
  }
  const int agent = pick();
  drop(agent);
  const int first = pick();

  for (int i = 1; i < sp.goal_length; ++i) {                    // goal path
    paint(kx[i - 1], ky[i - 1], i == sp.goal_length - 1 ? white : bw_color(goal_cols[i]));
    paint(kx[i - 1], ky[i - 1] + 1, bw_color(goal_cols[i - 1]));
    dic[(kx[i - 1] + 1) * S + ky[i - 1] + 2] = 1;
  }
  paint(first / w, first % w, bw_color(goal_cols[0]));          // loose first key
  for (int d = 0; d < sp.num_distractor; ++d) {                 // distractor branches
    const int base = ng + d * sp.distractor_length;
    if (sp.distractor_length <= 0) continue;
    paint(kx[base], ky[base] + 1, bw_color(goal_cols[dis_root[d]]));
    paint(kx[base], ky[base], bw_color(dis_cols[d][0]));
    dic[(kx[base] + 1) * S + ky[base] + 2] = 0;
    for (int k = 0; k + 1 < sp.distractor_length; ++k) {
      const int q = base + 1 + k;
      const int key_col = dis_cols[d][k == 0 ? sp.distractor_length - 1 : k - 1];   // python's [k-1] wrap
      paint(kx[q], ky[q], bw_color(key_col));
      paint(kx[q], ky[q] + 1, bw_color(dis_cols[d][k]));
      dic[(kx[q] + 1) * S + ky[q] + 2] = 0;
    }
  }
  paint(agent / w, agent % w, grey);
  pos[0] = agent / w + 1;
  pos[1] = agent % w + 1;
}

// ================================================================================================
// Step kernel: one warp per env.  Lane 0 resolves the move (a handful of byte reads / writes), then the
// warp streams the frame into the rollout slot for envs that continue; finished envs are handled by the
// reset kernel, which needs their global rank in env order.
// ================================================================================================
__device__ __forceinline__ bool px_is(const uint8_t* p, uint8_t g) { return p[0] == g && p[1] == g && p[2] == g; }

// Warp copy of a frame.  Loads of a batch (8 words per lane) are all issued before the first store: a plain
// load -> store loop pays one memory round trip per 128 bytes (the compiler cannot prove the pointers distinct).
__device__ __forceinline__ void copy_frame(const uint8_t* src, uint8_t* dst, int bytes, int lane, uint8_t* dst2 = nullptr) {
  if ((bytes & 3) == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(dst2);
    const int words = bytes >> 2;
    for (int base = 0; base < words; base += 256) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = base + lane + 32 * j;
        v[j] = i < words ? s[i] : 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = base + lane + 32 * j;
        if (i < words) {
          d[i] = v[j];
          if (d2) d2[i] = v[j];
        }
      }
    }
  } else {
    for (int i = lane; i < bytes; i += 32) {
      const uint8_t b = src[i];
      dst[i] = b;
      if (dst2) dst2[i] = b;
    }
  }
}

// One env's transition by ONE thread: every byte it needs (action, position, counters, the target cell and its two
// neighbours, the lock table entry, the owned key) is loaded before the move is resolved -- two dependent rounds of
// global loads per env, and as many envs in flight as there are threads (the warp-per-env form kept 31 of 32 lanes
// waiting behind lane 0's chain: 0.19 of HBM peak at 2^18 envs).  Returns the env's done flag.
__device__ __forceinline__ int step_env(const tpp_boxworld_state& st, int e, const int32_t* __restrict__ action,
                                        int32_t* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                                        int32_t* fin_ret, int32_t* fin_len, uint8_t* fin_solved) {
  const int S = st.n + 2, cells = S * S;
  uint8_t* W = st.world + (int64_t)e * cells * 3;
  int done = 0;
  const int8_t* D = st.world_dic + (int64_t)e * cells;
  const int a = action[e];
  const int pr = st.player_pos[2 * e], pc = st.player_pos[2 * e + 1];
  const int nr = pr + (a == 0 ? -1 : (a == 1 ? 1 : 0)), nc = pc + (a == 2 ? -1 : (a == 3 ? 1 : 0));
  const int steps = st.num_env_steps[e] + 1;
  int reward = 0, solved = 0;
  done = (steps == st.max_steps);
  auto clampi = [&](int v) { return v < 0 ? 0 : (v > st.n + 1 ? st.n + 1 : v); };
  const int ar = clampi(nr), ac = clampi(nc);
  const int at = ar * S + ac, left = ar * S + clampi(nc - 1), right = ar * S + clampi(nc + 1);
  const bool in_grid = nr > 0 && nc > 0 && nr <= st.n && nc <= st.n;
  const uint8_t here[3] = {W[at * 3], W[at * 3 + 1], W[at * 3 + 2]};
  const uint8_t lc[3] = {W[left * 3], W[left * 3 + 1], W[left * 3 + 2]};
  const bool empty = px_is(here, C_GRID);
  const bool left_clear = (nc == 1) || px_is(lc, C_GRID);
  const bool first_key = !empty && left_clear && (px_is(W + right * 3, C_GRID) || px_is(W + right * 3, C_AGENT));
  const int status = D[at];
  const bool is_lock = status != -1;
  uint8_t* own = st.owned_key + 4 * e;
  const bool key_fits = own[0] == here[0] && own[0] != C_GRID && own[1] == here[1] && own[1] != C_GRID &&
                        own[2] == here[2] && own[2] != C_GRID;
  const bool blocked = !(empty || first_key || is_lock) || (is_lock && !key_fits);
  if (in_grid && !blocked) {
    const int cur = pr * S + pc;
    const bool walk = empty, take = !walk && first_key, unlock = !walk && !take && is_lock && key_fits;
    if (walk || take || unlock) {
      W[cur * 3] = W[cur * 3 + 1] = W[cur * 3 + 2] = C_GRID;
      if (unlock) W[left * 3] = W[left * 3 + 1] = W[left * 3 + 2] = C_GRID;
      W[at * 3] = W[at * 3 + 1] = W[at * 3 + 2] = C_AGENT;
      st.player_pos[2 * e] = nr;
      st.player_pos[2 * e + 1] = nc;
      if (take) {
        W[0] = own[0] = here[0]; W[1] = own[1] = here[1]; W[2] = own[2] = here[2];
        reward += 1;
      }
      if (unlock) {
        W[0] = own[0] = lc[0]; W[1] = own[1] = lc[1]; W[2] = own[2] = lc[2];
        const bool goal = px_is(lc, C_GOAL);
        if (goal) { reward += 10; solved = 1; done = 1; }
        if (status == 1) reward += 1;
        if (status == 0) { reward -= 1; done = 1; }
      }
    }
  }
  const int ep = st.episode_reward[e] + reward;
  st.num_env_steps[e] = steps;
  st.episode_reward[e] = ep;
  reward_out[e] = reward;
  done_out[e] = (uint8_t)done;
  if (fin_ret) fin_ret[e] = done ? ep : 0;
  if (fin_len) fin_len[e] = done ? steps : 0;
  if (fin_solved) fin_solved[e] = (uint8_t)(done ? solved : 0);
  return done;
}

// An env's frame as the policy's next input row: channel-major integer pixel values as fp32 (TransposeFrame;
// ScaledFloatFrame's 1/255 lives in the policy's first layer) = tpp_frames_to_obs(raw) of that frame.
__device__ __forceinline__ void emit_obs_row(const uint8_t* W, float* __restrict__ dst, int cells, int ld_obs, int lane) {
  constexpr int OB = 20;                    // rows up to 640 floats: all byte loads first, then the stores
  if (ld_obs <= OB * 32) {
    uint8_t px[OB];
    int c = 0, p = lane;
#pragma unroll
    for (int i = 0; i < OB; ++i) {
      const int o = lane + 32 * i;
      px[i] = 0;
      if (o < cells * 3) {
        while (p >= cells) { p -= cells; ++c; }
        px[i] = W[p * 3 + c];
      }
      p += 32;
    }
#pragma unroll
    for (int i = 0; i < OB; ++i) {
      const int o = lane + 32 * i;
      if (o < ld_obs) dst[o] = (float)px[i];
    }
  } else {
    int c = 0, p = lane;
    for (int o = lane; o < ld_obs; o += 32) {
      float v = 0.0f;
      if (o < cells * 3) {
        while (p >= cells) { p -= cells; ++c; }
        v = (float)W[p * 3 + c];
      }
      p += 32;
      dst[o] = v;
    }
  }
}

// Level replacement of finished env e, the k-th finished env of this step in env order (its level seed follows the
// reference's sequential counter, box_world_env_vec.py:204-207,294-297), executed by one warp.
__device__ __forceinline__ void reset_env(const tpp_boxworld_state& st, int e, int64_t k, int64_t sc, int lane,
                                          uint32_t* mtbuf_w, uint8_t* __restrict__ frame_out) {
  const int S = st.n + 2, cells = S * S;
  int64_t seed = sc + k;
  if (st.n_levels > 0) seed = ((sc - st.start_seed + k) % st.n_levels) + st.start_seed;
  uint8_t* W = st.world + (int64_t)e * cells * 3;
  int8_t* D = st.world_dic + (int64_t)e * cells;
  if (st.n_levels > 0 && st.bank_world) {
    // level bank: frame, lock table and position loads are all in flight together; the frame goes to the env state
    // AND to the rollout slot from the same registers
    const int64_t li = seed - st.start_seed;
    const int8_t* bd = st.bank_dic + li * cells;
    int8_t dv[8];
    int pos = 0;
    if (cells <= 256) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[j] = (lane + 32 * j) < cells ? bd[lane + 32 * j] : (int8_t)0;
    }
    if (lane < 2) pos = st.bank_pos[2 * li + lane];
    copy_frame(st.bank_world + li * cells * 3, W, cells * 3, lane,
               frame_out ? frame_out + (int64_t)e * cells * 3 : nullptr);
    if (cells <= 256) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if ((lane + 32 * j) < cells) D[lane + 32 * j] = dv[j];
    } else {
      for (int i = lane; i < cells; i += 32) D[i] = bd[i];
    }
    if (lane < 2) st.player_pos[2 * e + lane] = pos;
    if (lane == 0) {
      st.num_env_steps[e] = 0;
      st.episode_reward[e] = 0;
      uint8_t* own = st.owned_key + 4 * e;
      own[0] = own[1] = own[2] = C_GRID;
    }
    __syncwarp();
    return;
  } else if (lane == 0) {
    const LevelSpec sp{st.n, st.goal_length, st.num_distractor, st.distractor_length};
    gen_level(sp, seed, mtbuf_w, W, D, st.player_pos + 2 * e);
  }
  if (lane == 0) {
    st.num_env_steps[e] = 0;
    st.episode_reward[e] = 0;
    uint8_t* own = st.owned_key + 4 * e;
    own[0] = own[1] = own[2] = C_GRID;
  }
  __syncwarp();
  if (frame_out) copy_frame(W, frame_out + (int64_t)e * cells * 3, cells * 3, lane);
}

// Step kernel.  A CTA owns GROUP consecutive envs: thread-per-env transitions, then the whole CTA copies the group's
// frames -- one contiguous span of the world array -- into the rollout slot with 16-byte accesses, every load of a
// batch in flight before the first store.  (Finished envs are copied too; the reset kernel overwrites their frames.)
// The group's finished-env count goes to scratch[group] for the reset kernel's in-order ranks.
template <int GROUP>
__global__ void __launch_bounds__(256) boxworld_step_kernel(tpp_boxworld_state st, const int32_t* __restrict__ action,
                                                            int32_t* __restrict__ reward_out,
                                                            uint8_t* __restrict__ done_out,
                                                            uint8_t* __restrict__ frame_out, int32_t* fin_ret,
                                                            int32_t* fin_len, uint8_t* fin_solved) {
  __shared__ int cta_done;
  if (threadIdx.x == 0) cta_done = 0;
  // programmatic dependent launch (rollout chain policy -> step -> reset -> policy): harmless when launched plainly
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  const int e0 = blockIdx.x * GROUP, e = e0 + threadIdx.x;
  const int cells = (st.n + 2) * (st.n + 2);
  if (threadIdx.x < GROUP && e < st.n_envs) {
    const int done = step_env(st, e, action, reward_out, done_out, fin_ret, fin_len, fin_solved);
    if (done) atomicAdd(&cta_done, 1);
  }
  __syncthreads();               // the group's cells are updated (same-SM stores are visible to the loads below)
  if (threadIdx.x == 0) st.scratch[blockIdx.x] = cta_done;
  if (!frame_out) return;
  const int n_here = min(GROUP, st.n_envs - e0);
  const int64_t bytes = (int64_t)n_here * cells * 3;
  const uint8_t* src = st.world + (int64_t)e0 * cells * 3;
  uint8_t* dst = frame_out + (int64_t)e0 * cells * 3;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)bytes) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const int n16 = (int)(bytes >> 4);
    for (int base = 0; base < n16; base += 256 * 4) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = base + threadIdx.x + 256 * j;
        if (i < n16) v[j] = s4[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = base + threadIdx.x + 256 * j;
        if (i < n16) d4[i] = v[j];
      }
    }
  } else if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)bytes) & 3) == 0) {
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
    for (int i = threadIdx.x; i < (int)(bytes >> 2); i += 256) d1[i] = s1[i];
  } else {
    for (int64_t i = threadIdx.x; i < bytes; i += 256) dst[i] = src[i];
  }
}

// Exclusive prefix sum of the per-group finished-env counts (single CTA; only launched when there are many groups, where
// letting every reset CTA re-sum its predecessors would be quadratic).  scratch: [0, n) counts, [n] ticket,
// [n+1, 2n+1) prefix.
__global__ void __launch_bounds__(1024) boxworld_scan_kernel(int32_t* scratch, int n) {
  __shared__ int warp_off[32];
  __shared__ int chunk_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? scratch[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_off[w] = incl;                 // warp totals
    __syncthreads();
    if (w == 0) {
      const int t = warp_off[lane];
      int inc2 = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += u;
      }
      warp_off[lane] = inc2 - t;                        // exclusive offset of each warp inside the chunk
      if (lane == 31) chunk_total = inc2;
    }
    __syncthreads();
    if (i < n) scratch[n + 1 + i] = carry + warp_off[w] + incl - v;
    carry += chunk_total;
    __syncthreads();
  }
}

// Reset kernel: same groups.  Rank of a finished env = (# finished envs in lower groups) + (# finished before it in its
// group); its level seed follows the reference's sequential counter.  Thread-per-env flags and ranks, then the CTA's
// warps replace the levels of the (few) finished envs from the compacted list.
template <int GROUP>
__global__ void __launch_bounds__(256) boxworld_reset_kernel(tpp_boxworld_state st, const uint8_t* __restrict__ done_in,
                                                             uint8_t* __restrict__ frame_out, int use_prefix,
                                                             float* __restrict__ obs_out, int ld_obs) {
  __shared__ int red[32];
  __shared__ int base_s, n_fin;
  __shared__ int warp_cnt[8];
  __shared__ int fin_env[GROUP], fin_rank[GROUP];
  extern __shared__ uint32_t mtbuf_dyn[];          // [8][624], only when levels are generated on the device
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ncta = gridDim.x;
  if (use_prefix) {          // exclusive prefix of the per-group counts was produced by boxworld_scan_kernel
    if (threadIdx.x == 0) base_s = st.scratch[ncta + 1 + blockIdx.x];
  } else {                   // few groups: sum the counts of the lower groups directly
    int part = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += blockDim.x) part += st.scratch[c];
    part = block_sum(part, red);
    if (threadIdx.x == 0) base_s = part;
  }
  const int64_t sc = *st.seed_counter;
  const int e0 = blockIdx.x * GROUP, e = e0 + threadIdx.x;
  const int S = st.n + 2, cells = S * S;
  const int mine = (threadIdx.x < GROUP && e < st.n_envs) ? (done_in[e] != 0) : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, mine);
  if (lane == 0) warp_cnt[wid] = __popc(bal);
  __syncthreads();
  int before = __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < wid; ++w) before += warp_cnt[w];
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += warp_cnt[w];
    n_fin = t;
  }
  if (mine) { fin_env[before] = e; fin_rank[before] = before; }
  __syncthreads();
  for (int k = wid; k < n_fin; k += 8)
    reset_env(st, fin_env[k], (int64_t)base_s + fin_rank[k], sc, lane, mtbuf_dyn + wid * 624, frame_out);
  if (obs_out) {          // every env: its (post-reset) frame as the policy's next input row
    __syncthreads();
    for (int k = wid; k < GROUP; k += 8) {
      const int ee = e0 + k;
      if (ee < st.n_envs) emit_obs_row(st.world + (int64_t)ee * cells * 3, obs_out + (int64_t)ee * ld_obs, cells, ld_obs, lane);
    }
  }
  // last CTA to finish advances the seed counter by the total number of finished envs
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(reinterpret_cast<unsigned int*>(st.scratch + ncta), 1u) == (unsigned)ncta - 1);
  __syncthreads();
  if (last) {
    int tot = 0;
    for (int c = threadIdx.x; c < ncta; c += blockDim.x) tot += st.scratch[c];
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      int64_t nsc = sc + tot;
      if (st.n_levels > 0) nsc = ((sc - st.start_seed + tot) % st.n_levels) + st.start_seed;
      *st.seed_counter = nsc;
      st.scratch[ncta] = 0;
    }
  }
}

__global__ void __launch_bounds__(BW_WARPS * 32) boxworld_emit_kernel(tpp_boxworld_state st, uint8_t* frame_out) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * BW_WARPS + (threadIdx.x >> 5);
  if (e >= st.n_envs) return;
  const int cells = (st.n + 2) * (st.n + 2);
  copy_frame(st.world + (int64_t)e * cells * 3, frame_out + (int64_t)e * cells * 3, cells * 3, lane);
}

__global__ void __launch_bounds__(128) boxworld_gen_kernel(tpp_boxworld_state st, const int32_t* env_ids,
                                                           const int64_t* seeds, int count) {
  __shared__ uint32_t mtbuf[4][624];
  // one level per warp-lane-0 (the generator is sequential); 4 levels per CTA
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + wid;
  if (i >= count || lane != 0) return;
  const int e = env_ids ? env_ids[i] : i;
  const int cells = (st.n + 2) * (st.n + 2);
  const LevelSpec sp{st.n, st.goal_length, st.num_distractor, st.distractor_length};
  gen_level(sp, seeds[i], mtbuf[wid], st.world + (int64_t)e * cells * 3, st.world_dic + (int64_t)e * cells,
            st.player_pos + 2 * e);
  st.num_env_steps[e] = 0;
  st.episode_reward[e] = 0;
  uint8_t* own = st.owned_key + 4 * e;
  own[0] = own[1] = own[2] = C_GRID;
  own[3] = 0;
}

// ================================================================================================
// VecNormalize (returns only): single CTA, float64 statistics like the reference
// ================================================================================================
__global__ void __launch_bounds__(1024) vecnormalize_kernel(double* ret, double* rms, const void* raw_rew, int raw_is_int,
                                                            const uint8_t* done, float* out_rew, int n, double gamma,
                                                            double cliprew, double epsilon) {
  __shared__ double red[32];
  __shared__ double bc[2];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = raw_is_int ? (double)reinterpret_cast<const int32_t*>(raw_rew)[i]
                                : (double)reinterpret_cast<const float*>(raw_rew)[i];
    const double v = ret[i] * gamma + r;
    ret[i] = v;
    s += v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) bc[0] = s / (double)n;
  __syncthreads();
  const double bmean = bc[0];
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = ret[i] - bmean;
    q += d * d;
  }
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    const double bvar = q / (double)n, bn = (double)n;
    const double mean = rms[0], var = rms[1], count = rms[2];
    const double delta = bmean - mean, tot = count + bn;
    const double new_mean = mean + delta * bn / tot;
    const double m2 = var * count + bvar * bn + delta * delta * count * bn / tot;
    rms[0] = new_mean;
    rms[1] = m2 / tot;
    rms[2] = tot;
    bc[1] = sqrt(m2 / tot + epsilon);
  }
  __syncthreads();
  const double sd = bc[1];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = raw_is_int ? (double)reinterpret_cast<const int32_t*>(raw_rew)[i]
                                : (double)reinterpret_cast<const float*>(raw_rew)[i];
    double o = r / sd;
    o = o < -cliprew ? -cliprew : (o > cliprew ? cliprew : o);
    out_rew[i] = (float)o;
    if (done[i]) ret[i] = 0.0;
  }
}

// The T steps of a finished rollout, one after the other, in ONE launch: per step the same three passes as above.  A
// thread keeps its envs' running returns in registers (4 x 1024 envs) and loads step t + 1's rewards / dones while the
// block reduces step t (the loop is otherwise one L2 round trip per pass).
__global__ void __launch_bounds__(1024) vecnormalize_rollout_kernel(double* ret, double* rms,
                                                                    const int32_t* __restrict__ raw_rew,
                                                                    const uint8_t* __restrict__ done,
                                                                    float* __restrict__ out_rew,
                                                                    float* __restrict__ out_raw, int T, int n, int64_t ld,
                                                                    double gamma, double cliprew, double epsilon) {
  constexpr int EPT = 4;                       // envs per thread (n <= 4096 takes this path)
  __shared__ double red[32];
  __shared__ double bc[2];
  double mean = rms[0], var = rms[1], count = rms[2];      // running moments: maintained by thread 0
  const double bn = (double)n;
  if (n <= EPT * 1024) {
    // Rewards are small integers: the 32 values -8 .. 23 are normalised once per step by the lanes of warp 0 (which
    // also keep the running moments, all lanes alike); every other reward value takes the division itself.
    __shared__ float tab[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double rr[EPT];
    int32_t rw_c[EPT];
    uint8_t dn_c[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int i = threadIdx.x + j * 1024;
      rr[j] = i < n ? ret[i] : 0.0;
      rw_c[j] = i < n ? raw_rew[i] : 0;
      dn_c[j] = i < n ? done[i] : 0;
    }
    for (int t = 0; t < T; ++t) {
      int32_t rw_n[EPT];
      uint8_t dn_n[EPT];
      const int tn = t + 1 < T ? t + 1 : t;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {          // next step's inputs: in flight during this step's reductions
        const int i = threadIdx.x + j * 1024;
        rw_n[j] = i < n ? raw_rew[(int64_t)tn * ld + i] : 0;
        dn_n[j] = i < n ? done[(int64_t)tn * ld + i] : 0;
      }
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int i = threadIdx.x + j * 1024;
        if (i < n) { rr[j] = rr[j] * gamma + (double)rw_c[j]; s += rr[j]; }
      }
      s = block_sum(s, red);
      if (threadIdx.x == 0) bc[0] = s / bn;
      __syncthreads();
      const double bmean = bc[0];
      double q = 0.0;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int i = threadIdx.x + j * 1024;
        if (i < n) { const double d = rr[j] - bmean; q += d * d; }
      }
      q = block_sum(q, red);
      if (wid == 0) {                          // Chan merge of the batch moments into the running ones
        q = __shfl_sync(0xffffffffu, q, 0);
        const double bvar = q / bn;
        const double delta = bmean - mean, tot = count + bn;
        const double m2 = var * count + bvar * bn + delta * delta * count * bn / tot;
        mean = mean + delta * bn / tot;
        var = m2 / tot;
        count = tot;
        const double sd = sqrt(var + epsilon);
        double v = (double)(lane - 8) / sd;
        v = v < -cliprew ? -cliprew : (v > cliprew ? cliprew : v);
        tab[lane] = (float)v;
        if (lane == 0) bc[1] = sd;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int i = threadIdx.x + j * 1024;
        if (i < n) {
          const int r = rw_c[j];
          float o;
          if ((unsigned)(r + 8) < 32u) {
            o = tab[r + 8];
          } else {
            double v = (double)r / bc[1];
            v = v < -cliprew ? -cliprew : (v > cliprew ? cliprew : v);
            o = (float)v;
          }
          out_rew[(int64_t)t * ld + i] = o;
          if (out_raw) out_raw[(int64_t)t * ld + i] = (float)r;
          if (dn_c[j]) rr[j] = 0.0;
        }
        rw_c[j] = rw_n[j];
        dn_c[j] = dn_n[j];
      }
      __syncthreads();      // bc[] / tab[] are rewritten by the next step
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int i = threadIdx.x + j * 1024;
      if (i < n) ret[i] = rr[j];
    }
  } else {
    for (int t = 0; t < T; ++t) {
      const int32_t* rw = raw_rew + (int64_t)t * ld;
      const uint8_t* dn = done + (int64_t)t * ld;
      double s = 0.0;
      for (int i = threadIdx.x; i < n; i += 1024) { const double v = ret[i] * gamma + (double)rw[i]; ret[i] = v; s += v; }
      s = block_sum(s, red);
      if (threadIdx.x == 0) bc[0] = s / bn;
      __syncthreads();
      const double bmean = bc[0];
      double q = 0.0;
      for (int i = threadIdx.x; i < n; i += 1024) { const double d = ret[i] - bmean; q += d * d; }
      q = block_sum(q, red);
      if (threadIdx.x == 0) {                  // Chan merge of the batch moments into the running ones (one thread)
        const double bvar = q / bn;
        const double delta = bmean - mean, tot = count + bn;
        const double m2 = var * count + bvar * bn + delta * delta * count * bn / tot;
        mean = mean + delta * bn / tot;
        var = m2 / tot;
        count = tot;
        bc[1] = sqrt(var + epsilon);
      }
      __syncthreads();
      const double sd = bc[1];
      for (int i = threadIdx.x; i < n; i += 1024) {
        const double r = (double)rw[i];
        double v = r / sd;
        v = v < -cliprew ? -cliprew : (v > cliprew ? cliprew : v);
        out_rew[(int64_t)t * ld + i] = (float)v;
        if (out_raw) out_raw[(int64_t)t * ld + i] = (float)r;
        if (dn[i]) ret[i] = 0.0;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) { rms[0] = mean; rms[1] = var; rms[2] = count; }
}

// ---- the same rollout normalisation as four short parallel launches (needs T*ld + 3T doubles of scratch) ------------
// The single-CTA kernel above walks the T steps one after the other with two block reductions per step (2.8 us per
// step, 720 us per 256-step rollout = 1.5 % of a Box-World iteration).  Only the Chan merge of the per-step batch
// moments is sequential in t; the discounted returns are T-step chains per ENV, the batch moments reductions per STEP.
__global__ void __launch_bounds__(128) vn_returns_kernel(double* __restrict__ ret, const int32_t* __restrict__ raw_rew,
                                                         const uint8_t* __restrict__ done, double* __restrict__ R,
                                                         int T, int n, int64_t ld, double gamma) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = ret[i];
  int t = 0;
  for (; t + 8 <= T; t += 8) {                 // the loads do not depend on the chain: 8 steps in flight
    int32_t rw[8];
    uint8_t dn[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      rw[u] = raw_rew[(int64_t)(t + u) * ld + i];
      dn[u] = done[(int64_t)(t + u) * ld + i];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      r = r * gamma + (double)rw[u];
      R[(int64_t)(t + u) * ld + i] = r;
      if (dn[u]) r = 0.0;
    }
  }
  for (; t < T; ++t) {
    r = r * gamma + (double)raw_rew[(int64_t)t * ld + i];
    R[(int64_t)t * ld + i] = r;
    if (done[(int64_t)t * ld + i]) r = 0.0;
  }
  ret[i] = r;
}

__global__ void __launch_bounds__(1024) vn_moments_kernel(const double* __restrict__ R, double* __restrict__ stats,
                                                          int n, int64_t ld) {
  // one CTA per step: batch mean, then the sum of squared deviations from it (np.var's two passes)
  __shared__ double red[32];
  __shared__ double bm;
  const double* row = R + (int64_t)blockIdx.x * ld;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += row[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) bm = s / (double)n;
  __syncthreads();
  const double bmean = bm;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = row[i] - bmean; q += d * d; }
  q = block_sum(q, red);
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = bmean; stats[2 * blockIdx.x + 1] = q; }
}

__global__ void vn_merge_kernel(double* rms, const double* __restrict__ stats, double* __restrict__ sd, int T, int n,
                                double epsilon) {
  // Chan merge of the T batch moments into the running ones, in order (RunningMeanStd.update, procgen_wrappers.py)
  if (threadIdx.x != 0) return;
  double mean = rms[0], var = rms[1], count = rms[2];
  const double bn = (double)n;
  for (int t = 0; t < T; ++t) {
    const double bmean = stats[2 * t], bvar = stats[2 * t + 1] / bn;
    const double delta = bmean - mean, tot = count + bn;
    const double m2 = var * count + bvar * bn + delta * delta * count * bn / tot;
    mean = mean + delta * bn / tot;
    var = m2 / tot;
    count = tot;
    sd[t] = sqrt(var + epsilon);
  }
  rms[0] = mean; rms[1] = var; rms[2] = count;
}

__global__ void __launch_bounds__(256) vn_apply_kernel(const int32_t* __restrict__ raw_rew, const double* __restrict__ sd,
                                                       float* __restrict__ out_rew, float* __restrict__ out_raw, int n,
                                                       int64_t ld, double cliprew) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t o = (int64_t)blockIdx.y * ld + i;
  const int r = raw_rew[o];
  double v = (double)r / sd[blockIdx.y];
  v = v < -cliprew ? -cliprew : (v > cliprew ? cliprew : v);
  out_rew[o] = (float)v;
  if (out_raw) out_raw[o] = (float)r;
}

}  // namespace tpp

static int bw_check(const tpp_boxworld_state* st) {
  if (!st || !st->world || !st->world_dic || !st->player_pos || !st->owned_key || !st->num_env_steps ||
      !st->episode_reward || !st->seed_counter || !st->scratch)
    return TPP_EINVAL;
  if (st->n_envs <= 0) return TPP_EINVAL;
  const tpp::LevelSpec sp{st->n, st->goal_length, st->num_distractor, st->distractor_length};
  if (!tpp::level_spec_ok(sp)) return TPP_ENOTSUP;
  if (st->n_levels > 0 && !(st->bank_world && st->bank_dic && st->bank_pos)) return TPP_EINVAL;
  return TPP_OK;
}

extern "C" int tpp_boxworld_step(const tpp_boxworld_state* st, const int32_t* action, int32_t* reward_out,
                                 uint8_t* done_out, uint8_t* frame_out, int32_t* fin_ret, int32_t* fin_len,
                                 uint8_t* fin_solved, float* obs_out, int32_t ld_obs, void* stream) {
  const int rc = bw_check(st);
  if (rc) return rc;
  TPP_CHECK_ARG(action && reward_out && done_out);
  TPP_CHECK_ARG(!obs_out || ld_obs >= 3 * (st->n + 2) * (st->n + 2));
  // scratch holds 2 N + 4096 ints (see tpp_boxworld_state)
  cudaStream_t s = tpp_stream(stream);
  // groups of 64 envs per CTA while that still fills the machine, 256 beyond (the C3-sized sweeps)
  const bool big = st->n_envs > 64 * 148 * 4;
  const int group = big ? 256 : 64;
  const int grid = tpp_ceil_div(st->n_envs, group);
  const int use_prefix = grid > 1024;
  const size_t dyn = (st->n_levels > 0 && st->bank_world) ? 0 : sizeof(uint32_t) * 8 * 624;
  if (big) {
    tpp::boxworld_step_kernel<256><<<grid, 256, 0, s>>>(*st, action, reward_out, done_out, frame_out, fin_ret, fin_len,
                                                        fin_solved);
    if (use_prefix) tpp::boxworld_scan_kernel<<<1, 1024, 0, s>>>(st->scratch, grid);
    tpp::boxworld_reset_kernel<256><<<grid, 256, dyn, s>>>(*st, done_out, frame_out, use_prefix, obs_out, ld_obs);
  } else {
    tpp::boxworld_step_kernel<64><<<grid, 256, 0, s>>>(*st, action, reward_out, done_out, frame_out, fin_ret, fin_len,
                                                       fin_solved);
    if (use_prefix) tpp::boxworld_scan_kernel<<<1, 1024, 0, s>>>(st->scratch, grid);
    tpp::boxworld_reset_kernel<64><<<grid, 256, dyn, s>>>(*st, done_out, frame_out, use_prefix, obs_out, ld_obs);
  }
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_boxworld_gen_levels_host(int32_t n, int32_t goal_length, int32_t num_distractor,
                                            int32_t distractor_length, int64_t seed0, int32_t count, uint8_t* world,
                                            int8_t* dic, int32_t* pos) {
  TPP_CHECK_ARG(world && dic && pos && count > 0);
  const tpp::LevelSpec sp{n, goal_length, num_distractor, distractor_length};
  if (!tpp::level_spec_ok(sp)) return TPP_ENOTSUP;
  const int cells = (n + 2) * (n + 2);
  uint32_t mtbuf[624];
  for (int i = 0; i < count; ++i)
    tpp::gen_level(sp, seed0 + i, mtbuf, world + (int64_t)i * cells * 3, dic + (int64_t)i * cells, pos + 2 * i);
  return TPP_OK;
}

extern "C" int tpp_boxworld_gen_levels_device(const tpp_boxworld_state* st, const int32_t* env_ids,
                                              const int64_t* seeds, int32_t count, void* stream) {
  const int rc = bw_check(st);
  if (rc) return rc;
  TPP_CHECK_ARG(seeds && count > 0);
  tpp::boxworld_gen_kernel<<<tpp_ceil_div(count, 4), 128, 0, tpp_stream(stream)>>>(*st, env_ids, seeds, count);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_boxworld_emit_frames(const tpp_boxworld_state* st, uint8_t* frame_out, void* stream) {
  const int rc = bw_check(st);
  if (rc) return rc;
  TPP_CHECK_ARG(frame_out);
  tpp::boxworld_emit_kernel<<<tpp_ceil_div(st->n_envs, tpp::BW_WARPS), tpp::BW_WARPS * 32, 0, tpp_stream(stream)>>>(
      *st, frame_out);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_vecnormalize_rollout(double* ret, double* rms, const int32_t* raw_rew, const uint8_t* done,
                                        float* out_rew, float* out_raw, int32_t T, int32_t n_envs, int64_t ld,
                                        double gamma, double cliprew, double epsilon, double* scratch,
                                        int64_t scratch_doubles, void* stream) {
  TPP_CHECK_ARG(ret && rms && raw_rew && done && out_rew && T > 0 && n_envs > 0 && ld >= n_envs);
  if (scratch && scratch_doubles >= (int64_t)T * ld + 3 * (int64_t)T && T <= 65535) {
    cudaStream_t s = tpp_stream(stream);
    double* R = scratch;
    double* stats = scratch + (int64_t)T * ld;
    double* sd = stats + 2 * (int64_t)T;
    tpp::vn_returns_kernel<<<tpp_ceil_div(n_envs, 128), 128, 0, s>>>(ret, raw_rew, done, R, T, n_envs, ld, gamma);
    tpp::vn_moments_kernel<<<T, n_envs >= 4096 ? 1024 : 256, 0, s>>>(R, stats, n_envs, ld);
    tpp::vn_merge_kernel<<<1, 32, 0, s>>>(rms, stats, sd, T, n_envs, epsilon);
    tpp::vn_apply_kernel<<<dim3(tpp_ceil_div(n_envs, 256), T), 256, 0, s>>>(raw_rew, sd, out_rew, out_raw, n_envs, ld,
                                                                           cliprew);
    TPP_LAUNCH_STATUS();
  }
  TPP_CHECK_ARG(n_envs <= 65536);
  tpp::vecnormalize_rollout_kernel<<<1, 1024, 0, tpp_stream(stream)>>>(ret, rms, raw_rew, done, out_rew, out_raw, T,
                                                                      n_envs, ld, gamma, cliprew, epsilon);
  TPP_LAUNCH_STATUS();
}

extern "C" int tpp_vecnormalize_step(double* ret, double* rms, const void* raw_rew, int raw_is_int,
                                     const uint8_t* done, float* out_rew, int32_t n_envs, double gamma, double cliprew,
                                     double epsilon, void* stream) {
  TPP_CHECK_ARG(ret && rms && raw_rew && done && out_rew && n_envs > 0 && n_envs <= 65536);
  tpp::vecnormalize_kernel<<<1, 1024, 0, tpp_stream(stream)>>>(ret, rms, raw_rew, raw_is_int, done, out_rew, n_envs,
                                                              gamma, cliprew, epsilon);
  TPP_LAUNCH_STATUS();
}
