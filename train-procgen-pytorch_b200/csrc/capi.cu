// Library-level entry points of the C-ABI (include/tpp_b200.h).
#include "tpp_common.cuh"

extern "C" int tpp_version(void) { return 3; }

extern "C" int tpp_device_sm_count(int* out_sm_count) {
  TPP_CHECK_ARG(out_sm_count);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(out_sm_count, cudaDevAttrMultiProcessorCount, dev);
  return (int)e;
}

extern "C" const char* tpp_error_string(int code) {
  switch (code) {
    case TPP_OK: return "ok";
    case TPP_EINVAL: return "TPP_EINVAL: bad argument (null pointer, size or alignment)";
    case TPP_ENOTSUP: return "TPP_ENOTSUP: shape or configuration not supported by this kernel";
    default: return cudaGetErrorString((cudaError_t)code);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// torch.randperm(n) on the default CPU generator, restated: MT19937 (ATen's at::mt19937: `left` counts down, refill
// when it reaches 0) driving the Fisher-Yates loop of ATen's randperm_cpu (`z = random() % (n - i)`, swap i and i + z;
// the form ATen uses for n < 2^32 / 20).  Same numbers, same generator state afterwards -- bit-exact minibatch indices
// (reference common/storage.py:87) -- at a fraction of the time: the loop runs on a 4-byte index array and writes the
// int64 result once.  Host code; state624 holds one 32-bit word per uint64 (torch.get_rng_state() layout).
// ---------------------------------------------------------------------------------------------------------------
#include <cstdint>
#include <vector>

namespace {
struct Mt {
  uint32_t s[624];
  int left;
  uint32_t next;
  static uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void refill() {
    uint32_t* p = s;
    left = 624;
    next = 0;
    for (int j = 624 - 397 + 1; --j; p++) *p = p[397] ^ twist(p[0], p[1]);
    for (int j = 397; --j; p++) *p = p[397 - 624] ^ twist(p[0], p[1]);
    *p = p[397 - 624] ^ twist(p[0], s[0]);
  }
  uint32_t operator()() {
    if (--left == 0) refill();
    uint32_t y = s[next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};
}  // namespace

// the Fisher-Yates loop on a 4-byte index array r[0..n)
static void randperm_core(Mt& mt, uint32_t* rp, int64_t n) {
  for (int64_t i = 0; i < n; ++i) rp[i] = (uint32_t)i;
  // the draws of a block first (the generator is the only sequential part), their swap partners prefetched: the
  // swaps themselves are random accesses into a multi-megabyte array
  constexpr int BLK = 64;
  for (int64_t i0 = 0; i0 < n - 1; i0 += BLK) {
    const int m = (int)((n - 1 - i0) < BLK ? (n - 1 - i0) : BLK);
    uint32_t tgt[BLK];
    for (int k = 0; k < m; ++k) {
      tgt[k] = (uint32_t)(i0 + k) + mt() % (uint32_t)(n - (i0 + k));
      __builtin_prefetch(rp + tgt[k], 1, 1);
    }
    for (int k = 0; k < m; ++k) {
      const uint32_t sav = rp[i0 + k];
      rp[i0 + k] = rp[tgt[k]];
      rp[tgt[k]] = sav;
    }
  }
}

static int randperm_any(uint64_t* state624, int32_t* left, uint64_t* next, int64_t n, int64_t* out64, int32_t* out32) {
  TPP_CHECK_ARG(state624 && left && next && (out64 || out32) && n > 0);
  if (n >= (int64_t)(0xFFFFFFFFu / 20)) return TPP_ENOTSUP;      // ATen switches to another algorithm there
  TPP_CHECK_ARG(*left >= 1 && *left <= 624 && *next <= 624);   // ATen refills at left == 0 and never stores it
  Mt mt;
  for (int i = 0; i < 624; ++i) mt.s[i] = (uint32_t)state624[i];
  mt.left = *left;
  mt.next = (uint32_t)*next;
  if (out32) {                       // n < 2^32 / 20 < 2^31: the indices fit int32; permuted in place
    randperm_core(mt, reinterpret_cast<uint32_t*>(out32), n);
  } else {
    std::vector<uint32_t> r((size_t)n);
    randperm_core(mt, r.data(), n);
    for (int64_t i = 0; i < n; ++i) out64[i] = (int64_t)r[(size_t)i];
  }
  for (int i = 0; i < 624; ++i) state624[i] = mt.s[i];
  *left = mt.left;
  *next = mt.next;
  return TPP_OK;
}

extern "C" int tpp_randperm_mt19937(uint64_t* state624, int32_t* left, uint64_t* next, int64_t n, int64_t* out) {
  return randperm_any(state624, left, next, n, out, nullptr);
}

extern "C" int tpp_randperm_mt19937_i32(uint64_t* state624, int32_t* left, uint64_t* next, int64_t n, int32_t* out) {
  return randperm_any(state624, left, next, n, nullptr, out);
}
