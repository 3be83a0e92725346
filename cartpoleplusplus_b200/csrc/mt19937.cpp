// MT19937 + numpy-legacy bounded integers (host).  Replaces np.random.randint(0, size, n) used by
// ReplayMemory.random_indexes (replay_memory.py:123-129); bit exact with numpy's RandomState stream.
#include <stdint.h>
#include <string.h>
#include "../../include/cartpolepp.h"

namespace cpp { void set_error(const char* fmt, ...); }

struct cpp_mt19937 {
  uint32_t key[624];
  int pos;
};

namespace {
constexpr int N = 624, M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;

void seed_state(cpp_mt19937* s, uint32_t seed) {
  s->key[0] = seed;
  for (int i = 1; i < N; ++i) s->key[i] = 1812433253u * (s->key[i - 1] ^ (s->key[i - 1] >> 30)) + (uint32_t)i;
  s->pos = N;
}

void twist(cpp_mt19937* s) {
  uint32_t* mt = s->key;
  int k = 0;
  for (; k < N - M; ++k) {
    const uint32_t y = (mt[k] & UPPER) | (mt[k + 1] & LOWER);
    mt[k] = mt[k + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; k < N - 1; ++k) {
    const uint32_t y = (mt[k] & UPPER) | (mt[k + 1] & LOWER);
    mt[k] = mt[k + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  const uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  s->pos = 0;
}

inline uint32_t next_u32(cpp_mt19937* s) {
  if (s->pos >= N) twist(s);
  uint32_t y = s->key[s->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
}  // namespace

extern "C" {

int cpp_mt_create(cpp_mt19937** out) {
  if (!out) { cpp::set_error("cpp_mt_create: null out"); return CPP_ERR_INVALID; }
  cpp_mt19937* s = new cpp_mt19937();
  seed_state(s, 5489u);
  *out = s;
  return CPP_OK;
}
int cpp_mt_destroy(cpp_mt19937* mt) { delete mt; return CPP_OK; }
int cpp_mt_seed(cpp_mt19937* mt, uint32_t seed) {
  if (!mt) { cpp::set_error("null generator"); return CPP_ERR_INVALID; }
  seed_state(mt, seed);
  return CPP_OK;
}
int cpp_mt_set_state(cpp_mt19937* mt, const uint32_t* key624, int32_t pos) {
  if (!mt || !key624 || pos < 0 || pos > N) { cpp::set_error("cpp_mt_set_state: bad state (pos=%d)", pos); return CPP_ERR_INVALID; }
  memcpy(mt->key, key624, sizeof(mt->key));
  mt->pos = pos;
  return CPP_OK;
}
int cpp_mt_get_state(const cpp_mt19937* mt, uint32_t* key624, int32_t* pos) {
  if (!mt || !key624 || !pos) { cpp::set_error("cpp_mt_get_state: null"); return CPP_ERR_INVALID; }
  memcpy(key624, mt->key, sizeof(mt->key));
  *pos = mt->pos;
  return CPP_OK;
}
int cpp_mt_randint(cpp_mt19937* mt, int64_t high, int64_t n, int64_t* out) {
  if (!mt || (!out && n > 0) || n < 0) { cpp::set_error("cpp_mt_randint: bad arguments"); return CPP_ERR_INVALID; }
  if (high <= 0) { cpp::set_error("cpp_mt_randint: high=%lld <= 0 (numpy raises ValueError: low >= high)", (long long)high); return CPP_ERR_INVALID; }
  const uint64_t rng = (uint64_t)(high - 1);
  if (rng > 0xffffffffull) { cpp::set_error("cpp_mt_randint: ranges above 2^32 are not used by the replay memory"); return CPP_ERR_INVALID; }
  if (rng == 0) {                 // numpy returns `low` without touching the stream
    for (int64_t i = 0; i < n; ++i) out[i] = 0;
    return CPP_OK;
  }
  uint32_t mask = (uint32_t)rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t v;
    do { v = next_u32(mt) & mask; } while (v > (uint32_t)rng);     // masked rejection
    out[i] = (int64_t)v;
  }
  return CPP_OK;
}

}  // extern "C"
