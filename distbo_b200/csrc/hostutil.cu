// Host-side helper of the mini-batch sampler: CPython's random.shuffle on an int64 array, bit for bit.
//
// The reference's stratified class mini-batches are drawn with the global `random.shuffle` (train/gp_utils.py:53,78).
// To follow that stream without Python-speed list shuffles on the training loop's host thread (0.7 ms per 3500 rows,
// dozens per epoch at the config-5 shape), the Mersenne-Twister state of a `random.Random(seed)` is handed over once
// and advanced here exactly as CPython does:
//   shuffle(x):      for i = len(x)-1 .. 1:  j = _randbelow(i + 1);  x[i], x[j] = x[j], x[i]        (Lib/random.py)
//   _randbelow(n):   k = n.bit_length();  r = getrandbits(k);  while r >= n: r = getrandbits(k)
//   getrandbits(k):  genrand_uint32() >> (32 - k)      for k <= 32                                 (_randommodule.c)
#include "../../include/distbo_b200.h"
#include "common.cuh"

namespace {
constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

inline uint32_t genrand_uint32(uint32_t* mt, int32_t* mti) {
  static const uint32_t mag01[2] = {0u, MATRIX_A};
  uint32_t y;
  if (*mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; ++kk) {
      y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; ++kk) {
      y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (mt[MT_N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    *mti = 0;
  }
  y = mt[(*mti)++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
}  // namespace

extern "C" int distbo_py_shuffle_i64(uint32_t* mt_state, int32_t* mt_index, int64_t* x, int64_t n) {
  if (!mt_state || !mt_index || (!x && n > 0) || n < 0 || n > 0x7fffffffLL || *mt_index < 0 || *mt_index > MT_N)
    return DISTBO_EARG;
  for (int64_t i = n - 1; i >= 1; --i) {
    const uint32_t bound = (uint32_t)(i + 1);
    int k = 0;
    for (uint32_t b = bound; b; b >>= 1) ++k;          // bit_length
    uint32_t r;
    do {
      r = genrand_uint32(mt_state, mt_index) >> (32 - k);
    } while (r >= bound);
    const int64_t t = x[i];
    x[i] = x[r];
    x[r] = t;
  }
  return DISTBO_OK;
}
