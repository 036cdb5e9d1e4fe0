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


// Mini-batch gather on the device: dst[r, :] = src[idx[r], :] for rows of `row_bytes` bytes (a multiple of 8; 16-byte
// pieces when the pitch allows).  Replaces torch.index_select in the captured training step (train/gp_utils.py:28-104
// draws the rows on the host; the copy of the selected rows is the device part).
namespace {
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, const int64_t* __restrict__ idx,
                                                          int64_t n_rows, int pieces, V* __restrict__ dst) {
  const int64_t total = n_rows * pieces;
  for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < total; g += (int64_t)gridDim.x * 256) {
    const int64_t r = g / pieces;
    const int c = (int)(g - r * pieces);
    dst[r * pieces + c] = src[idx[r] * pieces + c];
  }
}
}  // namespace

extern "C" int distbo_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int64_t n_rows, void* dst,
                                  void* stream) {
  if (!src || !idx || !dst || row_bytes <= 0 || row_bytes % 8 || n_rows < 0) return DISTBO_EARG;
  if (n_rows == 0) return DISTBO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v16 = row_bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0;
  const int pieces = (int)(row_bytes / (v16 ? 16 : 8));
  int64_t blocks = (n_rows * pieces + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (v16)
    gather_rows_kernel<uint4><<<(int)blocks, 256, 0, st>>>((const uint4*)src, idx, n_rows, pieces, (uint4*)dst);
  else
    gather_rows_kernel<unsigned long long><<<(int)blocks, 256, 0, st>>>((const unsigned long long*)src, idx, n_rows,
                                                                         pieces, (unsigned long long*)dst);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
