// Shared device helpers for the distbo_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#define DISTBO_OK 0
#define DISTBO_EARG (-1)
#define DISTBO_ECUDA (-2)

// Number of kernels this library has launched (distbo_launch_count); host-side bookkeeping only.
extern long long g_db_launches;

#define DB_CHECK_LAUNCH()                                                  \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) return DISTBO_ECUDA;                           \
    ++g_db_launches;                                                       \
  } while (0)

#define DB_CUDA(x)                                                         \
  do {                                                                     \
    cudaError_t e__ = (x);                                                 \
    if (e__ != cudaSuccess) return DISTBO_ECUDA;                           \
  } while (0)

#include <algorithm>
#include <vector>

namespace db {

// Recursive-halving tree of the triangular inverse over nb diagonal blocks of 128 rows: levels[h] = the nodes
// (first block, blocks in the top part, blocks in the bottom part) merged at height h (chol.cu, score.cu).
struct TriNode {
  int o, n1, n2;
};
inline int build_tri(int o, int cnt, std::vector<std::vector<TriNode>>& levels) {
  if (cnt == 1) return 0;
  int n1 = cnt / 2, n2 = cnt - n1;
  int l1 = build_tri(o, n1, levels), l2 = build_tri(o + n1, n2, levels);
  int lvl = std::max(l1, l2) + 1;
  if ((int)levels.size() <= lvl) levels.resize(lvl + 1);
  levels[lvl].push_back({o, n1, n2});
  return lvl;
}

// cudaFuncSetAttribute opt-ins and SM counts are per DEVICE: one-shot guards keyed by the current device ordinal
// (a process may drive several GPUs).  Thread safe: racing first calls repeat a harmless attribute write.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  int dev = 0;
  bool needed() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return !(done.load(std::memory_order_acquire) >> (dev & 63) & 1ull);
  }
  void mark() { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

inline int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = cache[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    cache[dev & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}

// Debug timeline (tools/potrf_timeline.py): when non-null, instrumented kernels record globaltimer stamps
// into tl[launch_id * 4 + {0: earliest CTA start, 1: latest CTA end, 2, 3: kernel specific}].
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_start(unsigned long long* tl, int id) {
  if (tl && threadIdx.x == 0) atomicMin(tl + id * 4 + 0, globaltimer_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* tl, int id) {
  if (tl && threadIdx.x == 0) atomicMax(tl + id * 4 + 1, globaltimer_ns());
}
__device__ __forceinline__ void tl_mark(unsigned long long* tl, int id, int slot) {
  if (tl && threadIdx.x == 0) atomicMax(tl + id * 4 + slot, globaltimer_ns());
}

constexpr double kSqrt3 = 1.7320508075688772935;
constexpr double kDistClamp = 1e-32;  // reference train/kernels.py:25
constexpr double kVarClamp = 1e-32;   // reference train/gp_regressor.py:197

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  SASS: DMMA.8x8x4.
// lane l holds a = A[l>>2][l&3], b = B[l&3][l>>2], c0/c1 = C[l>>2][2*(l&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Matern-3/2 from a squared distance, reference train/kernels.py:25-27.
__device__ __forceinline__ double matern32_from_d2(double d2) {
  double r = sqrt(fmax(d2, kDistClamp));
  double v = kSqrt3 * r;
  return (1.0 + v) * exp(-v);
}

// EI / UCB / LCB (bayes_opt/helpers.py:151-171); kind = DISTBO_ACQ_* of include/distbo_b200.h.
__device__ __forceinline__ double acquisition(int kind, double mean, double std, double y_max, double xi,
                                              double kappa) {
  if (kind == 0) {
    const double imp = mean - y_max - xi;
    const double z = imp / std;
    const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
    const double pdf = exp(-z * z / 2.0) / 2.5066282746310005024;
    return imp * cdf + std * pdf;
  }
  if (kind == 1) return mean + kappa * std;
  return mean - kappa * std;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace db
