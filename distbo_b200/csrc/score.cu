// K4: posterior mean/variance + acquisition + argmax over a large candidate set.
//   k_* = scale * matern32(theta, theta_c) * kvec ; A = L^-1 k_* = W k_* ; V = W (y - m)
//   mean = A^T V + m ; var = scale - colsum(A^2) ; std = sqrt(max(var, 1e-32))
// Reference: train/log_gaus_likelihood.py:128-218, train/gp_regressor.py:197,
//            bayes_opt/helpers.py:51-68,151-171.
//
// Per chunk of candidates (one 128-candidate tile per SM):
//   kst_kernel     : materialise k_* candidate-major, Kst[c][i] (i contiguous), zero on pad rows;
//   score_kernel   : CTA = (128-candidate tile, slice of the 128-row blocks of W); runs the DMMA main loop
//                    for k <= row block (W is lower triangular) and folds each accumulator tile into
//                    column sums of A^2 and A*V - A is never written; per-slice partial sums out;
//   score_finish   : adds a tile's slices in order, acquisition epilogue, per-tile (max value, lowest index);
//   merge_kernel   : deterministic (max value, min index) merge into the running best.
// From n_pad = 512 the contraction runs on the tcgen05 tensor cores instead (distbo_score_i8_f64): i8_kst_kernel
// writes the int8 digit planes of K_*, score_i8_kernel<0> (digit_planes.cuh) contracts them with the planes of W in
// tensor memory and carries the same two sums; finish and merge are shared.
#include "../../include/distbo_b200.h"
#include <stdlib.h>

#include <vector>

#include "gemm_core.cuh"
#include "digit_planes.cuh"

namespace db {

constexpr int SCORE_DMAX = 16;
constexpr int KST_CAND = 8;  // candidates per CTA in kst_kernel

template <int D>
__global__ void __launch_bounds__(256, 2) kst_kernel(const double* __restrict__ theta_s,
                                                     const double* __restrict__ sqnorm, int N, int n_pad, int d,
                                                     const double* __restrict__ log_bw,
                                                     const double* __restrict__ kvec,
                                                     const double* __restrict__ log_scale,
                                                     const double* __restrict__ theta_c,
                                                     const double* __restrict__ c_shift,
                                                     const double* __restrict__ c_scale, long long c_begin,
                                                     long long M, double* __restrict__ Kst) {
  __shared__ double ys[KST_CAND][D];
  __shared__ double ysq[KST_CAND];
  const int tid = threadIdx.x;
  const long long cl0 = (long long)blockIdx.x * KST_CAND;  // local candidate row in the chunk
  if (tid < KST_CAND * D) {
    const int q = tid / D, k = tid % D;
    const long long c = c_begin + cl0 + q;
    double v = 0.0;
    if (k < d && c < M) {
      v = theta_c[c * d + k];
      if (c_shift) v = (v - c_shift[k]) / c_scale[k];  // StandardScaler.transform, gp_regressor.py:161
      v = v / exp(log_bw[k]);
    }
    ys[q][k] = v;
  }
  __syncthreads();
  if (tid < KST_CAND) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) s += ys[tid][k] * ys[tid][k];
    ysq[tid] = s;
  }
  __syncthreads();
  const double scale = exp(*log_scale);
  for (int i = tid; i < n_pad; i += 256) {
    double xi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = (k < d) ? theta_s[(size_t)i * d + k] : 0.0;
    const double sqi = sqnorm[i];
    const double kv = (i < N) ? kvec[i] : 0.0;
#pragma unroll 1
    for (int q0 = 0; q0 < KST_CAND; q0 += 2) {   // two independent Matern chains at a time (128-register budget)
      double dot[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        dot[u] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) dot[u] += xi[k] * ys[q0 + u][k];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u;
        const double d2 = -2.0 * dot[u] + sqi + ysq[q];
        double v = (scale * matern32_from_d2(d2)) * kv;
        if (i >= N || c_begin + cl0 + q >= M) v = 0.0;
        Kst[(size_t)(cl0 + q) * n_pad + i] = v;
      }
    }
  }
}

constexpr int SCORE_SLICES = 16;  // CTAs that share one candidate tile (row blocks of W dealt out in snake order)

struct ScoreArgs {
  const double* W;
  const double* V;
  const double* Kst;
  int n_pad, slices;
  int tile_w;             // candidates per partial-sum tile: TN (DMMA path) or I8_TILE (tcgen05 path)
  const double* log_scale;
  const double* mean_param;
  long long c_begin, M, index_offset;
  int acq_kind;
  double y_max, xi, kappa;
  double* partial;        // [64-candidate tiles][slices][64][2]: (sum A^2, sum A V) of the slice's row blocks
  double* out_mean;
  double* out_std;
  double* out_acq;
  double* block_best_val;
  long long* block_best_idx;
};

// CTA = (64-candidate tile, slice), two CTAs resident per SM (gemm_core.cuh: 128 x 64 CTA tile, 3-stage
// pipeline).  Consecutive CTAs are the slices of ONE tile, so the tiles in flight at any time keep their K_*
// tiles (4 MB each at N = 8192) and the walked part of W in L2; a slice takes the row blocks ib with
// snake(ib) == slice, which gives every slice the same number of k-steps (work ~ ib + 1).
__global__ void __launch_bounds__(GEMM_THREADS, 2) score_kernel(ScoreArgs p) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double red2[4][TN], redv[4][TN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lk = lane & 3;
  const int S = p.slices;
  const int tile = blockIdx.x / S, slice = blockIdx.x % S;
  const int n0 = tile * TN;  // candidate tile (local rows of Kst)
  const int nb = p.n_pad / BM;

  double s2[4][2], sv[4][2];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) s2[ni][0] = s2[ni][1] = sv[ni][0] = sv[ni][1] = 0.0;

  for (int base = 0; base < nb; base += 2 * S) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ib = half == 0 ? base + slice : base + 2 * S - 1 - slice;
      if (ib >= nb) continue;
      double acc[4][4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      gemm_mainloop_t64<KMAJOR, KMAJOR>(acc, p.W, p.n_pad, p.Kst, p.n_pad, ib * BM, n0, 0, (ib + 1) * BM, smem);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const double v = p.V[ib * BM + wm * 32 + mi * 8 + lr];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          s2[ni][0] += acc[mi][ni][0] * acc[mi][ni][0];
          s2[ni][1] += acc[mi][ni][1] * acc[mi][ni][1];
          sv[ni][0] += acc[mi][ni][0] * v;
          sv[ni][1] += acc[mi][ni][1] * v;
        }
      }
    }
  }
  // reduce over the 8 row-lanes that share a column (lane bits 2..4), then over the four M-warps
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double a = s2[ni][e], b = sv[ni][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lr == 0) {
        const int col = wn * 32 + ni * 8 + lk * 2 + e;
        red2[wm][col] = a;
        redv[wm][col] = b;
      }
    }
  __syncthreads();
  if (tid < TN) {
    double* dst = p.partial + ((size_t)blockIdx.x * TN + tid) * 2;
    dst[0] = (red2[0][tid] + red2[1][tid]) + (red2[2][tid] + red2[3][tid]);
    dst[1] = (redv[0][tid] + redv[1][tid]) + (redv[2][tid] + redv[3][tid]);
  }
}

// One CTA per candidate tile: adds the slices in slice order, then posterior moments, acquisition and the
// tile's (max value, lowest index).
__global__ void __launch_bounds__(128) score_finish_kernel(ScoreArgs p) {
  __shared__ double bval[128];
  __shared__ long long bidx[128];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * BN;
  const long long c = p.c_begin + n0 + tid;
  double val = -INFINITY;
  long long idx = 0x7fffffffffffffffLL;
  if (c < p.M) {
    double sumsq = 0.0, dot = 0.0;
    const int t64 = (n0 + tid) / p.tile_w, l64 = (n0 + tid) % p.tile_w;
    for (int sl = 0; sl < p.slices; ++sl) {
      const double* src = p.partial + (((size_t)t64 * p.slices + sl) * p.tile_w + l64) * 2;
      sumsq += src[0];
      dot += src[1];
    }
    const double scale = exp(*p.log_scale);
    const double mean = dot + *p.mean_param;
    const double var = scale - sumsq;
    const double std = sqrt(fmax(var, kVarClamp));
    const double a = acquisition(p.acq_kind, mean, std, p.y_max, p.xi, p.kappa);
    if (p.out_mean) p.out_mean[c] = mean;
    if (p.out_std) p.out_std[c] = std;
    if (p.out_acq) p.out_acq[c] = a;
    val = a;
    idx = p.index_offset + c;
  }
  bval[tid] = val;
  bidx[tid] = idx;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (tid < s) {
      const double v2 = bval[tid + s];
      const long long i2 = bidx[tid + s];
      if (v2 > bval[tid] || (v2 == bval[tid] && i2 < bidx[tid])) {
        bval[tid] = v2;
        bidx[tid] = i2;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    p.block_best_val[blockIdx.x] = bval[0];
    p.block_best_idx[blockIdx.x] = bidx[0];
  }
}

// running (max value, min index) over the chunk's CTAs; `first` resets the running best
__global__ void merge_best_kernel(const double* __restrict__ bv, const long long* __restrict__ bi, int nblocks,
                                  int first, double* best_val, long long* best_idx) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v = first ? -INFINITY : *best_val;
  long long i = first ? 0x7fffffffffffffffLL : *best_idx;
  for (int b = 0; b < nblocks; ++b)
    if (bv[b] > v || (bv[b] == v && bi[b] < i)) {
      v = bv[b];
      i = bi[b];
    }
  *best_val = v;
  *best_idx = i;
}


// digit-plane kernels of the scoring path only (the shared machinery is in digit_planes.cuh)
// 2^(54 - f) and 2^(f - 60) with f = exponent of scale * max_i |kvec_i| (K_* = scale * matern * kvec, matern <= 1)
__global__ void __launch_bounds__(256) i8_kexp_kernel(const double* __restrict__ kvec, int N,
                                                      const double* __restrict__ log_scale, double* kexp) {
  __shared__ double red[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) m = fmax(m, fabs(kvec[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int f = 0;
    const double bound = exp(*log_scale) * red[0];
    if (bound > 0.0 && isfinite(bound)) frexp(bound, &f);
    kexp[0] = ldexp(1.0, I8_FRAC - f);
    kexp[1] = ldexp(1.0, f - 60);
  }
}

// digit planes of the K_* chunk: K8[s][c][i] (i contiguous), zero on pad rows / pad candidates.
// CTA = KST_CAND candidates; a thread takes four consecutive observations at a time (32-bit stores per plane).
template <int D>
#ifndef KST_MINB
#define KST_MINB 2
#endif
__global__ void __launch_bounds__(256, KST_MINB) i8_kst_kernel(const double* __restrict__ theta_s,
                                                        const double* __restrict__ sqnorm, int N, int n_pad, int d,
                                                        const double* __restrict__ log_bw,
                                                        const double* __restrict__ kvec,
                                                        const double* __restrict__ log_scale,
                                                        const double* __restrict__ theta_c,
                                                        const double* __restrict__ c_shift,
                                                        const double* __restrict__ c_scale, long long c_begin,
                                                        long long M, const double* __restrict__ kexp, int chunk_rows,
                                                        signed char* __restrict__ K8) {
  __shared__ double ys[KST_CAND][D];
  __shared__ double ysq[KST_CAND];
  const int tid = threadIdx.x;
  const long long cl0 = (long long)blockIdx.x * KST_CAND;
  if (tid < KST_CAND * D) {
    const int q = tid / D, k = tid % D;
    const long long c = c_begin + cl0 + q;
    double v = 0.0;
    if (k < d && c < M) {
      v = theta_c[c * d + k];
      if (c_shift) v = (v - c_shift[k]) / c_scale[k];
      v = v / exp(log_bw[k]);
    }
    ys[q][k] = v;
  }
  __syncthreads();
  if (tid < KST_CAND) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) s += ys[tid][k] * ys[tid][k];
    ysq[tid] = s;
  }
  __syncthreads();
  const double scale = exp(*log_scale);
  const double fx = kexp[0];
  const size_t plane = (size_t)chunk_rows * n_pad;
  for (int i4 = tid * 4; i4 < n_pad; i4 += 1024) {
    double sqi[4], kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sqi[u] = sqnorm[i4 + u];
      kv[u] = (i4 + u < N) ? kvec[i4 + u] : 0.0;
    }
    double xi[D <= 8 ? 4 : 1][D <= 8 ? D : 1];
    if (D <= 8) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < D; ++k) xi[D <= 8 ? u : 0][D <= 8 ? k : 0] = (k < d) ? theta_s[(size_t)(i4 + u) * d + k] : 0.0;
    }
#ifndef KST_UNROLL
#define KST_UNROLL 2
#endif
    constexpr int kst_unroll = KST_UNROLL;
#pragma unroll kst_unroll
    for (int q = 0; q < KST_CAND; ++q) {
      double dot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double yk = ys[q][k];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (D <= 8) dot[u] += xi[D <= 8 ? u : 0][D <= 8 ? k : 0] * yk;
          else if (k < d) dot[u] += theta_s[(size_t)(i4 + u) * d + k] * yk;
        }
      }
      const bool cand_ok = c_begin + cl0 + q < M;
      int dg[4][I8_NS];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double d2 = -2.0 * dot[u] + sqi[u] + ysq[q];
        double v = (scale * matern32_from_d2(d2)) * kv[u];
        if (i4 + u >= N || !cand_ok) v = 0.0;
        digits8(v * fx, dg[u]);
      }
      // the low bytes of the four observations' digits, one byte-permute tree per plane
#pragma unroll
      for (int s = 0; s < I8_NS; ++s) {
        const unsigned lo = __byte_perm((unsigned)dg[0][s], (unsigned)dg[1][s], 0x0040);
        const unsigned hi = __byte_perm((unsigned)dg[2][s], (unsigned)dg[3][s], 0x0040);
        *reinterpret_cast<unsigned*>(K8 + s * plane + (size_t)(cl0 + q) * n_pad + i4) = __byte_perm(lo, hi, 0x5410);
      }
    }
  }
}

// Optional per-launch timing of score_kernel (bench.py's roofline figure): CUDA events recorded on
// the launching stream around every score_kernel launch while enabled.
static bool g_timing = false;
static std::vector<cudaEvent_t> g_ev;   // pairs (begin, end)
static size_t g_ev_used = 0;

static bool timing_pair(cudaEvent_t* a, cudaEvent_t* b) {
  if (g_ev_used + 2 > g_ev.size()) {
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return false;
    g_ev.push_back(e0);
    g_ev.push_back(e1);
  }
  *a = g_ev[g_ev_used];
  *b = g_ev[g_ev_used + 1];
  g_ev_used += 2;
  return true;
}

}  // namespace db

using namespace db;

extern "C" int distbo_score_timing(int enable) {
  g_timing = enable != 0;
  g_ev_used = 0;
  return DISTBO_OK;
}

extern "C" int distbo_score_timing_read(double* total_ms, int64_t* launches) {
  if (!total_ms || !launches) return DISTBO_EARG;
  double sum = 0.0;
  for (size_t i = 0; i + 1 < g_ev_used; i += 2) {
    DB_CUDA(cudaEventSynchronize(g_ev[i + 1]));
    float ms = 0.f;
    DB_CUDA(cudaEventElapsedTime(&ms, g_ev[i], g_ev[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *launches = (int64_t)(g_ev_used / 2);
  g_ev_used = 0;
  return DISTBO_OK;
}

extern "C" int64_t distbo_score_workspace(int n_pad) {
  const int64_t chunk = distbo_score_chunk();
  return chunk * n_pad + chunk * SCORE_SLICES * 2;   // K_* chunk, then the per-slice partial sums
}

extern "C" int distbo_score_chunk(void) {
  return device_sm_count() * BN;  // one 128-candidate tile per SM of the current device: a single, balanced wave
}

extern "C" int distbo_score_f64(const double* W, const double* V, int N, int n_pad, const double* theta_s,
                                const double* sqnorm, int d, const double* log_bw, const double* kvec,
                                const double* log_scale, const double* mean_param, const double* theta_c,
                                const double* cand_shift, const double* cand_scale, int64_t M, int acq_kind, double y_max, double xi, double kappa, double* Kst,
                                double* out_mean, double* out_std, double* out_acq, double* block_best_val,
                                int64_t* block_best_idx, double* best_val, int64_t* best_idx,
                                int64_t index_offset, void* stream) {
  if (!W || !V || !theta_s || !sqnorm || !log_bw || !kvec || !log_scale || !mean_param || !theta_c || !Kst ||
      !block_best_val || !block_best_idx || !best_val || !best_idx)
    return DISTBO_EARG;
  if (N <= 0 || n_pad < N || n_pad % 128 || d <= 0 || d > SCORE_DMAX || M <= 0) return DISTBO_EARG;
  if (acq_kind < 0 || acq_kind > 2 || ((cand_shift == nullptr) != (cand_scale == nullptr))) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = distbo_score_chunk();
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TGEMM_SMEM_BYTES));
    once.mark();
  }
  int first = 1;
  for (int64_t c0 = 0; c0 < M; c0 += chunk) {
    const int64_t mc = (M - c0 < chunk) ? (M - c0) : chunk;
    const int tiles = (int)((mc + BN - 1) / BN);
    const int kgrid = tiles * BN / KST_CAND;
    if (d <= 4)
      kst_kernel<4><<<kgrid, 256, 0, st>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale, theta_c, cand_shift,
                                           cand_scale, (long long)c0, (long long)M, Kst);
    else if (d <= 8)
      kst_kernel<8><<<kgrid, 256, 0, st>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale, theta_c, cand_shift,
                                           cand_scale, (long long)c0, (long long)M, Kst);
    else
      kst_kernel<SCORE_DMAX><<<kgrid, 256, 0, st>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale, theta_c,
                                                    cand_shift, cand_scale, (long long)c0, (long long)M, Kst);
    DB_CHECK_LAUNCH();
    ScoreArgs a;
    a.W = W; a.V = V; a.Kst = Kst; a.n_pad = n_pad;
    a.slices = (n_pad / BM < SCORE_SLICES) ? n_pad / BM : SCORE_SLICES;
    a.partial = Kst + (size_t)chunk * n_pad;
    a.tile_w = TN;
    a.log_scale = log_scale; a.mean_param = mean_param;
    a.c_begin = c0; a.M = M; a.index_offset = index_offset;
    a.acq_kind = acq_kind; a.y_max = y_max; a.xi = xi; a.kappa = kappa;
    a.out_mean = out_mean; a.out_std = out_std; a.out_acq = out_acq;
    a.block_best_val = block_best_val; a.block_best_idx = (long long*)block_best_idx;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    const bool timed = g_timing && timing_pair(&t0, &t1);
    if (timed) cudaEventRecord(t0, st);
    score_kernel<<<2 * tiles * a.slices, GEMM_THREADS, TGEMM_SMEM_BYTES, st>>>(a);
    DB_CHECK_LAUNCH();
    if (timed) cudaEventRecord(t1, st);
    score_finish_kernel<<<tiles, 128, 0, st>>>(a);
    DB_CHECK_LAUNCH();
    merge_best_kernel<<<1, 32, 0, st>>>(block_best_val, (const long long*)block_best_idx, tiles, first, best_val,
                                        (long long*)best_idx);
    DB_CHECK_LAUNCH();
    first = 0;
  }
  return DISTBO_OK;
}

// ---- tcgen05 path, host side ---------------------------------------------------------------------------
namespace db {

// Row blocks (cost ~ k range + read-back) dealt to `slices` CTAs, longest first onto the least loaded slice.
static void i8_partition(int n_rb, int slices, I8Args& a) {
  std::vector<double> load(slices, 0.0);
  std::vector<std::vector<unsigned short>> lists(slices);
  for (int rb = n_rb - 1; rb >= 0; --rb) {
    int best = 0;
    for (int s = 1; s < slices; ++s)
      if (load[s] < load[best]) best = s;
    load[best] += (rb + 1) + 1.5;
    lists[best].push_back((unsigned short)rb);
  }
  int o = 0;
  for (int s = 0; s < slices; ++s) {
    a.off[s] = (unsigned short)o;
    for (unsigned short rb : lists[s]) a.order[o++] = rb;
  }
  a.off[slices] = (unsigned short)o;
}

// Slices x tiles in flight <= SMs.  Large N: 4 tiles in flight (their K_* planes, 8 MB each at N = 8192, stay in L2)
// and SMs / 4 slices.  With few row blocks the slice count is chosen for balance instead: every candidate count
// s <= min(SMs / 4, row blocks) that keeps the planes of the tiles in flight within L2 is scored by (mean load / max load of the longest-first deal) x (SMs used / SMs).
static void i8_shape(int n_pad, int* slices, int* tiles_in_flight) {
  const int sms = device_sm_count(), n_rb = n_pad / I8_RB;
  int smax = sms / 4;
  if (smax > I8_MAX_SLICES) smax = I8_MAX_SLICES;
  if (smax > n_rb) smax = n_rb;
  // the K_* planes of the tiles in flight must stay in L2: tiles in flight * 128 * n_pad * 8 bytes <= 40 MB
  int tf_max = (int)(40.0e6 / (1024.0 * n_pad));
  if (tf_max < 4) tf_max = 4;
  int smin = (sms + tf_max - 1) / tf_max;
  if (smin > smax) smin = smax;
  if (smin < 1) smin = 1;
  int best_s = smax;
  double best_score = -1.0;
  for (int sl = smax; sl >= smin; --sl) {
    std::vector<double> load(sl, 0.0);
    double total = 0.0;
    for (int rb = n_rb - 1; rb >= 0; --rb) {
      int b = 0;
      for (int q = 1; q < sl; ++q)
        if (load[q] < load[b]) b = q;
      load[b] += (rb + 1) + 1.5;
      total += (rb + 1) + 1.5;
    }
    double mx = 0.0;
    for (double v : load) mx = mx > v ? mx : v;
    const double score = (total / sl / mx) * ((double)(sl * (sms / sl)) / sms);
    if (score > best_score + 1e-9) {
      best_score = score;
      best_s = sl;
    }
  }
  *slices = best_s;
  *tiles_in_flight = sms / best_s;
}

}  // namespace db

static int64_t i8_meet_words(int n_pad) {
  // slices * waves * row blocks per slice <= (slices * waves) * n_rb; slices * tiles in flight <= SMs
  return (int64_t)(distbo_score_chunk() / I8_TILE) * (n_pad / I8_RB) + 2;
}

extern "C" int64_t distbo_score_i8_workspace(int n_pad) {
  const int64_t chunk = distbo_score_chunk();
  // digit planes of the K_* chunk (8 bytes per element), per-slice partial sums, the two exponent factors
  // + arrival counters of the row-block rendezvous (one per slice, wave and row block of the slice)
  // (the digit planes are double buffered: chunk i + 1 is generated while chunk i is contracted)
  return 2 * chunk * n_pad + (chunk / I8_TILE) * I8_MAX_SLICES * I8_TILE * 2 + 2 + i8_meet_words(n_pad) / 2 + 1;
}

extern "C" int distbo_score_i8_prepare(const double* W, int n_pad, int8_t* W8, double* rowscale, void* stream) {
  if (!W || !W8 || !rowscale || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB) return DISTBO_EARG;
  i8_wsplit_kernel<<<n_pad, 256, 0, (cudaStream_t)stream>>>(W, n_pad, (signed char*)W8, rowscale, nullptr);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_score_i8_f64(const int8_t* W8, const double* rowscale, const double* V, int N, int n_pad,
                                   const double* theta_s, const double* sqnorm, int d, const double* log_bw,
                                   const double* kvec, const double* log_scale, const double* mean_param,
                                   const double* theta_c, const double* cand_shift, const double* cand_scale,
                                   int64_t M, int acq_kind, double y_max, double xi, double kappa, double* ws,
                                   double* out_mean, double* out_std, double* out_acq, double* block_best_val,
                                   int64_t* block_best_idx, double* best_val, int64_t* best_idx,
                                   int64_t index_offset, void* stream) {
  if (!W8 || !rowscale || !V || !theta_s || !sqnorm || !log_bw || !kvec || !log_scale || !mean_param || !theta_c ||
      !ws || !block_best_val || !block_best_idx || !best_val || !best_idx)
    return DISTBO_EARG;
  if (N <= 0 || n_pad < N || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB || d <= 0 || d > SCORE_DMAX || M <= 0)
    return DISTBO_EARG;
  if (acq_kind < 0 || acq_kind > 2 || ((cand_shift == nullptr) != (cand_scale == nullptr))) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = distbo_score_chunk();
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_i8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    once.mark();
  }
  signed char* K8[2] = {reinterpret_cast<signed char*>(ws), reinterpret_cast<signed char*>(ws + (size_t)chunk * n_pad)};
  double* partial = ws + 2 * (size_t)chunk * n_pad;
  double* kexp = partial + (size_t)(chunk / I8_TILE) * I8_MAX_SLICES * I8_TILE * 2;
  CUtensorMap mapK[2], mapKhi[2], mapW;
  if (!tma::encode_u8_planes(&mapK[0], K8[0], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_A_LO) ||
      !tma::encode_u8_planes(&mapK[1], K8[1], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_A_LO) ||
      !tma::encode_u8_planes(&mapKhi[0], K8[0], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_NS - I8_A_LO) ||
      !tma::encode_u8_planes(&mapKhi[1], K8[1], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_NS - I8_A_LO) ||
      !tma::encode_u8_planes(&mapW, W8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS))
    return DISTBO_ECUDA;
  I8Args ia;
  ia.n_rb = n_pad / I8_RB;
  i8_shape(n_pad, &ia.slices, &ia.tiles_in_flight);
  i8_partition(ia.n_rb, ia.slices, ia);
  {
    const char* dbg = getenv("DISTBO_I8_DEBUG");
    ia.debug = dbg ? atoi(dbg) : 0;
  }
  ia.rowscale = rowscale; ia.kexp = kexp; ia.V = V; ia.partial = partial;
  ia.items = nullptr; ia.item_off = nullptr; ia.C = nullptr; ia.ldc = 0; ia.rowscale_a = nullptr; ia.alpha = 1.0; ia.beta = 0;
  unsigned* meet = reinterpret_cast<unsigned*>(kexp + 2);
  int max_rb = 0;
  for (int sl = 0; sl < ia.slices; ++sl) max_rb = max_rb > ia.off[sl + 1] - ia.off[sl] ? max_rb : ia.off[sl + 1] - ia.off[sl];
  ia.meet_stride = max_rb;
  ia.meet = getenv("DISTBO_I8_NOMEET") ? nullptr : meet;
  i8_kexp_kernel<<<1, 256, 0, st>>>(kvec, N, log_scale, kexp);
  DB_CHECK_LAUNCH();
  // The digit planes of chunk i + 1 (fp64 ALU work: Matern + digit extraction) are generated on a side stream
  // while chunk i is contracted on the tensor cores: the contraction kernel is enqueued first and takes one CTA
  // per SM, the generator's CTAs fill the registers that are left.  Stream and events live for this call only
  // (re-entrant; destruction is deferred by the runtime until the work has drained).
  const int64_t nchunks = (M + chunk - 1) / chunk;
  // (measured: 9 % faster at N = 600 .. 1230, 3 % slower at N = 8192 where the step is power-capped and the
  // generator takes clock from the tensor kernel - so only below n_pad = 4096)
  int overlap_below = 4096;
  if (const char* ev = getenv("DISTBO_I8_OVERLAP_BELOW")) overlap_below = atoi(ev);
  const bool overlap = nchunks > 1 && n_pad < overlap_below && !getenv("DISTBO_I8_NOOVERLAP");
  cudaStream_t s2 = st;
  cudaEvent_t evK[2] = {nullptr, nullptr}, evS[2] = {nullptr, nullptr}, ev0 = nullptr;
  struct SideGuard {   // releases the side stream and its events on every way out (the runtime defers the
    cudaStream_t* s;   // destruction until the enqueued work has drained)
    cudaStream_t main;
    cudaEvent_t *k, *sv, *e0;
    ~SideGuard() {
      for (int b = 0; b < 2; ++b) {
        if (k[b]) cudaEventDestroy(k[b]);
        if (sv[b]) cudaEventDestroy(sv[b]);
      }
      if (*e0) cudaEventDestroy(*e0);
      if (*s != main) cudaStreamDestroy(*s);
    }
  } side_guard{&s2, st, evK, evS, &ev0};
  if (overlap) {
    DB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      DB_CUDA(cudaEventCreateWithFlags(&evK[b], cudaEventDisableTiming));
      DB_CUDA(cudaEventCreateWithFlags(&evS[b], cudaEventDisableTiming));
    }
    DB_CUDA(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
    DB_CUDA(cudaEventRecord(ev0, st));
    DB_CUDA(cudaStreamWaitEvent(s2, ev0, 0));
  }
  auto gen_planes = [&](int64_t ci) -> int {
    const int64_t c0 = ci * chunk;
    const int64_t mc = (M - c0 < chunk) ? (M - c0) : chunk;
    const int tiles = (int)((mc + I8_TILE - 1) / I8_TILE);
    const int kgrid = tiles * I8_TILE / KST_CAND;
    signed char* dstK = K8[ci & 1];
    if (overlap && ci >= 2) DB_CUDA(cudaStreamWaitEvent(s2, evS[ci & 1], 0));   // chunk ci - 2 has left this buffer
    if (d <= 4)
      i8_kst_kernel<4><<<kgrid, 256, 0, s2>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale, theta_c,
                                               cand_shift, cand_scale, (long long)c0, (long long)M, kexp, chunk, dstK);
    else if (d <= 8)
      i8_kst_kernel<8><<<kgrid, 256, 0, s2>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale, theta_c,
                                               cand_shift, cand_scale, (long long)c0, (long long)M, kexp, chunk, dstK);
    else
      i8_kst_kernel<SCORE_DMAX><<<kgrid, 256, 0, s2>>>(theta_s, sqnorm, N, n_pad, d, log_bw, kvec, log_scale,
                                                        theta_c, cand_shift, cand_scale, (long long)c0,
                                                        (long long)M, kexp, chunk, dstK);
    DB_CHECK_LAUNCH();
    if (overlap) DB_CUDA(cudaEventRecord(evK[ci & 1], s2));
    return DISTBO_OK;
  };
  {
    const int rc0 = gen_planes(0);
    if (rc0 != DISTBO_OK) return rc0;
  }
  int first = 1;
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t c0 = ci * chunk;
    const int64_t mc = (M - c0 < chunk) ? (M - c0) : chunk;
    const int tiles = (int)((mc + I8_TILE - 1) / I8_TILE);
    if (overlap) DB_CUDA(cudaStreamWaitEvent(st, evK[ci & 1], 0));
    ia.ntiles = tiles;
    const int tf = ia.tiles_in_flight < tiles ? ia.tiles_in_flight : tiles;
    I8Args launch = ia;
    launch.tiles_in_flight = tf;
    launch.meet_waves = (tiles + tf - 1) / tf;
    if (launch.meet)
      DB_CUDA(cudaMemsetAsync(meet, 0, sizeof(unsigned) * (size_t)ia.slices * launch.meet_waves * max_rb, st));
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    const bool timed = g_timing && timing_pair(&t0, &t1);
    if (timed) cudaEventRecord(t0, st);
    score_i8_kernel<0><<<ia.slices * tf, I8_THREADS, I8_SMEM, st>>>(mapK[ci & 1], mapKhi[ci & 1], mapW, launch);
    DB_CHECK_LAUNCH();
    if (timed) cudaEventRecord(t1, st);
    if (overlap) DB_CUDA(cudaEventRecord(evS[ci & 1], st));
    if (ci + 1 < nchunks) {                      // enqueued AFTER the contraction: its CTAs are placed first
      const int rcn = gen_planes(ci + 1);
      if (rcn != DISTBO_OK) return rcn;
    }
    ScoreArgs a;
    a.W = nullptr; a.V = V; a.Kst = nullptr; a.n_pad = n_pad;
    a.slices = ia.slices; a.tile_w = I8_TILE; a.partial = partial;
    a.log_scale = log_scale; a.mean_param = mean_param;
    a.c_begin = c0; a.M = M; a.index_offset = index_offset;
    a.acq_kind = acq_kind; a.y_max = y_max; a.xi = xi; a.kappa = kappa;
    a.out_mean = out_mean; a.out_std = out_std; a.out_acq = out_acq;
    a.block_best_val = block_best_val; a.block_best_idx = (long long*)block_best_idx;
    score_finish_kernel<<<tiles, 128, 0, st>>>(a);
    DB_CHECK_LAUNCH();
    merge_best_kernel<<<1, 32, 0, st>>>(block_best_val, (const long long*)block_best_idx, tiles, first, best_val,
                                        (long long*)best_idx);
    DB_CHECK_LAUNCH();
    first = 0;
  }
  return DISTBO_OK;
}
