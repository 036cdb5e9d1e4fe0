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
#include "../../include/distbo_b200.h"
#include <vector>

#include "gemm_core.cuh"

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
    const int t64 = (n0 + tid) / TN, l64 = (n0 + tid) % TN;
    for (int sl = 0; sl < p.slices; ++sl) {
      const double* src = p.partial + (((size_t)t64 * p.slices + sl) * TN + l64) * 2;
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
