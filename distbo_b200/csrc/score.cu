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
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "gemm_core.cuh"
#include "tma.cuh"

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


// =========================================================================================================
// tcgen05 path: the same contraction A = K_* W^T on the 5th-generation tensor cores, exact in integers.
//
// fp64 has no tcgen05 kind, int8 has (kind::i8, s32 accumulators in tensor memory, 4x the DMMA rate per byte
// of operand).  Both operands are cut into I8_NS = 8 signed base-128 digits of a 56-bit fixed-point number
// (Ozaki splitting):   W[i,k]   = 2^(e_i - 55) sum_s d_s[i,k] 2^(49 - 7 s),   e_i = exponent of max_k |W[i,k]|
//                      K_*[c,k] = 2^(f - 55)   sum_t q_t[c,k] 2^(49 - 7 t),   f   = exponent of scale max|kvec|
// so that   A[c,i] = 2^(e_i + f - 61) sum_l 2^(49 - 7 l) acc_l[c,i],   acc_l = sum_{s+t=l} sum_k q_t[c,k] d_s[i,k].
// Every acc_l is an exact int32 dot product (|digit| <= 64: |acc_l| <= 8 * n_pad * 2^12 < 2^31 up to
// n_pad = 32768); the 36 digit pairs with l = s + t <= 7 are kept (the dropped ones are below 2^-52 of
// max|W_i| max|K_*| per term, the digit representation itself is exact to 2^-56).  Per (128-candidate tile,
// 64-row block of W) the eight level accumulators fill the SM's tensor memory (8 x 64 columns x 128 lanes);
// they are folded into fp64 only once, after the whole k range: two exact int64 Horner sums, two conversions,
// one fma.  A warp-specialised CTA per SM: one thread issues the TMA loads of the digit planes (64-byte
// swizzle, 96 KB per stage, two stages), one thread issues the 72 MMAs per stage, four warps read the
// accumulators back (tcgen05.ld) and carry sum A^2 and sum A V per candidate in registers.
// =========================================================================================================
constexpr int I8_NS = 8;          // digits per operand
constexpr int I8_TILE = 128;      // candidates per tile = MMA M
constexpr int I8_RB = 64;         // W rows per row block = MMA N
constexpr int I8_KC = 64;         // k per stage (bytes per smem row; two K = 32 MMA steps)
constexpr int I8_STAGES = 2;
constexpr int I8_STAGE_A = I8_NS * I8_TILE * I8_KC;   // 65536
constexpr int I8_STAGE_B = I8_NS * I8_RB * I8_KC;     // 32768
constexpr int I8_STAGE = I8_STAGE_A + I8_STAGE_B;
constexpr int I8_SMEM = I8_STAGES * I8_STAGE + 1024 /* alignment slack */ + 128 /* 12 barriers, tmem slot */;
constexpr int I8_THREADS = 192;   // warp 0: TMA, warp 1: MMA + tensor-memory owner, warps 2-5: read-back
constexpr int I8_MAX_RB = 512;    // n_pad <= 32768
constexpr int I8_MAX_SLICES = 64;
constexpr int I8_FRAC = 55;       // fractional bits of the fixed-point operands

struct I8Args {
  int n_rb, ntiles, slices, tiles_in_flight;
  int debug;                // measurement only: bit 0 = no MMAs, bit 1 = no TMA loads, bit 2 = no read-back
  const double* rowscale;   // [n_pad] 2^e_i
  const double* kexp;       // [2]: 2^(55 - f), 2^(f - 61)
  const double* V;
  double* partial;          // [ntiles][slices][128][2]
  const int4* items;        // MODE 1: (a_row, b_row, first k-chunk, end k-chunk) per work item
  const int* item_off;      // MODE 1: items of CTA b = items[item_off[b] .. item_off[b + 1])
  double* C;                // MODE 1: output matrix, C = alpha * A B^T
  int ldc;
  const double* rowscale_a; // MODE 1: 2^exponent per row of the A operand (rowscale: per row of the B operand)
  double alpha;
  unsigned* meet;           // [slices][waves][row blocks of the slice] arrival counters, zero on entry (or null)
  int meet_stride, meet_waves;   // counters per wave, waves per launch
  unsigned short off[I8_MAX_SLICES + 1];   // row blocks of slice s: order[off[s] .. off[s+1])
  unsigned short order[I8_MAX_RB];
};

// 56-bit fixed point -> 8 balanced base-128 digits, most significant first (d[0] in [-64, 64], others in [-64, 63]).
// The integer is split once into two balanced 28-bit halves; the digits then come from 32-bit arithmetic.
__device__ __forceinline__ void digits8(double x_scaled, int (&dg)[I8_NS]) {
  const long long v = __double2ll_rn(x_scaled);
  int lo = (int)((unsigned)(v + (1ll << 27)) & 0x0fffffffu) - (1 << 27);   // v = hi 2^28 + lo, lo in [-2^27, 2^27)
  int hi = (int)((v - lo) >> 28);
#pragma unroll
  for (int s = 3; s > 0; --s) {
    const int dl = ((lo + 64) & 127) - 64;
    dg[4 + s] = dl;
    lo = (lo - dl) >> 7;
    const int dh = ((hi + 64) & 127) - 64;
    dg[s] = dh;
    hi = (hi - dh) >> 7;
  }
  dg[4] = lo;   // in [-64, 64]
  dg[0] = hi;
}

// 2^(55 - f) and 2^(f - 61) with f = exponent of scale * max_i |kvec_i| (K_* = scale * matern * kvec, matern <= 1)
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
    kexp[1] = ldexp(1.0, f - 61);
  }
}

// digit planes of a lower-triangular matrix split per ROW: W8[s][i][k], rowscale[i] = 2^e_i.  One CTA per row.
// rr == null: the whole row (k <= i is data, k > i zero) - W = L^-1 once per fit.  Otherwise only the rows with a
// non-empty column range rr[i] = [k0, k1) are written, over that range (k > i zero): the bottom-right blocks W22
// of one level of the triangular inverse.
__global__ void __launch_bounds__(256) i8_wsplit_kernel(const double* __restrict__ W, int n_pad,
                                                        signed char* __restrict__ W8, double* __restrict__ rowscale,
                                                        const int2* __restrict__ rr) {
  __shared__ double red[256];
  const int i = blockIdx.x, tid = threadIdx.x;
  int k0 = 0, k1 = n_pad;
  if (rr) {
    const int2 r = rr[i];
    k0 = r.x;
    k1 = r.y;
    if (k1 <= k0) return;
  }
  const double* row = W + (size_t)i * n_pad;
  double m = 0.0;
  for (int k = k0 + tid; k <= i; k += 256) m = fmax(m, fabs(row[k]));
  red[tid] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmax(red[tid], red[tid + s]);
    __syncthreads();
  }
  int e = 0;
  if (red[0] > 0.0 && isfinite(red[0])) frexp(red[0], &e);
  if (tid == 0) rowscale[i] = ldexp(1.0, e);
  const double sc = ldexp(1.0, I8_FRAC - e);
  const size_t plane = (size_t)n_pad * n_pad;
  for (int k4 = k0 + tid * 4; k4 < k1; k4 += 1024) {
    unsigned pk[I8_NS];
#pragma unroll
    for (int s = 0; s < I8_NS; ++s) pk[s] = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k4 + u;
      int dg[I8_NS];
      digits8(k <= i ? row[k] * sc : 0.0, dg);
#pragma unroll
      for (int s = 0; s < I8_NS; ++s) pk[s] |= (unsigned)(dg[s] & 0xff) << (8 * u);
    }
#pragma unroll
    for (int s = 0; s < I8_NS; ++s)
      *reinterpret_cast<unsigned*>(W8 + s * plane + (size_t)i * n_pad + k4) = pk[s];
  }
}

// digit planes of the K_* chunk: K8[s][c][i] (i contiguous), zero on pad rows / pad candidates.
// CTA = KST_CAND candidates; a thread takes four consecutive observations at a time (32-bit stores per plane).
template <int D>
__global__ void __launch_bounds__(256, 2) i8_kst_kernel(const double* __restrict__ theta_s,
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
#pragma unroll 1
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
      unsigned pk[I8_NS];
#pragma unroll
      for (int s = 0; s < I8_NS; ++s) pk[s] = 0u;
      const bool cand_ok = c_begin + cl0 + q < M;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double d2 = -2.0 * dot[u] + sqi[u] + ysq[q];
        double v = (scale * matern32_from_d2(d2)) * kv[u];
        if (i4 + u >= N || !cand_ok) v = 0.0;
        int dg[I8_NS];
        digits8(v * fx, dg);
#pragma unroll
        for (int s = 0; s < I8_NS; ++s) pk[s] |= (unsigned)(dg[s] & 0xff) << (8 * u);
      }
#pragma unroll
      for (int s = 0; s < I8_NS; ++s)
        *reinterpret_cast<unsigned*>(K8 + s * plane + (size_t)(cl0 + q) * n_pad + i4) = pk[s];
    }
  }
}

// Work items of one CTA: (first row of the A tile, first row of the B block, k-chunk range).
//   MODE 0 (scoring): derived from (slice, tile wave): A = K_* planes of the candidate tile, B = W planes of row
//                     block rb, k-chunks 0 .. rb (W is lower triangular); the read-back reduces over the rows of W.
//   MODE 1 (LAUUM)  : explicit list p.items[p.item_off[cta] .. p.item_off[cta + 1]): A and B are both planes of
//                     T = W^T (row i of T = column i of W), C[i,j] = sum_k T[i,k] T[j,k] is stored as fp64.
template <int MODE, class F>
__device__ __forceinline__ void i8_for_items(const I8Args& p, int slice, int tl, F&& f) {
  if (MODE == 0) {
    // The row-block lists of the slices are only balanced to ~5 % (whole row blocks, longest first): a CTA takes
    // slice (slice + wave) mod slices in wave `wave`, so over a launch every CTA walks every list once.
    int wave = 0;
    for (int tile = tl; tile < p.ntiles; tile += p.tiles_in_flight, ++wave) {
      const int sl = (slice + wave) % p.slices;
      const int r_begin = p.off[sl], r_end = p.off[sl + 1];
      for (int r = r_begin; r < r_end; ++r) {
        const int rb = p.order[r];
        f(tile * I8_TILE, rb * I8_RB, 0, rb + 1, tile, wave, (sl << 16) | (r - r_begin), r == r_end - 1);
      }
    }
  } else {
    const int i0 = p.item_off[blockIdx.x], i1 = p.item_off[blockIdx.x + 1];
    for (int it = i0; it < i1; ++it) {
      const int4 w = __ldg(p.items + it);
      f(w.x, w.y, w.z, w.w, 0, 0, 0, false);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(I8_THREADS, 1)
    score_i8_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapW,
                    const __grid_constant__ I8Args p) {
  extern __shared__ unsigned char smem_raw[];
  const unsigned raw = tma::smem_u32(smem_raw);
  const unsigned base = (raw + 1023u) & ~1023u;                 // stage buffers: 1024-byte aligned
  // A stage is handed over in three loads of 32 KB with a barrier each: the W planes (all needed by digit 0), the
  // K_* planes 0-3 and the K_* planes 4-7 (waited for when digit 4 comes up, given back after digit 3 / digit 7).
  // One 96 KB unit per stage could not cover the L2 latency with two stages (measured: 0.3 us stall per stage);
  // sixteen 4-8 KB loads per stage halved the TMA throughput instead.
  const unsigned bars = base + I8_STAGES * I8_STAGE;   // per stage: fullW, fullA0, fullA1, emptyA0, emptyRest
  const unsigned bar_tfull = bars + 40 * I8_STAGES, bar_tempty = bar_tfull + 8;
  const unsigned slot = bar_tempty + 8;
  volatile unsigned* slot_ptr = reinterpret_cast<volatile unsigned*>(smem_raw + (slot - raw));
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler too
  // the CTAs that walk the same row blocks of W (one per tile in flight) are neighbours in the grid
  const int slice = MODE == 0 ? blockIdx.x / p.tiles_in_flight : 0;
  const int tl = MODE == 0 ? blockIdx.x % p.tiles_in_flight : 0;

  if (warp == 0 && lane == 0) {
    tma::prefetch_map(&mapK);
    tma::prefetch_map(&mapW);
#pragma unroll
    for (int s = 0; s < I8_STAGES * 5; ++s) tma::mbar_init(bars + 8 * s, 1);
    tma::mbar_init(bar_tfull, 1);
    tma::mbar_init(bar_tempty, 4);
    tma::fence_barrier_init();
  }
  if (warp == 1) tc5::alloc(slot, 512);
  tc5::fence_before();
  __syncthreads();
  tc5::fence_after();
  const unsigned tmem = *slot_ptr;

  // Warps 0 and 1 run their loops with all 32 lanes and issue under elect.sync: in a `lane == 0` branch the
  // compiler cannot prove the descriptors uniform and wraps every UTCIMMA / UTMALDG in a vote-and-broadcast
  // loop (~20 instructions each), which made the single issuing thread the bottleneck of the pipeline.
  if (warp == 0) {
    int stage = 0;
    unsigned phase = 0;
    i8_for_items<MODE>(p, slice, tl, [&](int a_row, int b_row, int kb, int ke, int tile, int wave, int ri, bool) {
      // Soft rendezvous of the CTAs that stream the same row block of W for different candidate tiles: they
      // start it within a few microseconds of each other, so that one of them pulls the digit planes from HBM
      // and the others hit L2 (without it they drift apart and W was read 3.6 times per wave: 35.7 GB of
      // DRAM reads per launch).  Performance only - the wait is short and bounded, never a dependency.
      if (MODE == 0 && p.meet && lane == 0) {
        unsigned* ctr = p.meet + ((size_t)(ri >> 16) * p.meet_waves + wave) * p.meet_stride + (ri & 0xffff);
        const int left = p.ntiles - wave * p.tiles_in_flight;
        const unsigned want = left < p.tiles_in_flight ? left : p.tiles_in_flight;
        atomicAdd(ctr, 1u);
        const unsigned long long t0 = globaltimer_ns();
        while (*(volatile unsigned*)ctr < want && globaltimer_ns() - t0 < 30000ull) __nanosleep(200);
      }
      __syncwarp();
      for (int kc = kb; kc < ke; ++kc) {
        const unsigned dst = base + stage * I8_STAGE, sb = bars + 40 * stage;
        tma::mbar_wait_bounded(sb + 24, phase ^ 1u);            // A planes 0-3 of this slot are free
        if (tma::elect_one()) {
          if (p.debug & 2) {
            tma::mbar_arrive(sb + 8);
          } else {
            tma::mbar_arrive_expect_tx(sb + 8, I8_STAGE_A / 2);
            tma::load_3d(dst, &mapK, kc * I8_KC, a_row, 0, sb + 8);
          }
        }
        tma::mbar_wait_bounded(sb + 32, phase ^ 1u);            // ... and the rest of it
        if (tma::elect_one()) {
          if (p.debug & 2) {
            tma::mbar_arrive(sb);
            tma::mbar_arrive(sb + 16);
          } else {
            tma::mbar_arrive_expect_tx(sb, I8_STAGE_B);
            tma::load_3d(dst + I8_STAGE_A, &mapW, kc * I8_KC, b_row, 0, sb);
            tma::mbar_arrive_expect_tx(sb + 16, I8_STAGE_A / 2);
            tma::load_3d(dst + I8_STAGE_A / 2, &mapK, kc * I8_KC, a_row, I8_NS / 2, sb + 16);
          }
        }
        __syncwarp();
        if (++stage == I8_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    });
  } else if (warp == 1) {
    int stage = 0;
    unsigned phase = 0, tphase = 0;
    i8_for_items<MODE>(p, slice, tl, [&](int, int, int kb, int ke, int, int, int, bool) {
      tma::mbar_wait_bounded(bar_tempty, tphase ^ 1u);          // the read-back of the previous item is done
      tc5::fence_after();
      for (int kc = kb; kc < ke; ++kc) {
        const unsigned sb = bars + 40 * stage;
        tma::mbar_wait_bounded(sb, phase);
        tma::mbar_wait_bounded(sb + 8, phase);
        tc5::fence_after();
        const unsigned long long da = tc5::desc_k_sw64(base + stage * I8_STAGE);
        const unsigned long long dbs = tc5::desc_k_sw64(base + stage * I8_STAGE + I8_STAGE_A);
        // Digit t of the A operand meets digits s = 0 .. 7 - t of the B operand at once: the B planes are
        // contiguous in the stage (64 rows x 64 bytes each), so planes s0 .. s0 + n - 1 form ONE K-major B operand
        // of N = 64 n rows, and because the level accumulators are contiguous in tensor memory too (level l at
        // columns 64 l), its product lands on levels t + s0 .. t + s0 + n - 1 in one MMA: 12 wide MMAs (N = 256 /
        // 192 / 128, one of 64) per k-step instead of 36 of N = 64, which were bound by re-reading the A tile from
        // shared memory (measured 45 % of the int8 rate).  Digits 0-3 need two MMAs: the second one takes the A
        // tile from the collector.  k-step outermost: consecutive MMAs write different accumulator columns (the
        // same columns come up again 12 MMAs later); digit-outermost serialised on the accumulate dependency.
        auto issue = [&](int ks, int t) {
          const unsigned long long ad = da + (unsigned long long)((t * I8_TILE * I8_KC + ks * 32) >> 4);
          const int planes = I8_NS - t;                                 // B digits 0 .. planes - 1
          const int n0 = planes > 4 ? (planes == 5 ? 3 : 4) : planes;   // 8=4+4 7=4+3 6=4+2 5=3+2
          const unsigned acc = (kc > kb || ks > 0 || t > 0) ? 1u : 0u;  // t = 0 covers all 512 columns
          if (planes > n0) {
            tc5::mma_i8<1>(tmem + t * I8_RB, ad, dbs + (unsigned long long)((ks * 32) >> 4),
                           tc5::idesc_i8(I8_TILE, n0 * I8_RB), acc);
            tc5::mma_i8<2>(tmem + (t + n0) * I8_RB, ad,
                           dbs + (unsigned long long)((n0 * I8_RB * I8_KC + ks * 32) >> 4),
                           tc5::idesc_i8(I8_TILE, (planes - n0) * I8_RB), acc);
          } else {
            tc5::mma_i8<0>(tmem + t * I8_RB, ad, dbs + (unsigned long long)((ks * 32) >> 4),
                           tc5::idesc_i8(I8_TILE, n0 * I8_RB), acc);
          }
        };
        if (!(p.debug & 1)) {
          if (tma::elect_one()) {
#pragma unroll
            for (int t = 0; t < I8_NS / 2; ++t) issue(0, t);
          }
          __syncwarp();
          tma::mbar_wait_bounded(sb + 16, phase);                       // A planes 4-7
          tc5::fence_after();
          if (tma::elect_one()) {
#pragma unroll
            for (int t = I8_NS / 2; t < I8_NS; ++t) issue(0, t);
#pragma unroll
            for (int t = 0; t < I8_NS / 2; ++t) issue(1, t);
            tc5::commit(sb + 24);                                       // A planes 0-3 are free
#pragma unroll
            for (int t = I8_NS / 2; t < I8_NS; ++t) issue(1, t);
            tc5::commit(sb + 32);                                       // ... and the rest of the stage
          }
          __syncwarp();
        } else {
          tma::mbar_wait_bounded(sb + 16, phase);
          if (tma::elect_one()) {
            tc5::commit(sb + 24);
            tc5::commit(sb + 32);
          }
          __syncwarp();
        }
        if (++stage == I8_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (tma::elect_one()) tc5::commit(bar_tfull);             // accumulators of this item are complete
      __syncwarp();
      tphase ^= 1u;
    });
  } else {
    const int q = warp & 3;                                     // tensor-memory lane quadrant of this warp
    const int c = q * 32 + lane;                                // row within the A tile
    const unsigned tq = tmem + ((unsigned)(q * 32) << 16);
    const double kx = MODE == 0 ? p.kexp[1] : 4.336808689942017736e-19;   // 2^(f - 61) / 2^-61
    unsigned tphase = 0;
    double s2 = 0.0, sv = 0.0;
    i8_for_items<MODE>(p, slice, tl, [&](int a_row, int b_row, int, int, int tile, int, int ri, bool last) {
      tma::mbar_wait_bounded(bar_tfull, tphase);
      tphase ^= 1u;
      tc5::fence_after();
      const double rsa = MODE == 1 ? __ldg(p.rowscale_a + a_row + c) * (kx * p.alpha) : kx;
      double* crow = MODE == 1 ? p.C + (size_t)(a_row + c) * p.ldc + b_row : nullptr;
#pragma unroll 1
      for (int j0 = 0; j0 < ((p.debug & 4) ? 0 : I8_RB); j0 += 8) {
        int a[I8_NS][8];
#pragma unroll
        for (int l = 0; l < I8_NS; ++l) tc5::ld_32x32b_x8(tq + l * I8_RB + j0, a[l]);
        tc5::wait_ld();
        double out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const long long hi = ((((long long)a[0][j] * 128 + a[1][j]) * 128 + a[2][j]) * 128) + a[3][j];
          const long long lo = ((((long long)a[4][j] * 128 + a[5][j]) * 128 + a[6][j]) * 128) + a[7][j];
          const int i = b_row + j0 + j;
          const double av = (__ldg(p.rowscale + i) * rsa) * fma((double)hi, 268435456.0, (double)lo);
          if (MODE == 0) {
            s2 = fma(av, av, s2);
            sv = fma(av, __ldg(p.V + i), sv);
          } else {
            out[j] = av;
          }
        }
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 8; j += 2)
            *reinterpret_cast<double2*>(crow + j0 + j) = make_double2(out[j], out[j + 1]);
        }
      }
      tc5::fence_before();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(bar_tempty);
      if (MODE == 0 && last) {
        double* dst = p.partial + (((size_t)tile * p.slices + (ri >> 16)) * I8_TILE + c) * 2;
        dst[0] = s2;
        dst[1] = sv;
        s2 = sv = 0.0;
      }
    });
  }
  tc5::fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc5::dealloc(tmem, 512);
  }
}

// digit planes of T = W^T with one exponent per row of T (= column of W): T8[s][i][k] = digit s of W[k][i] (k >= i)
// Row range of column j of the matrix that is split per COLUMN:
//   cr == null : [j, n_pad)            the lower triangle of W (LAUUM)
//   mode 0     : [j, cr[j].x)          the top-left blocks W11 of one level of the triangular inverse
//   mode 1     : [cr[j].x, cr[j].y)    the off-diagonal blocks T = L21 W11 of that level
// (cr[j] = (0, 0) on the columns that take no part).
__device__ __forceinline__ int2 i8_col_range(const int2* __restrict__ cr, int mode, int j, int n_pad) {
  if (!cr) return make_int2(j, n_pad);
  const int2 r = cr[j];
  return mode == 0 ? make_int2(j, r.x) : r;
}

// colmax pass: CTA = 64 columns x 512 rows (4 row lanes per column), merged with an integer atomicMax on the bit
// pattern (non-negative doubles order like integers); colmax must be zero on entry
__global__ void __launch_bounds__(256) i8_colmax_kernel(const double* __restrict__ W, int n_pad,
                                                        unsigned long long* __restrict__ colmax,
                                                        const int2* __restrict__ cr, int mode) {
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int2 rg = i8_col_range(cr, mode, i, n_pad);
  const int k0 = max((int)blockIdx.y * 512, rg.x), k1 = min(min(n_pad, (int)blockIdx.y * 512 + 512), rg.y);
  double m = 0.0;
  for (int k = k0 + (threadIdx.x >> 6); k < k1; k += 4) m = fmax(m, fabs(W[(size_t)k * n_pad + i]));
  if (m > 0.0) atomicMax(colmax + i, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256) i8_colscale_kernel(double* __restrict__ colscale, int n_pad) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_pad) return;
  const double m = colscale[i];   // the column maximum (bit pattern written by atomicMax)
  int e = 0;
  if (m > 0.0 && isfinite(m)) frexp(m, &e);
  colscale[i] = ldexp(1.0, e);
}

// CTA = 64 x 64 tile (k block, i block), transposed through shared memory.  cr == null: the lower-triangular tiles
// are enumerated in blockIdx.x; otherwise the grid is (k blocks, i blocks) and tiles outside the column ranges exit.
__global__ void __launch_bounds__(256) i8_wtsplit_kernel(const double* __restrict__ W, int n_pad,
                                                         const double* __restrict__ colscale,
                                                         signed char* __restrict__ T8, const int2* __restrict__ cr,
                                                         int mode) {
  __shared__ __align__(16) unsigned char sm[I8_NS][64][64];   // [plane][i][k]
  int kb, ib;
  if (!cr) {
    kb = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((kb + 1) * (kb + 2) / 2 <= (int)blockIdx.x) ++kb;
    while (kb * (kb + 1) / 2 > (int)blockIdx.x) --kb;
    ib = blockIdx.x - kb * (kb + 1) / 2;
  } else {
    kb = blockIdx.x;
    ib = blockIdx.y;
    const int2 r0 = i8_col_range(cr, mode, ib * 64, n_pad);   // the 64 columns of a tile share their block
    if (kb * 64 + 64 <= r0.x || kb * 64 >= r0.y) return;
  }
  const int tid = threadIdx.x;
  const int il = tid & 63;
  const int i = ib * 64 + il;
  const int2 rg = i8_col_range(cr, mode, i, n_pad);
  const double sc = ldexp(1.0, I8_FRAC) / colscale[i];          // colscale is a power of two: exact
  for (int kq = tid >> 6; kq < 16; kq += 4) {                   // four consecutive k per thread
    unsigned pk[I8_NS];
#pragma unroll
    for (int s = 0; s < I8_NS; ++s) pk[s] = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb * 64 + kq * 4 + u;
      int dg[I8_NS];
      digits8((k >= rg.x && k < rg.y) ? W[(size_t)k * n_pad + i] * sc : 0.0, dg);
#pragma unroll
      for (int s = 0; s < I8_NS; ++s) pk[s] |= (unsigned)(dg[s] & 0xff) << (8 * u);
    }
#pragma unroll
    for (int s = 0; s < I8_NS; ++s) *reinterpret_cast<unsigned*>(&sm[s][il][kq * 4]) = pk[s];
  }
  __syncthreads();
  const size_t plane = (size_t)n_pad * n_pad;
  for (int v = tid; v < I8_NS * 64 * 4; v += 256) {             // 16-byte pieces: [plane][i][4 pieces]
    const int s = v >> 8, r = (v >> 2) & 63, q = v & 3;
    *reinterpret_cast<uint4*>(T8 + s * plane + (size_t)(ib * 64 + r) * n_pad + kb * 64 + q * 16) =
        *reinterpret_cast<const uint4*>(&sm[s][r][q * 16]);
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
  CUtensorMap mapK[2], mapW;
  if (!tma::encode_u8_planes(&mapK[0], K8[0], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_NS / 2) ||
      !tma::encode_u8_planes(&mapK[1], K8[1], (uint64_t)n_pad, (uint64_t)chunk, I8_NS, I8_TILE, I8_NS / 2) ||
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
  ia.items = nullptr; ia.item_off = nullptr; ia.C = nullptr; ia.ldc = 0; ia.rowscale_a = nullptr; ia.alpha = 1.0;
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
  const bool overlap = nchunks > 1 && n_pad < 4096 && !getenv("DISTBO_I8_NOOVERLAP");
  cudaStream_t s2 = st;
  cudaEvent_t evK[2] = {nullptr, nullptr}, evS[2] = {nullptr, nullptr}, ev0 = nullptr;
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
    score_i8_kernel<0><<<ia.slices * tf, I8_THREADS, I8_SMEM, st>>>(mapK[ci & 1], mapW, launch);
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
  if (overlap) {
    for (int b = 0; b < 2; ++b) {
      cudaEventDestroy(evK[b]);
      cudaEventDestroy(evS[b]);
    }
    cudaEventDestroy(ev0);
    cudaStreamDestroy(s2);
  }
  return DISTBO_OK;
}

// ---- LAUUM on the tcgen05 tensor cores: Kinv = W^T W = T T^T with T = W^T in int8 digit planes ---------------
namespace db {
static int lauum_i8_items(int n_pad, std::vector<int4>* items, std::vector<int>* off) {
  const int sms = device_sm_count();
  const int nb = n_pad / I8_TILE, nk = n_pad / I8_KC;
  struct It { int4 w; double cost; };
  std::vector<It> all;
  for (int ib = 0; ib < nb; ++ib)
    for (int jb = 0; jb <= 2 * ib + 1; ++jb)
      all.push_back({make_int4(ib * I8_TILE, jb * I8_RB, 2 * ib, nk), (double)(nk - 2 * ib) + 2.0});
  // longest first onto the least loaded CTA (items are generated with decreasing k range already)
  std::vector<double> load(sms, 0.0);
  std::vector<std::vector<int4>> lists(sms);
  for (const It& it : all) {
    int best = 0;
    for (int c = 1; c < sms; ++c)
      if (load[c] < load[best]) best = c;
    load[best] += it.cost;
    lists[best].push_back(it.w);
  }
  if (items) {
    items->clear();
    off->assign(sms + 1, 0);
    for (int c = 0; c < sms; ++c) {
      (*off)[c] = (int)items->size();
      items->insert(items->end(), lists[c].begin(), lists[c].end());
    }
    (*off)[sms] = (int)items->size();
  }
  return (int)all.size();
}
}  // namespace db

extern "C" int64_t distbo_lauum_i8_sched_words(int n_pad) {
  if (n_pad <= 0 || n_pad % 128) return 0;
  return (int64_t)lauum_i8_items(n_pad, nullptr, nullptr) * 4 + device_sm_count() + 1 + 3;
}

extern "C" int distbo_lauum_i8_plan(int n_pad, int32_t* sched) {
  if (!sched || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB) return DISTBO_EARG;
  std::vector<int4> items;
  std::vector<int> off;
  lauum_i8_items(n_pad, &items, &off);
  // layout: [sms + 1 offsets, padded to a multiple of 4 words][items]
  const size_t off_words = (off.size() + 3) / 4 * 4;
  DB_CUDA(cudaMemcpy(sched, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  DB_CUDA(cudaMemcpy(sched + off_words, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return DISTBO_OK;
}

extern "C" int64_t distbo_lauum_i8_workspace(int n_pad) {
  return (int64_t)n_pad * n_pad + n_pad;   // doubles: 8 digit planes of W^T (8 bytes per element) + column exponents
}

extern "C" int distbo_lauum_i8_f64(const double* W, int n_pad, const int32_t* sched, double* ws, double* Kinv,
                                   void* stream) {
  if (!W || !sched || !ws || !Kinv || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    once.mark();
  }
  signed char* T8 = reinterpret_cast<signed char*>(ws);
  double* colscale = ws + (size_t)n_pad * n_pad;
  DB_CUDA(cudaMemsetAsync(colscale, 0, sizeof(double) * n_pad, st));
  i8_colmax_kernel<<<dim3(n_pad / 64, (n_pad + 511) / 512), 256, 0, st>>>(
      W, n_pad, reinterpret_cast<unsigned long long*>(colscale), nullptr, 0);
  DB_CHECK_LAUNCH();
  i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(colscale, n_pad);
  DB_CHECK_LAUNCH();
  const int nb64 = n_pad / 64;
  i8_wtsplit_kernel<<<nb64 * (nb64 + 1) / 2, 256, 0, st>>>(W, n_pad, colscale, T8, nullptr, 0);
  DB_CHECK_LAUNCH();
  CUtensorMap mapA, mapB;
  if (!tma::encode_u8_planes(&mapA, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS / 2) ||
      !tma::encode_u8_planes(&mapB, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS))
    return DISTBO_ECUDA;
  const int sms = device_sm_count();
  I8Args a;
  a.n_rb = n_pad / I8_RB; a.ntiles = 0; a.slices = 1; a.tiles_in_flight = 1;
  {
    const char* dbg = getenv("DISTBO_I8_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
  }
  a.rowscale = colscale; a.kexp = nullptr; a.V = nullptr; a.partial = nullptr;
  a.item_off = reinterpret_cast<const int*>(sched);
  a.items = reinterpret_cast<const int4*>(sched + (sms + 1 + 3) / 4 * 4);
  a.C = Kinv; a.ldc = n_pad; a.rowscale_a = colscale; a.alpha = 1.0;
  a.meet = nullptr; a.meet_stride = 0; a.meet_waves = 0;
  score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mapA, mapB, a);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

// ---- Top levels of the triangular inverse on the tcgen05 tensor cores ------------------------------------------
// Level h of the halving tree merges pairs of inverted diagonal blocks:  W21 = -W22 (L21 W11).  Both products run
// on the digit-plane kernel: T = L21 W11 from the row planes of L and the column planes of W11, then W21 = -W22 T
// from the row planes of W22 and the column planes of T.  Only the levels whose blocks have at least `min_rows`
// rows take this route (the GEMMs of the small levels are launch-bound; they stay on the DMMA tile GEMM).
namespace db {

struct TriI8Level {
  std::vector<int2> cr, rr;            // per column / per row ranges (see i8_col_range, i8_wsplit_kernel)
  std::vector<int4> items[2];          // work items of the two products, per CTA
  std::vector<int> off[2];
};

static void tri_i8_deal(const std::vector<int4>& all, int sms, std::vector<int4>* items, std::vector<int>* off) {
  std::vector<int4> sorted(all);
  std::stable_sort(sorted.begin(), sorted.end(),
                   [](const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); });
  std::vector<double> load(sms, 0.0);
  std::vector<std::vector<int4>> lists(sms);
  for (const int4& it : sorted) {
    int best = 0;
    for (int c = 1; c < sms; ++c)
      if (load[c] < load[best]) best = c;
    load[best] += (double)(it.w - it.z) + 2.0;
    lists[best].push_back(it);
  }
  items->clear();
  off->assign(sms + 1, 0);
  for (int c = 0; c < sms; ++c) {
    (*off)[c] = (int)items->size();
    items->insert(items->end(), lists[c].begin(), lists[c].end());
  }
  (*off)[sms] = (int)items->size();
}

static int tri_i8_levels_build(int n_pad, int min_rows, std::vector<TriI8Level>* out);

// cached per (SM count, n_pad, min_rows): the lists are rebuilt on the host only once (a training loop that is not
// replayed as a CUDA graph enqueues the inverse every epoch)
static int tri_i8_levels(int n_pad, int min_rows, std::vector<TriI8Level>* out) {
  if (!out) return tri_i8_levels_build(n_pad, min_rows, nullptr);
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, std::pair<int, std::vector<TriI8Level>>> cache;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(device_sm_count(), n_pad, min_rows);
  auto it = cache.find(key);
  if (it == cache.end()) {
    std::vector<TriI8Level> lv;
    const int first = tri_i8_levels_build(n_pad, min_rows, &lv);
    it = cache.emplace(key, std::make_pair(first, std::move(lv))).first;
  }
  *out = it->second.second;
  return it->second.first;
}

// returns the first height that takes the int8 route (levels.size() when none does) and fills `out`
static int tri_i8_levels_build(int n_pad, int min_rows, std::vector<TriI8Level>* out) {
  const int nb = n_pad / 128, sms = device_sm_count();
  std::vector<std::vector<TriNode>> levels;
  build_tri(0, nb, levels);
  int first = (int)levels.size();
  for (int h = 1; h < (int)levels.size(); ++h) {
    int small = 1 << 30;
    for (const TriNode& nd : levels[h]) small = std::min(small, std::min(nd.n1, nd.n2) * 128);
    if (small >= min_rows) {
      first = h;
      break;
    }
  }
  if (first < 1) first = 1;
  if (out) {
    out->clear();
    for (int h = first; h < (int)levels.size(); ++h) {
      TriI8Level lv;
      lv.cr.assign(n_pad, make_int2(0, 0));
      lv.rr.assign(n_pad, make_int2(0, 0));
      std::vector<int4> g1, g2;
      for (const TriNode& nd : levels[h]) {
        const int o = nd.o * 128, mid = (nd.o + nd.n1) * 128, end = (nd.o + nd.n1 + nd.n2) * 128;
        for (int j = o; j < mid; ++j) lv.cr[j] = make_int2(mid, end);
        for (int i = mid; i < end; ++i) lv.rr[i] = make_int2(mid, end);
        for (int a = mid; a < end; a += I8_TILE)
          for (int b = o; b < mid; b += I8_RB) {
            g1.push_back(make_int4(a, b, b / I8_KC, mid / I8_KC));                 // T = L21 W11 : k in [b, mid)
            g2.push_back(make_int4(a, b, mid / I8_KC, (a + I8_TILE) / I8_KC));     // W21 = -W22 T: k in [mid, a+128)
          }
      }
      tri_i8_deal(g1, sms, &lv.items[0], &lv.off[0]);
      tri_i8_deal(g2, sms, &lv.items[1], &lv.off[1]);
      out->push_back(std::move(lv));
    }
  }
  return first;
}

static size_t pad4(size_t w) { return (w + 3) / 4 * 4; }

}  // namespace db

extern "C" int distbo_trtri_i8_first_level(int n_pad, int min_rows) {
  if (n_pad <= 0 || n_pad % 128 || min_rows < 128) return DISTBO_EARG;
  return tri_i8_levels(n_pad, min_rows, nullptr);
}

extern "C" int64_t distbo_trtri_i8_sched_words(int n_pad, int min_rows) {
  if (n_pad <= 0 || n_pad % 128 || min_rows < 128) return 0;
  std::vector<TriI8Level> lv;
  tri_i8_levels(n_pad, min_rows, &lv);
  size_t w = 4;
  for (const TriI8Level& l : lv)
    w += 4 * (size_t)n_pad + pad4(l.off[0].size()) + 4 * l.items[0].size() + pad4(l.off[1].size()) + 4 * l.items[1].size();
  return (int64_t)w;
}

extern "C" int distbo_trtri_i8_plan(int n_pad, int min_rows, int32_t* sched) {
  if (!sched || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB || min_rows < 128) return DISTBO_EARG;
  std::vector<TriI8Level> lv;
  tri_i8_levels(n_pad, min_rows, &lv);
  std::vector<int32_t> host;
  auto put = [&](const void* src, size_t words, size_t padded) {
    const size_t at = host.size();
    host.resize(at + padded, 0);
    memcpy(host.data() + at, src, words * sizeof(int32_t));
  };
  for (const TriI8Level& l : lv) {
    put(l.cr.data(), 2 * (size_t)n_pad, 2 * (size_t)n_pad);
    put(l.rr.data(), 2 * (size_t)n_pad, 2 * (size_t)n_pad);
    for (int g = 0; g < 2; ++g) {
      put(l.off[g].data(), l.off[g].size(), pad4(l.off[g].size()));
      put(l.items[g].data(), 4 * l.items[g].size(), 4 * l.items[g].size());
    }
  }
  if (!host.empty()) DB_CUDA(cudaMemcpy(sched, host.data(), host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return DISTBO_OK;
}

extern "C" int64_t distbo_trtri_i8_workspace(int n_pad) {
  // doubles: four sets of digit planes (rows of L, columns of W11, rows of W22, columns of T) + their exponents
  return 4 * (int64_t)n_pad * n_pad + 4 * (int64_t)n_pad;
}

extern "C" int distbo_trtri_i8_f64(const double* L, double* W, double* T, int n_pad, int min_rows,
                                   const int32_t* sched, double* ws, void* stream) {
  if (!L || !W || !T || !sched || !ws || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB || min_rows < 128)
    return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    once.mark();
  }
  std::vector<TriI8Level> lv;
  tri_i8_levels(n_pad, min_rows, &lv);
  if (lv.empty()) return DISTBO_OK;
  const size_t nn = (size_t)n_pad * n_pad;
  signed char* L8 = reinterpret_cast<signed char*>(ws);
  signed char* C8 = reinterpret_cast<signed char*>(ws + nn);        // columns of W11
  signed char* R8 = reinterpret_cast<signed char*>(ws + 2 * nn);    // rows of W22
  signed char* T8 = reinterpret_cast<signed char*>(ws + 3 * nn);    // columns of T
  double* rsL = ws + 4 * nn;
  double* csW = rsL + n_pad;
  double* rsW = csW + n_pad;
  double* csT = rsW + n_pad;
  CUtensorMap mL, mC, mR, mT;
  if (!tma::encode_u8_planes(&mL, L8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS / 2) ||
      !tma::encode_u8_planes(&mC, C8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS) ||
      !tma::encode_u8_planes(&mR, R8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS / 2) ||
      !tma::encode_u8_planes(&mT, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS))
    return DISTBO_ECUDA;
  const int sms = device_sm_count();
  const dim3 cgrid(n_pad / 64, (n_pad + 511) / 512), tgrid(n_pad / 64, n_pad / 64);
  // rows of L, once (k <= i)
  i8_wsplit_kernel<<<n_pad, 256, 0, st>>>(L, n_pad, L8, rsL, nullptr);
  DB_CHECK_LAUNCH();
  const int32_t* cur = sched;
  for (const TriI8Level& l : lv) {
    const int2* cr = reinterpret_cast<const int2*>(cur);
    const int2* rr = reinterpret_cast<const int2*>(cur + 2 * (size_t)n_pad);
    cur += 4 * (size_t)n_pad;
    const int* off1 = cur;
    cur += pad4(l.off[0].size());
    const int4* it1 = reinterpret_cast<const int4*>(cur);
    cur += 4 * l.items[0].size();
    const int* off2 = cur;
    cur += pad4(l.off[1].size());
    const int4* it2 = reinterpret_cast<const int4*>(cur);
    cur += 4 * l.items[1].size();

    I8Args a;
    a.n_rb = n_pad / I8_RB; a.ntiles = 0; a.slices = 1; a.tiles_in_flight = 1; a.debug = 0;
    a.kexp = nullptr; a.V = nullptr; a.partial = nullptr;
    a.meet = nullptr; a.meet_stride = 0; a.meet_waves = 0;
    a.ldc = n_pad;
    // columns of W11, then T = L21 W11
    DB_CUDA(cudaMemsetAsync(csW, 0, sizeof(double) * n_pad, st));
    i8_colmax_kernel<<<cgrid, 256, 0, st>>>(W, n_pad, reinterpret_cast<unsigned long long*>(csW), cr, 0);
    DB_CHECK_LAUNCH();
    i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(csW, n_pad);
    DB_CHECK_LAUNCH();
    i8_wtsplit_kernel<<<tgrid, 256, 0, st>>>(W, n_pad, csW, C8, cr, 0);
    DB_CHECK_LAUNCH();
    a.item_off = off1; a.items = it1; a.C = T; a.alpha = 1.0; a.rowscale_a = rsL; a.rowscale = csW;
    score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mL, mC, a);
    DB_CHECK_LAUNCH();
    // rows of W22, columns of T, then W21 = -W22 T
    i8_wsplit_kernel<<<n_pad, 256, 0, st>>>(W, n_pad, R8, rsW, rr);
    DB_CHECK_LAUNCH();
    DB_CUDA(cudaMemsetAsync(csT, 0, sizeof(double) * n_pad, st));
    i8_colmax_kernel<<<cgrid, 256, 0, st>>>(T, n_pad, reinterpret_cast<unsigned long long*>(csT), cr, 1);
    DB_CHECK_LAUNCH();
    i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(csT, n_pad);
    DB_CHECK_LAUNCH();
    i8_wtsplit_kernel<<<tgrid, 256, 0, st>>>(T, n_pad, csT, T8, cr, 1);
    DB_CHECK_LAUNCH();
    a.item_off = off2; a.items = it2; a.C = W; a.alpha = -1.0; a.rowscale_a = rsW; a.rowscale = csT;
    score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mR, mT, a);
    DB_CHECK_LAUNCH();
  }
  return DISTBO_OK;
}
