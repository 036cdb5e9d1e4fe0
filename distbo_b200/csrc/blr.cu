// Bayesian-linear-regression surrogate on deep features (algorithm='BLR').
// Reference: train/log_gaus_likelihood_feature.py:13-225, train/feature_map.py:12-46.
//
//   F_i   = [ emb(task(i)) | theta_i ]                     (emb: dataset embedding / meta-features / one-hot task)
//   Phi_i = MLP(F_i): tanh(F W1 + b1) W2  [ -> tanh(. + b2) W3 ]
//   A     = I + (alpha / sigma^2) Phi^T Phi = L L^T,  e = L^-1 Phi^T y
//   loss  = (1/sigma^2)(|y|^2 - (alpha/sigma^2)|e|^2) + N log sigma^2 + 2 sum log L_ii
//   mean* = (alpha/sigma^2) e^T L^-1 phi*,   var* = alpha |L^-1 phi*|^2
//
// Everything is linear in N (and in the number of candidates); the D x D factorisation (D <= 64) lives in one CTA.
// The first layer is split as  F W1 = emb(task) W1[:Pe] + theta W1[Pe:]:  the embedding part has only T distinct
// values (one per task) and is evaluated once per task (c1), so the per-row work starts from d inputs.
// Row kernels work on 64-row tiles; every small matrix product runs through tile_mm (register-blocked FP64 FMA on
// shared-memory tiles, weights zero-padded to 64 columns).  All reductions over rows go through per-tile partial
// results that are added in tile order, so results are run-to-run deterministic.
#include "../../include/distbo_b200.h"

#include "common.cuh"

namespace db {

constexpr int BLR_R = 64;         // rows per tile
constexpr int BLR_W = 64;         // widest layer / feature width; weights are zero-padded to this many columns
constexpr int BLR_LD = 66;        // leading dimension of activation tiles in shared memory
constexpr int BLR_DMAX = 16;      // hyper-parameter dimensions
constexpr int BLR_THREADS = 256;

// acc[h][t][c]: tile row = lane + 32 h, column = 2 warp + 16 t + c.
// out(r, j) = sum_{k < K} A(r, k) B[k][j];  A(r, k) = a[r lda + k]  (AT: a[k lda + r]);  B is [K][BLR_W].
template <bool AT>
__device__ __forceinline__ void tile_mm(const double* __restrict__ a, int lda, const double* __restrict__ B, int K,
                                        double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[h][t][0] = acc[h][t][1] = 0.0;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const double a0 = AT ? a[k * lda + lane] : a[lane * lda + k];
    const double a1 = AT ? a[k * lda + lane + 32] : a[(lane + 32) * lda + k];
    const double2* b = reinterpret_cast<const double2*>(B + k * BLR_W + 2 * w);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double2 bv = b[8 * t];
      acc[0][t][0] = fma(a0, bv.x, acc[0][t][0]);
      acc[0][t][1] = fma(a0, bv.y, acc[0][t][1]);
      acc[1][t][0] = fma(a1, bv.x, acc[1][t][0]);
      acc[1][t][1] = fma(a1, bv.y, acc[1][t][1]);
    }
  }
}

template <class F>
__device__ __forceinline__ void tile_each(const double (&acc)[2][4][2], F f) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int c = 0; c < 2; ++c) f(lane + 32 * h, 2 * w + 16 * t + c, acc[h][t][c]);
}

// dst [rows][BLR_W] <- src [K][J] (row-major, ld = J), zero padded.  TR: dst[j][k] = src[k][j].
template <bool TR>
__device__ __forceinline__ void stage_weight(double* dst, const double* __restrict__ src, int K, int J) {
  for (int idx = threadIdx.x; idx < BLR_W * BLR_W; idx += BLR_THREADS) {
    const int r = idx / BLR_W, c = idx % BLR_W;
    double v = 0.0;
    if (!TR) {
      if (r < K && c < J) v = src[r * J + c];
    } else {
      if (c < K && r < J) v = src[c * J + r];
    }
    dst[idx] = v;
  }
}

// tile [BLR_R][BLR_LD] <- src rows row0.. (ld, width), zero outside.
__device__ __forceinline__ void load_tile(double* tile, const double* __restrict__ src, int ld, int width,
                                          long long row0, long long n_rows) {
  for (int idx = threadIdx.x; idx < BLR_R * BLR_W; idx += BLR_THREADS) {
    const int r = idx / BLR_W, c = idx % BLR_W;
    double v = 0.0;
    if (row0 + r < n_rows && c < width) v = src[(row0 + r) * ld + c];
    tile[r * BLR_LD + c] = v;
  }
}

__device__ __forceinline__ void store_tile(const double* tile, double* __restrict__ dst, long long row0,
                                           long long n_rows) {
  for (int idx = threadIdx.x; idx < BLR_R * BLR_W; idx += BLR_THREADS) {
    const int r = idx / BLR_W, c = idx % BLR_W;
    if (row0 + r < n_rows) dst[(row0 + r) * BLR_W + c] = tile[r * BLR_LD + c];
  }
}

struct BlrNet {
  const double *W1, *b1, *W2, *b2, *W3;
  int n_layers, Pe, d, h1, h2, D;
};

// ---- per-task half of the first layer: c1[t] = emb[t] W1[:Pe] + b1  (zero padded to BLR_W) ----------------------
__global__ void __launch_bounds__(BLR_W) blr_task_pre_kernel(BlrNet net, const double* __restrict__ E, int ldE,
                                                            double* __restrict__ c1) {
  const int t = blockIdx.x, h = threadIdx.x;
  double v = 0.0;
  if (h < net.h1) {
    v = net.b1[h];
    for (int p = 0; p < net.Pe; ++p) v = fma(E[(size_t)t * ldE + p], net.W1[p * net.h1 + h], v);
  }
  c1[t * BLR_W + h] = v;
}

// Shared-memory layout of the row kernels (doubles).
constexpr int BLR_SM_W1T = 0;                                   // [BLR_DMAX][BLR_W]
constexpr int BLR_SM_W2 = BLR_SM_W1T + BLR_DMAX * BLR_W;        // [BLR_W][BLR_W]
constexpr int BLR_SM_W3 = BLR_SM_W2 + BLR_W * BLR_W;            // [BLR_W][BLR_W]
constexpr int BLR_SM_W4 = BLR_SM_W3 + BLR_W * BLR_W;            // [BLR_W][BLR_W]  (scoring: L^-T, backward: Q)
constexpr int BLR_SM_T0 = BLR_SM_W4 + BLR_W * BLR_W;            // [BLR_R][BLR_LD]
constexpr int BLR_SM_T1 = BLR_SM_T0 + BLR_R * BLR_LD;
constexpr int BLR_SM_VEC = BLR_SM_T1 + BLR_R * BLR_LD;          // 4 vectors of BLR_W
constexpr int BLR_SM_DOUBLES = BLR_SM_VEC + 4 * BLR_W;
constexpr int BLR_SMEM_BYTES = BLR_SM_DOUBLES * 8;

__device__ __forceinline__ void stage_theta_layer(double* sm, const BlrNet& net) {
  // rows Pe.. of W1 [Pe + d][h1]
  for (int idx = threadIdx.x; idx < BLR_DMAX * BLR_W; idx += BLR_THREADS) {
    const int k = idx / BLR_W, c = idx % BLR_W;
    sm[BLR_SM_W1T + idx] = (k < net.d && c < net.h1) ? net.W1[(net.Pe + k) * net.h1 + c] : 0.0;
  }
}

// Forward of the feature map on one tile.  T0 holds the theta tile on entry; on exit the tile returned holds Phi,
// `h1_tile` / `h2_tile` tell the caller where the hidden activations were left (h2 is overwritten by Phi when
// keep == false).  c1(row, col) supplies the per-task first-layer term.
template <class C1>
__device__ __forceinline__ double* blr_forward_tile(double* sm, const BlrNet& net, C1 c1, double* H1g, double* H2g,
                                                    long long row0, long long n_rows) {
  double acc[2][4][2];
  double* T0 = sm + BLR_SM_T0;
  double* T1 = sm + BLR_SM_T1;
  tile_mm<false>(T0, BLR_LD, sm + BLR_SM_W1T, net.d, acc);
  tile_each(acc, [&](int r, int c, double v) { T1[r * BLR_LD + c] = tanh(v + c1(r, c)); });
  __syncthreads();
  if (H1g) store_tile(T1, H1g, row0, n_rows);
  tile_mm<false>(T1, BLR_LD, sm + BLR_SM_W2, net.h1, acc);
  __syncthreads();                                   // every warp is done reading T1 / T0
  if (net.n_layers == 2) {
    tile_each(acc, [&](int r, int c, double v) { T0[r * BLR_LD + c] = v; });
    __syncthreads();
    return T0;
  }
  const double* b2 = sm + BLR_SM_VEC;
  tile_each(acc, [&](int r, int c, double v) { T0[r * BLR_LD + c] = tanh(v + b2[c]); });
  __syncthreads();
  if (H2g) store_tile(T0, H2g, row0, n_rows);
  tile_mm<false>(T0, BLR_LD, sm + BLR_SM_W3, net.h2, acc);
  tile_each(acc, [&](int r, int c, double v) { T1[r * BLR_LD + c] = v; });
  __syncthreads();
  return T1;
}

__global__ void __launch_bounds__(BLR_THREADS) blr_rows_fwd_kernel(BlrNet net, const double* __restrict__ theta,
                                                                  int N, const double* __restrict__ c1,
                                                                  const int32_t* __restrict__ task_id,
                                                                  double* __restrict__ H1, double* __restrict__ H2,
                                                                  double* __restrict__ Phi) {
  extern __shared__ double sm[];
  __shared__ int trow[BLR_R];
  const long long row0 = (long long)blockIdx.x * BLR_R;
  stage_theta_layer(sm, net);
  stage_weight<false>(sm + BLR_SM_W2, net.W2, net.h1, net.h2);
  if (net.n_layers == 3) {
    stage_weight<false>(sm + BLR_SM_W3, net.W3, net.h2, net.D);
    if (threadIdx.x < BLR_W) sm[BLR_SM_VEC + threadIdx.x] = threadIdx.x < net.h2 ? net.b2[threadIdx.x] : 0.0;
  }
  load_tile(sm + BLR_SM_T0, theta, net.d, net.d, row0, N);
  if (threadIdx.x < BLR_R) {
    const long long i = row0 + threadIdx.x;
    trow[threadIdx.x] = (task_id && i < N) ? task_id[i] : 0;
  }
  __syncthreads();
  double* P = blr_forward_tile(
      sm, net, [&](int r, int c) { return c1[trow[r] * BLR_W + c]; }, H1, net.n_layers == 3 ? H2 : nullptr, row0, N);
  // rows beyond N carry tanh(c1) garbage: store_tile masks them, and the statistics pass reloads with zero fill
  store_tile(P, Phi, row0, N);
}

// ---- A^T B over the rows, one partial result per 64-row tile ------------------------------------------------------
struct AtbPair {
  const double* A;
  int lda, Ka;
  const double* B;
  int ldb, Jb;
  double* out;    // [tiles][Ka * Jb + Jb]: product, then the column sums of B
};
struct AtbArgs {
  AtbPair p[4];
  int N;
};

__global__ void __launch_bounds__(BLR_THREADS) blr_atb_kernel(AtbArgs a) {
  extern __shared__ double sm[];
  double* TA = sm;                      // [BLR_R][BLR_LD]
  double* TB = sm + BLR_R * BLR_LD;     // [BLR_R][BLR_W]
  const AtbPair p = a.p[blockIdx.y];
  const long long row0 = (long long)blockIdx.x * BLR_R;
  load_tile(TA, p.A, p.lda, p.Ka, row0, a.N);
  for (int idx = threadIdx.x; idx < BLR_R * BLR_W; idx += BLR_THREADS) {
    const int r = idx / BLR_W, c = idx % BLR_W;
    TB[idx] = (row0 + r < a.N && c < p.Jb) ? p.B[(row0 + r) * p.ldb + c] : 0.0;
  }
  __syncthreads();
  double acc[2][4][2];
  tile_mm<true>(TA, BLR_LD, TB, BLR_R, acc);
  const int rec = p.Ka * p.Jb + p.Jb;
  double* out = p.out + (size_t)blockIdx.x * rec;
  tile_each(acc, [&](int i, int j, double v) {
    if (i < p.Ka && j < p.Jb) out[i * p.Jb + j] = v;
  });
  if (threadIdx.x < p.Jb) {
    double s = 0.0;
    for (int r = 0; r < BLR_R; ++r) s += TB[r * BLR_W + threadIdx.x];
    out[p.Ka * p.Jb + threadIdx.x] = s;
  }
}

// dst[i] = sum over tiles (in tile order) of partial[tile * rec + off + i]
struct ReduceSeg {
  const double* partial;
  int rec, off, n;
  double* dst;
};
struct ReduceArgs {
  ReduceSeg s[6];
  int tiles;
};
__global__ void __launch_bounds__(256) blr_reduce_kernel(ReduceArgs a) {
  const ReduceSeg s = a.s[blockIdx.y];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (!s.dst || i >= s.n) return;
  const double* src = s.partial + s.off + i;
  double v = 0.0;
  int t = 0;
  for (; t + 8 <= a.tiles; t += 8) {      // eight loads in flight, added in tile order (fixed association)
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = src[(size_t)(t + u) * s.rec];
#pragma unroll
    for (int u = 0; u < 8; ++u) v += x[u];
  }
  for (; t < a.tiles; ++t) v += src[(size_t)t * s.rec];
  s.dst[i] = v;
}

// ---- D x D solve: A = I + c G, Cholesky, L^-1, e, loss and the seeds of the backward pass -------------------------
struct SolveArgs {
  const double *G, *b, *yy;          // Phi^T Phi [D*D], Phi^T y [D], y^T y (already summed over the tiles)
  int N, D;
  const double *log_sigma, *log_alpha;
  double *Linv, *evec;               // [BLR_W][BLR_W] (lower, zero padded), [BLR_W]
  double* loss;
  int32_t* info;
  double *Q, *rvec;                  // dPhi = Phi Q + y r^T           ([BLR_W][BLR_W] zero padded, [BLR_W])
  double *g_log_sigma, *g_log_alpha; // may be NULL (prediction only)
};

__global__ void __launch_bounds__(BLR_THREADS) blr_solve_kernel(SolveArgs a) {
  extern __shared__ double sm[];
  double* G = sm;                         // [D][D] dense, ld = BLR_W
  double* L = G + BLR_W * BLR_W;
  double* Li = L + BLR_W * BLR_W;
  double* Ai = Li + BLR_W * BLR_W;
  double* b = Ai + BLR_W * BLR_W;         // [BLR_W]
  double* e = b + BLR_W;
  double* w = e + BLR_W;
  double* red = w + BLR_W;                // [BLR_THREADS]
  __shared__ int bad;
  const int tid = threadIdx.x, D = a.D;
  const double sigma_sq = exp(2.0 * *a.log_sigma);
  const double s = 1.0 / sigma_sq;
  const double alpha = exp(*a.log_alpha);
  const double c = alpha * s;
  const double yy_s = *a.yy;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < BLR_W * BLR_W; idx += BLR_THREADS) {
    const int i = idx / BLR_W, j = idx % BLR_W;
    const double g = (i < D && j < D) ? a.G[i * D + j] : 0.0;
    G[idx] = g;
    L[idx] = (i < D && j < D) ? ((i == j ? 1.0 : 0.0) + c * g) : 0.0;
    Li[idx] = 0.0;
  }
  if (tid < BLR_W) b[tid] = tid < D ? a.b[tid] : 0.0;
  __syncthreads();
  // Right-looking Cholesky, one barrier per column: the trailing update uses the unscaled column,
  // A[i][j] -= A[i][k] A[j][k] / A[k][k]; the columns are scaled by 1 / sqrt(pivot) afterwards.
  for (int k = 0; k < D; ++k) {
    const double piv = L[k * BLR_W + k];
    if (tid == 0 && !(piv > 0.0) && !bad) bad = k + 1;
    const double inv = 1.0 / piv;
    const int m = D - k - 1;
    for (int idx = tid; idx < m * m; idx += BLR_THREADS) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      if (j <= i) L[i * BLR_W + j] -= L[i * BLR_W + k] * L[j * BLR_W + k] * inv;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < D * D; idx += BLR_THREADS) {
    const int i = idx / D, j = idx % D;
    if (j > i) L[i * BLR_W + j] = 0.0;
    else if (j < i) L[i * BLR_W + j] /= sqrt(L[j * BLR_W + j]);      // diagonal still holds the pivots
  }
  __syncthreads();
  if (tid < D) L[tid * BLR_W + tid] = sqrt(L[tid * BLR_W + tid]);
  __syncthreads();
  // L^-1: column j by forward substitution (thread = column)
  if (tid < D) {
    const int j = tid;
    for (int i = j; i < D; ++i) {
      double v = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) v -= L[i * BLR_W + k] * Li[k * BLR_W + j];
      Li[i * BLR_W + j] = v / L[i * BLR_W + i];
    }
  }
  __syncthreads();
  if (tid < BLR_W) {                      // e = L^-1 b
    double v = 0.0;
    if (tid < D)
      for (int k = 0; k <= tid; ++k) v = fma(Li[tid * BLR_W + k], b[k], v);
    e[tid] = v;
  }
  __syncthreads();
  if (tid < BLR_W) {                      // w = L^-T e = A^-1 b
    double v = 0.0;
    if (tid < D)
      for (int k = tid; k < D; ++k) v = fma(Li[k * BLR_W + tid], e[k], v);
    w[tid] = v;
  }
  for (int idx = tid; idx < BLR_W * BLR_W; idx += BLR_THREADS) {   // A^-1 = L^-T L^-1
    const int i = idx / BLR_W, j = idx % BLR_W;
    double v = 0.0;
    if (i < D && j < D)
      for (int k = (i > j ? i : j); k < D; ++k) v = fma(Li[k * BLR_W + i], Li[k * BLR_W + j], v);
    Ai[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < BLR_W * BLR_W; idx += BLR_THREADS) {
    const int i = idx / BLR_W, j = idx % BLR_W;
    a.Linv[idx] = Li[idx];
    if (a.Q) a.Q[idx] = (i < D && j < D) ? 2.0 * (s * c * c * w[i] * w[j] + c * Ai[idx]) : 0.0;
  }
  if (tid < BLR_W) {
    a.evec[tid] = e[tid];
    if (a.rvec) a.rvec[tid] = -2.0 * s * c * w[tid];
  }
  // scalars: |e|^2 = b.w, sum log L_ii, w^T G w, tr(A^-1 G)
  double part_wGw = 0.0, part_tr = 0.0;
  for (int idx = tid; idx < D * D; idx += BLR_THREADS) {
    const int i = idx / D, j = idx % D;
    part_wGw += w[i] * G[i * BLR_W + j] * w[j];
    part_tr += Ai[i * BLR_W + j] * G[i * BLR_W + j];
  }
  auto block_sum = [&](double v) {
    red[tid] = v;
    __syncthreads();
    for (int o = BLR_THREADS / 2; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  const double wGw = block_sum(part_wGw);
  const double trAG = block_sum(part_tr);
  if (tid == 0) {
    double ee = 0.0, bw = 0.0, logdet = 0.0;
    for (int k = 0; k < D; ++k) {
      ee += e[k] * e[k];
      bw += b[k] * w[k];
      logdet += log(L[k * BLR_W + k]);
    }
    *a.loss = s * (yy_s - c * ee) + log(sigma_sq) * (double)a.N + 2.0 * logdet;
    *a.info = bad;
    if (a.g_log_sigma) {
      const double g_c = -s * bw + s * c * wGw + trAG;       // d loss / d c   (c = alpha / sigma^2)
      const double g_s = yy_s - c * bw;                      // d loss / d (1 / sigma^2), c held fixed
      *a.g_log_alpha = c * g_c;
      *a.g_log_sigma = -2.0 * s * g_s - 2.0 * c * g_c + 2.0 * (double)a.N;
    }
  }
}

// ---- backward through the feature map on one tile -------------------------------------------------------------------
// dPhi = Phi Q + y r^T -> (3 layers: dz2 = (dPhi W3^T)(1 - h2^2)) -> da1 = (dz2 W2^T)(1 - h1^2); all three written out.
__global__ void __launch_bounds__(BLR_THREADS) blr_rows_bwd_kernel(BlrNet net, int N, const double* __restrict__ y,
                                                                  const double* __restrict__ Q,
                                                                  const double* __restrict__ rvec,
                                                                  const double* __restrict__ H1,
                                                                  const double* __restrict__ H2,
                                                                  const double* __restrict__ Phi,
                                                                  double* __restrict__ DPhi, double* __restrict__ DZ2,
                                                                  double* __restrict__ DA1) {
  extern __shared__ double sm[];
  const long long row0 = (long long)blockIdx.x * BLR_R;
  double* T0 = sm + BLR_SM_T0;
  double* T1 = sm + BLR_SM_T1;
  double* rv = sm + BLR_SM_VEC;
  double* ys = sm + BLR_SM_VEC + BLR_W;
  for (int idx = threadIdx.x; idx < BLR_W * BLR_W; idx += BLR_THREADS) sm[BLR_SM_W4 + idx] = Q[idx];
  stage_weight<true>(sm + BLR_SM_W2, net.W2, net.h1, net.h2);                            // W2^T [h2][h1]
  if (net.n_layers == 3) stage_weight<true>(sm + BLR_SM_W3, net.W3, net.h2, net.D);      // W3^T [D][h2]
  if (threadIdx.x < BLR_W) {
    rv[threadIdx.x] = rvec[threadIdx.x];
    ys[threadIdx.x] = (row0 + threadIdx.x < N) ? y[row0 + threadIdx.x] : 0.0;
  }
  load_tile(T0, Phi, BLR_W, BLR_W, row0, N);
  __syncthreads();
  double acc[2][4][2];
  tile_mm<false>(T0, BLR_LD, sm + BLR_SM_W4, net.D, acc);
  tile_each(acc, [&](int r, int c, double v) { T1[r * BLR_LD + c] = v + ys[r] * rv[c]; });
  __syncthreads();
  store_tile(T1, DPhi, row0, N);
  double* dz2 = T1;
  if (net.n_layers == 3) {
    tile_mm<false>(T1, BLR_LD, sm + BLR_SM_W3, net.D, acc);
    tile_each(acc, [&](int r, int c, double v) {
      const double h = (row0 + r < N) ? H2[(row0 + r) * BLR_W + c] : 0.0;
      T0[r * BLR_LD + c] = v * (1.0 - h * h);
    });
    __syncthreads();
    store_tile(T0, DZ2, row0, N);
    dz2 = T0;
  }
  double* da1 = (dz2 == T0) ? T1 : T0;
  tile_mm<false>(dz2, BLR_LD, sm + BLR_SM_W2, net.h2, acc);
  __syncthreads();
  tile_each(acc, [&](int r, int c, double v) {
    const double h = (row0 + r < N) ? H1[(row0 + r) * BLR_W + c] : 0.0;
    da1[r * BLR_LD + c] = v * (1.0 - h * h);
  });
  __syncthreads();
  store_tile(da1, DA1, row0, N);
}

// S1[t] = sum of DA1 over the rows of task t (row order);  dE[t] = S1[t] W1[:Pe]^T
__global__ void __launch_bounds__(BLR_W) blr_task_bwd_kernel(BlrNet net, const double* __restrict__ DA1,
                                                            const int32_t* __restrict__ task_off,
                                                            double* __restrict__ S1, double* __restrict__ dE, int ldE) {
  __shared__ double s1[BLR_W];
  const int t = blockIdx.x, h = threadIdx.x;
  double v = 0.0;
  int i = task_off[t];
  const int i_end = task_off[t + 1];
  for (; i + 8 <= i_end; i += 8) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = DA1[(size_t)(i + u) * BLR_W + h];
#pragma unroll
    for (int u = 0; u < 8; ++u) v += x[u];
  }
  for (; i < i_end; ++i) v += DA1[(size_t)i * BLR_W + h];
  s1[h] = v;
  S1[t * BLR_W + h] = v;
  __syncthreads();
  if (dE)
    for (int p = h; p < net.Pe; p += BLR_W) {
      double g = 0.0;
      for (int k = 0; k < net.h1; ++k) g = fma(s1[k], net.W1[p * net.h1 + k], g);
      dE[(size_t)t * ldE + p] = g;
    }
}

// gW1[p][h] = sum_t E[t][p] S1[t][h]   (p < Pe)
__global__ void __launch_bounds__(256) blr_task_finish_kernel(BlrNet net, const double* __restrict__ E, int ldE,
                                                             int T, const double* __restrict__ S1,
                                                             double* __restrict__ gW1) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= net.Pe * net.h1) return;
  const int p = idx / net.h1, h = idx % net.h1;
  double v = 0.0;
  int t = 0;
  for (; t + 8 <= T; t += 8) {            // loads in flight together, accumulated in task order
    double x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x[u] = E[(size_t)(t + u) * ldE + p];
      y[u] = S1[(t + u) * BLR_W + h];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) v = fma(x[u], y[u], v);
  }
  for (; t < T; ++t) v = fma(E[(size_t)t * ldE + p], S1[t * BLR_W + h], v);
  gW1[idx] = v;
}

// ---- scoring: feature map, u = L^-1 phi, posterior moments, acquisition, running best -------------------------------
struct BlrScoreArgs {
  BlrNet net;
  const double *cand, *shift, *scale;   // [M, d] raw candidates and the frozen scaler (may be NULL)
  long long M, index_offset;
  const double* emb;                    // [Pe] embedding of the candidates' task
  const double *Linv, *evec, *log_sigma, *log_alpha;
  int acq_kind;
  double y_max, xi, kappa;
  double *out_mean, *out_std, *out_acq;
  double* block_best_val;
  long long* block_best_idx;
};

__global__ void __launch_bounds__(BLR_THREADS) blr_score_kernel(BlrScoreArgs a) {
  extern __shared__ double sm[];
  __shared__ double bval[BLR_R];
  __shared__ long long bidx[BLR_R];
  const BlrNet& net = a.net;
  double* c1 = sm + BLR_SM_VEC + BLR_W;
  double* ev = sm + BLR_SM_VEC + 2 * BLR_W;
  stage_theta_layer(sm, net);
  stage_weight<false>(sm + BLR_SM_W2, net.W2, net.h1, net.h2);
  if (net.n_layers == 3) stage_weight<false>(sm + BLR_SM_W3, net.W3, net.h2, net.D);
  for (int idx = threadIdx.x; idx < BLR_W * BLR_W; idx += BLR_THREADS) {        // L^-T, so that u = phi L^-T
    const int k = idx / BLR_W, j = idx % BLR_W;
    sm[BLR_SM_W4 + idx] = a.Linv[j * BLR_W + k];
  }
  if (threadIdx.x < BLR_W) {
    const int h = threadIdx.x;
    sm[BLR_SM_VEC + h] = (net.n_layers == 3 && h < net.h2) ? net.b2[h] : 0.0;
    double v = 0.0;
    if (h < net.h1) {
      v = net.b1[h];
      for (int p = 0; p < net.Pe; ++p) v = fma(a.emb[p], net.W1[p * net.h1 + h], v);
    }
    c1[h] = v;
    ev[h] = a.evec[h];
  }
  const double sigma_sq = exp(2.0 * *a.log_sigma);
  const double alpha = exp(*a.log_alpha);
  const double coef = alpha / sigma_sq;
  double best = -INFINITY;
  long long best_i = 0x7fffffffffffffffLL;
  const long long tiles = (a.M + BLR_R - 1) / BLR_R;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row0 = tile * BLR_R;
    __syncthreads();
    double* T0 = sm + BLR_SM_T0;
    for (int idx = threadIdx.x; idx < BLR_R * BLR_DMAX; idx += BLR_THREADS) {
      const int r = idx / BLR_DMAX, k = idx % BLR_DMAX;
      double v = 0.0;
      if (row0 + r < a.M && k < net.d) {
        v = a.cand[(row0 + r) * net.d + k];
        if (a.shift) v = (v - a.shift[k]) / a.scale[k];
      }
      T0[r * BLR_LD + k] = v;
    }
    __syncthreads();
    double* P = blr_forward_tile(sm, net, [&](int, int c) { return c1[c]; }, nullptr, nullptr, row0, a.M);
    double* U = (P == T0) ? sm + BLR_SM_T1 : T0;
    double acc[2][4][2];
    tile_mm<false>(P, BLR_LD, sm + BLR_SM_W4, net.D, acc);
    tile_each(acc, [&](int r, int c, double v) { U[r * BLR_LD + c] = v; });
    __syncthreads();
    if (threadIdx.x < BLR_R) {
      const int r = threadIdx.x;
      const long long cidx = row0 + r;
      double val = -INFINITY;
      long long idx = 0x7fffffffffffffffLL;
      if (cidx < a.M) {
        double dot = 0.0, sq = 0.0;
        for (int j = 0; j < net.D; ++j) {
          const double u = U[r * BLR_LD + j];
          dot = fma(ev[j], u, dot);
          sq = fma(u, u, sq);
        }
        const double mean = coef * dot;
        const double std = sqrt(fmax(alpha * sq, kVarClamp));
        const double acqv = acquisition(a.acq_kind, mean, std, a.y_max, a.xi, a.kappa);
        if (a.out_mean) a.out_mean[cidx] = mean;
        if (a.out_std) a.out_std[cidx] = std;
        if (a.out_acq) a.out_acq[cidx] = acqv;
        val = acqv;
        idx = a.index_offset + cidx;
      }
      if (val > best || (val == best && idx < best_i)) {
        best = val;
        best_i = idx;
      }
    }
  }
  if (threadIdx.x < BLR_R) {
    bval[threadIdx.x] = best;
    bidx[threadIdx.x] = best_i;
  }
  __syncthreads();
  for (int s2 = BLR_R / 2; s2 > 0; s2 >>= 1) {
    if (threadIdx.x < s2) {
      const double v2 = bval[threadIdx.x + s2];
      const long long i2 = bidx[threadIdx.x + s2];
      if (v2 > bval[threadIdx.x] || (v2 == bval[threadIdx.x] && i2 < bidx[threadIdx.x])) {
        bval[threadIdx.x] = v2;
        bidx[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.block_best_val[blockIdx.x] = bval[0];
    a.block_best_idx[blockIdx.x] = bidx[0];
  }
}

__global__ void blr_merge_best_kernel(const double* __restrict__ bv, const long long* __restrict__ bi, int nblocks,
                                      double* best_val, long long* best_idx) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v = -INFINITY;
  long long i = 0x7fffffffffffffffLL;
  for (int b = 0; b < nblocks; ++b)
    if (bv[b] > v || (bv[b] == v && bi[b] < i)) {
      v = bv[b];
      i = bi[b];
    }
  *best_val = v;
  *best_idx = i;
}

static bool blr_net_ok(const BlrNet& n) {
  if (!n.W1 || !n.b1 || !n.W2) return false;
  if (n.n_layers != 2 && n.n_layers != 3) return false;
  if (n.n_layers == 3 && (!n.b2 || !n.W3)) return false;
  if (n.d <= 0 || n.d > BLR_DMAX || n.Pe < 0) return false;
  if (n.h1 <= 0 || n.h1 > BLR_W || n.h2 <= 0 || n.h2 > BLR_W || n.D <= 0 || n.D > BLR_W) return false;
  if (n.n_layers == 2 && n.D != n.h2) return false;
  return true;
}

static int blr_sms() { return device_sm_count(); }

static int blr_configure() {
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(blr_rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLR_SMEM_BYTES));
    DB_CUDA(cudaFuncSetAttribute(blr_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLR_SMEM_BYTES));
    DB_CUDA(cudaFuncSetAttribute(blr_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLR_SMEM_BYTES));
    DB_CUDA(cudaFuncSetAttribute(blr_atb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((BLR_R * BLR_LD + BLR_R * BLR_W) * sizeof(double))));
    DB_CUDA(cudaFuncSetAttribute(blr_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((4 * BLR_W * BLR_W + 3 * BLR_W + BLR_THREADS) * sizeof(double))));
    once.mark();
  }
  return DISTBO_OK;
}

}  // namespace db

using namespace db;

static inline int64_t blr_tiles(int64_t N) { return (N + BLR_R - 1) / BLR_R; }

// workspace layout (doubles); T >= 1
struct BlrWs {
  double *c1, *S1, *H1, *H2, *Phi, *DPhi, *DZ2, *DA1, *Q, *rvec, *G, *bvec, *pGG, *pGy, *pyy, *pW3, *pW2, *pW1;
  int64_t total;
};
static BlrWs blr_ws(double* base, int64_t N, int T, int d, int h1, int h2, int D) {
  BlrWs w;
  int64_t o = 0;
  auto take = [&](int64_t n) {
    double* p = base ? base + o : nullptr;
    o += (n + 1) & ~int64_t(1);
    return p;
  };
  const int64_t tiles = blr_tiles(N);
  w.c1 = take((int64_t)T * BLR_W);
  w.S1 = take((int64_t)T * BLR_W);
  w.H1 = take(N * BLR_W);
  w.H2 = take(N * BLR_W);
  w.Phi = take(N * BLR_W);
  w.DPhi = take(N * BLR_W);
  w.DZ2 = take(N * BLR_W);
  w.DA1 = take(N * BLR_W);
  w.Q = take(BLR_W * BLR_W);
  w.rvec = take(BLR_W);
  w.G = take(BLR_W * BLR_W);
  w.bvec = take(2 * BLR_W);
  w.pGG = take(tiles * ((int64_t)D * D + D));
  w.pGy = take(tiles * ((int64_t)D + 1));
  w.pyy = take(tiles * 2);
  w.pW3 = take(tiles * ((int64_t)h2 * D + D));
  w.pW2 = take(tiles * ((int64_t)h1 * h2 + h2));
  w.pW1 = take(tiles * ((int64_t)d * h1 + h1));
  w.total = o;
  return w;
}

extern "C" int64_t distbo_blr_workspace(int64_t N, int T, int d, int h1, int h2, int D) {
  if (N <= 0 || T <= 0 || d <= 0 || h1 <= 0 || h2 <= 0 || D <= 0) return -1;
  return blr_ws(nullptr, N, T, d, h1, h2, D).total;
}

extern "C" int distbo_blr_fit_f64(const double* theta, const double* y, int N, int d, const double* E, int ldE, int Pe,
                                  const int32_t* task_id, const int32_t* task_off, int T, const double* W1,
                                  const double* b1, const double* W2, const double* b2, const double* W3, int n_layers,
                                  int h1, int h2, int D, const double* log_sigma, const double* log_alpha, double* Linv,
                                  double* evec, double* loss, int32_t* info, double* gW1, double* gb1, double* gW2,
                                  double* gb2, double* gW3, double* g_log_sigma, double* g_log_alpha, double* dE,
                                  double* workspace, int want_grad, void* stream) {
  BlrNet net{W1, b1, W2, b2, W3, n_layers, Pe, d, h1, h2, D};
  if (!blr_net_ok(net) || !theta || !y || N <= 0 || !log_sigma || !log_alpha || !Linv || !evec || !loss || !info ||
      !workspace)
    return DISTBO_EARG;
  if (Pe > 0 && (!E || ldE < Pe || !task_id || !task_off || T <= 0)) return DISTBO_EARG;
  if (Pe == 0) T = 1;
  if (want_grad && (!gW1 || !gb1 || !gW2 || !g_log_sigma || !g_log_alpha || (n_layers == 3 && (!gb2 || !gW3))))
    return DISTBO_EARG;
  if (blr_configure() != DISTBO_OK) return DISTBO_ECUDA;
  cudaStream_t st = (cudaStream_t)stream;
  BlrWs w = blr_ws(workspace, N, T, d, h1, h2, D);
  const int tiles = (int)blr_tiles(N);

  blr_task_pre_kernel<<<T, BLR_W, 0, st>>>(net, E, ldE, w.c1);
  DB_CHECK_LAUNCH();
  blr_rows_fwd_kernel<<<tiles, BLR_THREADS, BLR_SMEM_BYTES, st>>>(net, theta, N, w.c1, Pe > 0 ? task_id : nullptr, w.H1,
                                                                  w.H2, w.Phi);
  DB_CHECK_LAUNCH();
  const size_t atb_smem = (BLR_R * BLR_LD + BLR_R * BLR_W) * sizeof(double);
  AtbArgs s{};
  s.N = N;
  s.p[0] = {w.Phi, BLR_W, D, w.Phi, BLR_W, D, w.pGG};
  s.p[1] = {w.Phi, BLR_W, D, y, 1, 1, w.pGy};
  s.p[2] = {y, 1, 1, y, 1, 1, w.pyy};
  blr_atb_kernel<<<dim3(tiles, 3), BLR_THREADS, atb_smem, st>>>(s);
  DB_CHECK_LAUNCH();
  ReduceArgs rs{};
  rs.tiles = tiles;
  rs.s[0] = {w.pGG, D * D + D, 0, D * D, w.G};
  rs.s[1] = {w.pGy, D + 1, 0, D, w.bvec};
  rs.s[2] = {w.pyy, 2, 0, 1, w.bvec + BLR_W};
  blr_reduce_kernel<<<dim3((BLR_W * BLR_W + 255) / 256, 3), 256, 0, st>>>(rs);
  DB_CHECK_LAUNCH();
  SolveArgs sa{w.G, w.bvec, w.bvec + BLR_W, N, D, log_sigma, log_alpha, Linv, evec, loss, info,
               want_grad ? w.Q : nullptr, want_grad ? w.rvec : nullptr, want_grad ? g_log_sigma : nullptr,
               want_grad ? g_log_alpha : nullptr};
  blr_solve_kernel<<<1, BLR_THREADS, (4 * BLR_W * BLR_W + 3 * BLR_W + BLR_THREADS) * sizeof(double), st>>>(sa);
  DB_CHECK_LAUNCH();
  if (!want_grad) return DISTBO_OK;

  blr_rows_bwd_kernel<<<tiles, BLR_THREADS, BLR_SMEM_BYTES, st>>>(net, N, y, w.Q, w.rvec, w.H1, w.H2, w.Phi, w.DPhi,
                                                                  w.DZ2, w.DA1);
  DB_CHECK_LAUNCH();
  // weight gradients: A^T B partials per tile, then the fixed-order sums
  const double* dz2 = n_layers == 3 ? w.DZ2 : w.DPhi;
  AtbArgs g{};
  g.N = N;
  g.p[0] = {w.H1, BLR_W, h1, dz2, BLR_W, h2, w.pW2};
  g.p[1] = {theta, d, d, w.DA1, BLR_W, h1, w.pW1};
  int pairs = 2;
  if (n_layers == 3) g.p[pairs++] = {w.H2, BLR_W, h2, w.DPhi, BLR_W, D, w.pW3};
  blr_atb_kernel<<<dim3(tiles, pairs), BLR_THREADS, atb_smem, st>>>(g);
  DB_CHECK_LAUNCH();
  ReduceArgs r{};
  r.tiles = tiles;
  r.s[0] = {w.pW2, h1 * h2 + h2, 0, h1 * h2, gW2};
  r.s[1] = {w.pW2, h1 * h2 + h2, h1 * h2, h2, n_layers == 3 ? gb2 : nullptr};
  r.s[2] = {w.pW1, d * h1 + h1, 0, d * h1, gW1 + (size_t)Pe * h1};
  r.s[3] = {w.pW1, d * h1 + h1, d * h1, h1, gb1};
  int segs = 4;
  if (n_layers == 3) r.s[segs++] = {w.pW3, h2 * D + D, 0, h2 * D, gW3};
  blr_reduce_kernel<<<dim3((BLR_W * BLR_W + 255) / 256, segs), 256, 0, st>>>(r);
  DB_CHECK_LAUNCH();
  if (Pe > 0) {
    blr_task_bwd_kernel<<<T, BLR_W, 0, st>>>(net, w.DA1, task_off, w.S1, dE, ldE);
    DB_CHECK_LAUNCH();
    blr_task_finish_kernel<<<(Pe * h1 + 255) / 256, 256, 0, st>>>(net, E, ldE, T, w.S1, gW1);
    DB_CHECK_LAUNCH();
  }
  return DISTBO_OK;
}

extern "C" int distbo_blr_score_blocks(void) { return 2 * blr_sms(); }

extern "C" int distbo_blr_score_f64(const double* cand, const double* cand_shift, const double* cand_scale, int64_t M,
                                    int d, const double* emb, int Pe, const double* W1, const double* b1,
                                    const double* W2, const double* b2, const double* W3, int n_layers, int h1, int h2,
                                    int D, const double* Linv, const double* evec, const double* log_sigma,
                                    const double* log_alpha, int acq_kind, double y_max, double xi, double kappa,
                                    double* out_mean, double* out_std, double* out_acq, double* block_best_val,
                                    int64_t* block_best_idx, double* best_val, int64_t* best_idx, int64_t index_offset,
                                    void* stream) {
  BlrNet net{W1, b1, W2, b2, W3, n_layers, Pe, d, h1, h2, D};
  if (!blr_net_ok(net) || !cand || M <= 0 || (Pe > 0 && !emb) || !Linv || !evec || !log_sigma || !log_alpha ||
      !block_best_val || !block_best_idx || !best_val || !best_idx)
    return DISTBO_EARG;
  if (acq_kind < 0 || acq_kind > 2 || ((cand_shift == nullptr) != (cand_scale == nullptr))) return DISTBO_EARG;
  if (blr_configure() != DISTBO_OK) return DISTBO_ECUDA;
  cudaStream_t st = (cudaStream_t)stream;
  BlrScoreArgs a{net, cand, cand_shift, cand_scale, (long long)M, (long long)index_offset, emb, Linv, evec, log_sigma,
                 log_alpha, acq_kind, y_max, xi, kappa, out_mean, out_std, out_acq, block_best_val,
                 (long long*)block_best_idx};
  const long long tiles = blr_tiles(M);
  int grid = blr_sms();             // one persistent CTA per SM (170 KB of shared memory each)
  if (tiles < grid) grid = (int)tiles;
  blr_score_kernel<<<grid, BLR_THREADS, BLR_SMEM_BYTES, st>>>(a);
  DB_CHECK_LAUNCH();
  blr_merge_best_kernel<<<1, 32, 0, st>>>(block_best_val, (const long long*)block_best_idx, grid, best_val,
                                          (long long*)best_idx);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
