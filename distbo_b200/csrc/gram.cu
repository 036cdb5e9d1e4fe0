// K2: Matern-3/2 product Gram over (theta, dataset embedding), its hyper-parameter gradient
// contraction, the task-level (T x T) embedding Gram and its backward, and the TF-style Adam step.
// Reference: train/kernels.py:19-28,58-88; train/log_gaus_likelihood.py:86-122; train/train.py:31-40.
#include "../../include/distbo_b200.h"
#include "common.cuh"

namespace db {

constexpr int DMAX = 16;  // max hyper-parameter dimension handled in registers

// ------------------------------------------------------------------ theta / bandwidth scaling
__global__ void scale_theta_kernel(const double* __restrict__ theta, int N, int d,
                                   const double* __restrict__ log_bw, double* __restrict__ theta_s,
                                   double* __restrict__ sqnorm, int n_rows_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows_out) return;
  double s = 0.0;
  for (int k = 0; k < d; ++k) {
    double v = 0.0;
    if (i < N) v = theta[(size_t)i * d + k] / exp(log_bw[k]);  // tf.divide(X, stddev), kernels.py:20
    theta_s[(size_t)i * d + k] = v;
    s += v * v;
  }
  sqnorm[i] = s;
}

// ------------------------------------------------------------------ task-level Gram  KE [T,T]
// one warp per (t,t') pair, expanded-distance form as kernels.py:22-25
__global__ void __launch_bounds__(256) task_gram_kernel(const double* __restrict__ E, int lde, int T,
                                                        int P, const double* __restrict__ log_bw_embed,
                                                        double* __restrict__ KE) {
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= T * T) return;
  const int t = pair / T, u = pair % T;
  double xy = 0.0, xx = 0.0, yy = 0.0;
  for (int c = lane; c < P; c += 32) {
    const double s = exp(log_bw_embed[c]);
    const double a = E[(size_t)t * lde + c] / s, b = E[(size_t)u * lde + c] / s;
    xy += a * b;
    xx += a * a;
    yy += b * b;
  }
  xy = warp_sum(xy);
  xx = warp_sum(xx);
  yy = warp_sum(yy);
  if (lane == 0) KE[pair] = matern32_from_d2(-2.0 * xy + xx + yy);
}

// ------------------------------------------------------------------ N x N Gram, lower 128-tiles
constexpr int GT = 128;  // tile edge
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ theta_s,
                                                   const double* __restrict__ sqnorm, int N, int n_pad,
                                                   int d, const int32_t* __restrict__ task_id,
                                                   const double* __restrict__ KE, int T,
                                                   const double* __restrict__ log_scale,
                                                   const double* __restrict__ log_sigma, double jitter,
                                                   double* __restrict__ K) {
  extern __shared__ __align__(16) double sm[];
  // lower-triangular tile index -> (bi, bj), bj <= bi
  int bi = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  double* sI = sm;                 // [GT][d]   rows of the i block
  double* sJt = sm + GT * d;       // [d][GT]   transposed rows of the j block
  double* sqI = sJt + GT * d;      // [GT]
  double* sqJ = sqI + GT;          // [GT]
  int* tI = (int*)(sqJ + GT);      // [GT]
  int* tJ = tI + GT;               // [GT]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < GT * d; idx += 256) {
    const int r = idx / d, k = idx % d;
    sI[idx] = theta_s[(size_t)(bi * GT + r) * d + k];
    sJt[k * GT + r] = theta_s[(size_t)(bj * GT + r) * d + k];
  }
  for (int r = tid; r < GT; r += 256) {
    const int gi = bi * GT + r, gj = bj * GT + r;
    sqI[r] = sqnorm[gi];
    sqJ[r] = sqnorm[gj];
    tI[r] = (gi < N && task_id) ? task_id[gi] : 0;
    tJ[r] = (gj < N && task_id) ? task_id[gj] : 0;
  }
  __syncthreads();
  const double scale = exp(*log_scale);
  const double diag_add = exp(2.0 * (*log_sigma)) + jitter;
  const int lane = tid & 31;  // columns lane, lane+32, lane+64, lane+96: conflict-free shared reads, and a warp
  const int w = tid >> 5;     // stores four 256-byte row segments
  for (int r = w; r < GT; r += 8) {
    const int gi = bi * GT + r;
    double dot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < d; ++k) {
      const double xi = sI[r * d + k];
#pragma unroll
      for (int q = 0; q < 4; ++q) dot[q] += xi * sJt[k * GT + lane + 32 * q];
    }
    double* dst = K + (size_t)gi * n_pad + bj * GT;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q, gj = bj * GT + c;
      double v;
      if (gi < N && gj < N) {
        const double d2 = -2.0 * dot[q] + sqI[r] + sqJ[c];
        v = scale * matern32_from_d2(d2);
        if (KE) v *= KE[(size_t)tI[r] * T + tJ[c]];
        if (gi == gj) v += diag_add;
      } else {
        v = (gi == gj) ? 1.0 : 0.0;
      }
      dst[c] = v;
    }
  }
}

// ------------------------------------------------------------------ gradient contraction (K2g)
// One warp per row i (longest rows first).  With g_ij = Kinv_ij - alpha_i alpha_j (= 2 G_ij) and
// weight 1 for j < i, 1/2 for j == i, the symmetric sums over all (i,j) are recovered.
//   rowH [n_pad][T]   : R[i][t'] = sum_{j in t', j<=i} w g_ij scale ktheta_ij
//   rowbw[n_pad][d]   : sum_j w g_ij scale KE_ij 3 exp(-v) delta_k^2
//   rowsc[n_pad][3]   : {alpha_i, Kinv_ii, alpha_i^2}
template <int D>
__global__ void __launch_bounds__(256, 2) gram_grad_rows_kernel(
    const double* __restrict__ Kinv, const double* __restrict__ alpha, const double* __restrict__ theta_s,
    const double* __restrict__ sqnorm, int N, int n_pad, int d, const int32_t* __restrict__ task_id,
    const int32_t* __restrict__ task_off, const double* __restrict__ KE, int T,
    const double* __restrict__ log_scale, double* __restrict__ rowH, double* __restrict__ rowbw,
    double* __restrict__ rowsc) {
  // CTA = 8 consecutive rows (one per warp, longest rows first); the columns are walked in chunks of 128 whose
  // theta rows are staged TRANSPOSED in shared memory once per CTA (conflict-free reads, no strided global
  // loads); lane = columns lane, lane+32, lane+64, lane+96 of the chunk.  rowH must be zero on entry.
  __shared__ double sJt[D][128];
  __shared__ double sSq[128], sAl[128];
  __shared__ int sTj[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i_hi = N - 1 - (int)blockIdx.x * 8;
  const int i = i_hi - warp;
  const bool row_ok = i >= 0;
  const double scale = exp(*log_scale);
  const int ti = (row_ok && task_id) ? task_id[i] : 0;
  const double ai = row_ok ? alpha[i] : 0.0, sqi = row_ok ? sqnorm[i] : 0.0;
  double xi[D], gbw[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    xi[k] = (row_ok && k < d) ? theta_s[(size_t)i * d + k] : 0.0;
    gbw[k] = 0.0;
  }
  // rows k >= d of the staged tile are never filled below but enter the unrolled dot product with xi[k] = 0:
  // they must hold finite numbers (0 * NaN left over in shared memory by an earlier kernel is NaN)
  for (int idx = threadIdx.x; idx < (D - d) * 128; idx += 256) sJt[d + idx / 128][idx % 128] = 0.0;
  const double* krow = Kinv + (size_t)(row_ok ? i : 0) * n_pad;
  int tcur = -1;      // lane 0: running (task, partial sum) of rowH[i][.]
  double hcur = 0.0;

  // Kinv row values of the next chunk are fetched one chunk ahead (they stream from HBM exactly once)
  double kq[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) kq[u] = (row_ok && lane + 32 * u <= i) ? krow[lane + 32 * u] : 0.0;
  for (int jb = 0; jb <= i_hi; jb += 128) {
    double kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      kv[u] = kq[u];
      const int jn = jb + 128 + lane + 32 * u;
      kq[u] = (row_ok && jn <= i) ? krow[jn] : 0.0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 128 * d; idx += 256) {
      const int r = idx / d, k = idx - r * d;
      sJt[k][r] = (jb + r < N) ? theta_s[(size_t)(jb + r) * d + k] : 0.0;
    }
    if (threadIdx.x < 128) {
      const int j = jb + threadIdx.x;
      sSq[threadIdx.x] = (j < N) ? sqnorm[j] : 0.0;
      sAl[threadIdx.x] = (j < N) ? alpha[j] : 0.0;
      sTj[threadIdx.x] = (j < N && task_id) ? task_id[j] : 0;
    }
    __syncthreads();
    if (!row_ok || jb > i) continue;
    double hc[4], f[4];
    int tj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = lane + 32 * u, j = jb + c;
      const bool live = j <= i;
      double g = live ? kv[u] - ai * sAl[c] : 0.0;
      if (j == i) g *= 0.5;
      double dot = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) dot += xi[k] * sJt[k][c];
      const double d2 = -2.0 * dot + sqi + sSq[c];
      const double r = sqrt(fmax(d2, kDistClamp));
      const double v = kSqrt3 * r;
      const double e = exp(-v);
      tj[u] = sTj[c];
      const double gs = g * scale;
      hc[u] = gs * ((1.0 + v) * e);
      const double ke = (KE && live) ? KE[(size_t)ti * T + tj[u]] : 1.0;
      // tf.maximum routes the gradient to the constant when the distance is clamped
      f[u] = (d2 > kDistClamp) ? gs * ke * 3.0 * e : 0.0;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double t = xi[k] - sJt[k][lane + 32 * u];
        gbw[k] += f[u] * t * t;
      }
    }
    // block sums by the columns' task (tasks are contiguous column ranges): one masked warp sum per task
    const int t_lo = sTj[0], t_hi = sTj[min(127, i - jb)];
    for (int tp = t_lo; tp <= t_hi; ++tp) {
      double sacc = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) sacc += (tj[u] == tp) ? hc[u] : 0.0;
      sacc = warp_sum(sacc);
      if (lane == 0) {
        if (tp == tcur) {
          hcur += sacc;
        } else {
          if (tcur >= 0) rowH[(size_t)i * T + tcur] = hcur;
          tcur = tp;
          hcur = sacc;
        }
      }
    }
  }
  if (!row_ok) return;
  if (lane == 0 && tcur >= 0) rowH[(size_t)i * T + tcur] = hcur;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double s = warp_sum(gbw[k]);
    if (lane == 0 && k < d) rowbw[(size_t)i * d + k] = s;
  }
  if (lane == 0) {
    rowsc[(size_t)i * 3 + 0] = ai;
    rowsc[(size_t)i * 3 + 1] = krow[i];
    rowsc[(size_t)i * 3 + 2] = ai * ai;
  }
}

// fixed-order column sums of a [rows x cols] row-major matrix over row range, one CTA per column
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q];
  }
  __syncthreads();
  return s;  // valid in thread 0
}

// grid: T*T blocks (t,t') -> Hsym[t][t'] = sum_{i in t} rowH[i][t'] (t >= t'), then symmetrise later
__global__ void __launch_bounds__(256) gram_grad_reduceH_kernel(const double* __restrict__ rowH, int T,
                                                                const int32_t* __restrict__ task_off,
                                                                int N, double* __restrict__ Hsym) {
  __shared__ double red[8];
  const int t = blockIdx.x / T, u = blockIdx.x % T;
  double acc = 0.0;
  if (u <= t) {
    const int i0 = task_off ? task_off[t] : 0, i1 = task_off ? task_off[t + 1] : N;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) acc += rowH[(size_t)i * T + u];
  }
  const double s = block_sum_256(acc, red);
  if (threadIdx.x == 0) Hsym[(size_t)t * T + u] = s;
}

// single block: finishes everything. H (full, symmetric) from Hsym; g_log_bw; scalars.
__global__ void __launch_bounds__(256) gram_grad_finish_kernel(
    const double* __restrict__ Hsym, const double* __restrict__ rowbw, const double* __restrict__ rowsc,
    const double* __restrict__ KE, int T, int N, int d, const double* __restrict__ log_sigma,
    double* __restrict__ H, double* __restrict__ g_log_bw, double* __restrict__ g_scalars) {
  __shared__ double red[8];
  // H[t][t] = Hsym[t][t]; H[t][u] = H[u][t] = 0.5 Hsym[max][min]
  double gs = 0.0;
  for (int idx = threadIdx.x; idx < T * T; idx += 256) {
    const int t = idx / T, u = idx % T;
    const int a = max(t, u), b = min(t, u);
    const double h = (t == u) ? Hsym[(size_t)t * T + t] : 0.5 * Hsym[(size_t)a * T + b];
    H[idx] = h;
    gs += h * (KE ? KE[idx] : 1.0);
  }
  gs = block_sum_256(gs, red);
  if (threadIdx.x == 0) g_scalars[0] = gs;  // d/dlog_scale = sum_ij G_ij K_ij
  for (int k = 0; k < d; ++k) {
    double a = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) a += rowbw[(size_t)i * d + k];
    a = block_sum_256(a, red);
    if (threadIdx.x == 0) g_log_bw[k] = a;
  }
  double sa = 0.0, tr = 0.0, aa = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    sa += rowsc[(size_t)i * 3 + 0];
    tr += rowsc[(size_t)i * 3 + 1];
    aa += rowsc[(size_t)i * 3 + 2];
  }
  sa = block_sum_256(sa, red);
  tr = block_sum_256(tr, red);
  aa = block_sum_256(aa, red);
  if (threadIdx.x == 0) {
    g_scalars[1] = -sa;                                       // d/dmean
    g_scalars[2] = exp(2.0 * (*log_sigma)) * (tr - aa);       // d/dlog_sigma = 2 sigma^2 tr(G)
  }
}

// ------------------------------------------------------------------ task Gram backward
// One CTA per task row t.  ev[u] = exp(-v_tu) (0 where the distance was clamped: tf.maximum routes the
// gradient to the constant).  Warp w handles embedding columns c = w, w+8, ...; lanes split u and are
// combined by a fixed-order butterfly, so the result is deterministic.
//   dE[t][c]      = sum_u (H[t][u] + H[u][t]) (-1.5 ev_tu) 2 (E_tc - E_uc) / s_c^2
//   gpart[t][c]   = sum_u  H[t][u] 3 ev_tu (E_tc - E_uc)^2 / s_c^2        (summed over t by the finish kernel)
__global__ void __launch_bounds__(256) task_gram_bwd_rows_kernel(const double* __restrict__ E, int lde, int T,
                                                                 int P, const double* __restrict__ log_bw_embed,
                                                                 const double* __restrict__ H,
                                                                 double* __restrict__ gpart,
                                                                 double* __restrict__ dE) {
  extern __shared__ __align__(16) double ev[];  // [T]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x;
  for (int u = warp; u < T; u += 8) {
    double xy = 0.0, xx = 0.0, yy = 0.0;
    for (int c = lane; c < P; c += 32) {
      const double s = exp(log_bw_embed[c]);
      const double a = E[(size_t)t * lde + c] / s, b = E[(size_t)u * lde + c] / s;
      xy += a * b;
      xx += a * a;
      yy += b * b;
    }
    xy = warp_sum(xy);
    xx = warp_sum(xx);
    yy = warp_sum(yy);
    if (lane == 0) {
      const double d2 = -2.0 * xy + xx + yy;
      ev[u] = (d2 > kDistClamp) ? exp(-kSqrt3 * sqrt(d2)) : 0.0;
    }
  }
  __syncthreads();
  for (int c = warp; c < P; c += 8) {
    const double s = exp(log_bw_embed[c]);
    const double inv_s2 = 1.0 / (s * s);
    const double et = E[(size_t)t * lde + c];
    double de = 0.0, gbw = 0.0;
    for (int u = lane; u < T; u += 32) {
      const double diff = et - E[(size_t)u * lde + c];
      const double h = H[(size_t)t * T + u], e = ev[u];
      de += (h + H[(size_t)u * T + t]) * (-1.5 * e) * 2.0 * diff * inv_s2;
      gbw += h * 3.0 * e * diff * diff * inv_s2;
    }
    de = warp_sum(de);
    gbw = warp_sum(gbw);
    if (lane == 0) {
      dE[(size_t)t * lde + c] = de;
      gpart[(size_t)t * P + c] = gbw;
    }
  }
}

__global__ void __launch_bounds__(256) task_gram_bwd_finish_kernel(const double* __restrict__ gpart, int T, int P,
                                                                   double* __restrict__ g_log_bw_embed) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= P) return;
  double s = 0.0;
  for (int t = 0; t < T; ++t) s += gpart[(size_t)t * P + c];
  g_log_bw_embed[c] = s;
}

// ------------------------------------------------------------------ Adam (TF 1.7 form)
__global__ void __launch_bounds__(1024) adam_kernel(double* __restrict__ params,
                                                    const double* __restrict__ grads,
                                                    double* __restrict__ m, double* __restrict__ v,
                                                    long long n, double lr_t_host,
                                                    const double* __restrict__ lr_t_dev, double b1, double b2,
                                                    double eps, double clip_norm, double* __restrict__ ctrl,
                                                    const double* __restrict__ nll,
                                                    const int32_t* __restrict__ info,
                                                    double* __restrict__ loss_hist, long long hist_cap) {
  __shared__ double red[32];
  __shared__ double gscale;
  // Training-loop control on the device (train/train.py:43-65), so that the host never reads the loss back
  // between epochs: once the early-stopping rule has fired (ctrl[STOP] = 1) or a factorisation has failed
  // (= 2, TF would have raised inside sess.run before applying the update) every later launch is a no-op, and
  // the parameters stay those of the epoch the reference breaks at.
  if (ctrl) {
    if (ctrl[DISTBO_CTRL_STOP] != 0.0) return;
    const int32_t inf = info ? *info : 0;
    if (inf != 0) {
      if (threadIdx.x == 0) {
        ctrl[DISTBO_CTRL_STOP] = 2.0;
        ctrl[DISTBO_CTRL_INFO] = (double)inf;
      }
      return;
    }
  }
  const double lr_t = lr_t_dev ? *lr_t_dev : lr_t_host;
  if (clip_norm > 0.0) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += grads[i] * grads[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      s = warp_sum(red[threadIdx.x]);
      if (threadIdx.x == 0) {
        const double gn = sqrt(s);
        gscale = clip_norm / fmax(gn, clip_norm);  // tf.clip_by_global_norm
      }
    }
    __syncthreads();
  } else if (threadIdx.x == 0) {
    gscale = 1.0;
  }
  __syncthreads();
  const double gsc = gscale;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const double g = grads[i] * gsc;
    const double mi = b1 * m[i] + (1.0 - b1) * g;
    const double vi = b2 * v[i] + (1.0 - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    params[i] -= lr_t * mi / (sqrt(vi) + eps);
  }
  if (ctrl && threadIdx.x == 0) {
    // this epoch's loss (evaluated BEFORE the update, as sess.run([optimize_step, loss]) returns it) and the
    // early-stopping rule of train/train.py:53-65: cur_min / countdown start at +inf
    const double loss = nll ? *nll : 0.0;
    const long long epoch = (long long)ctrl[DISTBO_CTRL_EPOCHS];
    if (loss_hist && hist_cap > 0) loss_hist[epoch % hist_cap] = loss;
    ctrl[DISTBO_CTRL_LAST_LOSS] = loss;
    ctrl[DISTBO_CTRL_EPOCHS] = (double)(epoch + 1);
    if ((double)epoch >= ctrl[DISTBO_CTRL_FIRST_EPOCH]) {
      double cur_min = ctrl[DISTBO_CTRL_CUR_MIN], countdown = ctrl[DISTBO_CTRL_COUNTDOWN];
      if ((cur_min - loss) > ctrl[DISTBO_CTRL_THRESHOLD]) {
        countdown = ctrl[DISTBO_CTRL_PATIENCE];
        cur_min = loss;
      } else {
        countdown -= 1.0;
      }
      ctrl[DISTBO_CTRL_CUR_MIN] = cur_min;
      ctrl[DISTBO_CTRL_COUNTDOWN] = countdown;
      if (countdown <= 0.0) ctrl[DISTBO_CTRL_STOP] = 1.0;
    }
  }
}


// ---- multi-task kernel (train/kernels.py:32-55; log_gaus_likelihood.py:68-75) ---------------------------
// Task covariance C = D^-1/2 (L L^T) D^-1/2, L = tril(fill_triangular(exp(log_tri))), D = diag(L L^T).
// T is the number of tasks (tens), so one CTA does all of it; the T x T temporaries live in the caller's
// workspace (ws: 3 T^2 doubles) and the phases are separated by __syncthreads().

// Index into the length T(T+1)/2 parameter vector of lower-triangle element (i, j >= ... j <= i), following
// tf.contrib.distributions.fill_triangular: the n*n row-major matrix is concat(x[n:], reverse(x)).
__device__ __forceinline__ int tri_param_index(int i, int j, int n) {
  const int m = n * (n + 1) / 2;
  const int k = i * n + j;
  return k < m - n ? n + k : 2 * m - n - 1 - k;
}

__device__ void task_tri_build(const double* __restrict__ log_tri, int T, double* L, double* Bm) {
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e % T;
    L[e] = j <= i ? exp(log_tri[tri_param_index(i, j, T)]) : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e % T;
    const int kmax = min(i, j);
    double acc = 0.0;
    for (int k = 0; k <= kmax; ++k) acc = fma(L[i * T + k], L[j * T + k], acc);
    Bm[e] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) task_tri_fwd_kernel(const double* __restrict__ log_tri, int T,
                                                          double* __restrict__ KE, double* __restrict__ ws) {
  double* L = ws;
  double* Bm = ws + (size_t)T * T;
  task_tri_build(log_tri, T, L, Bm);
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e % T;
    KE[e] = (Bm[e] / sqrt(Bm[i * T + i])) / sqrt(Bm[j * T + j]);   // rows first, then columns (kernels.py:39-40)
  }
}

// H = dNLL/dKE [T,T] (all ordered pairs) -> g_log_tri.
__global__ void __launch_bounds__(256) task_tri_bwd_kernel(const double* __restrict__ log_tri, int T,
                                                          const double* __restrict__ H,
                                                          double* __restrict__ g_log_tri, double* __restrict__ ws) {
  double* L = ws;
  double* Bm = ws + (size_t)T * T;
  double* dB = ws + 2 * (size_t)T * T;
  task_tri_build(log_tri, T, L, Bm);
  // direct term dB_ij = H_ij / (s_i s_j)
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e % T;
    dB[e] = H[e] / (sqrt(Bm[i * T + i]) * sqrt(Bm[j * T + j]));
  }
  __syncthreads();
  // through the normalisers s_i = sqrt(B_ii): dB_ii += ds_i / (2 s_i),
  // ds_i = -(1 / s_i^2) sum_j (H_ij + H_ji) B_ij / s_j
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const double si = sqrt(Bm[i * T + i]);
    double acc = 0.0;
    for (int j = 0; j < T; ++j) acc += (H[i * T + j] + H[j * T + i]) * Bm[i * T + j] / sqrt(Bm[j * T + j]);
    dB[i * T + i] += -acc / (si * si) / (2.0 * si);
  }
  __syncthreads();
  // B = L L^T  =>  dL = (dB + dB^T) L ;  L_ik = exp(x)  =>  dx = dL_ik L_ik
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, k = e % T;
    if (k > i) continue;
    double acc = 0.0;
    for (int j = k; j < T; ++j) acc = fma(dB[i * T + j] + dB[j * T + i], L[j * T + k], acc);
    g_log_tri[tri_param_index(i, k, T)] = acc * L[i * T + k];
  }
}

}  // namespace db

using namespace db;

extern "C" int distbo_scale_theta_f64(const double* theta, int N, int d, const double* log_bw,
                                      double* theta_s, double* sqnorm, int n_rows_out, void* stream) {
  if (!theta || !log_bw || !theta_s || !sqnorm || N <= 0 || d <= 0 || n_rows_out < N) return DISTBO_EARG;
  scale_theta_kernel<<<(n_rows_out + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta, N, d, log_bw, theta_s,
                                                                                 sqnorm, n_rows_out);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_task_gram_f64(const double* E, int lde, int T, int P, const double* log_bw_embed,
                                    double* KE, void* stream) {
  if (!E || !log_bw_embed || !KE || T <= 0 || P <= 0 || lde < P) return DISTBO_EARG;
  task_gram_kernel<<<(T * T + 7) / 8, 256, 0, (cudaStream_t)stream>>>(E, lde, T, P, log_bw_embed, KE);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_gram_f64(const double* theta_s, const double* sqnorm, int N, int n_pad, int d,
                               const int32_t* task_id, const double* KE, int T, const double* log_scale,
                               const double* log_sigma, double jitter, double* K, void* stream) {
  if (!theta_s || !sqnorm || !log_scale || !log_sigma || !K || N <= 0 || n_pad < N || n_pad % 128 || d <= 0 ||
      d > 64)
    return DISTBO_EARG;
  if (KE && (!task_id || T <= 0)) return DISTBO_EARG;
  const int nb = n_pad / GT;
  const int smem = (2 * GT * d + 2 * GT) * (int)sizeof(double) + 2 * GT * (int)sizeof(int);
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    once.mark();
  }
  gram_kernel<<<nb * (nb + 1) / 2, 256, smem, (cudaStream_t)stream>>>(theta_s, sqnorm, N, n_pad, d, task_id, KE, T,
                                                                     log_scale, log_sigma, jitter, K);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_gram_grad_f64(const double* Kinv, const double* alpha, const double* theta_s,
                                    const double* sqnorm, int N, int n_pad, int d, const int32_t* task_id,
                                    const int32_t* task_off, const double* KE, int T, const double* log_scale,
                                    const double* log_sigma, double* g_log_bw, double* g_scalars, double* H,
                                    double* workspace, void* stream) {
  if (!Kinv || !alpha || !theta_s || !sqnorm || !log_scale || !log_sigma || !g_log_bw || !g_scalars || !H ||
      !workspace || N <= 0 || n_pad < N || d <= 0 || d > DMAX || T <= 0)
    return DISTBO_EARG;
  if (T > 1 && (!task_id || !task_off)) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* rowH = workspace;                          // [n_pad][T]
  double* rowbw = rowH + (size_t)n_pad * T;          // [n_pad][d]
  double* rowsc = rowbw + (size_t)n_pad * d;         // [n_pad][3]
  double* Hsym = rowsc + (size_t)n_pad * 3;          // [T*T]; workspace = n_pad * (T + d + 3) + T * T doubles
  const int grid = (N + 7) / 8;
  DB_CUDA(cudaMemsetAsync(rowH, 0, (size_t)n_pad * T * sizeof(double), st));   // rows only write the tasks they meet
  if (d <= 4)
    gram_grad_rows_kernel<4><<<grid, 256, 0, st>>>(Kinv, alpha, theta_s, sqnorm, N, n_pad, d, task_id, task_off, KE,
                                                   T, log_scale, rowH, rowbw, rowsc);
  else if (d <= 8)
    gram_grad_rows_kernel<8><<<grid, 256, 0, st>>>(Kinv, alpha, theta_s, sqnorm, N, n_pad, d, task_id, task_off, KE,
                                                   T, log_scale, rowH, rowbw, rowsc);
  else
    gram_grad_rows_kernel<DMAX><<<grid, 256, 0, st>>>(Kinv, alpha, theta_s, sqnorm, N, n_pad, d, task_id, task_off,
                                                      KE, T, log_scale, rowH, rowbw, rowsc);
  DB_CHECK_LAUNCH();
  gram_grad_reduceH_kernel<<<T * T, 256, 0, st>>>(rowH, T, task_off, N, Hsym);
  DB_CHECK_LAUNCH();
  gram_grad_finish_kernel<<<1, 256, 0, st>>>(Hsym, rowbw, rowsc, KE, T, N, d, log_sigma, H, g_log_bw, g_scalars);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_task_gram_bwd_f64(const double* E, int lde, int T, int P, const double* log_bw_embed,
                                        const double* H, double* g_log_bw_embed, double* dE, double* workspace,
                                        void* stream) {
  if (!E || !log_bw_embed || !H || !g_log_bw_embed || !dE || !workspace || T <= 0 || P <= 0 || lde < P)
    return DISTBO_EARG;
  if ((size_t)T * sizeof(double) > 48 * 1024) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  task_gram_bwd_rows_kernel<<<T, 256, (size_t)T * sizeof(double), st>>>(E, lde, T, P, log_bw_embed, H, workspace, dE);
  DB_CHECK_LAUNCH();
  task_gram_bwd_finish_kernel<<<(P + 255) / 256, 256, 0, st>>>(workspace, T, P, g_log_bw_embed);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_task_tri_f64(const double* log_tri, int T, double* KE, double* workspace, void* stream) {
  if (!log_tri || !KE || !workspace || T <= 0) return DISTBO_EARG;
  task_tri_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(log_tri, T, KE, workspace);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_task_tri_bwd_f64(const double* log_tri, int T, const double* H, double* g_log_tri,
                                       double* workspace, void* stream) {
  if (!log_tri || !H || !g_log_tri || !workspace || T <= 0) return DISTBO_EARG;
  task_tri_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(log_tri, T, H, g_log_tri, workspace);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

extern "C" int distbo_adam_f64(double* params, const double* grads, double* m, double* v, int64_t n, double lr,
                               double beta1, double beta2, double eps, int64_t step, double clip_norm,
                               const double* lr_t_dev, double* ctrl, const double* nll, const int32_t* info,
                               double* loss_hist, int64_t hist_cap, void* stream) {
  if (!params || !grads || !m || !v || n <= 0 || (step <= 0 && !lr_t_dev)) return DISTBO_EARG;
  if (ctrl && (!nll || !info || hist_cap < 0)) return DISTBO_EARG;
  const double lr_t =
      lr_t_dev ? 0.0 : lr * sqrt(1.0 - pow(beta2, (double)step)) / (1.0 - pow(beta1, (double)step));
  adam_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grads, m, v, (long long)n, lr_t, lr_t_dev, beta1, beta2,
                                                    eps, clip_norm, ctrl, nll, info, loss_hist, (long long)hist_cap);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
