// K1: kernel-mean-embedding network.  Fused tanh-MLP feature maps phi_x / phi_y, per-sample outer
// product and per-dataset (segment) mean, forward and backward.
// Reference: train/distbo_net.py:7-48 (nn_net, joint / classification branches),
//            train/base.py:9-43 (mean pooling by a sparse [T,S] matrix with 1/size entries).
//
// Forward: one thread-block CLUSTER of EMB_CL CTAs per dataset.  Each CTA streams its share of the
// dataset's rows through shared memory in k-chunks (coalesced loads), one thread per sample keeps
// the hidden activations in registers, a second phase accumulates the outer products, and the
// CTA partials are combined in rank order through distributed shared memory (deterministic, no
// global workspace, no atomics).
#include <cooperative_groups.h>

#include "../../include/distbo_b200.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace db {

constexpr int EMB_THREADS = 128;  // samples per batch (one per thread)
constexpr int EMB_CL = 4;         // CTAs per dataset (cluster size)
constexpr int EMB_KC = 32;        // input columns staged per chunk
constexpr int EMB_HMAX = 32;      // max hidden width
constexpr int EMB_PMAX = 16;      // max feature-map output width (p, q)
constexpr int EMB_ACC = 4;        // output columns per thread: P' <= EMB_ACC * EMB_THREADS
constexpr int EMB_BWD_GRID = 296;

template <typename T>
__device__ __forceinline__ T tanh_t(T x);
template <>
__device__ __forceinline__ float tanh_t<float>(float x) { return tanhf(x); }
template <>
__device__ __forceinline__ double tanh_t<double>(double x) { return tanh(x); }

struct NetDims {
  int din, h, p;
};

// shared-memory layout helpers (element counts)
__host__ __device__ inline int net_w_elems(int din, int h, int p) { return din * h + h + h * p; }

// Computes out[0..p) = tanh(x W1 + b1) W2 for this thread's sample; x streamed from global in
// k-chunks through sX [EMB_THREADS][EMB_KC+1].  hid[] (tanh activations) is returned for backward.
template <typename T>
__device__ __forceinline__ void mlp_forward(const T* __restrict__ Xg, long long row0, int nrows, int din,
                                            const T* sW1, const T* sb1, const T* sW2, int h, int p, T* sX,
                                            T (&hid)[EMB_HMAX], T (&out)[EMB_PMAX]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < EMB_HMAX; ++j) hid[j] = (j < h) ? sb1[j] : (T)0;
  for (int k0 = 0; k0 < din; k0 += EMB_KC) {
    const int kc = min(EMB_KC, din - k0);
    __syncthreads();
    for (int idx = tid; idx < EMB_THREADS * kc; idx += EMB_THREADS) {
      const int r = idx / kc, k = idx % kc;
      sX[r * (EMB_KC + 1) + k] = (r < nrows) ? Xg[(row0 + r) * din + k0 + k] : (T)0;
    }
    __syncthreads();
    for (int k = 0; k < kc; ++k) {
      const T x = sX[tid * (EMB_KC + 1) + k];
      const T* w = sW1 + (size_t)(k0 + k) * h;
#pragma unroll
      for (int j = 0; j < EMB_HMAX; ++j)
        if (j < h) hid[j] += x * w[j];
    }
  }
#pragma unroll
  for (int j = 0; j < EMB_HMAX; ++j) hid[j] = (j < h) ? tanh_t<T>(hid[j]) : (T)0;
#pragma unroll
  for (int i = 0; i < EMB_PMAX; ++i) out[i] = (T)0;
#pragma unroll
  for (int j = 0; j < EMB_HMAX; ++j)
    if (j < h) {
      const T hj = hid[j];
#pragma unroll
      for (int i = 0; i < EMB_PMAX; ++i)
        if (i < p) out[i] += hj * sW2[j * p + i];
    }
}

template <typename T>
struct EmbedArgs {
  const T* X;
  const T* Y;
  const int32_t* seg_off;
  int T_, dx, dy;
  const T* W1x; const T* b1x; const T* W2x; int hx, p;
  const T* W1y; const T* b1y; const T* W2y; int hy, q;
  int y_mode;
  T* E;
  int lde;
};

template <typename T>
__global__ void __launch_bounds__(EMB_THREADS) embed_fwd_kernel(EmbedArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int seg = blockIdx.x / EMB_CL;
  const int tid = threadIdx.x;

  const int gq = (a.y_mode == 0) ? 1 : (a.y_mode == 1 ? a.q : a.dy);  // width of the y-side factor
  const int pout = (a.y_mode == 0) ? a.p : a.p * gq + (a.y_mode == 2 ? a.dy : 0);

  T* sWx = reinterpret_cast<T*>(smem_raw);
  T* sWy = sWx + net_w_elems(a.dx, a.hx, a.p);
  T* sX = sWy + ((a.y_mode == 1) ? net_w_elems(a.dy, a.hy, a.q) : 0);
  T* sF = sX + EMB_THREADS * (EMB_KC + 1);                   // [EMB_THREADS][p + gq]
  T* sOut = sF + EMB_THREADS * (a.p + gq);                   // [pout]

  for (int i = tid; i < a.dx * a.hx; i += EMB_THREADS) sWx[i] = a.W1x[i];
  for (int i = tid; i < a.hx; i += EMB_THREADS) sWx[a.dx * a.hx + i] = a.b1x[i];
  for (int i = tid; i < a.hx * a.p; i += EMB_THREADS) sWx[a.dx * a.hx + a.hx + i] = a.W2x[i];
  if (a.y_mode == 1) {
    for (int i = tid; i < a.dy * a.hy; i += EMB_THREADS) sWy[i] = a.W1y[i];
    for (int i = tid; i < a.hy; i += EMB_THREADS) sWy[a.dy * a.hy + i] = a.b1y[i];
    for (int i = tid; i < a.hy * a.q; i += EMB_THREADS) sWy[a.dy * a.hy + a.hy + i] = a.W2y[i];
  }
  __syncthreads();

  const long long s0 = a.seg_off[seg], s1 = a.seg_off[seg + 1];
  const long long len = s1 - s0;
  const long long per = (len + EMB_CL - 1) / EMB_CL;
  const long long r_begin = s0 + min(len, per * rank), r_end = s0 + min(len, per * (rank + 1));

  T accv[EMB_ACC];
#pragma unroll
  for (int c = 0; c < EMB_ACC; ++c) accv[c] = (T)0;

  for (long long row0 = r_begin; row0 < r_end; row0 += EMB_THREADS) {
    const int nrows = (int)min((long long)EMB_THREADS, r_end - row0);
    T hid[EMB_HMAX], f[EMB_PMAX];
    mlp_forward<T>(a.X, row0, nrows, a.dx, sWx, sWx + a.dx * a.hx, sWx + a.dx * a.hx + a.hx, a.hx, a.p, sX, hid, f);
    T g[EMB_PMAX];
    if (a.y_mode == 1) {
      mlp_forward<T>(a.Y, row0, nrows, a.dy, sWy, sWy + a.dy * a.hy, sWy + a.dy * a.hy + a.hy, a.hy, a.q, sX, hid, g);
    } else if (a.y_mode == 2) {
#pragma unroll
      for (int j = 0; j < EMB_PMAX; ++j) g[j] = (j < a.dy && tid < nrows) ? a.Y[(row0 + tid) * a.dy + j] : (T)0;
    } else {
      g[0] = (T)1;
    }
    __syncthreads();
    const bool live = tid < nrows;
#pragma unroll
    for (int i = 0; i < EMB_PMAX; ++i)
      if (i < a.p) sF[tid * (a.p + gq) + i] = live ? f[i] : (T)0;
#pragma unroll
    for (int j = 0; j < EMB_PMAX; ++j)
      if (j < gq) sF[tid * (a.p + gq) + a.p + j] = live ? g[j] : (T)0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < EMB_ACC; ++c) {
      const int o = tid + c * EMB_THREADS;
      if (o < pout) {
        T s = (T)0;
        if (o < a.p * gq) {
          const int i = o / gq, j = o % gq;   // row-major flatten i*q + j (distbo_net.py:40-42)
          for (int r = 0; r < EMB_THREADS; ++r) s += sF[r * (a.p + gq) + i] * sF[r * (a.p + gq) + a.p + j];
        } else {
          const int j = o - a.p * gq;         // class ratios: mean of y (distbo_net.py:45-48)
          for (int r = 0; r < EMB_THREADS; ++r) s += sF[r * (a.p + gq) + a.p + j];
        }
        accv[c] += s;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < EMB_ACC; ++c) {
    const int o = tid + c * EMB_THREADS;
    if (o < pout) sOut[o] = accv[c];
  }
  cluster.sync();
  if (rank == 0) {
    const T inv = (T)1 / (T)len;
    for (int o = tid; o < pout; o += EMB_THREADS) {
      T s = (T)0;
      for (int r = 0; r < EMB_CL; ++r) s += *cluster.map_shared_rank(sOut + o, r);
      a.E[(size_t)seg * a.lde + o] = s * inv;
    }
  }
  cluster.sync();  // keep remote shared memory alive until rank 0 has read it
}

template <typename T>
static int embed_fwd_launch(const EmbedArgs<T>& a, cudaStream_t st) {
  if (!a.X || !a.seg_off || !a.W1x || !a.b1x || !a.W2x || !a.E || a.T_ <= 0 || a.dx <= 0) return DISTBO_EARG;
  if (a.hx <= 0 || a.hx > EMB_HMAX || a.p <= 0 || a.p > EMB_PMAX) return DISTBO_EARG;
  if (a.y_mode < 0 || a.y_mode > 2) return DISTBO_EARG;
  if (a.y_mode == 1 && (!a.Y || !a.W1y || !a.b1y || !a.W2y || a.hy <= 0 || a.hy > EMB_HMAX || a.q <= 0 ||
                        a.q > EMB_PMAX || a.dy <= 0))
    return DISTBO_EARG;
  if (a.y_mode == 2 && (!a.Y || a.dy <= 0 || a.dy > EMB_PMAX)) return DISTBO_EARG;
  const int gq = (a.y_mode == 0) ? 1 : (a.y_mode == 1 ? a.q : a.dy);
  const int pout = (a.y_mode == 0) ? a.p : a.p * gq + (a.y_mode == 2 ? a.dy : 0);
  if (pout > EMB_ACC * EMB_THREADS || a.lde < pout) return DISTBO_EARG;
  size_t elems = net_w_elems(a.dx, a.hx, a.p) + ((a.y_mode == 1) ? net_w_elems(a.dy, a.hy, a.q) : 0) +
                 EMB_THREADS * (EMB_KC + 1) + EMB_THREADS * (a.p + gq) + pout;
  size_t smem = elems * sizeof(T) + 16;
  if (smem > 220 * 1024) return DISTBO_EARG;
  DB_CUDA(cudaFuncSetAttribute(embed_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.T_ * EMB_CL, 1, 1);
  cfg.blockDim = dim3(EMB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = EMB_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DB_CUDA(cudaLaunchKernelEx(&cfg, embed_fwd_kernel<T>, a));
  ++g_db_launches;
  return DISTBO_OK;
}

// ------------------------------------------------------------------ backward (fp64)
struct EmbedBwdArgs {
  EmbedArgs<double> f;
  const double* dE;
  double* partial;  // [EMB_BWD_GRID][nparams]
  int nparams;
};

// accumulate G[k][j] += sum_r A[r][k] * B[r][j] over the batch, G in shared (owner-exclusive)
__global__ void __launch_bounds__(EMB_THREADS) embed_bwd_kernel(EmbedBwdArgs b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EmbedArgs<double>& a = b.f;
  const int tid = threadIdx.x;
  const int gq = (a.y_mode == 0) ? 1 : (a.y_mode == 1 ? a.q : a.dy);
  const int nx = net_w_elems(a.dx, a.hx, a.p);
  const int ny = (a.y_mode == 1) ? net_w_elems(a.dy, a.hy, a.q) : 0;

  double* sWx = reinterpret_cast<double*>(smem_raw);
  double* sWy = sWx + nx;
  double* sG = sWy + ny;                                  // [nx + ny] gradient accumulators
  double* sX = sG + nx + ny;                              // [EMB_THREADS][EMB_KC+1]
  double* sH = sX + EMB_THREADS * (EMB_KC + 1);           // [EMB_THREADS][hmax] tanh activations
  const int hmax = max(a.hx, a.hy);
  const int pmax = max(a.p, gq);
  double* sDZ = sH + EMB_THREADS * hmax;                  // [EMB_THREADS][hmax] dL/dpreact
  double* sDF = sDZ + EMB_THREADS * hmax;                 // [EMB_THREADS][pmax] dL/dout

  for (int i = tid; i < a.dx * a.hx; i += EMB_THREADS) sWx[i] = a.W1x[i];
  for (int i = tid; i < a.hx; i += EMB_THREADS) sWx[a.dx * a.hx + i] = a.b1x[i];
  for (int i = tid; i < a.hx * a.p; i += EMB_THREADS) sWx[a.dx * a.hx + a.hx + i] = a.W2x[i];
  if (a.y_mode == 1) {
    for (int i = tid; i < a.dy * a.hy; i += EMB_THREADS) sWy[i] = a.W1y[i];
    for (int i = tid; i < a.hy; i += EMB_THREADS) sWy[a.dy * a.hy + i] = a.b1y[i];
    for (int i = tid; i < a.hy * a.q; i += EMB_THREADS) sWy[a.dy * a.hy + a.hy + i] = a.W2y[i];
  }
  for (int i = tid; i < nx + ny; i += EMB_THREADS) sG[i] = 0.0;
  __syncthreads();

  const long long S = a.seg_off[a.T_];
  const long long nbatch = (S + EMB_THREADS - 1) / EMB_THREADS;
  for (long long bt = blockIdx.x; bt < nbatch; bt += gridDim.x) {
    const long long row0 = bt * EMB_THREADS;
    const int nrows = (int)min((long long)EMB_THREADS, S - row0);
    const bool live = tid < nrows;
    // segment of this thread's sample (binary search in seg_off)
    int seg = 0;
    double inv_len = 0.0;
    if (live) {
      const long long r = row0 + tid;
      int lo = 0, hi = a.T_;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.seg_off[mid] <= r) lo = mid; else hi = mid;
      }
      seg = lo;
      inv_len = 1.0 / (double)(a.seg_off[seg + 1] - a.seg_off[seg]);
    }
    double hidx[EMB_HMAX], f[EMB_PMAX], hidy[EMB_HMAX], g[EMB_PMAX];
    mlp_forward<double>(a.X, row0, nrows, a.dx, sWx, sWx + a.dx * a.hx, sWx + a.dx * a.hx + a.hx, a.hx, a.p, sX,
                        hidx, f);
    if (a.y_mode == 1) {
      mlp_forward<double>(a.Y, row0, nrows, a.dy, sWy, sWy + a.dy * a.hy, sWy + a.dy * a.hy + a.hy, a.hy, a.q, sX,
                          hidy, g);
    } else if (a.y_mode == 2) {
#pragma unroll
      for (int j = 0; j < EMB_PMAX; ++j) g[j] = (j < a.dy && live) ? a.Y[(row0 + tid) * a.dy + j] : 0.0;
    } else {
      g[0] = 1.0;
    }
    // upstream gradients of the pooled outer product
    double df[EMB_PMAX], dg[EMB_PMAX];
#pragma unroll
    for (int i = 0; i < EMB_PMAX; ++i) df[i] = dg[i] = 0.0;
    if (live) {
      const double* de = b.dE + (size_t)seg * a.lde;
#pragma unroll
      for (int i = 0; i < EMB_PMAX; ++i)
        if (i < a.p) {
#pragma unroll
          for (int j = 0; j < EMB_PMAX; ++j)
            if (j < gq) {
              const double w = de[i * gq + j] * inv_len;
              df[i] += w * g[j];
              dg[j] += w * f[i];
            }
        }
    }
    // ---- x-net: second layer gradients and dz
    for (int net = 0; net < ((a.y_mode == 1) ? 2 : 1); ++net) {
      const int din = net ? a.dy : a.dx, h = net ? a.hy : a.hx, po = net ? a.q : a.p;
      const double* sW = net ? sWy : sWx;
      const double* sW2 = sW + din * h + h;
      double* G = sG + (net ? nx : 0);
      const double* Xg = net ? a.Y : a.X;
      const double* hid = net ? hidy : hidx;
      const double* dout = net ? dg : df;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < EMB_HMAX; ++j)
        if (j < h) {
          double dh = 0.0;
#pragma unroll
          for (int i = 0; i < EMB_PMAX; ++i)
            if (i < po) dh += sW2[j * po + i] * dout[i];
          const double hj = hid[j];
          sH[tid * hmax + j] = live ? hj : 0.0;
          sDZ[tid * hmax + j] = live ? dh * (1.0 - hj * hj) : 0.0;
        }
#pragma unroll
      for (int i = 0; i < EMB_PMAX; ++i)
        if (i < po) sDF[tid * pmax + i] = live ? dout[i] : 0.0;
      __syncthreads();
      // dW2[j][i] += sum_r H[r][j] DF[r][i] ; db1[j] += sum_r DZ[r][j]
      for (int e = tid; e < h * po; e += EMB_THREADS) {
        const int j = e / po, i = e % po;
        double s = 0.0;
        for (int r = 0; r < EMB_THREADS; ++r) s += sH[r * hmax + j] * sDF[r * pmax + i];
        G[din * h + h + e] += s;
      }
      for (int j = tid; j < h; j += EMB_THREADS) {
        double s = 0.0;
        for (int r = 0; r < EMB_THREADS; ++r) s += sDZ[r * hmax + j];
        G[din * h + j] += s;
      }
      // dW1[k][j] += sum_r X[r][k] DZ[r][j], X re-staged in k-chunks
      for (int k0 = 0; k0 < din; k0 += EMB_KC) {
        const int kc = min(EMB_KC, din - k0);
        __syncthreads();
        for (int idx = tid; idx < EMB_THREADS * kc; idx += EMB_THREADS) {
          const int r = idx / kc, k = idx % kc;
          sX[r * (EMB_KC + 1) + k] = (r < nrows) ? Xg[(row0 + r) * din + k0 + k] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < kc * h; e += EMB_THREADS) {
          const int k = e / h, j = e % h;
          double s = 0.0;
          for (int r = 0; r < EMB_THREADS; ++r) s += sX[r * (EMB_KC + 1) + k] * sDZ[r * hmax + j];
          G[(size_t)(k0 + k) * h + j] += s;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < nx + ny; i += EMB_THREADS) b.partial[(size_t)blockIdx.x * b.nparams + i] = sG[i];
}

__global__ void embed_bwd_reduce_kernel(const double* __restrict__ partial, int nparts, int nparams, int nx,
                                        int dxh, int hx, double* gW1x, double* gb1x, double* gW2x, int dyh,
                                        int hy, double* gW1y, double* gb1y, double* gW2y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparams) return;
  double s = 0.0;
  for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * nparams + i];
  if (i < nx) {
    if (i < dxh) gW1x[i] = s;
    else if (i < dxh + hx) gb1x[i - dxh] = s;
    else gW2x[i - dxh - hx] = s;
  } else {
    const int k = i - nx;
    if (k < dyh) gW1y[k] = s;
    else if (k < dyh + hy) gb1y[k - dyh] = s;
    else gW2y[k - dyh - hy] = s;
  }
}

}  // namespace db

using namespace db;

extern "C" int distbo_embed_fwd_f64(const double* X, const double* Y, const int32_t* seg_off, int T, int dx,
                                    int dy, const double* W1x, const double* b1x, const double* W2x, int hx,
                                    int p, const double* W1y, const double* b1y, const double* W2y, int hy,
                                    int q, int y_mode, double* E, int lde, void* stream) {
  EmbedArgs<double> a{X, Y, seg_off, T, dx, dy, W1x, b1x, W2x, hx, p, W1y, b1y, W2y, hy, q, y_mode, E, lde};
  return embed_fwd_launch<double>(a, (cudaStream_t)stream);
}

extern "C" int distbo_embed_fwd_f32(const float* X, const float* Y, const int32_t* seg_off, int T, int dx, int dy,
                                    const float* W1x, const float* b1x, const float* W2x, int hx, int p,
                                    const float* W1y, const float* b1y, const float* W2y, int hy, int q,
                                    int y_mode, float* E, int lde, void* stream) {
  EmbedArgs<float> a{X, Y, seg_off, T, dx, dy, W1x, b1x, W2x, hx, p, W1y, b1y, W2y, hy, q, y_mode, E, lde};
  return embed_fwd_launch<float>(a, (cudaStream_t)stream);
}

extern "C" int64_t distbo_embed_bwd_workspace(int64_t S, int dx, int dy, int hx, int p, int hy, int q) {
  (void)S;
  return (int64_t)EMB_BWD_GRID * (net_w_elems(dx, hx, p) + ((hy > 0 && q > 0) ? net_w_elems(dy, hy, q) : 0));
}

extern "C" int distbo_embed_bwd_f64(const double* X, const double* Y, const int32_t* seg_off, int T, int dx,
                                    int dy, const double* W1x, const double* b1x, const double* W2x, int hx,
                                    int p, const double* W1y, const double* b1y, const double* W2y, int hy,
                                    int q, int y_mode, const double* dE, int lde, double* gW1x, double* gb1x,
                                    double* gW2x, double* gW1y, double* gb1y, double* gW2y, double* workspace,
                                    int64_t workspace_elems, void* stream) {
  if (!X || !seg_off || !W1x || !b1x || !W2x || !dE || !gW1x || !gb1x || !gW2x || !workspace || T <= 0)
    return DISTBO_EARG;
  if (hx <= 0 || hx > EMB_HMAX || p <= 0 || p > EMB_PMAX || y_mode < 0 || y_mode > 2) return DISTBO_EARG;
  if (y_mode == 1 && (!Y || !W1y || !b1y || !W2y || !gW1y || !gb1y || !gW2y || hy <= 0 || hy > EMB_HMAX ||
                      q <= 0 || q > EMB_PMAX))
    return DISTBO_EARG;
  if (y_mode == 2 && (!Y || dy <= 0 || dy > EMB_PMAX)) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  EmbedBwdArgs b;
  b.f = EmbedArgs<double>{X, Y, seg_off, T, dx, dy, W1x, b1x, W2x, hx, p, W1y, b1y, W2y, hy, q, y_mode, nullptr, lde};
  b.dE = dE;
  const int nx = net_w_elems(dx, hx, p), ny = (y_mode == 1) ? net_w_elems(dy, hy, q) : 0;
  b.nparams = nx + ny;
  if (workspace_elems < (int64_t)EMB_BWD_GRID * b.nparams) return DISTBO_EARG;
  b.partial = workspace;
  const int gq = (y_mode == 0) ? 1 : (y_mode == 1 ? q : dy);
  const int hmax = hx > hy ? hx : hy, pmax = p > gq ? p : gq;
  size_t elems = (size_t)2 * (nx + ny) + EMB_THREADS * (EMB_KC + 1) + (size_t)2 * EMB_THREADS * hmax +
                 (size_t)EMB_THREADS * pmax;
  size_t smem = elems * sizeof(double) + 16;
  if (smem > 220 * 1024) return DISTBO_EARG;
  DB_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = EMB_BWD_GRID;  // grid-stride over 128-sample batches; S = seg_off[T] is read on device
  embed_bwd_kernel<<<grid, EMB_THREADS, smem, st>>>(b);
  DB_CHECK_LAUNCH();
  embed_bwd_reduce_kernel<<<(b.nparams + 255) / 256, 256, 0, st>>>(workspace, grid, b.nparams, nx, dx * hx, hx, gW1x,
                                                                   gb1x, gW2x, dy * hy, hy, gW1y, gb1y, gW2y);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
