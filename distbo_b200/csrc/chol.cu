// K3: GP fit on FP64 tensor cores.
//   * distbo_potrf_f64 : blocked right-looking Cholesky, 128-wide panels.  The diagonal block is
//     factorised (and inverted) by one CTA in shared memory; the panel solve is a DMMA GEMM with the
//     inverted diagonal block; the trailing SYRK (lower tiles only) is a DMMA GEMM.  A high-priority
//     panel stream runs panel k+1 while the rest of trailing update k is still in flight (look-ahead).
//   * distbo_trtri_f64 : W = L^-1 by a level-batched recursive scheme (two DMMA GEMM launches per
//     tree level, all nodes of a level in one launch).
//   * distbo_lauum_f64 : Kinv = W^T W, lower tiles, one launch.
//   * distbo_nll_f64   : V = W (y-m), alpha = W^T V, negative log marginal likelihood.
// Replaces tf.cholesky / tf.matrix_triangular_solve / TF autodiff through them
// (reference train/log_gaus_likelihood.py:123,211; train/gp_utils.py:118-147).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../../include/distbo_b200.h"
#include "gemm_core.cuh"
#include "potf2.cuh"

namespace db {

// ------------------------------------------------------------------ plan
struct Range {
  int off, cnt;
};
struct TriNode {
  int o, n1, n2;
};

struct Plan {
  int n = 0, nb = 0;
  int4* d_tiles = nullptr;
  std::vector<Range> trsm, syrk_first, syrk_rest, tri1, tri2;
  Range lauum{0, 0};
  cudaStream_t panel_stream = nullptr;
  std::vector<cudaEvent_t> ev_panel, ev_first;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  bool lookahead = true;
};

static int build_tri(int o, int cnt, std::vector<std::vector<TriNode>>& levels) {
  if (cnt == 1) return 0;
  int n1 = cnt / 2, n2 = cnt - n1;
  int l1 = build_tri(o, n1, levels), l2 = build_tri(o + n1, n2, levels);
  int lvl = std::max(l1, l2) + 1;
  if ((int)levels.size() <= lvl) levels.resize(lvl + 1);
  levels[lvl].push_back({o, n1, n2});
  return lvl;
}

static bool longer_k(const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); }

}  // namespace db

using namespace db;

long long g_db_launches = 0;

extern "C" int distbo_version(void) { return 1000; }

extern "C" int64_t distbo_launch_count(int reset) {
  const long long v = g_db_launches;
  if (reset) g_db_launches = 0;
  return (int64_t)v;
}

extern "C" int distbo_plan_create(int n_pad, void** plan_out) {
  if (n_pad <= 0 || n_pad % 128 != 0 || !plan_out) return DISTBO_EARG;
  Plan* p = new Plan();
  p->n = n_pad;
  p->nb = n_pad / 128;
  const int nb = p->nb;
  std::vector<int4> tiles;
  auto T = [](int i) { return i * 128; };
  for (int kb = 0; kb < nb; ++kb) {
    Range r{(int)tiles.size(), 0};
    for (int i = kb + 1; i < nb; ++i) tiles.push_back(make_int4(T(i), T(kb), T(kb), T(kb + 1)));
    r.cnt = (int)tiles.size() - r.off;
    p->trsm.push_back(r);
    r = Range{(int)tiles.size(), 0};
    if (kb + 1 < nb)
      for (int i = kb + 1; i < nb; ++i) tiles.push_back(make_int4(T(i), T(kb + 1), T(kb), T(kb + 1)));
    r.cnt = (int)tiles.size() - r.off;
    p->syrk_first.push_back(r);
    r = Range{(int)tiles.size(), 0};
    for (int i = kb + 2; i < nb; ++i)
      for (int j = kb + 2; j <= i; ++j) tiles.push_back(make_int4(T(i), T(j), T(kb), T(kb + 1)));
    r.cnt = (int)tiles.size() - r.off;
    p->syrk_rest.push_back(r);
  }
  std::vector<std::vector<TriNode>> levels;
  build_tri(0, nb, levels);
  for (size_t l = 1; l < levels.size(); ++l) {
    std::vector<int4> t1, t2;
    for (const TriNode& nd : levels[l]) {
      const int mid = nd.o + nd.n1;
      for (int i = mid; i < mid + nd.n2; ++i)
        for (int j = nd.o; j < mid; ++j) {
          t1.push_back(make_int4(T(i), T(j), T(j), T(mid)));
          t2.push_back(make_int4(T(i), T(j), T(mid), T(i + 1)));
        }
    }
    std::stable_sort(t1.begin(), t1.end(), longer_k);
    std::stable_sort(t2.begin(), t2.end(), longer_k);
    p->tri1.push_back(Range{(int)tiles.size(), (int)t1.size()});
    tiles.insert(tiles.end(), t1.begin(), t1.end());
    p->tri2.push_back(Range{(int)tiles.size(), (int)t2.size()});
    tiles.insert(tiles.end(), t2.begin(), t2.end());
  }
  p->lauum.off = (int)tiles.size();
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j <= i; ++j) tiles.push_back(make_int4(T(i), T(j), T(i), T(nb)));
  p->lauum.cnt = (int)tiles.size() - p->lauum.off;

  if (cudaMalloc(&p->d_tiles, std::max<size_t>(1, tiles.size()) * sizeof(int4)) != cudaSuccess) {
    delete p;
    return DISTBO_ECUDA;
  }
  if (!tiles.empty() &&
      cudaMemcpy(p->d_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(p->d_tiles);
    delete p;
    return DISTBO_ECUDA;
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&p->panel_stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
    cudaFree(p->d_tiles);
    delete p;
    return DISTBO_ECUDA;
  }
  p->ev_panel.resize(nb);
  p->ev_first.resize(nb);
  for (int i = 0; i < nb; ++i) {
    cudaEventCreateWithFlags(&p->ev_panel[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->ev_first[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&p->ev_begin, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_end, cudaEventDisableTiming);
  const char* la = getenv("DISTBO_LOOKAHEAD");
  p->lookahead = !(la && la[0] == '0');
  cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
  *plan_out = p;
  return DISTBO_OK;
}

extern "C" int distbo_plan_destroy(void* plan) {
  if (!plan) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  cudaDeviceSynchronize();
  for (auto e : p->ev_panel) cudaEventDestroy(e);
  for (auto e : p->ev_first) cudaEventDestroy(e);
  if (p->ev_begin) cudaEventDestroy(p->ev_begin);
  if (p->ev_end) cudaEventDestroy(p->ev_end);
  if (p->panel_stream) cudaStreamDestroy(p->panel_stream);
  if (p->d_tiles) cudaFree(p->d_tiles);
  delete p;
  return DISTBO_OK;
}

extern "C" int distbo_potrf_f64(void* plan, double* A, double* W, int32_t* info, void* stream) {
  if (!plan || !A || !W || !info) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  cudaStream_t U = (cudaStream_t)stream;
  const int n = p->n, nb = p->nb;
  DB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), U));

  GemmArgs trsm{};  // A[i-blk, kb] <- A[i-blk, kb] * inv(L_kk)^T   (in place; B = W diagonal block)
  trsm.A = A; trsm.B = W; trsm.C = A; trsm.Cin = A;
  trsm.lda = trsm.ldb = trsm.ldc = trsm.ldcin = n;
  trsm.alpha = 1.0; trsm.beta = 0.0;
  GemmArgs syrk{};  // C -= P P^T
  syrk.A = A; syrk.B = A; syrk.C = A; syrk.Cin = A;
  syrk.lda = syrk.ldb = syrk.ldc = syrk.ldcin = n;
  syrk.alpha = -1.0; syrk.beta = 1.0;

  const bool la = p->lookahead && nb > 2;
  cudaStream_t P = la ? p->panel_stream : U;
  if (la) {
    DB_CUDA(cudaEventRecord(p->ev_begin, U));
    DB_CUDA(cudaStreamWaitEvent(P, p->ev_begin, 0));
  }
  for (int kb = 0; kb < nb; ++kb) {
    potf2_inv_kernel<<<1, POTF2_THREADS, POTF2_SMEM, P>>>(A, n, W, kb, info);
    DB_CHECK_LAUNCH();
    if (p->trsm[kb].cnt) {
      trsm.tiles = p->d_tiles + p->trsm[kb].off;
      DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(trsm, p->trsm[kb].cnt, P)));
    }
    if (la) {
      DB_CUDA(cudaEventRecord(p->ev_panel[kb], P));
      DB_CUDA(cudaStreamWaitEvent(U, p->ev_panel[kb], 0));
    }
    if (p->syrk_first[kb].cnt) {
      syrk.tiles = p->d_tiles + p->syrk_first[kb].off;
      DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(syrk, p->syrk_first[kb].cnt, U)));
    }
    if (la) {
      DB_CUDA(cudaEventRecord(p->ev_first[kb], U));
      DB_CUDA(cudaStreamWaitEvent(P, p->ev_first[kb], 0));
    }
    if (p->syrk_rest[kb].cnt) {
      syrk.tiles = p->d_tiles + p->syrk_rest[kb].off;
      DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(syrk, p->syrk_rest[kb].cnt, U)));
    }
  }
  if (la) {
    DB_CUDA(cudaEventRecord(p->ev_end, P));
    DB_CUDA(cudaStreamWaitEvent(U, p->ev_end, 0));
  }
  return DISTBO_OK;
}

extern "C" int distbo_trtri_f64(void* plan, const double* L, double* W, double* T, void* stream) {
  if (!plan || !L || !W || !T) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = p->n;
  for (size_t l = 0; l < p->tri1.size(); ++l) {
    GemmArgs g1{};  // T = L21 * W11
    g1.A = L; g1.B = W; g1.C = T; g1.Cin = T;
    g1.lda = g1.ldb = g1.ldc = g1.ldcin = n;
    g1.alpha = 1.0; g1.beta = 0.0;
    g1.tiles = p->d_tiles + p->tri1[l].off;
    DB_CUDA((launch_gemm_tiles<KMAJOR, XMAJOR>(g1, p->tri1[l].cnt, st)));
    GemmArgs g2{};  // W21 = -W22 * T
    g2.A = W; g2.B = T; g2.C = W; g2.Cin = W;
    g2.lda = g2.ldb = g2.ldc = g2.ldcin = n;
    g2.alpha = -1.0; g2.beta = 0.0;
    g2.tiles = p->d_tiles + p->tri2[l].off;
    DB_CUDA((launch_gemm_tiles<KMAJOR, XMAJOR>(g2, p->tri2[l].cnt, st)));
  }
  return DISTBO_OK;
}

extern "C" int distbo_lauum_f64(void* plan, const double* W, double* Kinv, void* stream) {
  if (!plan || !W || !Kinv) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  GemmArgs g{};
  g.A = W; g.B = W; g.C = Kinv; g.Cin = Kinv;
  g.lda = g.ldb = g.ldc = g.ldcin = p->n;
  g.alpha = 1.0; g.beta = 0.0;
  g.tiles = p->d_tiles + p->lauum.off;
  DB_CUDA((launch_gemm_tiles<XMAJOR, XMAJOR>(g, p->lauum.cnt, (cudaStream_t)stream)));
  return DISTBO_OK;
}

// ------------------------------------------------------------------ vectors and the likelihood
namespace db {

// V[i] = sum_{k<=i} W[i][k] * (y[k] - mean)   (one warp per row)
__global__ void __launch_bounds__(256) gemv_lower_kernel(const double* __restrict__ W, int n_pad, int N,
                                                         const double* __restrict__ y,
                                                         const double* __restrict__ mean,
                                                         double* __restrict__ V) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  const double m = *mean;
  const double* w = W + (size_t)row * n_pad;
  double acc = 0.0;
  const int kmax = min(row, N - 1);
  for (int k = lane; k <= kmax; k += 32) acc += w[k] * (y[k] - m);
  acc = warp_sum(acc);
  if (lane == 0) V[row] = acc;
}

// alpha[j] = sum_{i>=j} W[i][j] * V[i] in two coalesced passes with a fixed summation order:
//   partial[b][j] = sum_{i in rows of block b} W[i][j] V[i]   (block b = GVT_ROWS rows; thread = column, so a
//                   warp reads 256 contiguous bytes of each row; entries above the diagonal inside the
//                   diagonal 128-block are stored zeros)
//   alpha[j]      = sum_b partial[b][j] over the blocks whose 128-block row is >= j / 128.
constexpr int GVT_ROWS = 32;
__global__ void __launch_bounds__(256) gemvT_partial_kernel(const double* __restrict__ W, int n_pad,
                                                            const double* __restrict__ V,
                                                            double* __restrict__ partial) {
  __shared__ double sv[GVT_ROWS];
  const int i0 = blockIdx.x * GVT_ROWS;
  if (threadIdx.x < GVT_ROWS) sv[threadIdx.x] = V[i0 + threadIdx.x];
  __syncthreads();
  const int jmax = (i0 / 128 + 1) * 128;
  for (int j = threadIdx.x; j < jmax; j += 256) {
    const double* w = W + (size_t)i0 * n_pad + j;
    double acc = 0.0;
#pragma unroll 8
    for (int r = 0; r < GVT_ROWS; ++r) acc += w[(size_t)r * n_pad] * sv[r];
    partial[(size_t)blockIdx.x * n_pad + j] = acc;
  }
}

__global__ void __launch_bounds__(256) gemvT_reduce_kernel(const double* __restrict__ partial, int n_pad,
                                                           double* __restrict__ alpha) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n_pad) return;
  const int nblk = n_pad / GVT_ROWS;
  double s = 0.0;
  for (int b = (j / 128) * (128 / GVT_ROWS); b < nblk; ++b) s += partial[(size_t)b * n_pad + j];
  alpha[j] = s;
}

// nll = 0.5 |V|^2 + 0.5 N log(2 pi) + sum_i log L_ii    (reference train/gp_utils.py:141-147)
__global__ void __launch_bounds__(1024) nll_kernel(const double* __restrict__ L, int n_pad, int N,
                                                   const double* __restrict__ V, double* __restrict__ nll) {
  __shared__ double r1[32], r2[32];
  double q = 0.0, ld = 0.0;
  for (int i = threadIdx.x; i < N; i += 1024) {
    const double v = V[i];
    q += v * v;
    ld += log(L[(size_t)i * n_pad + i]);
  }
  q = warp_sum(q);
  ld = warp_sum(ld);
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = q;
    r2[threadIdx.x >> 5] = ld;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    q = warp_sum(r1[threadIdx.x]);
    ld = warp_sum(r2[threadIdx.x]);
    if (threadIdx.x == 0) *nll = 0.5 * q + 0.5 * (double)N * 1.8378770664093454836 + ld;
  }
}

}  // namespace db

extern "C" int64_t distbo_nll_workspace(int n_pad) { return (int64_t)(n_pad / GVT_ROWS) * n_pad; }

extern "C" int distbo_nll_f64(const double* L, const double* W, int N, int n_pad, const double* y,
                              const double* mean, double* V, double* alpha, double* nll, double* workspace,
                              void* stream) {
  if (!L || !W || !y || !mean || !V || !alpha || !nll || !workspace || N <= 0 || n_pad < N || n_pad % 128)
    return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  gemv_lower_kernel<<<n_pad / 8, 256, 0, st>>>(W, n_pad, N, y, mean, V);
  DB_CHECK_LAUNCH();
  gemvT_partial_kernel<<<n_pad / GVT_ROWS, 256, 0, st>>>(W, n_pad, V, workspace);
  DB_CHECK_LAUNCH();
  gemvT_reduce_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(workspace, n_pad, alpha);
  DB_CHECK_LAUNCH();
  nll_kernel<<<1, 1024, 0, st>>>(L, n_pad, N, V, nll);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
