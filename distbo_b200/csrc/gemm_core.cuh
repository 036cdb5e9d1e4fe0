// FP64 tensor-core (DMMA.8x8x4) tile GEMM core shared by the Cholesky / inverse / scoring stages.
//
// One CTA (256 threads, 8 warps as 2x4, warp tile 64x32) produces one 128x128 output tile
//     C[m0:m0+128, n0:n0+128] = alpha * sum_{k in [k_lo,k_hi)} opA(m,k) * opB(k,n) + beta * Cin
// from a tile descriptor (m0, n0, k_lo, k_hi) in GLOBAL matrix coordinates.  Operands are staged
// global -> shared with a 4-stage cp.async (LDGSTS) pipeline; shared rows are padded (+4 doubles)
// so that every m8n8k4 fragment load is bank-conflict free.  All dimensions are multiples of the
// tile (the host pads every matrix to a multiple of 128 with an identity block), so the hot loop
// has no bounds checks.  Triangular structure (SYRK lower tiles, TRMM/TRTRI/LAUUM k-ranges) is
// expressed purely through the host-generated tile list.
#pragma once
#include "common.cuh"

namespace db {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int STAGES = 4;
constexpr int LDK = BK + 4;             // k-contiguous tile: [128][20]
constexpr int LDX = BM + 4;             // row-contiguous tile: [16][132]
constexpr int STAGE_ELEMS = BM * LDK;   // 2560 doubles (>= BK*LDX = 2112)
constexpr int GEMM_SMEM_BYTES = STAGES * 2 * STAGE_ELEMS * (int)sizeof(double);  // 163840

// Operand layouts in global memory (both row-major matrices with leading dimension ld):
//   KMAJOR: element (x, k) at g[x*ld + k]   (x = m for A, n for B)  -> "A" / "B^T" style
//   XMAJOR: element (x, k) at g[k*ld + x]                           -> "A^T" / "B" style
enum { KMAJOR = 0, XMAJOR = 1 };

template <int LAY>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, int ld, int x0,
                                          int k0, int tid) {
  if (LAY == KMAJOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * GEMM_THREADS;
      int row = c >> 3, kc = c & 7;
      cp_async16(s + row * LDK + kc * 2, g + (size_t)(x0 + row) * ld + k0 + kc * 2);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * GEMM_THREADS;
      int kr = c >> 6, xc = c & 63;
      cp_async16(s + kr * LDX + xc * 2, g + (size_t)(k0 + kr) * ld + x0 + xc * 2);
    }
  }
}

template <int LAY>
__device__ __forceinline__ double frag_ld(const double* s, int x, int k) {
  return LAY == KMAJOR ? s[x * LDK + k] : s[k * LDX + x];
}

// acc[mi][ni][0..1]: warp tile 64x32 as 8x4 m8n8 sub-tiles.
template <int ALAY, int BLAY>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[8][4][2], const double* __restrict__ A,
                                              int lda, const double* __restrict__ B, int ldb, int m0,
                                              int n0, int k_lo, int k_hi, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  double* sA = smem;
  double* sB = smem + STAGES * STAGE_ELEMS;
  const int nk = (k_hi - k_lo) / BK;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      load_tile<ALAY>(sA + s * STAGE_ELEMS, A, lda, m0, k_lo + s * BK, tid);
      load_tile<BLAY>(sB + s * STAGE_ELEMS, B, ldb, n0, k_lo + s * BK, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int kn = kt + STAGES - 1;
      if (kn < nk) {
        int s = kn & (STAGES - 1);
        load_tile<ALAY>(sA + s * STAGE_ELEMS, A, lda, m0, k_lo + kn * BK, tid);
        load_tile<BLAY>(sB + s * STAGE_ELEMS, B, ldb, n0, k_lo + kn * BK, tid);
      }
      cp_async_commit();
    }
    const double* a = sA + (kt & (STAGES - 1)) * STAGE_ELEMS;
    const double* b = sB + (kt & (STAGES - 1)) * STAGE_ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) af[mi] = frag_ld<ALAY>(a, wm * 64 + mi * 8 + lr, k4 * 4 + lk);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = frag_ld<BLAY>(b, wn * 32 + ni * 8 + lr, k4 * 4 + lk);
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // all warps done with smem: safe to reuse it / to overwrite an in-place operand
}

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  const double* Cin;  // may alias C; ignored when beta == 0
  int lda, ldb, ldc, ldcin;
  double alpha, beta;
  const int4* tiles;  // (m0, n0, k_lo, k_hi) of 128 x 128 macro tiles
  unsigned long long* tl = nullptr;  // debug timeline (common.cuh), normally null
  int tl_id = 0;
};

// ---------------------------------------------------------------------------------------------------------
// Tile-list GEMM.  Each 128 x 128 macro tile of the list is computed by TWO CTAs (128 x 64 halves); a CTA
// has 8 warps as 4 x 2 with 32 x 32 warp tiles (64 accumulator registers per thread) and a 3-stage cp.async
// pipeline of 90 KB, so that two CTAs are resident per SM: while one CTA is fetching its C tile, filling its
// pipeline or storing its result, the other one keeps the DMMA pipe busy.  That hides the per-tile fixed cost,
// which dominates the k = 128 trailing updates of the Cholesky, and keeps the hardware's dynamic CTA
// scheduling (needed by the look-ahead: SMs become free for the high-priority panel every few microseconds).
constexpr int TN = 64;                       // CTA tile width
constexpr int TSTAGES = 3;
constexpr int LDXB = TN + 4;                 // row-contiguous B tile: [16][68]
constexpr int TSTAGE_A = BM * LDK;           // 2560 doubles (>= BK * LDX)
constexpr int TSTAGE_B = TN * LDK;           // 1280 doubles (>= BK * LDXB = 1088)
constexpr int TGEMM_SMEM_BYTES = TSTAGES * (TSTAGE_A + TSTAGE_B) * (int)sizeof(double);  // 92160

template <int LAY>
__device__ __forceinline__ void load_tile_b64(double* s, const double* __restrict__ g, int ld, int x0, int k0,
                                              int tid) {
  if (LAY == KMAJOR) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int row = c >> 3, kc = c & 7;
      cp_async16(s + row * LDK + kc * 2, g + (size_t)(x0 + row) * ld + k0 + kc * 2);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int kr = c >> 5, xc = c & 31;
      cp_async16(s + kr * LDXB + xc * 2, g + (size_t)(k0 + kr) * ld + x0 + xc * 2);
    }
  }
}

template <int LAY>
__device__ __forceinline__ double frag_ld_b64(const double* s, int x, int k) {
  return LAY == KMAJOR ? s[x * LDK + k] : s[k * LDXB + x];
}

// acc (32 x 32 warp tile of the 128 x 64 CTA tile at (m0, n0)) += sum_{k in [k_lo, k_hi)} opA(m, k) opB(k, n).
// smem: TGEMM_SMEM_BYTES.  Ends with a barrier (the buffers may be reused).
template <int ALAY, int BLAY>
__device__ __forceinline__ void gemm_mainloop_t64(double (&acc)[4][4][2], const double* __restrict__ A, int lda,
                                                  const double* __restrict__ B, int ldb, int m0, int n0, int k_lo,
                                                  int k_hi, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lk = lane & 3;
  double* sA = smem;
  double* sB = smem + TSTAGES * TSTAGE_A;
  const int nk = (k_hi - k_lo) / BK;
#pragma unroll
  for (int s = 0; s < TSTAGES - 1; ++s) {
    if (s < nk) {
      load_tile<ALAY>(sA + s * TSTAGE_A, A, lda, m0, k_lo + s * BK, tid);
      load_tile_b64<BLAY>(sB + s * TSTAGE_B, B, ldb, n0, k_lo + s * BK, tid);
    }
    cp_async_commit();
  }
  int cs = 0, ps = TSTAGES - 1;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<TSTAGES - 2>();
    __syncthreads();
    {
      const int kn = kt + TSTAGES - 1;
      if (kn < nk) {
        load_tile<ALAY>(sA + ps * TSTAGE_A, A, lda, m0, k_lo + kn * BK, tid);
        load_tile_b64<BLAY>(sB + ps * TSTAGE_B, B, ldb, n0, k_lo + kn * BK, tid);
      }
      cp_async_commit();
      ps = (ps + 1 == TSTAGES) ? 0 : ps + 1;
    }
    const double* a = sA + cs * TSTAGE_A;
    const double* b = sB + cs * TSTAGE_B;
    cs = (cs + 1 == TSTAGES) ? 0 : cs + 1;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[4], bf[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) af[mi] = frag_ld<ALAY>(a, wm * 32 + mi * 8 + lr, k4 * 4 + lk);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = frag_ld_b64<BLAY>(b, wn * 32 + ni * 8 + lr, k4 * 4 + lk);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

template <int ALAY, int BLAY>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_tiles_kernel(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  tl_start(p.tl, p.tl_id);
  const int4 t = p.tiles[blockIdx.x >> 1];
  const int m0 = t.x, n0 = t.y + TN * (blockIdx.x & 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lk = lane & 3;
  // out = alpha * (acc0 + A B) with acc0 = (beta / alpha) Cin: the C tile is requested before the pipeline
  // prologue, so its latency overlaps the first operand stages
  double acc[4][4][2];
  if (p.beta != 0.0) {
    const double f = p.beta / p.alpha;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int row = m0 + wm * 32 + mi * 8 + lr;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int col = n0 + wn * 32 + ni * 8 + lk * 2;
        const double2 c = *reinterpret_cast<const double2*>(p.Cin + (size_t)row * p.ldcin + col);
        acc[mi][ni][0] = f * c.x;
        acc[mi][ni][1] = f * c.y;
      }
    }
  } else {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }
  gemm_mainloop_t64<ALAY, BLAY>(acc, p.A, p.lda, p.B, p.ldb, m0, n0, t.z, t.w, smem);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int row = m0 + wm * 32 + mi * 8 + lr;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int col = n0 + wn * 32 + ni * 8 + lk * 2;
      *reinterpret_cast<double2*>(p.C + (size_t)row * p.ldc + col) =
          make_double2(p.alpha * acc[mi][ni][0], p.alpha * acc[mi][ni][1]);
    }
  }
  tl_end(p.tl, p.tl_id);
}

template <int ALAY, int BLAY>
inline cudaError_t launch_gemm_tiles(const GemmArgs& p, int ntiles, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  static PerDeviceOnce once;
  if (once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tiles_kernel<ALAY, BLAY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, TGEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    once.mark();
  }
  gemm_tiles_kernel<ALAY, BLAY><<<2 * ntiles, GEMM_THREADS, TGEMM_SMEM_BYTES, st>>>(p);
  ++g_db_launches;
  return cudaGetLastError();
}

}  // namespace db
