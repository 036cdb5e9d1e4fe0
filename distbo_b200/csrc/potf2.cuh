// Diagonal block of the blocked Cholesky: potf2 + inverse of a 128x128 block by ONE CTA.
//   factor : right-looking over four 32-wide sub-panels; the 32x32 diagonal sub-block is factorised
//            and inverted by one warp entirely in registers (lane = row / column, operands exchanged
//            by warp shuffles); the sub-panel solve (x Dinv^T) and the trailing update run on DMMA.
//   inverse: block forward substitution over the 4x4 grid of 32x32 blocks on DMMA,
//            W_ij = -W_ii * sum_{k=j}^{i-1} L_ik W_kj, by block distance (3 dependent levels).
// Shared layout: sA [128][132] holds L in its lower part; the strictly upper 32-blocks (j,i) receive
// W_ij^T; the four diagonal inverses live in sDinv [4][32][36]; sS [3][32][36] is the S scratch.
// Row strides 132 / 36 (== 4 mod 16 doubles) make every m8n8k4 fragment load conflict free.
#pragma once
#include "common.cuh"

namespace db {

constexpr int PB = 128;
constexpr int PLD = 132;
constexpr int SB = 32;
constexpr int DLD = 36;
constexpr int POTF2_THREADS = 256;
constexpr int POTF2_SMEM = (PB * PLD + 4 * SB * DLD + 3 * SB * DLD + SB) * (int)sizeof(double);

// Branch-free 1/sqrt(d) for d > 0 in the float range: MUFU.RSQ seed + two Newton steps (~1 ulp).
// No slow-path branches, so the surrounding warp stays provably convergent (no WARPSYNC per shuffle).
__device__ __forceinline__ double rsqrt_nr(double d) {
  double y = (double)rsqrtf((float)d);
  const double h = 0.5 * d;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// Factor the 32x32 block at sA[c0.., c0..] (lower) entirely in registers: lane = row, operands exchanged
// by warp shuffles.  Executed by EVERY warp of the CTA redundantly so that the shuffles sit in fully
// convergent code (a warp-id guard would cost a WARPSYNC/ENDCOLLECTIVE pair per shuffle); only warp 0
// stores.  a[] returns row `lane` of L, myr = 1/L[lane][lane].  Returns the 1-based local index of the
// first non-positive pivot (0 = ok), identical in all lanes.
__device__ __forceinline__ int factor_32(double (&a)[SB], double& myr, int lane) {
  int bad = 0;
  myr = 0.0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    bad = (!(d > 0.0) && bad == 0) ? j + 1 : bad;
    double y = rsqrt_nr(d);
    double sq = d * y;
    sq = fma(0.5 * y, fma(-sq, sq, d), sq);   // sqrt(d), one correction step
    y = fma(y, fma(-sq, y, 1.0), y);          // 1/sqrt(d) consistent with sq
    a[j] = (lane == j) ? sq : a[j] * y;       // lanes < j hold junk in a[j]; never used
    myr = (lane == j) ? y : myr;
#pragma unroll
    for (int k = 0; k < SB; ++k) {
      if (k > j) {  // folds at compile time after full unrolling
        const double lkj = __shfl_sync(0xffffffffu, a[j], k);
        a[k] = (lane >= k) ? fma(-a[j], lkj, a[k]) : a[k];
      }
    }
  }
  return bad;
}

// Variant used by default: warp 0 alone factors the block (lane = row).  Column j is exchanged through a
// double-buffered 32-entry shared-memory line (one store + broadcast 128-bit loads) instead of 62 shuffle
// instructions, there is no per-entry select (entries above the diagonal hold junk that is never read), and
// the only dependent chain per column is  pivot -> 1/d (float seed + two Newton steps) -> t = a_j / d ->
// fma into the next pivot;  sqrt / rsqrt for the stored column run off that chain.
//   update:  a[k] -= (a[j] / d) * c_k,  c_k = unscaled column entry of row k  (= L_ij L_kj)
__device__ __forceinline__ double rcp_nr(double d) {
  double x = (double)__frcp_rn((float)d);
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

__device__ __forceinline__ int factor_32_smem(double (&a)[SB], double& myr, int lane, double* sCol /*[2][SB]*/) {
  int bad = 0;
  myr = 0.0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    double* line = sCol + (j & 1) * SB;
    line[lane] = a[j];
    __syncwarp();
    const double d = line[j];
    bad = (!(d > 0.0) && bad == 0) ? j + 1 : bad;
    const double t = a[j] * rcp_nr(d);
#pragma unroll
    for (int k2 = 0; k2 < SB; k2 += 2) {
      if (k2 + 1 > j) {  // folds at compile time after full unrolling
        const double2 c = *reinterpret_cast<const double2*>(line + k2);
        if (k2 > j) a[k2] = fma(-t, c.x, a[k2]);
        a[k2 + 1] = fma(-t, c.y, a[k2 + 1]);
      }
    }
    // off the critical chain: the stored column L[.][j] = a[j] / sqrt(d), L[j][j] = sqrt(d)
    double y = rsqrt_nr(d);
    double sq = d * y;
    sq = fma(0.5 * y, fma(-sq, sq, d), sq);
    y = fma(y, fma(-sq, y, 1.0), y);
    a[j] = (lane == j) ? sq : a[j] * y;
    myr = (lane == j) ? y : myr;
  }
  return bad;
}

// Default since round 2: the pivot chain with 2x2 block pivots.  Columns (j, j+1) are eliminated together:
//   D = [p q; q r] (rows j, j+1 of the two current columns),  det = p r - q^2,
//   u = (r c1 - q c2) / det,  v = (p c2 - q c1) / det,   a[k] -= u C1[k] + v C2[k]   for k >= j + 2,
// i.e. ONE reciprocal per two columns on the dependent chain (measured with tools/factor32_bench.cu on B200: a
// dependent DFMA costs 13.3 cycles; 1x1 pivots 289 cycles per column, 2x2 pivots 178), and nothing else on it:
// warp 0 leaves the UNSCALED columns in sC [32][32] (sC[j][k] = entry of row k in column j before the pair step,
// column j+1 not yet updated by column j) and the whole CTA forms L from them afterwards (scale_2x2):
//   l_j = c1 / sqrt(p),   l_{j+1} = (c2 - (q/p) c1) / sqrt(r - q^2/p).
// Algebraically the same factor; rounding differs from the column-by-column form at the 1e-16 level.
// Returns the 1-based local index of the first non-positive pivot (0 = ok).
__device__ __forceinline__ int chain_2x2(double (&a)[SB], int lane, double* sC /*[SB][SB]*/, double* rinv /*[SB]*/,
                                         double* tq /*[SB/2]*/) {
  int bad = 0;
  double kp = 1.0, kq = 0.0, kr = 1.0;   // lane L keeps the 2x2 pivot block of pair L (columns 2L, 2L+1)
#pragma unroll
  for (int j = 0; j < SB; j += 2) {
    double* l1 = sC + j * SB;
    double* l2 = l1 + SB;
    l1[lane] = a[j];
    l2[lane] = a[j + 1];
    __syncwarp();
    const double p = l1[j], q = l1[j + 1], r = l2[j + 1];
    const double det = fma(p, r, -q * q);
    bad = (bad == 0 && !(p > 0.0)) ? j + 1 : bad;
    bad = (bad == 0 && !(det > 0.0)) ? j + 2 : bad;
    const double id = rcp_nr(det);
    const double u = fma(r, a[j], -q * a[j + 1]) * id;
    const double v = fma(p, a[j + 1], -q * a[j]) * id;
    const bool mine = lane == (j >> 1);
    kp = mine ? p : kp;
    kq = mine ? q : kq;
    kr = mine ? r : kr;
#pragma unroll
    for (int k2 = 0; k2 < SB; k2 += 2) {
      if (k2 > j + 1) {  // folds at compile time after full unrolling
        const double2 c1 = *reinterpret_cast<const double2*>(l1 + k2);
        const double2 c2 = *reinterpret_cast<const double2*>(l2 + k2);
        a[k2] = fma(-v, c2.x, fma(-u, c1.x, a[k2]));
        a[k2 + 1] = fma(-v, c2.y, fma(-u, c1.y, a[k2 + 1]));
      }
    }
  }
  // off the chain: the scaling constants of every pair, one lane each (scale_2x2 applies them):
  //   isp = 1 / sqrt(p),  t = q isp = L[j+1][j],  iss = 1 / sqrt(r - t^2)
  if (lane < SB / 2) {
    const double isp = rsqrt_nr(kp);
    const double t = kq * isp;
    const double iss = rsqrt_nr(fma(-t, t, kr));
    rinv[2 * lane] = isp;
    rinv[2 * lane + 1] = iss;
    tq[lane] = t;
  }
  return bad;
}

// L (lower 32x32 block at sA[c0.., c0..]) from the unscaled columns chain_2x2 left in sC and its per-pair constants
// (rinv[j] = 1 / L[j][j], tq[j/2] = L[j+1][j]); executed by all POTF2_THREADS threads: one (row, column pair) item
// each, two rounds, two multiplications per entry.  The diagonal comes out as p / sqrt(p) and s / sqrt(s).
__device__ __forceinline__ void scale_2x2(const double* sC, double* sA, int c0, const double* rinv, const double* tq,
                                          int tid) {
  for (int it = tid; it < SB * (SB / 2); it += POTF2_THREADS) {
    const int row = it & (SB - 1), j = (it >> 5) * 2;
    const double isp = rinv[j], iss = rinv[j + 1], t = tq[j >> 1];
    const double l1 = sC[j * SB + row] * isp;
    const double l2 = fma(-t, l1, sC[(j + 1) * SB + row]) * iss;
    if (row >= j) sA[(c0 + row) * PLD + c0 + j] = l1;
    if (row >= j + 1) sA[(c0 + row) * PLD + c0 + j + 1] = l2;
  }
}

// Inverse of the 32x32 lower-triangular block just stored at sA[c0.., c0..]: lane j owns column j of
// Dinv; L[i][k] and 1/L[i][i] come from shared memory as broadcast loads (no shuffles: may run under a
// warp-id guard).  dinv is row-major [32][DLD] with explicit zeros above the diagonal.
__device__ __forceinline__ void invert_32(const double* sA, int c0, double* dinv, const double* rinv, int lane) {
  double w[SB];
#pragma unroll
  for (int i = 0; i < SB; ++i) {
    // four independent accumulation chains: a dependent DFMA costs ~13 cycles on B200, the row sum is the chain
    double acc0 = (lane == i) ? 1.0 : 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
    for (int k = 0; k < SB; ++k) {
      if (k < i) {
        const double lik = sA[(c0 + i) * PLD + c0 + k];
        if ((k & 3) == 0) acc0 = fma(-lik, w[k], acc0);
        else if ((k & 3) == 1) acc1 = fma(-lik, w[k], acc1);
        else if ((k & 3) == 2) acc2 = fma(-lik, w[k], acc2);
        else acc3 = fma(-lik, w[k], acc3);
      }
    }
    w[i] = ((acc0 + acc1) + (acc2 + acc3)) * rinv[i];
  }
#pragma unroll
  for (int k = 0; k < SB; ++k) dinv[k * DLD + lane] = w[k];  // Dinv[k][lane]
}

#define PT_STAMP(dbg) do { if ((dbg) && threadIdx.x == 0) *(dbg)++ = clock64(); } while (0)

// Factor + invert the 128x128 block held in shared memory (sA [128][PLD], lower part valid, upper zero).
// On return: lower sA = L; strictly upper 32-blocks (j,i) = W_ij^T; sDinv = the four diagonal inverses.
// row_base = global row of the block's first row (for `info`).  dbg (optional) receives clock64 stamps.

// Block forward substitution over the 4x4 grid of 32x32 blocks: fills the strictly upper 32-blocks (j,i) of sA
// with W_ij^T given L in the lower part of sA and the four diagonal inverses in sDinv.
__device__ __forceinline__ void potf2_assemble_inverse(double* s, long long*& dbg) {
  double* sA = s;
  double* sDinv = s + PB * PLD;
  double* sS = sDinv + 4 * SB * DLD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  (void)tid;
  // ---- inverse by block distance
  for (int dist = 1; dist < 4; ++dist) {
    const int nblk = 4 - dist;
    for (int t = warp; t < nblk * 16; t += 8) {
      const int b = t >> 4, tr = (t >> 2) & 3, tc = t & 3;
      const int j = b, i = b + dist;
      double c0v = 0.0, c1v = 0.0;
      for (int kb = j; kb < i; ++kb) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const double av = sA[(SB * i + tr * 8 + lr) * PLD + SB * kb + kk * 4 + lk];
          const double bv = (kb == j) ? sDinv[j * SB * DLD + (kk * 4 + lk) * DLD + tc * 8 + lr]
                                      : sA[(SB * j + tc * 8 + lr) * PLD + SB * kb + kk * 4 + lk];
          dmma884(c0v, c1v, av, bv);
        }
      }
      *reinterpret_cast<double2*>(&sS[b * SB * DLD + (tr * 8 + lr) * DLD + tc * 8 + 2 * lk]) =
          make_double2(c0v, c1v);
    }
    __syncthreads();
    for (int t = warp; t < nblk * 16; t += 8) {
      const int b = t >> 4, tr = (t >> 2) & 3, tc = t & 3;
      const int j = b, i = b + dist;
      double c0v = 0.0, c1v = 0.0;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        dmma884(c0v, c1v, sDinv[i * SB * DLD + (tr * 8 + lr) * DLD + kk * 4 + lk],
                sS[b * SB * DLD + (kk * 4 + lk) * DLD + tc * 8 + lr]);
      // W_ij[r][c] -> transposed into the upper block (j,i)
      const int r = tr * 8 + lr, c = tc * 8 + 2 * lk;
      sA[(SB * j + c) * PLD + SB * i + r] = -c0v;
      sA[(SB * j + c + 1) * PLD + SB * i + r] = -c1v;
    }
    __syncthreads();
    PT_STAMP(dbg);
  }
}

// Optional early publication of the sub-steps (used by the Cholesky panel kernel): in sub-step jb, as soon as the
// diagonal 32-block is factored and inverted, the final rows [32 jb, 32 jb + 32) of L and Dinv_jb go to their
// final places in global memory (Ab / Wb, leading dimension ld) and flags[jb] is set.  (Publishing the block
// column below separately, so that consumers can start step s+1 earlier, was measured slower: +1.5 us per
// sub-step on the diagonal CTA against -2 us on the consumers' tail.)
struct Potf2Publish {
  double* Ab;
  double* Wb;
  int ld;
  int* flags;   // [4] sub-step flags; null: no publication
};

// Sub-panel solve of sub-step jb for the 8-row groups [rg_lo, rg_hi) below the diagonal sub-block:
// P <- P * Dinv_jb^T on DMMA (each warp owns whole row groups: in place is safe).  `w` of `nw` warps take part.
__device__ __forceinline__ void potf2_solve_groups(double* sA, const double* sDinvJ, int c0, int rg_lo, int rg_hi,
                                                   int w, int nw, int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  for (int rg = rg_lo + w; rg < rg_hi; rg += nw) {
    const int r0 = c0 + SB + rg * 8;
    double af[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) af[kk] = sA[(r0 + lr) * PLD + c0 + kk * 4 + lk];
    double acc[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        dmma884(acc[nt][0], acc[nt][1], af[kk], sDinvJ[(nt * 8 + lr) * DLD + kk * 4 + lk]);
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<double2*>(&sA[(r0 + lr) * PLD + c0 + nt * 8 + 2 * lk]) = make_double2(acc[nt][0], acc[nt][1]);
  }
}

// Trailing update C -= P P^T of sub-step jb for the lower 8x8 tiles t in [t_lo, t_hi) (row-major enumeration of the
// triangle below / right of the diagonal sub-block: t = ri (ri + 1) / 2 + ci).
__device__ __forceinline__ void potf2_update_tiles(double* sA, int c0, int t_lo, int t_hi, int w, int nw, int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  for (int t = t_lo + w; t < t_hi; t += nw) {
    int ri = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
    while (ri * (ri + 1) / 2 > t) --ri;
    const int ci = t - ri * (ri + 1) / 2;
    const int r0 = c0 + SB + ri * 8, q0 = c0 + SB + ci * 8;
    double2 c = *reinterpret_cast<double2*>(&sA[(r0 + lr) * PLD + q0 + 2 * lk]);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
      dmma884(c.x, c.y, -sA[(r0 + lr) * PLD + c0 + kk * 4 + lk], sA[(q0 + lr) * PLD + c0 + kk * 4 + lk]);
    __syncwarp();
    *reinterpret_cast<double2*>(&sA[(r0 + lr) * PLD + q0 + 2 * lk]) = c;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Publication of sub-step jb by the threads [t0, t0 + nt) of the CTA: rows [32 jb, 32 jb + 32) of L (zeros above the
// diagonal) and Dinv_jb go to global memory, then flags[jb] is set.  bar_id: named barrier of exactly those threads.
__device__ __forceinline__ void potf2_publish(const double* sA, const double* sDinv, const Potf2Publish& pub, int jb,
                                              int t0, int nt, int bar_id) {
  const int c0 = jb * SB;
  const int rel = threadIdx.x - t0;
  for (int idx = rel; idx < SB * PB; idx += nt) {
    const int r = idx >> 7, c = idx & 127;
    pub.Ab[(size_t)(c0 + r) * pub.ld + c] = (c <= c0 + r) ? sA[(c0 + r) * PLD + c] : 0.0;
  }
  for (int idx = rel; idx < SB * SB; idx += nt) {
    const int r = idx >> 5, c = idx & 31;
    pub.Wb[(size_t)(c0 + r) * pub.ld + c0 + c] = sDinv[jb * SB * DLD + r * DLD + c];
  }
  if (nt == POTF2_THREADS) __syncthreads();
  else named_bar_sync(bar_id, nt);
  if (rel == 0) {   // cumulative fence: the barrier ordered every participating thread's stores before this one
    __threadfence();
    atomicExch(pub.flags + jb, 1);
  }
}

// Factor + invert the 128x128 block in shared memory.  Round 2: the dependent chain of a 32-column sub-step is
//   chain_2x2 (warp 0) -> scale_2x2 -> invert_32 -> solve + update of the NEXT 32 rows / columns only,
// and everything else of the sub-step - its publication, the solve of the rows further below and their updates -
// runs on warps 1..7 WHILE warp 0 is already in the next sub-step's chain (look-ahead inside the block).
__device__ __forceinline__ void potf2_factor_invert(double* s, int row_base, int32_t* info, long long*& dbg,
                                                    Potf2Publish pub = Potf2Publish{nullptr, nullptr, 0, nullptr},
                                                    bool assemble = true) {
  double* sA = s;
  double* sDinv = s + PB * PLD;
  double* sS = sDinv + 4 * SB * DLD;
  double* sR = sS + 3 * SB * DLD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef DB_FACTOR_1X1
  const int lr = lane >> 2, lk = lane & 3;
  (void)lr; (void)lk;
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * SB;
    {
      double a[SB], myr;
#pragma unroll
      for (int k = 0; k < SB; ++k) a[k] = sA[(c0 + lane) * PLD + c0 + k];
      __syncthreads();  // every warp has read the block before warp 0 overwrites it
      if (warp == 0) {
        const int bad = factor_32_smem(a, myr, lane, sS);
#pragma unroll
        for (int k = 0; k < SB; ++k)
          if (k <= lane) sA[(c0 + lane) * PLD + c0 + k] = a[k];
        sR[lane] = myr;
        if (bad && lane == 0 && *info == 0) *info = row_base + c0 + bad;
      }
    }
    __syncthreads();
    PT_STAMP(dbg);
    if (warp == 0) invert_32(sA, c0, sDinv + jb * SB * DLD, sR, lane);
    __syncthreads();
    PT_STAMP(dbg);
    if (pub.flags) potf2_publish(sA, sDinv, pub, jb, 0, POTF2_THREADS, 0);
    const int nrg = (PB - c0 - SB) / 8;  // 8-row groups below the diagonal sub-block
    potf2_solve_groups(sA, sDinv + jb * SB * DLD, c0, 0, nrg, warp, 8, lane);
    __syncthreads();
    PT_STAMP(dbg);
    potf2_update_tiles(sA, c0, 0, nrg * (nrg + 1) / 2, warp, 8, lane);
    __syncthreads();
    PT_STAMP(dbg);
  }
#else
  constexpr int NEAR_RG = SB / 8;                      // row groups of the next 32-row block
  constexpr int NEAR_TILES = NEAR_RG * (NEAR_RG + 1) / 2;
  constexpr int FAR_THREADS = POTF2_THREADS - 32;      // warps 1..7
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * SB;
    if (warp == 0) {
      // the pivot chain of sub-step jb: warp 0 alone, 2x2 block pivots, unscaled columns -> sS (used as sC)
      double a[SB];
#pragma unroll
      for (int k = 0; k < SB; ++k) a[k] = sA[(c0 + lane) * PLD + c0 + k];
      const int bad = chain_2x2(a, lane, sS, sR, sS + SB * SB);
      if (bad && lane == 0 && *info == 0) *info = row_base + c0 + bad;
    } else if (jb > 0) {
      // meanwhile, warps 1..7: the rest of sub-step jb - 1 (nothing the chain above reads or writes)
      const int pj = jb - 1, pc0 = pj * SB;
      const int nrg = (PB - pc0 - SB) / 8;
      if (pub.flags) potf2_publish(sA, sDinv, pub, pj, 32, FAR_THREADS, 1);
      potf2_solve_groups(sA, sDinv + pj * SB * DLD, pc0, NEAR_RG, nrg, warp - 1, 7, lane);
      named_bar_sync(1, FAR_THREADS);
      potf2_update_tiles(sA, pc0, NEAR_TILES, nrg * (nrg + 1) / 2, warp - 1, 7, lane);
    }
    __syncthreads();
    PT_STAMP(dbg);
    scale_2x2(sS, sA, c0, sR, sS + SB * SB, tid);   // all warps: the L block from the unscaled columns
    __syncthreads();
    if (warp == 0) invert_32(sA, c0, sDinv + jb * SB * DLD, sR, lane);
    __syncthreads();
    PT_STAMP(dbg);
    if (jb < 3) {
      // what the next chain waits for: the next 32 rows of this sub-panel and the next diagonal sub-block
      potf2_solve_groups(sA, sDinv + jb * SB * DLD, c0, 0, NEAR_RG, warp, 8, lane);
      __syncthreads();
      potf2_update_tiles(sA, c0, 0, NEAR_TILES, warp, 8, lane);
      __syncthreads();
      PT_STAMP(dbg);
    }
  }
  if (pub.flags) potf2_publish(sA, sDinv, pub, 3, 0, POTF2_THREADS, 0);   // last sub-step: nothing below it
  PT_STAMP(dbg);
#endif

  if (assemble) potf2_assemble_inverse(s, dbg);
}

// W block (inverse, lower) from the shared-memory result to global memory.
__device__ __forceinline__ void potf2_store_W(const double* s, double* __restrict__ Wb, int ld) {
  const double* sA = s;
  const double* sDinv = s + PB * PLD;
  const int tid = threadIdx.x;
#pragma unroll 8
  for (int it = 0; it < PB * PB / POTF2_THREADS; ++it) {
    const int idx = tid + it * POTF2_THREADS;
    const int r = idx >> 7, c = idx & 127;
    const int rb = r >> 5, cb = c >> 5;
    double wv = 0.0;
    if (rb == cb) wv = sDinv[rb * SB * DLD + (r & 31) * DLD + (c & 31)];
    else if (rb > cb) wv = sA[c * PLD + r];
    Wb[(size_t)r * ld + c] = wv;
  }
}

// L block (lower, zeros above the diagonal) to global memory.
__device__ __forceinline__ void potf2_store_L(const double* s, double* __restrict__ Ab, int ld) {
  const double* sA = s;
  const int tid = threadIdx.x;
#pragma unroll 8
  for (int it = 0; it < PB * PB / POTF2_THREADS; ++it) {
    const int idx = tid + it * POTF2_THREADS;
    const int r = idx >> 7, c = idx & 127;
    Ab[(size_t)r * ld + c] = (c <= r) ? sA[r * PLD + c] : 0.0;
  }
}

// Stand-alone form: one CTA factors and inverts diagonal block `blk` of A in place.
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, int ld, double* __restrict__ W, int blk, int32_t* info, long long* dbg) {
  extern __shared__ __align__(16) double s[];
  double* sA = s;
  const int tid = threadIdx.x;
  double* Ab = A + (size_t)blk * PB * ld + (size_t)blk * PB;
  double* Wb = W + (size_t)blk * PB * ld + (size_t)blk * PB;
  PT_STAMP(dbg);
#pragma unroll 16
  for (int it = 0; it < PB * PB / POTF2_THREADS; ++it) {
    const int idx = tid + it * POTF2_THREADS;
    const int r = idx >> 7, c = idx & 127;
    sA[r * PLD + c] = (c <= r) ? Ab[(size_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  PT_STAMP(dbg);
  potf2_factor_invert(s, blk * PB, info, dbg);
  potf2_store_W(s, Wb, ld);
  potf2_store_L(s, Ab, ld);
  PT_STAMP(dbg);
}

}  // namespace db
