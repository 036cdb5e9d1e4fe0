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
#include "digit_planes.cuh"

namespace db {

// ------------------------------------------------------------------ plan
struct Range {
  int off, cnt;
};
struct Plan {
  int n = 0, nb = 0;
  int4* d_tiles = nullptr;
  std::vector<Range> tri1, tri2;   // TRTRI levels
  // POTRF schedule.  A trailing update ("rest") applies the panels [p_lo, p_hi) to every tile (i, j) with
  // j >= col_lo; it is issued after panel p_hi-1 and must be complete before the first panel whose column
  // it touches.  Early on a rest covers TWO panels (k = 256: twice the arithmetic intensity of the rank-128
  // update, which is what the trailing GEMM is short of while the matrix is large); once the panel chain is
  // the longer of the two it covers one.  pend_lo[kb] = first panel the panel kernel still has to apply to
  // column kb itself (its update step runs over k in [pend_lo*128, kb*128)).
  struct Rest {
    Range tiles;
    int p_lo, p_hi, col_lo;
  };
  std::vector<Rest> rests;
  std::vector<int> rest_after;   // [nb] index of the rest issued after panel kb, or -1
  std::vector<int> wait_rest;    // [nb] index of the rest panel kb must wait for, or -1
  std::vector<int> pend_lo;      // [nb]
  // Near / far schedule (DISTBO_POTRF_FAR = W > 0): after panel kb the "near" update brings the next W block
  // columns fully up to date (rank-128 steps, or a longer catch-up for a column that just entered the window);
  // the columns far behind the window are updated only every G panels, with k = 128 G in one pass (higher
  // arithmetic intensity, fewer passes over the trailing matrix), on a low-priority stream next to the
  // panel / near chain.  Every tile carries its own k range, so both kinds are plain tile lists.
  // The near launch is split in two: "next" = column kb+2 alone (the only one panel kb+2 waits for: <= 62 tiles, so
  // it is short and never the bound of the panel chain) and "window" = columns kb+4 .. kb+W on a stream of its own
  // (column kb+3 is left to the next step's "next" launch, which catches it up with k = 256).
  struct Step {
    Range next{0, 0}, window{0, 0}, far{0, 0};
    // the far launch on the tcgen05 digit-plane kernel (far_i8): word offset of its work lists in d_far_sched, and
    // the column panel of L it multiplies (rows >= far_row0, columns [far_k0, far_k1))
    int far_i8_off = -1, far_row0 = 0, far_k0 = 0, far_k1 = 0;
    int next_wait_window = -1, next_wait_far = -1;   // earlier launches (step index) on the same tiles
    int window_wait_far = -1;
    double near_work = 0.0, far_work = 0.0;          // tiles x panels
  };
  std::vector<Step> steps;         // [nb] (empty: uniform schedule above)
  std::vector<int> wait_next, wait_window, wait_far;   // [nb] step whose launch panel kb waits for, or -1
  std::vector<cudaEvent_t> ev_next, ev_window, ev_far;
  cudaStream_t near_stream = nullptr, window_stream = nullptr, far_stream = nullptr;
  cudaEvent_t ev_near_end = nullptr, ev_window_end = nullptr, ev_far_end = nullptr;
  // chained panels: even / odd column blocks on two streams, dependencies between consecutive panels through
  // device flags (PanelArgs::done / relay_ready), so potf2(kb+1) starts as soon as the relay product is there
  // instead of after the last consumer of panel kb has left
  bool chain = false;
  cudaStream_t panel_stream2 = nullptr;
  cudaEvent_t ev_end2 = nullptr;
  int* d_done = nullptr;           // [nb * nb] done flags followed by [nb + 1] relay_ready flags
  Range lauum{0, 0};
  cudaStream_t panel_stream = nullptr;
  std::vector<cudaEvent_t> ev_panel, ev_first;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  bool lookahead = true;
  int* d_flags = nullptr;    // [nb][NFLAG] sub-step flags of the panel kernels, then [nb] "second relay CTA has its copy"
  double* d_relay = nullptr; // [2][128*128] relay products (see panel_kernel)
  int sms = 148;
  bool relay_on = true;      // DISTBO_POTRF_RELAY=0 switches the relay CTA off (A/B measurements)
  bool relay2_on = true;     // DISTBO_POTRF_RELAY2=0: the relay product stays on one CTA (A/B measurements)
  bool small_ok = false;     // the whole-matrix cooperative kernel can be used (n_pad <= 2048, all CTAs resident)
  int* d_small_flags = nullptr;   // done [nb*nb] | sub [nb*4] | xf [nb*4]
  // Far updates on the tcgen05 digit-plane kernel (DISTBO_POTRF_FAR_I8, default on with the far schedule): the
  // rank-512 updates of the columns far behind the window are C -= P P^T with P the last four column panels of L;
  // P is cut into int8 digit planes (i8_panel_split_kernel) and the product runs on far_ctas persistent CTAs with
  // a read-modify-write epilogue, at ~2.5x the DMMA rate and on a fixed share of the SMs (the panel and the near
  // updates keep the others).
  bool far_i8 = false;
  int far_ctas = 64;
  int32_t* d_far_sched = nullptr;     // per far step: [far_ctas + 1 offsets, padded to 4][items]
  signed char* d_far_planes = nullptr;   // [I8_NS][n][n] (only the panel columns of the current far step are live)
  double* d_far_rowscale = nullptr;   // [n]
  CUtensorMap far_map_a, far_map_ahi, far_map_b;
};

// ------------------------------------------------------------------ fused panel kernel
// One launch per column block kb, CTA c -> row block i = kb + c:
//   1. acc = A[i,kb] - L[i, pend..kb) L[kb, pend..kb)^T   (the updates of column kb still pending: the last one
//                                                    to three panels; older ones came from the trailing kernels)
//   2. c == 0: the diagonal tile goes to shared memory and is factorised there in four 32-column sub-steps
//              (potf2.cuh); after each sub-step the finished rows of L_kk and the 32x32 diagonal inverse are
//              written to A / W and published through a flag.
//      c  > 0: the tile stays in shared memory; the CTA follows the sub-step flags and solves 32 columns at a
//              time, X_s = (B_s - X_{<s} L_kk[s,<s]^T) Dinv_s^T on DMMA, result -> A[i,kb].  CTA 1 (row block
//              kb+1, the "relay") also accumulates L[kb+1,kb] L[kb+1,kb]^T for the next launch's diagonal tile.
// The chain update -> potf2 -> solve therefore costs one launch, no global round trip of the tile, and the
// solve / next-diagonal update overlap the factorisation.  Waiting CTAs only ever wait for block 0 of their
// own grid (<= 64 CTAs, all resident on 148 SMs).
//
// Every inter-CTA wait is BOUNDED: the producers a CTA waits for are co-resident in practice (lower block
// indices of the same grid, dispatched first, everything fits on the device), but a plain launch does not
// guarantee that.  If a flag does not arrive within SPIN_LIMIT polls (>= 64 ns each, i.e. >= ~0.5 s - five orders
// of magnitude above the longest legitimate wait) the waiter records DISTBO_INFO_TIMEOUT in `info` and goes on;
// the factor is then garbage, and the host raises on the non-zero info exactly as for a failed pivot, so a
// scheduling surprise is an error instead of a hung GPU.
constexpr int SPIN_LIMIT = 1 << 23;
__device__ __noinline__ void spin_until_set(int* f, int32_t* info) {
  int polls = 0;
  while (atomicAdd(f, 0) == 0) {
    __nanosleep(64);
    if (++polls > SPIN_LIMIT) {
      if (info) atomicExch(info, DISTBO_INFO_TIMEOUT);
      break;
    }
  }
  __threadfence();
}
__device__ __forceinline__ void wait_flag(int* f, int32_t* info) {
  if (threadIdx.x == 0) spin_until_set(f, info);
  __syncthreads();
}

struct PanelArgs {
  double* A;
  double* W;
  int ld, kb, nb;
  int pend_lo;   // first panel whose update of column kb is still pending (applied by this kernel)
  int32_t* info;
  int* flags;               // [nb][NFLAG]: sub-step flags of the diagonal factor (potf2.cuh)
  double* relay;            // [2][128*128]: L[kb+1,kb] L[kb+1,kb]^T from the relay CTA of the previous launch
  int use_relay;            // this launch's diagonal tile takes panel kb-1's update from `relay`
  unsigned long long* tl;   // debug timeline, normally null
  // Chained mode (consecutive panels on alternating streams, overlapping in time): what panel kb needs from panel
  // kb-1 arrives through flags instead of the launch boundary.  done[q * nb + i] = L[i, q] is final;
  // relay_ready[kb] = the relay product for diagonal tile kb is in `relay`.  Null: launches are serialised.
  int* done;
  int* relay_ready;
  int relay2;               // 1: the LAST CTA of the grid is the second relay CTA (half 1 of the relay product)
  int* r2_loaded;           // relay2: set by the second relay CTA once it holds its copy of tile (kb+1, kb)
};

constexpr int PANEL_SMEM = (PB * PLD + STAGES * STAGE_ELEMS) * (int)sizeof(double);  // 217088 >= POTF2_SMEM
constexpr int NFLAG = 4;   // per column block: the four sub-step flags of the diagonal factor (potf2.cuh)

// Solve of one off-diagonal 128x128 tile B (in shared memory, sT = smem [128][PLD]) against the diagonal block
// L_kk of block column n0 / 128, following the diagonal CTA's four 32-column sub-steps as they are published
// through sub[0..3]:  X_s = (B_s - X_{<s} L_kk[s,<s]^T) Dinv_s^T  on DMMA, so L[i,kb] is complete a few
// microseconds after the last sub-step; the assembled 128x128 inverse is not needed.  X_s is written to the
// tile in shared memory and to Ab (global, leading dimension ld); if xflags != null, xflags[s] is set once
// column group s is visible device-wide.  RELAY: also accumulates P = X X^T, one rank-32 update per sub-step,
// and stores it to Pout [128][128] (the next launch's diagonal tile takes its update from there).
// Relay product P = X X^T (X = the solved tile of row block kb+1, 128 x 128): the next launch's diagonal tile only
// reads its LOWER triangle, i.e. 10 of the 16 32x32 blocks, and one SM needs 17 us for the full 128^3 product - as
// long as the whole panel chain of round 1.  The 10 blocks are therefore split over TWO CTAs ("halves", 80 8x8
// tiles each; both solve the same row block, only the first stores X):
//   half 0: blocks (0,0) (1,0) (1,1) (2,0) (2,1)        half 1: blocks (2,2) (3,0) (3,1) (3,2) (3,3)
// Warp w of a half owns a strip of 8 tiles in one 8-row tile row (A fragment shared) plus 2 tiles of the half's
// leftover 32x32 block: 12 shared-memory fragment loads per 10 DMMAs and k-step.
struct RelayMap {
  int mr0, mc0;   // main strip: rows mr0..mr0+7, columns mc0 .. mc0+63 (8 tiles)
  int er0, ec0;   // extra pair: rows er0..er0+7, columns ec0 .. ec0+15 (2 tiles)
};
__device__ __forceinline__ RelayMap relay_map(int half, int w) {
  RelayMap m;
  if (half == 0) {
    m.mr0 = 32 + w * 8; m.mc0 = 0; m.er0 = (w >> 1) * 8; m.ec0 = (w & 1) * 16;
  } else {
    m.mr0 = 96 + (w >> 1) * 8; m.mc0 = (w & 1) * 64; m.er0 = 64 + (w >> 1) * 8; m.ec0 = 64 + (w & 1) * 16;
  }
  return m;
}
// pr[0..7] main strip, pr[8..9] extra pair;  P += X[:, c0 .. c0+31] X[:, c0 .. c0+31]^T on the map's tiles
__device__ __forceinline__ void relay_accumulate(double (&pr)[10][2], const RelayMap& m, const double* sT, int c0,
                                                 int lr, int lk) {
#pragma unroll
  for (int k = 0; k < SB; k += 4) {
    const double am = sT[(m.mr0 + lr) * PLD + c0 + k + lk];
    const double ae = sT[(m.er0 + lr) * PLD + c0 + k + lk];
#pragma unroll
    for (int t = 0; t < 8; ++t) dmma884(pr[t][0], pr[t][1], am, sT[(m.mc0 + t * 8 + lr) * PLD + c0 + k + lk]);
#pragma unroll
    for (int t = 0; t < 2; ++t) dmma884(pr[8 + t][0], pr[8 + t][1], ae, sT[(m.ec0 + t * 8 + lr) * PLD + c0 + k + lk]);
  }
}
__device__ __forceinline__ void relay_store(const double (&pr)[10][2], const RelayMap& m, double* Pout, int lr, int lk) {
#pragma unroll
  for (int t = 0; t < 8; ++t)
    *reinterpret_cast<double2*>(Pout + (size_t)(m.mr0 + lr) * PB + m.mc0 + t * 8 + 2 * lk) = make_double2(pr[t][0], pr[t][1]);
#pragma unroll
  for (int t = 0; t < 2; ++t)
    *reinterpret_cast<double2*>(Pout + (size_t)(m.er0 + lr) * PB + m.ec0 + t * 8 + 2 * lk) =
        make_double2(pr[8 + t][0], pr[8 + t][1]);
}

// RELAY: 0 = plain consumer; 1 = one half of the relay product (`half`); 2 = both halves (no second relay CTA).
// store_x: write the solved column groups to Ab (false on the second relay CTA, which only duplicates the solve).
// gate: the tile is solved IN PLACE in global memory, and the second relay CTA reads the unsolved tile when it
// starts - the storing CTA therefore waits for *gate (set by the second CTA once its copy is in shared memory)
// before its first global store.
template <int RELAY>
__device__ __forceinline__ void stepwise_solve(double* smem, const double* A, const double* W, int ld, int n0,
                                               const int* sub, double* Ab, int* xflags, double* Pout,
                                               int32_t* info, int half = 0, bool store_x = true,
                                               int* gate = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  double* sT = smem;
  double* sLr = smem + PB * PLD;            // [32][PLD] rows 32s..32s+31 of L_kk (columns 0..32s-1)
  double* sDi = sLr + SB * PLD;             // [32][DLD] Dinv_s
  const double* Lkk = A + (size_t)n0 * ld + n0;
  const double* Wkk = W + (size_t)n0 * ld + n0;
  const int r0 = warp * 16;                 // this warp's 16 rows of the tile
  double pr[10][2], pr2[10][2];             // dead (and removed) in the instantiations that do not use them
  const RelayMap map = relay_map(RELAY == 2 ? 0 : half, warp), map2 = relay_map(1, warp);
  if constexpr (RELAY != 0) {
#pragma unroll
    for (int t = 0; t < 10; ++t) pr[t][0] = pr[t][1] = pr2[t][0] = pr2[t][1] = 0.0;
  }
  const int sr = tid >> 3, sc = (tid & 7) * 2;   // staging: thread = (row of the 32-row slab, 16-byte column slot)
  for (int sb = 0; sb < 4; ++sb) {
    const int c0 = sb * SB;
    if (tid == 0) spin_until_set(const_cast<int*>(sub) + sb, info);
    __syncthreads();                        // flag seen; previous step's staging buffers are free
    for (int c = sc; c < c0; c += 16)
      *reinterpret_cast<double2*>(&sLr[sr * PLD + c]) =
          __ldcg(reinterpret_cast<const double2*>(Lkk + (size_t)(c0 + sr) * ld + c));
    for (int c = sc; c < SB; c += 16)
      *reinterpret_cast<double2*>(&sDi[sr * DLD + c]) =
          __ldcg(reinterpret_cast<const double2*>(Wkk + (size_t)(c0 + sr) * ld + c0 + c));
    __syncthreads();
    double xa[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const double2 v = *reinterpret_cast<const double2*>(&sT[(r0 + mi * 8 + lr) * PLD + c0 + ni * 8 + 2 * lk]);
        xa[mi][ni][0] = v.x;
        xa[mi][ni][1] = v.y;
      }
    for (int k = 0; k < c0; k += 4) {       // B_s - X_{<s} L_kk[s,<s]^T
      double af[2], bf[4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) af[mi] = -sT[(r0 + mi * 8 + lr) * PLD + k + lk];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = sLr[(ni * 8 + lr) * PLD + k + lk];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(xa[mi][ni][0], xa[mi][ni][1], af[mi], bf[ni]);
    }
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        *reinterpret_cast<double2*>(&sT[(r0 + mi * 8 + lr) * PLD + c0 + ni * 8 + 2 * lk]) =
            make_double2(xa[mi][ni][0], xa[mi][ni][1]);
    __syncwarp();
    double xo[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) xo[mi][ni][0] = xo[mi][ni][1] = 0.0;
#pragma unroll
    for (int k = 0; k < SB; k += 4) {       // (.) Dinv_s^T
      double af[2], bf[4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) af[mi] = sT[(r0 + mi * 8 + lr) * PLD + c0 + k + lk];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = sDi[(ni * 8 + lr) * DLD + k + lk];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(xo[mi][ni][0], xo[mi][ni][1], af[mi], bf[ni]);
    }
    __syncwarp();
    if (gate && sb == 0) wait_flag(gate, info);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = r0 + mi * 8 + lr, c = c0 + ni * 8 + 2 * lk;
        const double2 v = make_double2(xo[mi][ni][0], xo[mi][ni][1]);
        *reinterpret_cast<double2*>(&sT[r * PLD + c]) = v;
        if (store_x) *reinterpret_cast<double2*>(Ab + (size_t)r * ld + c) = v;
      }
    if (xflags) {
      __syncthreads();
      if (tid == 0) {   // cumulative fence after the barrier, then publish column group sb of this tile
        __threadfence();
        atomicExch(xflags + sb, 1);
      }
    }
    if constexpr (RELAY != 0) {
      __syncthreads();                      // columns [c0, c0+32) of the tile are complete in sT
      relay_accumulate(pr, map, sT, c0, lr, lk);
      if constexpr (RELAY == 2) relay_accumulate(pr2, map2, sT, c0, lr, lk);
    }
  }
  if constexpr (RELAY != 0) {
    relay_store(pr, map, Pout, lr, lk);
    if constexpr (RELAY == 2) relay_store(pr2, map2, Pout, lr, lk);
  }
  __syncthreads();                          // sT and the staging buffers may be overwritten
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) panel_kernel(PanelArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  const int kb = p.kb, ld = p.ld, nb = p.nb;
  const int n0 = kb * BN;
  double* sT = smem;  // [128][PLD] tile
  tl_start(p.tl, 2 * kb);

  // CTA 0 owns the diagonal block; CTA c = 1 .. G-1 owns row blocks kb + c, kb + c + (G-1), ...; with relay2 the
  // extra CTA G duplicates the solve of row block kb + 1 and forms the second half of the relay product
  const int n_solve = (int)gridDim.x - 1 - p.relay2;
  const bool second_relay = p.relay2 && (int)blockIdx.x == (int)gridDim.x - 1;
  const int stride = (blockIdx.x == 0 || second_relay) ? nb : n_solve;
  for (int i = second_relay ? kb + 1 : kb + (int)blockIdx.x; i < nb; i += stride) {
    const int m0 = i * BM;
    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
      const int row = m0 + wm * 64 + mi * 8 + lr;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int col = n0 + wn * 32 + ni * 8 + lk * 2;
        const double2 c = *reinterpret_cast<const double2*>(p.A + (size_t)row * ld + col);
        acc[mi][ni][0] = -c.x;   // tile = -(acc) = C - A B after the main loop
        acc[mi][ni][1] = -c.y;
      }
    }
    const bool diag = blockIdx.x == 0;
    // the diagonal tile may take the update of panel kb-1 from the relay product of the previous launch
    const int k_hi = (diag && p.use_relay) ? n0 - BN : n0;
    if (p.done && kb >= 1 && p.pend_lo <= kb - 1 && k_hi == n0) {
      // panel kb-1 may still be running (older panels are complete: same stream): its rows i and kb
      wait_flag(p.done + (kb - 1) * nb + i, p.info);
      if (!diag) wait_flag(p.done + (kb - 1) * nb + kb, p.info);
    }
    if (p.pend_lo * BN < k_hi) gemm_mainloop<KMAJOR, KMAJOR>(acc, p.A, ld, p.A, ld, m0, n0, p.pend_lo * BN, k_hi, smem);
    if (diag && p.use_relay) {
      if (p.relay_ready) wait_flag(p.relay_ready + kb, p.info);
      const double* P = p.relay + (size_t)(kb & 1) * PB * PB;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int r = wm * 64 + mi * 8 + lr;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int c = wn * 32 + ni * 8 + lk * 2;
          const double2 v = __ldcg(reinterpret_cast<const double2*>(P + (size_t)r * PB + c));   // L2: may be fresh
          acc[mi][ni][0] += v.x;   // acc holds -(tile): tile = A - pending - P
          acc[mi][ni][1] += v.y;
        }
      }
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
      const int r = wm * 64 + mi * 8 + lr;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int c = wn * 32 + ni * 8 + lk * 2;
        double2 v = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        if (diag) {  // diagonal tile: potf2 expects zeros above the diagonal
          if (c > r) v.x = 0.0;
          if (c + 1 > r) v.y = 0.0;
        }
        *reinterpret_cast<double2*>(&sT[r * PLD + c]) = v;
      }
    }
    __syncthreads();

    double* Ab = p.A + (size_t)m0 * ld + n0;
    if (diag) {
      double* Wb = p.W + (size_t)m0 * ld + n0;
      long long* dbg = nullptr;
      tl_mark(p.tl, 2 * kb, 2);   // update done, potf2 starts
      // The four sub-steps publish the final rows of L_kk and the 32x32 diagonal inverses as they complete; the
      // assembly of the full 128x128 inverse (needed by TRTRI only) is deferred to diag_inverse_kernel, which
      // does all diagonal blocks in parallel instead of one per step on the critical chain.
      potf2_factor_invert(smem, m0, p.info, dbg, Potf2Publish{Ab, Wb, ld, p.flags + kb * NFLAG}, false);
      tl_mark(p.tl, 2 * kb, 3);   // last sub-step published
      break;
    }

    // ---- off-diagonal row block: stepwise solve (see stepwise_solve); CTA 1's first row block is the relay
    {
      const bool relay = p.relay && blockIdx.x == 1 && i == kb + 1;
      const int* sub = p.flags + kb * NFLAG;
      double* Pout = p.relay + (size_t)((kb + 1) & 1) * PB * PB;
      if (second_relay) {
        if (tid == 0) atomicExch(p.r2_loaded, 1);   // after the barrier above: every thread's part of the tile is in sT
        stepwise_solve<1>(smem, p.A, p.W, ld, n0, sub, Ab, nullptr, Pout, p.info, 1, false);
      } else if (relay && p.relay2) {
        stepwise_solve<1>(smem, p.A, p.W, ld, n0, sub, Ab, nullptr, Pout, p.info, 0, true, p.r2_loaded);
      }
      else if (relay) stepwise_solve<2>(smem, p.A, p.W, ld, n0, sub, Ab, nullptr, Pout, p.info);
      else stepwise_solve<0>(smem, p.A, p.W, ld, n0, sub, Ab, nullptr, nullptr, p.info);
      if (p.done && !second_relay && tid == 0) {   // stepwise_solve ended with a barrier: cumulative fence, then publish
        __threadfence();
        atomicExch(p.done + kb * nb + i, 1);
        if (relay) atomicExch(p.relay_ready + kb + 1, 1);
      }
    }
  }
  tl_end(p.tl, 2 * kb);
}

// ------------------------------------------------------------------ whole-matrix kernel for small problems
// For n_pad <= 16 * 128 the entire factorisation runs in ONE cooperative launch: CTA = one lower tile (r, c),
// all nbt (nbt + 1) / 2 CTAs resident at once (cooperative launch), dependencies through flags in global memory:
//   done[r][c]   tile L[r,c] is final (set by its owner)
//   sub[c][s]    sub-step s of the diagonal factor of column block c is published (potf2.cuh)
//   xf[c][s]     column group s of L[c+1,c] is final (set by the owner of tile (c+1, c))
// Owner of (r, c): accumulates A[r,c] - sum_{c'<c} L[r,c'] L[c,c']^T as the operand tiles become final
// (left-looking, DMMA), then solves against L_cc following the sub-steps (off-diagonal) or factors the tile
// (diagonal).  A diagonal owner takes its LAST update, L[c,c-1] L[c,c-1]^T, 32 columns at a time as its neighbour
// publishes them, so potf2(c) starts a few microseconds after potf2(c-1) ends: no launch boundary, no trailing
// kernel, no relay buffer on the chain.  This is what the reference's shipped configurations need (N <= 1.3 k,
// 10^4 optimiser steps per fit): the step time there is the latency of this chain.
struct SmallArgs {
  double* A;
  double* W;
  int ld, nbt;
  int32_t* info;
  int* done;   // [nbt][nbt]
  int* sub;    // [nbt][4]
  int* xf;     // [nbt][4]
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_small_kernel(SmallArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  const int nbt = p.nbt, ld = p.ld;
  int idx = blockIdx.x, c = 0;            // column-major enumeration: earlier columns get the lower block ids
  while (idx >= nbt - c) {
    idx -= nbt - c;
    ++c;
  }
  const int r = c + idx;
  const int m0 = r * BM, n0 = c * BN;
  const bool diag = r == c;
  double* sT = smem;

  double acc[8][4][2];                    // -(tile): tile = A - sum L L^T
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int row = m0 + wm * 64 + mi * 8 + lr;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int col = n0 + wn * 32 + ni * 8 + lk * 2;
      const double2 v = *reinterpret_cast<const double2*>(p.A + (size_t)row * ld + col);
      acc[mi][ni][0] = -v.x;
      acc[mi][ni][1] = -v.y;
    }
  }
  const int c_hi = (diag && c >= 1) ? c - 1 : c;   // a diagonal owner takes panel c-1 progressively (below)
  for (int cp = 0; cp < c_hi; ++cp) {
    wait_flag(p.done + r * nbt + cp, p.info);
    if (!diag) wait_flag(p.done + c * nbt + cp, p.info);
    gemm_mainloop<KMAJOR, KMAJOR>(acc, p.A, ld, p.A, ld, m0, n0, cp * BN, (cp + 1) * BN, smem);
  }
  if (diag && c >= 1) {
    double* sX = smem;                    // [128][DLD] column group of L[c,c-1]
    const double* Lrow = p.A + (size_t)m0 * ld + (size_t)(c - 1) * BN;
    for (int sb = 0; sb < 4; ++sb) {
      wait_flag(p.xf + (c - 1) * 4 + sb, p.info);
      for (int q = tid; q < PB * SB; q += GEMM_THREADS) {
        const int rr = q >> 5, cc = q & 31;
        sX[rr * DLD + cc] = __ldcg(Lrow + (size_t)rr * ld + sb * SB + cc);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SB; k += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) af[mi] = sX[(wm * 64 + mi * 8 + lr) * DLD + k + lk];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = sX[(wn * 32 + ni * 8 + lr) * DLD + k + lk];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
      __syncthreads();                    // sX is refilled by the next sub-step
    }
  }
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int rr = wm * 64 + mi * 8 + lr;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int cc = wn * 32 + ni * 8 + lk * 2;
      double2 v = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
      if (diag) {
        if (cc > rr) v.x = 0.0;
        if (cc + 1 > rr) v.y = 0.0;
      }
      *reinterpret_cast<double2*>(&sT[rr * PLD + cc]) = v;
    }
  }
  __syncthreads();
  double* Ab = p.A + (size_t)m0 * ld + n0;
  if (diag) {
    double* Wb = p.W + (size_t)m0 * ld + n0;
    long long* dbg = nullptr;
    potf2_factor_invert(smem, m0, p.info, dbg, Potf2Publish{Ab, Wb, ld, p.sub + c * 4}, false);
    return;
  }
  stepwise_solve<0>(smem, p.A, p.W, ld, n0, p.sub + c * 4, Ab, (r == c + 1) ? p.xf + c * 4 : nullptr, nullptr, p.info);
  if (tid == 0) {   // stepwise_solve ended with a barrier: every store of the tile is ordered before this fence
    __threadfence();
    atomicExch(p.done + r * nbt + c, 1);
  }
}

// inv(L_kk) for every diagonal block at once (one CTA each): L_kk and its four 32x32 diagonal inverses, left in
// A / W by the panel kernels, are brought into shared memory, the off-diagonal 32-blocks of the inverse are
// assembled by block forward substitution (potf2.cuh) and the 128x128 block is written to W.
__global__ void __launch_bounds__(POTF2_THREADS, 1) diag_inverse_kernel(const double* __restrict__ A,
                                                                        double* __restrict__ W, int ld) {
  extern __shared__ __align__(16) double s[];
  double* sA = s;
  double* sDinv = s + PB * PLD;
  const int tid = threadIdx.x;
  const size_t off = (size_t)blockIdx.x * PB * ld + (size_t)blockIdx.x * PB;
  const double* Ab = A + off;
  double* Wb = W + off;
  for (int idx = tid; idx < PB * PB; idx += POTF2_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    sA[r * PLD + c] = (c <= r) ? Ab[(size_t)r * ld + c] : 0.0;
    if ((r >> 5) == (c >> 5)) sDinv[(r >> 5) * SB * DLD + (r & 31) * DLD + (c & 31)] = Wb[(size_t)r * ld + c];
  }
  __syncthreads();
  long long* dbg = nullptr;
  potf2_assemble_inverse(s, dbg);
  potf2_store_W(s, Wb, ld);
}

// Number of CTAs of the panel kernel for column block kb: as few as keep the panel (update -> potf2 ->
// solve chain over its row blocks) shorter than the concurrent trailing update, which then gets the
// remaining SMs; all row blocks in parallel once the trailing update is the shorter of the two.
static int panel_ctas(int nb, int kb, int sms, int rest_tiles_prev, int depth, int share) {
  // depth = pending panels the kernel applies itself; share = panels that run next to the same trailing update
  const int rows = nb - kb;            // row blocks incl. the diagonal
  if (rows <= 2) return rows;
  const double t_tile = 21.0, t_upd = 6.0 + 17.0 * depth, t_potf2 = 52.0, t_solve = 18.0;   // microseconds, measured
  for (int G = 2; G < rows; ++G) {
    const int per = (rows - 1 + G - 2) / (G - 1);
    const double panel = t_upd + t_potf2 + t_solve + (per - 1) * (t_upd + t_solve);
    const double rest = rest_tiles_prev * t_tile / (double)(sms - G);
    if (panel * share <= 0.85 * rest) return G;
  }
  return rows;
}

static bool longer_k(const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); }

}  // namespace db

using namespace db;

long long g_db_launches = 0;
static unsigned long long* g_timeline = nullptr;

// Debug aid (tools/potrf_timeline.py): device buffer of 3*nb*4 uint64 (init: slot 0 = ~0, others 0) that the
// POTRF kernels fill with globaltimer stamps; pass NULL to switch it off.  Not part of the product path.
extern "C" int distbo_debug_timeline(void* buf) {
  g_timeline = (unsigned long long*)buf;
  return DISTBO_OK;
}

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
  std::vector<int32_t> far_sched;   // work lists of the far updates on the digit-plane kernel
  auto T = [](int i) { return i * 128; };
  {
    // number of double-depth steps (env DISTBO_POTRF_DOUBLE).  Measured on B200 at N = 8192 (profiles/
    // r01_potrf_schedule.txt): the k = 256 update is only ~14 % faster per flop than two k = 128 ones while the
    // panels next to it get slower (longer pending updates), so the default is 0 = single depth throughout.
    int js = 0;
    const char* env = getenv("DISTBO_POTRF_DOUBLE");
    if (env) js = atoi(env);
    if (js < 0) js = 0;
    if (2 * js + 2 > nb) js = (nb - 2) / 2 > 0 ? (nb - 2) / 2 : 0;
    p->rest_after.assign(nb, -1);
    p->wait_rest.assign(nb, -1);
    p->pend_lo.assign(nb, 0);
    auto add_rest = [&](int p_lo, int p_hi, int col_lo) {
      Plan::Rest r;
      r.p_lo = p_lo; r.p_hi = p_hi; r.col_lo = col_lo;
      r.tiles.off = (int)tiles.size();
      for (int i = col_lo; i < nb; ++i)
        for (int j = col_lo; j <= i; ++j) tiles.push_back(make_int4(T(i), T(j), T(p_lo), T(p_hi)));
      r.tiles.cnt = (int)tiles.size() - r.tiles.off;
      if (r.tiles.cnt == 0) return;
      p->rest_after[p_hi - 1] = (int)p->rests.size();
      p->rests.push_back(r);
    };
    for (int J = 0; J < js; ++J) add_rest(2 * J, 2 * J + 2, 2 * J + 4);
    for (int kb = 2 * js; kb < nb; ++kb) add_rest(kb, kb + 1, kb + 2);
    // per column: which panels are already applied by rests, and the last rest touching it
    for (int c = 0; c < nb; ++c) {
      int applied = 0, last = -1;
      for (size_t q = 0; q < p->rests.size(); ++q)
        if (p->rests[q].col_lo <= c && p->rests[q].p_hi <= c) {
          applied = p->rests[q].p_hi;   // rests are contiguous in panel order
          last = (int)q;
        }
      p->pend_lo[c] = applied;
      p->wait_rest[c] = last;
    }
  }
  {
    // Measured on B200 (profiles/r01_potrf_schedule.txt): Gram + POTRF as a CUDA graph at N = 8192 8.97 -> 8.31 ms
    // with W = 6, G = 4 (W = 4..8 within 1 %, W = 12 or G = 2 slower); N <= 4096 is faster with the uniform
    // schedule, whose panel chain is the bound there.  Default: on from 56 column blocks (N >= 7168);
    // DISTBO_POTRF_FAR=0 switches it off.
    // With the far updates on the tcgen05 digit-plane kernel (below) the schedule already pays from 40 column blocks
    // (N = 5120: 2.86 -> 2.70 ms, 6144: 4.24 -> 3.71 ms; at 4096 the uniform schedule stays faster: 1.91 vs 2.18 ms).
    const bool far_i8_wanted = !(getenv("DISTBO_POTRF_FAR_I8") && atoi(getenv("DISTBO_POTRF_FAR_I8")) == 0) &&
                               n_pad / I8_RB <= I8_MAX_RB;
    // (window of 5 instead of 6 column blocks once the far updates are cheap: 6.35 -> 6.20 ms at N = 8192)
    int W = nb >= (far_i8_wanted ? 40 : 56) ? (far_i8_wanted ? 5 : 6) : 0, G = 4;
    if (const char* env = getenv("DISTBO_POTRF_FAR")) W = atoi(env);
    if (const char* env = getenv("DISTBO_POTRF_FARG")) G = atoi(env);
    if (W >= 2 && G >= 2 && nb >= 3 * G + W) {
      {
        const char* env = getenv("DISTBO_POTRF_FAR_I8");
        p->far_i8 = !(env && atoi(env) == 0) && n_pad / I8_RB <= I8_MAX_RB;
        if (const char* c = getenv("DISTBO_POTRF_FAR_I8_CTAS")) p->far_ctas = std::max(8, std::min(148, atoi(c)));
      }
      std::vector<int> applied(nb, 0), last_next(nb, -1), last_window(nb, -1), last_far(nb, -1);
      p->steps.assign(nb, Plan::Step());
      p->wait_next.assign(nb, -1);
      p->wait_window.assign(nb, -1);
      p->wait_far.assign(nb, -1);
      auto column_tiles = [&](std::vector<int4>& out, int c, int upto) {
        for (int i = c; i < nb; ++i) out.push_back(make_int4(T(i), T(c), T(applied[c]), T(upto)));
        return (double)(nb - c) * (upto - applied[c]);
      };
      // Tail: once the panel chain is the bound (the rank-128 update of everything that is left takes less time
      // than one panel), nothing is deferred any more - from step `tail` on the window covers every remaining
      // column (W = infinity, no far launches); its first launch there catches the far columns up.
      // Measured at N = 8192 (nb = 64, eager / graph ms): no tail 8.02 / 8.22, tail from 20: 7.87 / 7.99, 24: 7.71 / 7.84,
      // 28: 7.63 / 7.73, 32: 7.56 / 7.74, 36: 7.66 / 7.88.
      int tail = nb - 34;
      if (const char* env = getenv("DISTBO_POTRF_TAIL")) tail = atoi(env);
      for (int kb = 0; kb < nb; ++kb) {
        const int Wk = (tail >= 0 && kb >= tail) ? nb : W;
        p->pend_lo[kb] = applied[kb];          // the panel kernel applies [applied, kb) itself
        p->wait_next[kb] = last_next[kb];
        p->wait_window[kb] = last_window[kb];
        p->wait_far[kb] = last_far[kb];
        Plan::Step& st = p->steps[kb];
        if (kb + 2 < nb) {
          const int c = kb + 2;
          std::vector<int4> nt;
          st.near_work += column_tiles(nt, c, kb + 1);
          applied[c] = kb + 1;
          st.next_wait_window = last_window[c];
          st.next_wait_far = last_far[c];
          last_next[c] = kb;
          st.next = Range{(int)tiles.size(), (int)nt.size()};
          tiles.insert(tiles.end(), nt.begin(), nt.end());
        }
        {
          std::vector<int4> wt;
          for (int c = kb + 4; c < nb && c <= kb + Wk; ++c) {
            st.near_work += column_tiles(wt, c, kb + 1);
            applied[c] = kb + 1;
            st.window_wait_far = std::max(st.window_wait_far, last_far[c]);
            last_window[c] = kb;
          }
          std::stable_sort(wt.begin(), wt.end(), longer_k);
          st.window = Range{(int)tiles.size(), (int)wt.size()};
          tiles.insert(tiles.end(), wt.begin(), wt.end());
        }
        if ((kb + 1) % G == 0) {
          // columns that no near launch of the next G steps reaches
          std::vector<int4> ft;
          for (int c = kb + 1 + G + Wk; c < nb; ++c) {
            if (applied[c] >= kb + 1) continue;
            st.far_work += column_tiles(ft, c, kb + 1);
            applied[c] = kb + 1;
            last_far[c] = kb;
          }
          st.far = Range{(int)tiles.size(), (int)ft.size()};
          tiles.insert(tiles.end(), ft.begin(), ft.end());
          if (p->far_i8 && !ft.empty()) {
            // the same tiles as work items of the digit-plane kernel: (128-row tile, 64-row block, k-chunks of 64)
            std::vector<int4> items;
            int row0 = 1 << 30, k0 = 1 << 30, k1 = 0;
            for (const int4& t : ft) {
              items.push_back(make_int4(t.x, t.y, t.z / I8_KC, t.w / I8_KC));
              items.push_back(make_int4(t.x, t.y + I8_RB, t.z / I8_KC, t.w / I8_KC));
              row0 = std::min(row0, t.y);
              k0 = std::min(k0, t.z);
              k1 = std::max(k1, t.w);
            }
            std::stable_sort(items.begin(), items.end(),
                             [](const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); });
            std::vector<double> load(p->far_ctas, 0.0);
            std::vector<std::vector<int4>> lists(p->far_ctas);
            for (const int4& it : items) {
              int best = 0;
              for (int c = 1; c < p->far_ctas; ++c)
                if (load[c] < load[best]) best = c;
              load[best] += (double)(it.w - it.z) + 3.0;
              lists[best].push_back(it);
            }
            st.far_i8_off = (int)far_sched.size();
            st.far_row0 = row0; st.far_k0 = k0; st.far_k1 = k1;
            const size_t offw = ((size_t)p->far_ctas + 1 + 3) / 4 * 4;
            far_sched.resize(far_sched.size() + offw, 0);
            int run = 0;
            for (int c = 0; c < p->far_ctas; ++c) {
              far_sched[st.far_i8_off + c] = run;
              run += (int)lists[c].size();
            }
            far_sched[st.far_i8_off + p->far_ctas] = run;
            for (int c = 0; c < p->far_ctas; ++c)
              for (const int4& it : lists[c]) {
                far_sched.push_back(it.x); far_sched.push_back(it.y); far_sched.push_back(it.z); far_sched.push_back(it.w);
              }
          }
        }
      }
    }
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
  if (p->far_i8 && !far_sched.empty()) {
    const size_t planes = (size_t)I8_NS * n_pad * n_pad;
    bool ok = cudaMalloc(&p->d_far_sched, far_sched.size() * sizeof(int32_t)) == cudaSuccess &&
              cudaMemcpy(p->d_far_sched, far_sched.data(), far_sched.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc(&p->d_far_planes, planes) == cudaSuccess &&
              cudaMalloc(&p->d_far_rowscale, (size_t)n_pad * sizeof(double)) == cudaSuccess &&
              cudaMemset(p->d_far_planes, 0, planes) == cudaSuccess &&
              tma::encode_u8_planes(&p->far_map_a, p->d_far_planes, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE,
                                    I8_A_LO) &&
              tma::encode_u8_planes(&p->far_map_ahi, p->d_far_planes, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS,
                                    I8_TILE, I8_NS - I8_A_LO) &&
              tma::encode_u8_planes(&p->far_map_b, p->d_far_planes, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB,
                                    I8_NS) &&
              cudaFuncSetAttribute(score_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM) ==
                  cudaSuccess;
    if (!ok) {   // not fatal: the far updates stay on the DMMA tile GEMM
      cudaGetLastError();
      p->far_i8 = false;
    }
  } else {
    p->far_i8 = false;
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&p->panel_stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
    cudaFree(p->d_tiles);
    delete p;
    return DISTBO_ECUDA;
  }
  if (!p->steps.empty()) {
    const int mid = (lo + hi) / 2;       // priorities: panel (hi) > near (mid) > far (lo)
    const int wprio = mid + (lo > hi ? 1 : -1) * (mid != lo ? 1 : 0);   // one step towards the low end
    if (cudaStreamCreateWithPriority(&p->near_stream, cudaStreamNonBlocking, mid) != cudaSuccess ||
        cudaStreamCreateWithPriority(&p->window_stream, cudaStreamNonBlocking, wprio) != cudaSuccess ||
        cudaStreamCreateWithPriority(&p->far_stream, cudaStreamNonBlocking, lo) != cudaSuccess) {
      cudaFree(p->d_tiles);
      delete p;
      return DISTBO_ECUDA;
    }
    p->ev_next.resize(nb);
    p->ev_window.resize(nb);
    p->ev_far.resize(nb);
    for (int i = 0; i < nb; ++i) {
      cudaEventCreateWithFlags(&p->ev_next[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&p->ev_window[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&p->ev_far[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&p->ev_near_end, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->ev_window_end, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->ev_far_end, cudaEventDisableTiming);
    // measured: no gain at N = 8192 (8.31 ms either way, the panels' stream-level waits decide when they start),
    // 2-3 % at N = 3000-4096 where this schedule is not the default; off unless DISTBO_POTRF_CHAIN=1
    const char* ch = getenv("DISTBO_POTRF_CHAIN");
    p->chain = ch ? ch[0] != '0' : false;
    if (p->chain) {
      if (cudaStreamCreateWithPriority(&p->panel_stream2, cudaStreamNonBlocking, hi) != cudaSuccess ||
          cudaMalloc(&p->d_done, (size_t)(nb * nb + nb + 1) * sizeof(int)) != cudaSuccess)
        p->chain = false;
      else
        cudaEventCreateWithFlags(&p->ev_end2, cudaEventDisableTiming);
    }
  }
  p->ev_panel.resize(nb);
  p->ev_first.resize(p->rests.size());   // "rest q is complete"
  for (int i = 0; i < nb; ++i) cudaEventCreateWithFlags(&p->ev_panel[i], cudaEventDisableTiming);
  for (size_t i = 0; i < p->ev_first.size(); ++i) cudaEventCreateWithFlags(&p->ev_first[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_begin, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_end, cudaEventDisableTiming);
  const char* la = getenv("DISTBO_LOOKAHEAD");
  p->lookahead = !(la && la[0] == '0');
  const char* rl = getenv("DISTBO_POTRF_RELAY");
  p->relay_on = !(rl && rl[0] == '0');
  const char* rl2 = getenv("DISTBO_POTRF_RELAY2");
  p->relay2_on = !(rl2 && rl2[0] == '0');
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (cudaMalloc(&p->d_flags, (size_t)nb * (NFLAG + 1) * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&p->d_relay, (size_t)2 * PB * PB * sizeof(double)) != cudaSuccess) {
    cudaFree(p->d_tiles);
    if (p->d_flags) cudaFree(p->d_flags);
    delete p;
    return DISTBO_ECUDA;
  }
  cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
  cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
  cudaFuncSetAttribute(diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
  {
    const char* sm = getenv("DISTBO_POTRF_SMALL");
    int coop = 0, per_sm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaFuncSetAttribute(potrf_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, potrf_small_kernel, GEMM_THREADS, PANEL_SMEM);
    const int ctas = nb * (nb + 1) / 2;
    p->small_ok = !(sm && sm[0] == '0') && coop && nb <= 16 && ctas <= per_sm * p->sms;
    if (p->small_ok && cudaMalloc(&p->d_small_flags, (size_t)(nb * nb + 8 * nb) * sizeof(int)) != cudaSuccess)
      p->small_ok = false;
  }
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
  for (auto e : p->ev_next) cudaEventDestroy(e);
  for (auto e : p->ev_window) cudaEventDestroy(e);
  for (auto e : p->ev_far) cudaEventDestroy(e);
  if (p->ev_near_end) cudaEventDestroy(p->ev_near_end);
  if (p->ev_window_end) cudaEventDestroy(p->ev_window_end);
  if (p->window_stream) cudaStreamDestroy(p->window_stream);
  if (p->ev_far_end) cudaEventDestroy(p->ev_far_end);
  if (p->panel_stream2) cudaStreamDestroy(p->panel_stream2);
  if (p->ev_end2) cudaEventDestroy(p->ev_end2);
  if (p->d_done) cudaFree(p->d_done);
  if (p->near_stream) cudaStreamDestroy(p->near_stream);
  if (p->far_stream) cudaStreamDestroy(p->far_stream);
  if (p->panel_stream) cudaStreamDestroy(p->panel_stream);
  if (p->d_tiles) cudaFree(p->d_tiles);
  if (p->d_flags) cudaFree(p->d_flags);
  if (p->d_relay) cudaFree(p->d_relay);
  if (p->d_small_flags) cudaFree(p->d_small_flags);
  if (p->d_far_sched) cudaFree(p->d_far_sched);
  if (p->d_far_planes) cudaFree(p->d_far_planes);
  if (p->d_far_rowscale) cudaFree(p->d_far_rowscale);
  delete p;
  return DISTBO_OK;
}

extern "C" int distbo_potrf_f64(void* plan, double* A, double* W, int32_t* info, void* stream) {
  if (!plan || !A || !W || !info) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  cudaStream_t U = (cudaStream_t)stream;
  const int n = p->n, nb = p->nb;
  DB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), U));

  if (p->small_ok) {
    DB_CUDA(cudaMemsetAsync(p->d_small_flags, 0, (size_t)(nb * nb + 8 * nb) * sizeof(int), U));
    SmallArgs sa{A, W, n, nb, info, p->d_small_flags, p->d_small_flags + nb * nb, p->d_small_flags + nb * nb + 4 * nb};
    // Plain launch: a CTA only ever waits for CTAs with a LOWER block index (earlier column, or the diagonal of
    // its own column), blocks are dispatched in index order and all nb (nb + 1) / 2 <= 136 of them fit on an idle
    // device at once.  CUDA does not promise either for a non-cooperative launch (the cooperative API does, but made
    // CUDA-graph replays of the training step ~1 ms slower per launch), so the waits are bounded (spin_until_set):
    // if the device is shared and a producer is not resident, info = DISTBO_INFO_TIMEOUT and the host raises.
    potrf_small_kernel<<<nb * (nb + 1) / 2, GEMM_THREADS, PANEL_SMEM, U>>>(sa);
    DB_CHECK_LAUNCH();
    return DISTBO_OK;
  }

  if (!p->steps.empty() && p->lookahead) {
    DB_CUDA(cudaMemsetAsync(p->d_flags, 0, (size_t)nb * (NFLAG + 1) * sizeof(int), U));
    GemmArgs upd{};   // C -= L[i, k range) L[j, k range)^T, k range per tile
    upd.A = A; upd.B = A; upd.C = A; upd.Cin = A;
    upd.lda = upd.ldb = upd.ldc = upd.ldcin = n;
    upd.alpha = -1.0; upd.beta = 1.0;
    upd.tl = g_timeline;
    cudaStream_t Pst[2] = {p->panel_stream, p->chain ? p->panel_stream2 : p->panel_stream};
    cudaStream_t Nn = p->near_stream, Nw = p->window_stream, F = p->far_stream;
    int* done = p->chain ? p->d_done : nullptr;
    int* relay_ready = p->chain ? p->d_done + nb * nb : nullptr;
    if (p->chain) DB_CUDA(cudaMemsetAsync(p->d_done, 0, (size_t)(nb * nb + nb + 1) * sizeof(int), U));
    DB_CUDA(cudaEventRecord(p->ev_begin, U));
    DB_CUDA(cudaStreamWaitEvent(Pst[0], p->ev_begin, 0));
    if (p->chain) DB_CUDA(cudaStreamWaitEvent(Pst[1], p->ev_begin, 0));
    DB_CUDA(cudaStreamWaitEvent(Nn, p->ev_begin, 0));
    DB_CUDA(cudaStreamWaitEvent(Nw, p->ev_begin, 0));
    DB_CUDA(cudaStreamWaitEvent(F, p->ev_begin, 0));
    double running = 0.0, far_share = 0.0;   // update work (tiles x panels) that runs next to the panel
    int g_prev = 0;
    for (int kb = 0; kb < nb; ++kb) {
      const int g_this = panel_ctas(nb, kb, p->sms, (int)(running + far_share), kb - p->pend_lo[kb], 1);
      // While the trailing updates are the bound, a panel runs on few SMs for a long time and its successor would
      // only spin next to it: consecutive panels overlap (alternate streams) once every row block has its own CTA,
      // i.e. when the panel chain is the bound.  Panel kb-2 is complete either way (same stream, or through the
      // "next" launch of step kb-2 that panel kb waits for).
      cudaStream_t P = (g_this == nb - kb) ? Pst[kb & 1] : Pst[0];
      if (p->wait_next[kb] >= 0) DB_CUDA(cudaStreamWaitEvent(P, p->ev_next[p->wait_next[kb]], 0));
      if (p->wait_window[kb] >= 0) DB_CUDA(cudaStreamWaitEvent(P, p->ev_window[p->wait_window[kb]], 0));
      if (p->wait_far[kb] >= 0) DB_CUDA(cudaStreamWaitEvent(P, p->ev_far[p->wait_far[kb]], 0));
      const int use_relay = (p->relay_on && kb >= 1 && g_prev >= 2 && p->pend_lo[kb] <= kb - 1) ? 1 : 0;
      // second relay CTA: once the panel chain is the bound (every row block has its own CTA), launches serialised
      const int relay2 = (p->relay_on && p->relay2_on && !p->chain && g_this >= 2 && g_this == nb - kb) ? 1 : 0;
      PanelArgs pa{A, W, n, kb, nb, p->pend_lo[kb], info, p->d_flags, p->relay_on ? p->d_relay : nullptr, use_relay,
                   g_timeline, done, relay_ready, relay2, p->d_flags + nb * NFLAG + kb};
      g_prev = g_this;
      panel_kernel<<<g_this + relay2, GEMM_THREADS, PANEL_SMEM, P>>>(pa);
      DB_CHECK_LAUNCH();
      const Plan::Step& st = p->steps[kb];
      if (st.next.cnt > 0 || st.window.cnt > 0 || st.far.cnt > 0) DB_CUDA(cudaEventRecord(p->ev_panel[kb], P));
      if (st.next.cnt > 0) {
        DB_CUDA(cudaStreamWaitEvent(Nn, p->ev_panel[kb], 0));
        if (st.next_wait_window >= 0) DB_CUDA(cudaStreamWaitEvent(Nn, p->ev_window[st.next_wait_window], 0));
        if (st.next_wait_far >= 0) DB_CUDA(cudaStreamWaitEvent(Nn, p->ev_far[st.next_wait_far], 0));
        upd.tiles = p->d_tiles + st.next.off;
        upd.tl_id = 2 * kb + 1;
        DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(upd, st.next.cnt, Nn)));
        DB_CUDA(cudaEventRecord(p->ev_next[kb], Nn));
      }
      if (st.window.cnt > 0) {
        DB_CUDA(cudaStreamWaitEvent(Nw, p->ev_panel[kb], 0));
        if (st.window_wait_far >= 0) DB_CUDA(cudaStreamWaitEvent(Nw, p->ev_far[st.window_wait_far], 0));
        upd.tiles = p->d_tiles + st.window.off;
        upd.tl_id = 2 * kb + 1;
        DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(upd, st.window.cnt, Nw)));
        DB_CUDA(cudaEventRecord(p->ev_window[kb], Nw));
      }
      running = st.near_work;
      static const bool skip_far = getenv("DISTBO_POTRF_SKIPFAR") != nullptr;   // timing experiment only (wrong L)
      if (st.far.cnt > 0 && skip_far) {
        DB_CUDA(cudaStreamWaitEvent(F, p->ev_panel[kb], 0));
        DB_CUDA(cudaEventRecord(p->ev_far[kb], F));
      } else if (st.far.cnt > 0 && p->far_i8 && st.far_i8_off >= 0) {
        DB_CUDA(cudaStreamWaitEvent(F, p->ev_panel[kb], 0));
        i8_panel_split_kernel<<<n - st.far_row0, 128, 0, F>>>(A, n, st.far_row0, st.far_k0, st.far_k1, p->d_far_planes,
                                                            p->d_far_rowscale);
        DB_CHECK_LAUNCH();
        I8Args fa;
        fa.n_rb = n / I8_RB; fa.ntiles = 0; fa.slices = 1; fa.tiles_in_flight = 1; fa.debug = 0;
        fa.rowscale = p->d_far_rowscale; fa.rowscale_a = p->d_far_rowscale; fa.kexp = nullptr; fa.V = nullptr;
        fa.partial = nullptr; fa.meet = nullptr; fa.meet_stride = 0; fa.meet_waves = 0;
        fa.item_off = p->d_far_sched + st.far_i8_off;
        fa.items = reinterpret_cast<const int4*>(p->d_far_sched + st.far_i8_off + ((size_t)p->far_ctas + 1 + 3) / 4 * 4);
        fa.C = A; fa.ldc = n; fa.alpha = -1.0; fa.beta = 1;
        score_i8_kernel<1><<<p->far_ctas, I8_THREADS, I8_SMEM, F>>>(p->far_map_a, p->far_map_ahi, p->far_map_b, fa);
        DB_CHECK_LAUNCH();
        DB_CUDA(cudaEventRecord(p->ev_far[kb], F));
        static const double far_div = getenv("DISTBO_POTRF_FAR_I8_DIV") ? atof(getenv("DISTBO_POTRF_FAR_I8_DIV")) : 8.0;
        far_share = st.far_work / far_div;
      } else if (st.far.cnt > 0) {
        DB_CUDA(cudaStreamWaitEvent(F, p->ev_panel[kb], 0));
        upd.tiles = p->d_tiles + st.far.off;
        upd.tl_id = 2 * nb + kb;          // far launches have their own timeline rows (buffer: 3 nb rows)
        DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(upd, st.far.cnt, F)));
        DB_CUDA(cudaEventRecord(p->ev_far[kb], F));
        far_share = st.far_work / 4.0;
      }
    }
    DB_CUDA(cudaEventRecord(p->ev_end, Pst[0]));
    if (p->chain) {
      DB_CUDA(cudaEventRecord(p->ev_end2, Pst[1]));
      DB_CUDA(cudaStreamWaitEvent(U, p->ev_end2, 0));
    }
    DB_CUDA(cudaEventRecord(p->ev_near_end, Nn));
    DB_CUDA(cudaEventRecord(p->ev_window_end, Nw));
    DB_CUDA(cudaEventRecord(p->ev_far_end, F));
    DB_CUDA(cudaStreamWaitEvent(U, p->ev_end, 0));
    DB_CUDA(cudaStreamWaitEvent(U, p->ev_near_end, 0));
    DB_CUDA(cudaStreamWaitEvent(U, p->ev_window_end, 0));
    DB_CUDA(cudaStreamWaitEvent(U, p->ev_far_end, 0));
    return DISTBO_OK;
  }

  {
    DB_CUDA(cudaMemsetAsync(p->d_flags, 0, (size_t)nb * (NFLAG + 1) * sizeof(int), U));
    GemmArgs rest{};  // C -= P P^T on the tiles (i, j), j >= kb + 2
    rest.A = A; rest.B = A; rest.C = A; rest.Cin = A;
    rest.lda = rest.ldb = rest.ldc = rest.ldcin = n;
    rest.alpha = -1.0; rest.beta = 1.0;
    const bool fork = p->lookahead && nb > 2;
    cudaStream_t P = fork ? p->panel_stream : U;
    if (fork) {
      DB_CUDA(cudaEventRecord(p->ev_begin, U));
      DB_CUDA(cudaStreamWaitEvent(P, p->ev_begin, 0));
    }
    int running_rest_tiles = 0;   // tiles of the trailing update that runs next to the panel being launched
    int running_share = 1;        // panels that run next to that same update
    int g_prev = 0;
    for (int kb = 0; kb < nb; ++kb) {
      // column kb must have received every trailing update that touches it
      if (fork && p->wait_rest[kb] >= 0) DB_CUDA(cudaStreamWaitEvent(P, p->ev_first[p->wait_rest[kb]], 0));
      const int g_this = fork ? panel_ctas(nb, kb, p->sms, running_rest_tiles, kb - p->pend_lo[kb], running_share) : nb - kb;
      // the previous launch had a relay CTA (row block kb) whenever it had more than one CTA
      const int use_relay = (p->relay_on && kb >= 1 && g_prev >= 2 && p->pend_lo[kb] <= kb - 1) ? 1 : 0;
      const int relay2 = (p->relay_on && p->relay2_on && g_this >= 2 && g_this == nb - kb) ? 1 : 0;
      PanelArgs pa{A, W, n, kb, nb, p->pend_lo[kb], info, p->d_flags, p->relay_on ? p->d_relay : nullptr, use_relay,
                   g_timeline, nullptr, nullptr, relay2, p->d_flags + nb * NFLAG + kb};
      g_prev = g_this;
      panel_kernel<<<g_this + relay2, GEMM_THREADS, PANEL_SMEM, P>>>(pa);
      DB_CHECK_LAUNCH();
      const int q = p->rest_after[kb];
      if (q >= 0) {
        const Plan::Rest& r = p->rests[q];
        if (fork) {
          DB_CUDA(cudaEventRecord(p->ev_panel[kb], P));
          DB_CUDA(cudaStreamWaitEvent(U, p->ev_panel[kb], 0));
        }
        rest.tiles = p->d_tiles + r.tiles.off;
        rest.tl = g_timeline;
        rest.tl_id = 2 * kb + 1;
        DB_CUDA((launch_gemm_tiles<KMAJOR, KMAJOR>(rest, r.tiles.cnt, U)));
        if (fork) DB_CUDA(cudaEventRecord(p->ev_first[q], U));
        running_rest_tiles = r.tiles.cnt * (r.p_hi - r.p_lo);
        running_share = r.p_hi - r.p_lo;
      }
    }
    if (fork) {
      DB_CUDA(cudaEventRecord(p->ev_end, P));
      DB_CUDA(cudaStreamWaitEvent(U, p->ev_end, 0));
    }
    return DISTBO_OK;
  }

  return DISTBO_EARG;  // unreachable: the fused panel path above is the only one
}

// Levels [level_begin, level_end) of the triangular inverse: level 0 = the 128-row diagonal blocks, level h >= 1 = the
// merges at height h of the halving tree (two DMMA tile-GEMM launches each).  distbo_trtri_f64 = all of them.
extern "C" int distbo_trtri_levels(void* plan) {
  if (!plan) return DISTBO_EARG;
  return (int)((Plan*)plan)->tri1.size() + 1;
}

extern "C" int distbo_trtri_range_f64(void* plan, const double* L, double* W, double* T, int level_begin,
                                      int level_end, void* stream) {
  if (!plan || !L || !W || !T || level_begin < 0) return DISTBO_EARG;
  Plan* p = (Plan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = p->n;
  if (level_begin == 0 && level_end > 0) {
    diag_inverse_kernel<<<p->nb, POTF2_THREADS, POTF2_SMEM, st>>>(L, W, n);
    DB_CHECK_LAUNCH();
  }
  for (int h = std::max(level_begin, 1); h < level_end && h <= (int)p->tri1.size(); ++h) {
    const size_t l = (size_t)h - 1;
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

extern "C" int distbo_trtri_f64(void* plan, const double* L, double* W, double* T, void* stream) {
  if (!plan) return DISTBO_EARG;
  return distbo_trtri_range_f64(plan, L, W, T, 0, (int)((Plan*)plan)->tri1.size() + 1, stream);
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
  // 32 columns per CTA, 8 row groups striding the row blocks; fixed-order combine through shared memory
  __shared__ double red[8][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  const int nblk = n_pad / GVT_ROWS;
  const int b0 = (j / 128) * (128 / GVT_ROWS);
  double s = 0.0;
  for (int b = b0 + g; b < nblk; b += 8) s += partial[(size_t)b * n_pad + j];
  red[g][c] = s;
  __syncthreads();
  if (g == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][c];
    alpha[j] = t;
  }
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
  gemvT_reduce_kernel<<<n_pad / 32, 256, 0, st>>>(workspace, n_pad, alpha);
  DB_CHECK_LAUNCH();
  nll_kernel<<<1, 1024, 0, st>>>(L, n_pad, N, V, nll);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}
