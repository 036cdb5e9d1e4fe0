// The fp64 contractions of this library on the tcgen05 tensor cores: operands as int8 digit planes, exact s32
// accumulation in tensor memory, one fold to fp64.  Shared by the scoring path (score.cu) and the
// inverse products of the fit (inverse_i8.cu): constants, kernel arguments, the split kernels and the
// warp-specialised TMA / tcgen05 contraction kernel.  Kernels have internal linkage (one copy per translation unit).
#pragma once
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace db {

// =========================================================================================================
// tcgen05 path: the same contraction A = K_* W^T on the 5th-generation tensor cores, exact in integers.
//
// fp64 has no tcgen05 kind, int8 has (kind::i8, s32 accumulators in tensor memory, 4x the DMMA rate per byte
// of operand).  Both operands are cut into I8_NS = 7 signed base-256 digits of a 55-bit fixed-point number
// (Ozaki splitting):   W[i,k]   = 2^(e_i - 54) sum_s d_s[i,k] 2^(48 - 8 s),   e_i = exponent of max_k |W[i,k]|
//                      K_*[c,k] = 2^(f - 54)   sum_t q_t[c,k] 2^(48 - 8 t),   f   = exponent of scale max|kvec|
// so that   A[c,i] = 2^(e_i + f - 60) sum_l 2^(48 - 8 l) acc_l[c,i],   acc_l = sum_{s+t=l} sum_k q_t[c,k] d_s[i,k].
// Every acc_l is an exact int32 dot product (|digit| <= 128: |acc_l| <= 7 * n_pad * 2^14 < 2^31 up to
// n_pad = 16384); the 28 digit pairs with l = s + t <= 6 are kept (the dropped ones are below 2^-52 of
// max|W_i| max|K_*| per term, the digit representation itself is exact to 2^-55).  (Round 2 started with 8 digits
// of 7 bits = 36 pairs; 7 digits of 8 bits carry the same 55-56 bits in 28 pairs - 22 % fewer MMAs and 12.5 % fewer
// bytes for the same error.)  Per (128-candidate tile, 64-row block of W) the seven level accumulators take
// 448 of the SM's 512 tensor-memory columns (7 x 64 columns x 128 lanes);
// they are folded into fp64 only once, after the whole k range: two exact int64 Horner sums, two conversions,
// one fma.  A warp-specialised CTA per SM: one thread issues the TMA loads of the digit planes (64-byte
// swizzle, 84 KB per stage, two stages), one thread issues the 20 MMAs per stage, four warps read the
// accumulators back (tcgen05.ld) and carry sum A^2 and sum A V per candidate in registers.
// =========================================================================================================
constexpr int I8_NS = 7;          // digits per operand (base 256, balanced)
constexpr int I8_A_LO = 4;        // A planes of the first load / waited for at once; the rest when digit 4 comes up
constexpr int I8_TILE = 128;      // candidates per tile = MMA M
constexpr int I8_RB = 64;         // W rows per row block = MMA N
constexpr int I8_KC = 64;         // k per stage (bytes per smem row; two K = 32 MMA steps)
constexpr int I8_STAGES = 2;
constexpr int I8_STAGE_A = I8_NS * I8_TILE * I8_KC;   // 57344
constexpr int I8_STAGE_B = I8_NS * I8_RB * I8_KC;     // 28672
constexpr int I8_STAGE = I8_STAGE_A + I8_STAGE_B;
constexpr int I8_SMEM = I8_STAGES * I8_STAGE + 1024 /* alignment slack */ + 128 /* 12 barriers, tmem slot */;
constexpr int I8_THREADS = 192;   // warp 0: TMA, warp 1: MMA + tensor-memory owner, warps 2-5: read-back
constexpr int I8_MAX_RB = 256;    // n_pad <= 16384: 7 pairs * n_pad * 2^14 must stay below 2^31
constexpr int I8_MAX_SLICES = 64;
constexpr int I8_FRAC = 54;       // fractional bits of the fixed-point operands (|x| < 1 -> |integer| < 2^54)

struct I8Args {
  int n_rb, ntiles, slices, tiles_in_flight;
  int debug;                // measurement only: bit 0 = no MMAs, bit 1 = no TMA loads, bit 2 = no read-back
  const double* rowscale;   // [n_pad] 2^e_i
  const double* kexp;       // [2]: 2^(54 - f), 2^(f - 60)
  const double* V;
  double* partial;          // [ntiles][slices][128][2]
  const int4* items;        // MODE 1: (a_row, b_row, first k-chunk, end k-chunk) per work item
  const int* item_off;      // MODE 1: items of CTA b = items[item_off[b] .. item_off[b + 1])
  double* C;                // MODE 1: output matrix, C = alpha * A B^T
  int ldc;
  const double* rowscale_a; // MODE 1: 2^exponent per row of the A operand (rowscale: per row of the B operand)
  double alpha;
  int beta;                 // MODE 1: 0 = C is overwritten, 1 = C += alpha * A B^T (read-modify-write of the tile)
  unsigned* meet;           // [slices][waves][row blocks of the slice] arrival counters, zero on entry (or null)
  int meet_stride, meet_waves;   // counters per wave, waves per launch
  unsigned short off[I8_MAX_SLICES + 1];   // row blocks of slice s: order[off[s] .. off[s+1])
  unsigned short order[I8_MAX_RB];
};

// 55-bit fixed point -> 7 balanced base-256 digits, most significant first: d[1..6] in [-128, 127], d[0] in
// [-65, 65].  The low three digits come from the low 24 bits (carry into the rest), all in 32-bit arithmetic.
__device__ __forceinline__ void digits8(double x_scaled, int (&dg)[I8_NS]) {
  const long long v = __double2ll_rn(x_scaled);          // |v| <= 2^54
  int lo = (int)((unsigned)v & 0x00ffffffu);             // v = (v >> 24) 2^24 + lo, 0 <= lo < 2^24
  int hi = (int)(v >> 24);                               // floor; |hi| <= 2^30
#pragma unroll
  for (int s = 6; s > 3; --s) {
    const int d = ((lo + 128) & 255) - 128;
    dg[s] = d;
    lo = (lo - d) >> 8;
  }
  hi += lo;                                              // carry of the low group: 0 or 1
#pragma unroll
  for (int s = 3; s > 0; --s) {
    const int d = ((hi + 128) & 255) - 128;
    dg[s] = d;
    hi = (hi - d) >> 8;
  }
  dg[0] = hi;
}

// digit planes of a lower-triangular matrix split per ROW: W8[s][i][k], rowscale[i] = 2^e_i.  One CTA per row.
// rr == null: the whole row (k <= i is data, k > i zero) - W = L^-1 once per fit.  Otherwise only the rows with a
// non-empty column range rr[i] = [k0, k1) are written, over that range (k > i zero): the bottom-right blocks W22
// of one level of the triangular inverse.
static __global__ void __launch_bounds__(256) i8_wsplit_kernel(const double* __restrict__ W, int n_pad,
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
    score_i8_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapKhi,
                    const __grid_constant__ CUtensorMap mapW, const __grid_constant__ I8Args p) {
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
    tma::prefetch_map(&mapKhi);
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
            tma::mbar_arrive_expect_tx(sb + 8, I8_A_LO * I8_TILE * I8_KC);
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
            tma::mbar_arrive_expect_tx(sb + 16, (I8_NS - I8_A_LO) * I8_TILE * I8_KC);
            tma::load_3d(dst + I8_A_LO * I8_TILE * I8_KC, &mapKhi, kc * I8_KC, a_row, I8_A_LO, sb + 16);
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
        // columns 64 l), its product lands on levels t + s0 .. t + s0 + n - 1 in one MMA: 10 wide MMAs (N = 256 /
        // 192 / 128, one of 64) per k-step instead of 28 of N = 64, which were bound by re-reading the A tile from
        // shared memory (measured 45 % of the int8 rate).  Digits 0-2 need two MMAs: the second one takes the A
        // tile from the collector.  k-step outermost: consecutive MMAs write different accumulator columns (the
        // same columns come up again 10 MMAs later); digit-outermost serialised on the accumulate dependency.
        auto issue = [&](int ks, int t) {
          const unsigned long long ad = da + (unsigned long long)((t * I8_TILE * I8_KC + ks * 32) >> 4);
          const int planes = I8_NS - t;                                 // B digits 0 .. planes - 1
          const int n0 = planes > 4 ? (planes == 5 ? 3 : 4) : planes;   // 7=4+3 6=4+2 5=3+2
          const unsigned acc = (kc > kb || ks > 0 || t > 0) ? 1u : 0u;  // t = 0 covers all 448 columns
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
            for (int t = 0; t < I8_A_LO; ++t) issue(0, t);
          }
          __syncwarp();
          tma::mbar_wait_bounded(sb + 16, phase);                       // A planes 4-7
          tc5::fence_after();
          if (tma::elect_one()) {
#pragma unroll
            for (int t = I8_A_LO; t < I8_NS; ++t) issue(0, t);
#pragma unroll
            for (int t = 0; t < I8_A_LO; ++t) issue(1, t);
            tc5::commit(sb + 24);                                       // A planes 0-3 are free
#pragma unroll
            for (int t = I8_A_LO; t < I8_NS; ++t) issue(1, t);
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
    const double kx = MODE == 0 ? p.kexp[1] : 8.673617379884035472e-19;   // 2^(f - 60) / 2^-60
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
          // sum_l acc_l 2^(48 - 8 l) = hi 2^32 + lo; |hi| < 2^47 converts exactly, |lo| < 2^55 to 2^-53 of itself
          const long long hi = (((long long)a[0][j] * 256 + a[1][j]) * 256) + a[2][j];
          const long long lo = ((((long long)a[3][j] * 256 + a[4][j]) * 256 + a[5][j]) * 256) + a[6][j];
          const int i = b_row + j0 + j;
          const double av = (__ldg(p.rowscale + i) * rsa) * fma((double)hi, 4294967296.0, (double)lo);
          if (MODE == 0) {
            s2 = fma(av, av, s2);
            sv = fma(av, __ldg(p.V + i), sv);
          } else {
            out[j] = av;
          }
        }
        if (MODE == 1) {
          if (p.beta) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const double2 c = *reinterpret_cast<const double2*>(crow + j0 + j);
              out[j] += c.x;
              out[j + 1] += c.y;
            }
          }
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
// digit planes of a column panel of L: rows [row0, n_pad), columns [k0, k1) (all below the diagonal), one exponent
// per row over that column range; one CTA per row.  Feeds the rank-(k1 - k0) trailing updates of the Cholesky.
static __global__ void __launch_bounds__(128) i8_panel_split_kernel(const double* __restrict__ L, int n_pad, int row0,
                                                                    int k0, int k1, signed char* __restrict__ P8,
                                                                    double* __restrict__ rowscale) {
  __shared__ double red[128];
  const int i = row0 + blockIdx.x, tid = threadIdx.x;
  const double* row = L + (size_t)i * n_pad;
  double m = 0.0;
  for (int k = k0 + tid; k < k1; k += 128) m = fmax(m, fabs(row[k]));
  red[tid] = m;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmax(red[tid], red[tid + s]);
    __syncthreads();
  }
  int e = 0;
  if (red[0] > 0.0 && isfinite(red[0])) frexp(red[0], &e);
  if (tid == 0) rowscale[i] = ldexp(1.0, e);
  const double sc = ldexp(1.0, I8_FRAC - e);
  const size_t plane = (size_t)n_pad * n_pad;
  for (int k4 = k0 + tid * 4; k4 < k1; k4 += 512) {
    unsigned pk[I8_NS];
#pragma unroll
    for (int s = 0; s < I8_NS; ++s) pk[s] = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int dg[I8_NS];
      digits8(row[k4 + u] * sc, dg);
#pragma unroll
      for (int s = 0; s < I8_NS; ++s) pk[s] |= (unsigned)(dg[s] & 0xff) << (8 * u);
    }
#pragma unroll
    for (int s = 0; s < I8_NS; ++s)
      *reinterpret_cast<unsigned*>(P8 + s * plane + (size_t)i * n_pad + k4) = pk[s];
  }
}

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
static __global__ void __launch_bounds__(256) i8_colmax_kernel(const double* __restrict__ W, int n_pad,
                                                        unsigned long long* __restrict__ colmax,
                                                        const int2* __restrict__ cr, int mode) {
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int2 rg = i8_col_range(cr, mode, i, n_pad);
  const int k0 = max((int)blockIdx.y * 512, rg.x), k1 = min(min(n_pad, (int)blockIdx.y * 512 + 512), rg.y);
  double m = 0.0;
  for (int k = k0 + (threadIdx.x >> 6); k < k1; k += 4) m = fmax(m, fabs(W[(size_t)k * n_pad + i]));
  if (m > 0.0) atomicMax(colmax + i, (unsigned long long)__double_as_longlong(m));
}

static __global__ void __launch_bounds__(256) i8_colscale_kernel(double* __restrict__ colscale, int n_pad) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_pad) return;
  const double m = colscale[i];   // the column maximum (bit pattern written by atomicMax)
  int e = 0;
  if (m > 0.0 && isfinite(m)) frexp(m, &e);
  colscale[i] = ldexp(1.0, e);
}

// CTA = 64 x 64 tile (k block, i block), transposed through shared memory.  cr == null: the lower-triangular tiles
// are enumerated in blockIdx.x; otherwise the grid is (k blocks, i blocks) and tiles outside the column ranges exit.
static __global__ void __launch_bounds__(256) i8_wtsplit_kernel(const double* __restrict__ W, int n_pad,
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

}  // namespace db
