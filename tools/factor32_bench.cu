// Micro-benchmark of the 32x32 diagonal factor on the critical path of the blocked Cholesky (potf2.cuh).
// One warp, clock64 around each variant, same SPD block; reports cycles per column and the max error of the
// reconstructed factor against a host Cholesky.
//   A  factor_32_smem   (round-1 default: column line through shared memory, sqrt / rsqrt in the same warp)
//   A' factor_32        (shuffle form)
//   B  chain_1x1        pivots + UNSCALED columns only (the sqrt scaling is left to other warps)
//   B2 chain_1x1_shfl   as B, pivot fetched by one shuffle so that the reciprocal starts before the LDS round trip
//   C  chain_2x2        two columns per step: one reciprocal (of the 2x2 determinant) per pair
//   D  dfma latency     dependent DFMA chain, cycles per instruction (context for the numbers above)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/factor32_bench tools/factor32_bench.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include "../distbo_b200/csrc/potf2.cuh"
using namespace db;

// B: lane = row.  Column j (unscaled: c_k = a_kj after the updates of columns < j) goes to sC[j][.] and stays
// there; piv[j] = d_j.  L[k][j] = c_k / sqrt(d_j) is NOT formed here.
__device__ __forceinline__ int chain_1x1(double (&a)[SB], int lane, double* sC /*[SB][SB]*/) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    double* line = sC + j * SB;
    line[lane] = a[j];
    __syncwarp();
    const double d = line[j];
    bad = (!(d > 0.0) && bad == 0) ? j + 1 : bad;
    const double t = a[j] * rcp_nr(d);
#pragma unroll
    for (int k2 = 0; k2 < SB; k2 += 2) {
      if (k2 + 1 > j) {
        const double2 c = *reinterpret_cast<const double2*>(line + k2);
        if (k2 > j) a[k2] = fma(-t, c.x, a[k2]);
        a[k2 + 1] = fma(-t, c.y, a[k2 + 1]);
      }
    }
  }
  return bad;
}

__device__ __forceinline__ int chain_1x1_shfl(double (&a)[SB], int lane, double* sC) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    double* line = sC + j * SB;
    line[lane] = a[j];
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    __syncwarp();
    bad = (!(d > 0.0) && bad == 0) ? j + 1 : bad;
    const double t = a[j] * rcp_nr(d);
#pragma unroll
    for (int k2 = 0; k2 < SB; k2 += 2) {
      if (k2 + 1 > j) {
        const double2 c = *reinterpret_cast<const double2*>(line + k2);
        if (k2 > j) a[k2] = fma(-t, c.x, a[k2]);
        a[k2 + 1] = fma(-t, c.y, a[k2 + 1]);
      }
    }
  }
  return bad;
}

// C: columns (j, j+1) together.  D = [p q; q r] (rows j, j+1 of the two unscaled columns), det = p r - q^2,
// u = (r c1 - q c2) / det, v = (p c2 - q c1) / det, a[k] -= u C1[k] + v C2[k] for k >= j + 2.
// sC[j] = column j as it was BEFORE the pair step (c1), sC[j+1] = column j+1 before the pair step (c2, i.e. not yet
// updated by column j); the scaling pass forms l_j = c1 / sqrt(p), l_{j+1} = (c2 - (q/p) c1) / sqrt(r - q^2 / p).
__device__ __forceinline__ int chain_2x2_bench(double (&a)[SB], int lane, double* sC) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < SB; j += 2) {
    double* l1 = sC + j * SB;
    double* l2 = l1 + SB;
    l1[lane] = a[j];
    l2[lane] = a[j + 1];
    __syncwarp();
    const double p = l1[j], q = l1[j + 1], r = l2[j + 1];
    const double det = fma(p, r, -q * q);
    bad = ((!(p > 0.0) || !(det > 0.0)) && bad == 0) ? j + 1 + (p > 0.0 ? 1 : 0) : bad;
    const double id = rcp_nr(det);
    const double u = fma(r, a[j], -q * a[j + 1]) * id;
    const double v = fma(p, a[j + 1], -q * a[j]) * id;
#pragma unroll
    for (int k2 = 0; k2 < SB; k2 += 2) {
      if (k2 > j + 1) {
        const double2 c1 = *reinterpret_cast<const double2*>(l1 + k2);
        const double2 c2 = *reinterpret_cast<const double2*>(l2 + k2);
        a[k2] = fma(-v, c2.x, fma(-u, c1.x, a[k2]));
        a[k2 + 1] = fma(-v, c2.y, fma(-u, c1.y, a[k2 + 1]));
      }
    }
  }
  return bad;
}

__global__ void bench_kernel(const double* __restrict__ A, double* __restrict__ out, long long* cyc, int variant) {
  __shared__ double sC[SB * SB];
  __shared__ double sS[2 * SB];
  const int lane = threadIdx.x;
  double a[SB], myr = 0.0;
#pragma unroll
  for (int k = 0; k < SB; ++k) a[k] = A[lane * SB + k];
  __syncwarp();
  const long long t0 = clock64();
  int bad = 0;
  if (variant == 0) bad = factor_32_smem(a, myr, lane, sS);
  else if (variant == 1) bad = factor_32(a, myr, lane);
  else if (variant == 2) bad = chain_1x1(a, lane, sC);
  else if (variant == 3) bad = chain_1x1_shfl(a, lane, sC);
  else if (variant == 4) bad = chain_2x2_bench(a, lane, sC);
  else {
    double x = A[lane], y = A[lane + 32];
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-3);
    a[0] = x;
  }
  const long long t1 = clock64();
  if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = bad; }
  __syncwarp();
  // reconstruct L from the variant's output so that every variant is checked the same way
  if (variant <= 1) {
#pragma unroll
    for (int k = 0; k < SB; ++k) out[lane * SB + k] = (k <= lane) ? a[k] : 0.0;
  } else if (variant == 2 || variant == 3) {
    for (int j = 0; j < SB; ++j) out[lane * SB + j] = (j <= lane) ? sC[j * SB + lane] / sqrt(sC[j * SB + j]) : 0.0;
  } else if (variant == 4) {
    for (int j = 0; j < SB; j += 2) {
      const double p = sC[j * SB + j], q = sC[j * SB + j + 1], r = sC[(j + 1) * SB + j + 1];
      const double c1 = sC[j * SB + lane], c2 = sC[(j + 1) * SB + lane];
      out[lane * SB + j] = (j <= lane) ? c1 / sqrt(p) : 0.0;
      out[lane * SB + j + 1] = (j + 1 <= lane) ? (c2 - (q / p) * c1) / sqrt(r - q * q / p) : 0.0;
    }
  } else {
    out[lane] = a[0];
  }
}

int main() {
  const int n = SB;
  std::vector<double> h(n * n), L(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) h[i * n + j] = exp(-0.05 * abs(i - j)) + (i == j ? 0.5 : 0.0);
  for (int j = 0; j < n; ++j) {
    double d = h[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    L[j * n + j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = h[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / L[j * n + j];
    }
  }
  double *A, *out;
  long long* cyc;
  cudaMalloc(&A, n * n * 8);
  cudaMalloc(&out, n * n * 8);
  cudaMalloc(&cyc, 16);
  cudaMemcpy(A, h.data(), n * n * 8, cudaMemcpyHostToDevice);
  const char* names[] = {"A  factor_32_smem (default)", "A' factor_32 shuffle", "B  chain_1x1 (unscaled cols)",
                         "B2 chain_1x1 + pivot shuffle", "C  chain_2x2", "D  256 dependent DFMA"};
  for (int v = 0; v < 6; ++v) {
    long long best = 1LL << 60, hc[2];
    std::vector<double> g(n * n);
    for (int rep = 0; rep < 5; ++rep) {
      bench_kernel<<<1, 32>>>(A, out, cyc, v);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
      if (hc[0] < best) best = hc[0];
    }
    cudaMemcpy(g.data(), out, n * n * 8, cudaMemcpyDeviceToHost);
    double err = 0;
    if (v < 5) for (int i = 0; i < n * n; ++i) err = fmax(err, fabs(g[i] - L[i]));
    if (v < 5) printf("%-32s %6lld cycles = %5.1f per column, info %lld, max |L - L_host| %.2e\n", names[v], best,
                      best / 32.0, hc[1], err);
    else printf("%-32s %6lld cycles = %5.2f per DFMA\n", names[v], best, best / 256.0);
  }
  return 0;
}
