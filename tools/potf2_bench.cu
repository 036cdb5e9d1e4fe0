// Micro-benchmark + check of potf2_inv_kernel on one 128x128 SPD block: per-phase clock64 stamps, total time,
// and max error of L and W against a host Cholesky / inverse.  Build twice (with and without -DDB_FACTOR_SHUFFLE).
#include <cmath>
#include <cstdio>
#include <vector>
#include "../distbo_b200/csrc/potf2.cuh"
using namespace db;
int main() {
  const int n = 128;
  std::vector<double> h(n * n), L(n * n, 0.0), Wi(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) h[i * n + j] = exp(-0.05 * abs(i - j)) + (i == j ? 0.5 : 0.0);
  for (int j = 0; j < n; ++j) {   // host reference
    double d = h[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    L[j * n + j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = h[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / L[j * n + j];
    }
  }
  for (int c = 0; c < n; ++c)
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i * n + k] * Wi[k * n + c];
      Wi[i * n + c] = s / L[i * n + i];
    }
  double *A, *W; int* info; long long* dbg;
  cudaMalloc(&A, n * n * 8); cudaMalloc(&W, n * n * 8); cudaMalloc(&info, 4); cudaMalloc(&dbg, 64 * 8);
  cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(A, h.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaMemset(info, 0, 4);
    cudaEventRecord(e0);
    potf2_inv_kernel<<<1, POTF2_THREADS, POTF2_SMEM>>>(A, n, W, 0, info, dbg);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hd[64]; cudaMemcpy(hd, dbg, 64 * 8, cudaMemcpyDeviceToHost);
    std::vector<double> gl(n * n), gw(n * n);
    cudaMemcpy(gl.data(), A, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(gw.data(), W, n * n * 8, cudaMemcpyDeviceToHost);
    double el = 0, ew = 0;
    for (int i = 0; i < n * n; ++i) { el = fmax(el, fabs(gl[i] - L[i])); ew = fmax(ew, fabs(gw[i] - Wi[i])); }
    printf("rep %d err=%d time %.1f us maxerr L %.2e W %.2e; phase cycles:", rep, (int)e, ms * 1000, el, ew);
    for (int i = 1; i < 23; ++i) printf(" %lld", hd[i] - hd[i - 1]);
    printf("\n");
  }
  return 0;
}
