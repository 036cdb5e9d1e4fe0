// Micro-benchmark: per-phase clock64 stamps of potf2_inv_kernel on one 128x128 SPD block.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../distbo_b200/csrc/potf2.cuh"
using namespace db;
int main() {
  const int n = 128;
  std::vector<double> h(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) h[i * n + j] = exp(-0.05 * abs(i - j)) + (i == j ? 0.5 : 0.0);
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
    printf("rep %d err=%d time %.1f us; stamps (cycles since start):", rep, (int)e, ms * 1000);
    for (int i = 1; i < 22; ++i) printf(" %lld", hd[i] - hd[i - 1]);
    printf("\n");
  }
  return 0;
}
