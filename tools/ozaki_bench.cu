// tools/ozaki_bench.cu -- measuring stick + accuracy study for the INT8 (Ozaki-split) FP64 contraction.
// Builds a synthetic precision matrix shaped like the benchmark model's (P = L L^T, banded L + 1 % fill),
// residual rows, runs (a) the tcgen05 int8 kernel with S digit planes, (b) cublasDgemm, and compares both
// with a long-double dot product on sampled entries.
// Usage: ozaki_bench [K=1997] [B=8192] [iters=10] [S=7] [seed=1]
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../mcmc-date_b200/csrc/gemm_i8_ozaki.cuh"

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                           \
    }                                                                                    \
  } while (0)

using namespace mcd;

template <int S>
int run(int K, int B, int iters, unsigned seed) {
  const int ld8 = (K + OZ_KB - 1) / OZ_KB * OZ_KB;
  const int Mp = (K + OZ_N - 1) / OZ_N * OZ_N;
  const int Bp = (B + OZ_M - 1) / OZ_M * OZ_M;
  const int ldy = Mp;
  printf("== S=%d K=%d B=%d ld8=%d Mp=%d Bp=%d smem=%zu\n", S, K, B, ld8, Mp, Bp, oz_smem_bytes<S>());
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> nrm(0.0, 1.0);
  std::exponential_distribution<double> expd(1.0 / 0.05);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  // L: diag 1/sigma_k, band 64 + 1 % fill
  std::vector<double> L((size_t)K * K, 0.0), sig(K);
  for (int k = 0; k < K; ++k) {
    sig[k] = 0.1 * (expd(rng) + 1e-4) + 1e-3;
    L[(size_t)k * K + k] = 1.0 / sig[k];
    for (int j = std::max(0, k - 64); j < k; ++j) L[(size_t)k * K + j] = 0.05 / sig[k] * nrm(rng);
    for (int j = 0; j < k - 64; ++j)
      if (uni(rng) < 0.01) L[(size_t)k * K + j] = 0.05 / sig[k] * nrm(rng);
  }
  std::vector<double> X((size_t)B * K);
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < K; ++k) X[(size_t)b * K + k] = sig[k] * nrm(rng) * (uni(rng) < 0.05 ? 1e-3 : 1.0);

  cublasHandle_t cb;
  cublasCreate(&cb);
  double *dL, *dP, *dX, *dYr, *dY, *dsa, *dsb;
  CK(cudaMalloc(&dL, (size_t)K * K * 8));
  CK(cudaMalloc(&dP, (size_t)K * K * 8));
  CK(cudaMalloc(&dX, (size_t)B * K * 8));
  CK(cudaMalloc(&dYr, (size_t)B * K * 8));
  CK(cudaMalloc(&dY, (size_t)Bp * ldy * 8));
  CK(cudaMalloc(&dsa, (size_t)Bp * 8));
  CK(cudaMalloc(&dsb, (size_t)Mp * 8));
  CK(cudaMemset(dsa, 0, (size_t)Bp * 8));
  CK(cudaMemset(dsb, 0, (size_t)Mp * 8));
  CK(cudaMemcpy(dL, L.data(), (size_t)K * K * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), (size_t)B * K * 8, cudaMemcpyHostToDevice));
  const double one = 1.0, zero = 0.0;
  // row-major L is column-major L^T: P = L L^T (symmetric, so the storage order does not matter)
  cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, K, K, K, &one, dL, K, dL, K, &zero, dP, K);
  std::vector<double> P((size_t)K * K);
  CK(cudaMemcpy(P.data(), dP, (size_t)K * K * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < i; ++j) P[(size_t)i * K + j] = P[(size_t)j * K + i];  // exactly symmetric
  CK(cudaMemcpy(dP, P.data(), (size_t)K * K * 8, cudaMemcpyHostToDevice));

  // digit planes
  signed char *pA, *pB;
  const size_t strideA = (size_t)Bp * ld8, strideB = (size_t)Mp * ld8;
  CK(cudaMalloc(&pA, strideA * S));
  CK(cudaMalloc(&pB, strideB * S));
  CK(cudaMemset(pA, 0, strideA * S));
  CK(cudaMemset(pB, 0, strideB * S));
  CUtensorMap tmA, tmB;
  if (oz_make_plane_map(&tmA, pA, (size_t)S * Bp, ld8, OZ_M) || oz_make_plane_map(&tmB, pB, (size_t)S * Mp, ld8, OZ_N)) {
    fprintf(stderr, "tensor map failed\n");
    return 1;
  }
  CK(gemm_i8_ozaki_configure<S>());
  oz_split_rows_kernel<S><<<(K + 7) / 8, 256>>>(dP, K, K, K, pB, ld8, strideB, dsb, ldexp(1.0, -16));
  CK(cudaGetLastError());
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventCreate(&e2);
  float ms_split = 0, ms_mm = 0;
  for (int it = 0; it < iters + 2; ++it) {
    if (it == 2) CK(cudaEventRecord(e0));
    oz_split_rows_kernel<S><<<(B + 7) / 8, 256>>>(dX, K, B, K, pA, ld8, strideA, dsa, 1.0);
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms_split, e0, e1);
  for (int it = 0; it < iters + 2; ++it) {
    if (it == 2) CK(cudaEventRecord(e1));
    CK(gemm_i8_ozaki_launch<S>(tmA, tmB, dsa, dsb, dY, Mp, Bp, ld8, ldy, Bp, 0));
  }
  CK(cudaEventRecord(e2));
  CK(cudaEventSynchronize(e2));
  cudaEventElapsedTime(&ms_mm, e1, e2);
  ms_split /= iters;
  ms_mm /= iters;
  const double pairs = S * (S + 1) / 2.0;
  printf("split %.3f ms   int8 contraction %.3f ms  (%.1f TFLOP/s FP64-equivalent, %.0f TOP/s int8 executed)\n", ms_split,
         ms_mm, 2.0 * K * K * B / ms_mm / 1e9, pairs * 2.0 * ld8 * Mp * Bp / ms_mm / 1e9);
  // cublas reference: Yr[b][m] = sum_k P[m][k] X[b][k]; column-major view: Yr^T (K x B) = P^T (as col-major P) ...
  float ms_cb = 0;
  for (int it = 0; it < 3; ++it) {
    if (it == 1) CK(cudaEventRecord(e0));
    cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, K, B, K, &one, dP, K, dX, K, &zero, dYr, K);
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms_cb, e0, e1);
  printf("cublasDgemm %.3f ms (%.1f TFLOP/s)\n", ms_cb / 2, 2.0 * K * K * B / (ms_cb / 2) / 1e9);
  std::vector<double> Y((size_t)Bp * ldy), Yr((size_t)B * K);
  CK(cudaMemcpy(Y.data(), dY, Y.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Yr.data(), dYr, Yr.size() * 8, cudaMemcpyDeviceToHost));
  // sampled long-double truth
  double e_i8_abs = 0, e_cb_abs = 0, e_i8_rel = 0, e_cb_rel = 0, e_i8_rms = 0, e_cb_rms = 0, e_i8_max = 0;
  const int ns = 4000;
  std::uniform_int_distribution<int> ub(0, B - 1), um(0, K - 1);
  for (int s = 0; s < ns; ++s) {
    const int b = s < 8 ? (s & 1 ? B - 1 : 0) : ub(rng), m = s < 8 ? (s & 2 ? K - 1 : 0) : um(rng);
    long double acc = 0, mag = 0, pmax = 0, xmax = 0;
    for (int k = 0; k < K; ++k) {
      const long double t = (long double)P[(size_t)m * K + k] * (long double)X[(size_t)b * K + k];
      acc += t;
      mag += fabsl(t);
      pmax = std::max(pmax, fabsl((long double)P[(size_t)m * K + k]));
      xmax = std::max(xmax, fabsl((long double)X[(size_t)b * K + k]));
    }
    const double yi = Y[(size_t)b * ldy + m], yc = Yr[(size_t)b * K + m];
    const double ei = (double)fabsl(yi - acc), ec = (double)fabsl(yc - acc);
    e_i8_abs = std::max(e_i8_abs, ei / (double)mag);
    e_cb_abs = std::max(e_cb_abs, ec / (double)mag);
    e_i8_rel = std::max(e_i8_rel, ei / (double)fabsl(acc));
    e_cb_rel = std::max(e_cb_rel, ec / (double)fabsl(acc));
    e_i8_max = std::max(e_i8_max, ei / (double)(pmax * xmax));
    e_i8_rms += (ei / (double)mag) * (ei / (double)mag);
    e_cb_rms += (ec / (double)mag) * (ec / (double)mag);
  }
  printf("error / sum|P||x| : int8 max %.3e rms %.3e | dgemm max %.3e rms %.3e\n", e_i8_abs, std::sqrt(e_i8_rms / ns), e_cb_abs,
         std::sqrt(e_cb_rms / ns));
  printf("error / |y|       : int8 max %.3e | dgemm max %.3e ;  int8 error / (max|P| max|x|) max %.3e\n", e_i8_rel, e_cb_rel, e_i8_max);
  // whole-matrix comparison with dgemm
  double dmax = 0, ymax = 0;
  for (int b = 0; b < B; ++b)
    for (int m = 0; m < K; ++m) {
      dmax = std::max(dmax, std::fabs(Y[(size_t)b * ldy + m] - Yr[(size_t)b * K + m]));
      ymax = std::max(ymax, std::fabs(Yr[(size_t)b * K + m]));
    }
  printf("whole matrix: max |Y_i8 - Y_dgemm| = %.3e (max |Y| = %.3e)  -> %s\n", dmax, ymax, dmax <= 1e-9 * ymax ? "OK" : "MISMATCH");
  cudaFree(dL); cudaFree(dP); cudaFree(dX); cudaFree(dYr); cudaFree(dY); cudaFree(dsa); cudaFree(dsb); cudaFree(pA); cudaFree(pB);
  cublasDestroy(cb);
  return dmax <= 1e-9 * ymax ? 0 : 2;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 1997;
  const int B = argc > 2 ? atoi(argv[2]) : 8192;
  const int iters = argc > 3 ? atoi(argv[3]) : 10;
  const int S = argc > 4 ? atoi(argv[4]) : 7;
  const unsigned seed = argc > 5 ? (unsigned)atoi(argv[5]) : 1u;
  switch (S) {
    case 5: return run<5>(K, B, iters, seed);
    case 6: return run<6>(K, B, iters, seed);
    case 7: return run<7>(K, B, iters, seed);
    default: fprintf(stderr, "S must be 5..7\n"); return 1;
  }
}
