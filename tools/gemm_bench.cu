// tools/gemm_bench.cu -- measuring stick for the FP64 contraction (not part of the product).
// Times (a) the product DMMA kernel, (b) a plain-DFMA register-tiled kernel with the same tiling,
// (c) cublasDgemm, on the benchmark shape, and checks (a),(b) against (c).
// Build: see tools/Makefile.   Usage: gemm_bench [K=1997] [B=8192] [iters=10]
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mcmc-date_b200/csrc/gemm_f64.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

using namespace mcd;

// ---- local helpers for the DFMA comparison kernel (cp.async ring, padded smem rows)
constexpr int DF_LD = GEMM_BK + 2;
constexpr int DF_STAGES = 4;
constexpr size_t DF_SMEM = (size_t)DF_STAGES * (GEMM_BT + GEMM_PR) * DF_LD * sizeof(double);
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Same tiles and smem layout as the DMMA kernel, but vector FP64 FMAs (8x8 outputs / thread).
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f64_dfma_kernel(const double* __restrict__ P, const double* __restrict__ DX,
                     double* __restrict__ Y, int ldk, int ldy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + (size_t)DF_STAGES * GEMM_BT * DF_LD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lm = lane >> 2, ln = lane & 3;
  const int wbt = (warp & 1) * 64, wpr = (warp >> 1) * 32;
  const int pr0 = blockIdx.x * GEMM_PR, bt0 = blockIdx.y * GEMM_BT;
  const double* gA = DX + (size_t)bt0 * ldk;
  const double* gB = P + (size_t)pr0 * ldk;
  const int nk = ldk / GEMM_BK;
  auto load_stage = [&](int stage, int kt) {
    double* dA = sA + (size_t)stage * GEMM_BT * DF_LD;
    double* dB = sB + (size_t)stage * GEMM_PR * DF_LD;
    const int k0 = kt * GEMM_BK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * GEMM_THREADS;
      int row = c >> 3, ch = c & 7;
      cp_async16(dA + row * DF_LD + ch * 2, gA + (size_t)row * ldk + k0 + ch * 2);
      cp_async16(dB + row * DF_LD + ch * 2, gB + (size_t)row * ldk + k0 + ch * 2);
    }
  };
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
#pragma unroll
  for (int s = 0; s < DF_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<DF_STAGES - 2>();
    __syncthreads();
    int nxt = kt + DF_STAGES - 1;
    if (nxt < nk) load_stage(nxt % DF_STAGES, nxt);
    cp_async_commit();
    const double* cA = sA + (size_t)(kt % DF_STAGES) * GEMM_BT * DF_LD;
    const double* cB = sB + (size_t)(kt % DF_STAGES) * GEMM_PR * DF_LD;
#pragma unroll
    for (int k = 0; k < GEMM_BK; k += 2) {
      double2 a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = *reinterpret_cast<const double2*>(cA + (wbt + lm + 8 * j) * DF_LD + k);
        b[j] = *reinterpret_cast<const double2*>(cB + (wpr + ln + 4 * j) * DF_LD + k);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fma(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fma(a[i].y, b[j].y, acc[i][j]);
        }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      Y[(size_t)(bt0 + wbt + lm + 8 * i) * ldy + pr0 + wpr + ln + 4 * j] = acc[i][j];
}

static double maxdiff(const std::vector<double>& a, const std::vector<double>& b, double* scale) {
  double m = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    m = fmax(m, fabs(a[i] - b[i]));
    s = fmax(s, fabs(b[i]));
  }
  *scale = s;
  return m;
}

int main(int argc, char** argv) {
  int K = argc > 1 ? atoi(argv[1]) : 1997;
  int B = argc > 2 ? atoi(argv[2]) : 8192;
  int iters = argc > 3 ? atoi(argv[3]) : 10;
  int ldk = (K + 15) / 16 * 16, Mp = (K + 127) / 128 * 128, Bp = (B + 127) / 128 * 128, ldy = Mp;
  printf("K=%d B=%d  padded: Mp=%d ldk=%d Bp=%d\n", K, B, Mp, ldk, Bp);
  size_t nP = (size_t)Mp * ldk, nX = (size_t)Bp * ldk, nY = (size_t)Bp * ldy;
  std::vector<double> hP(nP, 0.0), hX(nX, 0.0);
  srand(1234);
  for (int m = 0; m < K; ++m)
    for (int k = 0; k <= m; ++k) {
      double v = (rand() / (double)RAND_MAX - 0.5) * (m == k ? 50.0 : 1.0);
      hP[(size_t)m * ldk + k] = v;
      hP[(size_t)k * ldk + m] = v;
    }
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < K; ++k) hX[(size_t)b * ldk + k] = rand() / (double)RAND_MAX - 0.5;
  double *dP, *dX, *dY, *dYref;
  CK(cudaMalloc(&dP, nP * 8));
  CK(cudaMalloc(&dX, nX * 8));
  CK(cudaMalloc(&dY, nY * 8));
  CK(cudaMalloc(&dYref, nY * 8));
  CK(cudaMemcpy(dP, hP.data(), nP * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, hX.data(), nX * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double flops = 2.0 * Mp * (double)Bp * ldk;          // executed (padded)
  const double flops_alg = 2.0 * K * (double)K * (double)B;  // algorithmic 2K^2 per chain
  float ms;

  // (c) cuBLAS
  cublasHandle_t h;
  cublasCreate(&h);
  double one = 1.0, zero = 0.0;
  auto run_cublas = [&]() {
    cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, Mp, Bp, ldk, &one, dP, ldk, dX, ldk, &zero, dYref, ldy);
  };
  for (int i = 0; i < 3; ++i) run_cublas();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) run_cublas();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("cublasDgemm      : %8.3f ms  %6.2f TFLOP/s executed  %6.2f algorithmic\n", ms / iters,
         flops / (ms / iters) * 1e-9, flops_alg / (ms / iters) * 1e-9);
  std::vector<double> yref(nY), y(nY);
  CK(cudaMemcpy(yref.data(), dYref, nY * 8, cudaMemcpyDeviceToHost));

  // (a) DMMA product kernel
  CK(gemm_f64_dmma_configure());
  CUtensorMap tmP, tmX;
  if (make_tile_map(&tmP, dP, Mp, ldk) || make_tile_map(&tmX, dX, Bp, ldk)) { fprintf(stderr, "tensor map failed\n"); return 1; }
  for (int i = 0; i < 3; ++i) CK(gemm_f64_dmma_launch(tmP, tmX, dY, Mp, Bp, ldk, ldy, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(gemm_f64_dmma_launch(tmP, tmX, dY, Mp, Bp, ldk, ldy, 0));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("dmma tma (product): %8.3f ms  %6.2f TFLOP/s executed  %6.2f algorithmic\n", ms / iters,
         flops / (ms / iters) * 1e-9, flops_alg / (ms / iters) * 1e-9);
  CK(cudaMemcpy(y.data(), dY, nY * 8, cudaMemcpyDeviceToHost));
  double sc, md = maxdiff(y, yref, &sc);
  printf("  max|dmma - cublas| = %.3e (max|ref| %.3e)\n", md, sc);

  // (b) DFMA kernel
  CK(cudaFuncSetAttribute(gemm_f64_dfma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)DF_SMEM));
  dim3 grid(Mp / GEMM_PR, Bp / GEMM_BT);
  CK(cudaMemset(dY, 0, nY * 8));
  for (int i = 0; i < 3; ++i)
    gemm_f64_dfma_kernel<<<grid, GEMM_THREADS, DF_SMEM>>>(dP, dX, dY, ldk, ldy);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i)
    gemm_f64_dfma_kernel<<<grid, GEMM_THREADS, DF_SMEM>>>(dP, dX, dY, ldk, ldy);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("dfma 8x8/thread  : %8.3f ms  %6.2f TFLOP/s executed  %6.2f algorithmic\n", ms / iters,
         flops / (ms / iters) * 1e-9, flops_alg / (ms / iters) * 1e-9);
  CK(cudaMemcpy(y.data(), dY, nY * 8, cudaMemcpyDeviceToHost));
  md = maxdiff(y, yref, &sc);
  printf("  max|dfma - cublas| = %.3e\n", md);

  // host spot check of the reference itself (first chain, first 4 rows)
  for (int m = 0; m < 4; ++m) {
    double s = 0;
    for (int k = 0; k < K; ++k) s += hP[(size_t)m * ldk + k] * hX[k];
    printf("  host y[0][%d]=%.15g  cublas=%.15g\n", m, s, yref[m]);
  }
  return 0;
}
