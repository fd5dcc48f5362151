// cholesky.cuh -- blocked right-looking Cholesky factorisation P = L L^T on the device (FP64), used once per model for the
// value-only path's triangular contraction quad = |L^T dx|^2 (north_star's "fixed Cholesky factor of the precision matrix").
// The reference has no counterpart: hmatrix multiplies with the full inverse covariance (app/Probability.hs:166-173).
//
// Work matrix W [K][ldw] row-major, lower triangle; block size 64:
//   for every diagonal block j:  chol_diag_kernel   L_jj          (one CTA, the block in shared memory)
//                                chol_panel_kernel  W[i][j] <- W[i][j] L_jj^-T   for the blocks i below (one CTA per 64 rows)
//                                chol_update_kernel W[i][c] -= W[i][j] W[c][j]^T for j < c <= i (one CTA per 64 x 64 tile)
// ~K^3/3 flops in 3 K / 64 small launches: milliseconds at K = 2000 where the scalar host loop took seconds.
// A non-positive pivot sets *flag (the matrix is not positive definite; the caller keeps the symmetric product).
#pragma once
#include <cuda_runtime.h>

namespace mcd {

constexpr int CH_NB = 64;

__global__ void __launch_bounds__(256)
chol_diag_kernel(double* __restrict__ W, int ldw, int K, int j0, int* __restrict__ flag) {
  __shared__ double a[CH_NB][CH_NB + 1];
  const int nb = min(CH_NB, K - j0), tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += 256) a[e / nb][e % nb] = W[(size_t)(j0 + e / nb) * ldw + j0 + e % nb];
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    const double d = a[c][c];
    if (!(d > 0.0)) {  // uniform over the CTA
      if (tid == 0) *flag = 1;
      return;
    }
    const double sd = sqrt(d);
    __syncthreads();
    if (tid == 0) a[c][c] = sd;
    for (int r = c + 1 + tid; r < nb; r += 256) a[r][c] = a[r][c] / sd;
    __syncthreads();
    // trailing update of the block: a[r][c2] -= a[r][c] a[c2][c], c < c2 <= r
    const int m = nb - c - 1;
    for (int e = tid; e < m * m; e += 256) {
      const int r = c + 1 + e / m, c2 = c + 1 + e % m;
      if (c2 <= r) a[r][c2] = fma(-a[r][c], a[c2][c], a[r][c2]);
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += 256) {
    const int r = e / nb, c = e % nb;
    W[(size_t)(j0 + r) * ldw + j0 + c] = c <= r ? a[r][c] : 0.0;
  }
}

// rows [i0, i0 + 64) of the panel below the diagonal block: X L_jj^T = B, one thread per row
__global__ void __launch_bounds__(64)
chol_panel_kernel(double* __restrict__ W, int ldw, int K, int j0) {
  extern __shared__ __align__(16) unsigned char smem_ch[];
  double (*l)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(smem_ch);
  double (*b)[CH_NB + 1] = l + CH_NB;
  const int nb = min(CH_NB, K - j0), tid = threadIdx.x;
  const int i0 = j0 + CH_NB + blockIdx.x * CH_NB, rows = min(CH_NB, K - i0);
  for (int e = tid; e < nb * nb; e += 64) l[e / nb][e % nb] = W[(size_t)(j0 + e / nb) * ldw + j0 + e % nb];
  for (int e = tid; e < rows * nb; e += 64) b[e / nb][e % nb] = W[(size_t)(i0 + e / nb) * ldw + j0 + e % nb];
  __syncthreads();
  if (tid < rows) {
    for (int c = 0; c < nb; ++c) {
      double s = b[tid][c];
      for (int t = 0; t < c; ++t) s = fma(-b[tid][t], l[c][t], s);
      b[tid][c] = s / l[c][c];
    }
  }
  __syncthreads();
  for (int e = tid; e < rows * nb; e += 64) W[(size_t)(i0 + e / nb) * ldw + j0 + e % nb] = b[e / nb][e % nb];
}

// trailing update: tile (ti, tc) of the blocks below / right of block j, tc <= ti:  W[i][c] -= sum_t W[i][j0 + t] W[c][j0 + t]
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ W, int ldw, int K, int j0) {
  const int ti = blockIdx.y, tc = blockIdx.x;
  if (tc > ti) return;
  extern __shared__ __align__(16) unsigned char smem_ch[];
  double (*pa)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(smem_ch);
  double (*pb)[CH_NB + 1] = pa + CH_NB;
  const int base = j0 + CH_NB, i0 = base + ti * CH_NB, c0 = base + tc * CH_NB;
  const int ri = min(CH_NB, K - i0), rc = min(CH_NB, K - c0), tid = threadIdx.x;
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int r = e / CH_NB, t = e % CH_NB;
    pa[r][t] = r < ri ? W[(size_t)(i0 + r) * ldw + j0 + t] : 0.0;
    pb[r][t] = r < rc ? W[(size_t)(c0 + r) * ldw + j0 + t] : 0.0;
  }
  __syncthreads();
  const int r0 = (tid / 16) * 4, q0 = (tid % 16) * 4;  // 4 x 4 outputs per thread
  double acc[4][4] = {};
#pragma unroll 8
  for (int t = 0; t < CH_NB; ++t) {
    double x[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { x[u] = pa[r0 + u][t]; y[u] = pb[q0 + u][t]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fma(x[u], y[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = r0 + u, c = q0 + v;
      if (r < ri && c < rc && c0 + c <= i0 + r) W[(size_t)(i0 + r) * ldw + c0 + c] -= acc[u][v];
    }
}

// U[m][k] = L[k][m] for k >= m (zero elsewhere) into the padded operand of the contraction
__global__ void __launch_bounds__(256)
chol_transpose_kernel(const double* __restrict__ W, int ldw, int K, double* __restrict__ U, int ldu) {
  const int m = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
  if (m >= K || k >= K) return;
  U[(size_t)m * ldu + k] = k >= m ? W[(size_t)k * ldw + m] : 0.0;
}

// W: lower triangle of the symmetric matrix on entry, L on exit (strictly upper part of the diagonal blocks zeroed)
constexpr int CH_SMEM = 2 * CH_NB * (CH_NB + 1) * 8;
inline cudaError_t cholesky_device(double* W, int ldw, int K, int* d_flag, cudaStream_t st) {
  cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
  cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
  for (int j0 = 0; j0 < K; j0 += CH_NB) {
    chol_diag_kernel<<<1, 256, 0, st>>>(W, ldw, K, j0, d_flag);
    const int below = K - j0 - CH_NB;
    if (below > 0) {
      const int nt = (below + CH_NB - 1) / CH_NB;
      chol_panel_kernel<<<nt, 64, CH_SMEM, st>>>(W, ldw, K, j0);
      chol_update_kernel<<<dim3(nt, nt), 256, CH_SMEM, st>>>(W, ldw, K, j0);
    }
  }
  return cudaGetLastError();
}

}  // namespace mcd
