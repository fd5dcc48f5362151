// hmc_kernels.cuh -- leapfrog trajectories resident on the device (first half).
//
// The reference's Hamiltonian proposal (app/Hamiltonian.hs:95-104 -> third-party `mcmc`: leapfrog
// integrator over the masked position vector of toVector / fromVectorWith) asks the host for one
// gradient per leapfrog step.  Here the positions, momenta and gradients of all chains stay in HBM for
// the whole trajectory; per step only the three evaluation kernels and one of these elementwise kernels
// run, and PCIe is touched at the two ends.
//
//   H(theta, p) = -ln post(theta) + 1/2 p^T M^-1 p          (M diagonal)
//   kick:  p += c * eps_b * grad        drift:  theta += eps_b * M^-1 p
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcd {

constexpr int HMC_THREADS = 256;

// p += kick * eps[b] * g ; if (drift) theta += eps[b] * inv_mass[t] * p.  Also folds the step's status
// word into the trajectory's (thread t == 0 of each chain).
__global__ void __launch_bounds__(HMC_THREADS)
leapfrog_update_kernel(double* __restrict__ theta, double* __restrict__ mom, const double* __restrict__ gtheta,
                       const double* __restrict__ inv_mass, const double* __restrict__ eps, double kick, int drift,
                       const int* __restrict__ status, int* __restrict__ status_acc, int D, int B) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * HMC_THREADS + threadIdx.x;
  if (b >= B || t >= D) return;
  const size_t o = (size_t)b * D + t;
  const double e = eps[b];
  const double p = mom[o] + kick * e * gtheta[o];
  mom[o] = p;
  if (drift) theta[o] = theta[o] + e * inv_mass[t] * p;
  if (t == 0 && status != nullptr) status_acc[b] |= status[b];
}

// energy[b][which] = -ln post + 1/2 sum_t p_t^2 inv_mass[t]; one warp per chain, fixed order (deterministic)
__global__ void __launch_bounds__(HMC_THREADS)
hamiltonian_kernel(const double* __restrict__ mom, const double* __restrict__ inv_mass, const double* __restrict__ out,
                   double* __restrict__ energy, int which, int D, int B) {
  const int b = blockIdx.x * (HMC_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const double* p = mom + (size_t)b * D;
  double s = 0.0;
  for (int t = lane; t < D; t += 32) s = fma(p[t] * p[t], inv_mass[t], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) energy[(size_t)b * 2 + which] = -out[(size_t)b * 8 + 6] + 0.5 * s;
}

}  // namespace mcd
