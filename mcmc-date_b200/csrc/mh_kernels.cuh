// mh_kernels.cuh -- Metropolis-Hastings proposals on chains that live in HBM (first part).
//
// The reference's default sampler cycles through thousands of single-node proposals per iteration
// (app/Definitions.hs:145-278), each followed by a full prior + likelihood evaluation.  Here the chains' states
// stay resident on the device: one call applies one proposal to every chain (in place, with an undo log), the
// batched value-only evaluation scores the proposed states, and an accept kernel keeps or restores them.
//
// Proposals restated (all first-party code of the reference, no third-party sampling involved):
//   MH_SLIDE_NODE      slideNodeAtUltrametric      lib/Mcmc/Tree/Proposal/Ultrametric.hs:50-62
//                      h' ~ truncated normal(mean h, sd t*s) on (max child height, parent height); |J| = 1
//   MH_SCALE_SUBTREE   scaleSubTreeAtUltrametric   lib/Mcmc/Tree/Proposal/Ultrametric.hs:126-147
//                      h' ~ truncated normal(mean h, sd t*s) on (0, parent height); every height of the sub tree is
//                      scaled by xi = h'/h; |J| = xi^(n_inner - 1)
// with truncatedNormalSample / the truncated normal of lib/Mcmc/Tree/Proposal/Internal.hs:100-137 and
// lib/Statistics/Distribution/TruncatedNormal.hs:61-131:
//   z(m) = Phi((b-m)/s') - Phi((a-m)/s'),  quantile(p) = erfinv(2 (p z + Phi(alpha)) - 1) sqrt(2) s' + m,
//   Hastings factor q = density_{h'}(h) / density_{h}(h') = z(h) / z(h').
// Uniform random numbers: Philox4x32-10 (hmc_kernels.cuh), counter (chain, iteration, draw, 2).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "hmc_kernels.cuh"

namespace mcd {

enum { MH_SLIDE_NODE = 0, MH_SCALE_SUBTREE = 1 };
enum { MH_ST_OK = 0, MH_ST_INVALID = 1 };  // invalid: the reference's truncatedNormalDistr would call `error`

__device__ __forceinline__ double mh_uniform(uint64_t seed, uint32_t chain, uint32_t iteration, uint32_t draw) {
  uint32_t c[4] = {chain, iteration, draw, 2u};
  Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}(c);
  const uint64_t k = ((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6);
  return ((double)k + 0.5) * 1.1102230246251565e-16;
}
__device__ __forceinline__ double mh_phi2(double x) { return 0.5 * (1.0 + erf(x * 0.70710678118654752440)); }

// Apply the proposal to node `node` (same node for all chains; node < 0: every chain draws its own inner non-root
// node uniformly) of every chain IN PLACE.  One CTA per chain.  undo[b][0..] keeps the old heights of the modified
// range, meta[b] = (first modified node, number of modified nodes, validity, -), lq[b] = ln(q |J|).
__global__ void __launch_bounds__(256)
mh_propose_kernel(double* __restrict__ states, double* __restrict__ undo, int4* __restrict__ meta, double* __restrict__ lq,
                  const int* __restrict__ parent /*leaf bit 31*/, const int* __restrict__ child1, const int* __restrict__ sub_size,
                  const int* __restrict__ sub_inner, const int* __restrict__ inner_list, int n_inner_nonroot, int kind, int node,
                  double sd_tuned, uint64_t seed, uint32_t iteration, int S, int N, int B) {
  __shared__ double sh_xi;
  __shared__ int sh_j, sh_cnt;
  const int b = blockIdx.x;
  if (b >= B) return;
  double* h = states + (size_t)b * S + 3;
  if (threadIdx.x == 0) {
    int j = node;
    if (j < 0) {
      const double un = mh_uniform(seed, (uint32_t)b, iteration, 2u);
      int pick = (int)(un * (double)n_inner_nonroot);
      if (pick >= n_inner_nonroot) pick = n_inner_nonroot - 1;
      j = inner_list[pick];
    }
    const double hj = h[j], hP = h[parent[j] & 0x7fffffff];
    double a, bb;
    if (kind == MH_SLIDE_NODE) {
      a = fmax(h[j + 1], h[child1[j]]);  // hbdMaximumChildrenHeight
      bb = hP;
    } else {
      a = 0.0;
      bb = hP;
    }
    const double s = sd_tuned;
    // truncatedNormalDistr's error conditions (TruncatedNormal.hs:61-79); NaNs fail the comparisons too
    const bool ok = (s > 0.0) && (a < bb) && !(a > hj) && !(bb < hj) && (hj == hj);
    double xi = 1.0, lnq = 0.0, hnew = hj;
    int cnt = 0;
    if (ok) {
      const double p = mh_uniform(seed, (uint32_t)b, iteration, 0u);
      const double phiA = mh_phi2((a - hj) / s), z = mh_phi2((bb - hj) / s) - phiA;
      hnew = erfinv(2.0 * (p * z + phiA) - 1.0) * 1.41421356237309504880 * s + hj;
      // truncatedNormalSample: a sample outside [a, b] is a numerical failure (`error` in the reference)
      if (!(a > hnew || bb < hnew) && hnew == hnew && z > 0.0) {
        const double z2 = mh_phi2((bb - hnew) / s) - mh_phi2((a - hnew) / s);
        lnq = log(z) - log(z2);  // q = qYX / qXY = z(h) / z(h')
        if (kind == MH_SCALE_SUBTREE) {
          xi = hnew / hj;
          lnq += (double)(sub_inner[j] - 1) * log(xi);
          cnt = sub_size[j];
        } else {
          cnt = 1;
        }
      }
    }
    sh_xi = xi;
    sh_j = j;
    sh_cnt = cnt;
    meta[b] = make_int4(j, cnt, cnt > 0 ? MH_ST_OK : MH_ST_INVALID, 0);
    lq[b] = lnq;
    if (cnt > 0) {
      undo[(size_t)b * N] = hj;
      h[j] = hnew;  // the sub tree's root gets the sampled height itself (scaleUltrametricTreeF)
    }
  }
  __syncthreads();
  const int j = sh_j, cnt = sh_cnt;
  const double xi = sh_xi;
  // scale the rest of the sub tree (pre-order: nodes j+1 .. j+cnt-1); leaves stay at 0 * xi = 0
  for (int i = 1 + threadIdx.x; i < cnt; i += 256) {
    const double old = h[j + i];
    undo[(size_t)b * N + i] = old;
    h[j + i] = old * xi;
  }
}

// ln r = ln post(y) - ln post(x) + ln(q |J|); accept iff ln u < ln r.  post = prior * likelihood (* root-branch
// Jacobian for proposals lifted with jacobianRootBranch, app/Definitions.hs:145-150).  Rejected chains get their
// heights back from the undo log.  One CTA per chain.
__global__ void __launch_bounds__(256)
mh_accept_kernel(double* __restrict__ states, const double* __restrict__ undo, const int4* __restrict__ meta,
                 const double* __restrict__ lq, double* __restrict__ cur_out, const double* __restrict__ new_out,
                 int* __restrict__ cur_status, const int* __restrict__ new_status, int* __restrict__ accepted,
                 int use_root_jacobian, uint64_t seed, uint32_t iteration, int S, int N, int B) {
  __shared__ int sh_acc;
  const int b = blockIdx.x;
  if (b >= B) return;
  const int4 m = meta[b];
  if (threadIdx.x == 0) {
    int acc = 0;
    if (m.y > 0) {
      const double* o1 = new_out + (size_t)b * 8;
      const double* o0 = cur_out + (size_t)b * 8;
      double lr = (o1[3] + o1[4]) - (o0[3] + o0[4]) + lq[b];
      if (use_root_jacobian) lr += o1[5] - o0[5];
      const double u = mh_uniform(seed, (uint32_t)b, iteration, 1u);
      acc = log(u) < lr;  // false for NaN and for -inf
    }
    sh_acc = acc;
    if (accepted) accepted[b] = m.z == MH_ST_INVALID ? -1 : acc;
  }
  __syncthreads();
  if (sh_acc) {
    if (threadIdx.x < 8) cur_out[(size_t)b * 8 + threadIdx.x] = new_out[(size_t)b * 8 + threadIdx.x];
    if (threadIdx.x == 8) cur_status[b] = new_status[b];
  } else {
    double* h = states + (size_t)b * S + 3;
    for (int i = threadIdx.x; i < m.y; i += 256) h[m.x + i] = undo[(size_t)b * N + i];
  }
}

}  // namespace mcd
