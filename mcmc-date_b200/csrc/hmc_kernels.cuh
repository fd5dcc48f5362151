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

// =====================================================================================================
// Batched No-U-Turn sampler (second half).
//
// The reference's Hamiltonian proposal is `nuts` of the third-party `mcmc` package (app/Hamiltonian.hs:95-104;
// source not vendored): the "efficient NUTS" of Hoffman & Gelman (2014), Algorithm 3 -- slice variable
// u ~ U(0, exp(ln post - kinetic)), repeated doubling of the trajectory in a random direction, a point of the
// new half is kept with probability n'/n, stop at the first U-turn of any balanced sub-trajectory or at a
// divergence (energy error > 1000).  Here every chain builds its own tree, all chains advance in lockstep, one
// leapfrog step per tick: the recursion of BuildTree is unrolled into the usual checkpoint scheme (the first
// point of every still-open balanced sub-trajectory is kept; a leaf with odd index closes as many
// sub-trajectories as it has trailing one bits), the uniform choice among the valid points of the new half is
// reservoir sampling (same distribution as the recursive n''/(n'+n'') rule).  Positions, momenta, gradients,
// checkpoints and candidates of all chains stay in HBM; per tick only the three evaluation kernels and
// nuts_leaf_kernel run.
//
// Random numbers: Philox4x32-10, key = seed, counter = (chain, iteration, draw index, stream), so a host
// restatement draws bit-identical uniforms (tests/nuts_ref.py).

struct Philox {
  uint32_t k0, k1;
  __host__ __device__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ void operator()(uint32_t (&c)[4]) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, a, b);
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
  }
};
// uniform in (0, 1): 53 random bits, (k + 0.5) 2^-53
__device__ __forceinline__ double nuts_uniform(uint64_t seed, uint32_t chain, uint32_t iteration, uint32_t draw) {
  uint32_t c[4] = {chain, iteration, draw, 0u};
  Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}(c);
  const uint64_t k = ((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6);
  return ((double)k + 0.5) * 1.1102230246251565e-16;
}
// two standard normals (Box-Muller), stream 1: used when the caller supplies no momenta
__device__ __forceinline__ void nuts_normal2(uint64_t seed, uint32_t chain, uint32_t iteration, uint32_t pair, double* z0, double* z1) {
  uint32_t c[4] = {chain, iteration, pair, 1u};
  Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}(c);
  const double u0 = ((double)(((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6)) + 0.5) * 1.1102230246251565e-16;
  const double u1 = ((double)(((uint64_t)(c[2] >> 5) << 26) | (uint64_t)(c[3] >> 6)) + 0.5) * 1.1102230246251565e-16;
  const double rad = sqrt(-2.0 * log(u0));
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);
  *z0 = rad * cs;
  *z1 = rad * sn;
}

// per-chain integer / real scalars
enum { NI_ACTIVE = 0, NI_DEPTH, NI_LEAF, NI_DIR, NI_N, NI_NSUB, NI_NALPHA, NI_NLEAP, NI_STATUS, NI_DRAW, NI_DIVERGED, NI_TURNED,
       NI_SLOT,  // row of the compacted evaluation batch this chain's next leaf is evaluated in
       NI_COLS = 16 };
enum { NR_LOGU = 0, NR_H0NEG, NR_ALPHA, NR_EPS, NR_COLS = 4 };
constexpr double NUTS_DELTA_MAX = 1000.0;

struct NutsBuffers {
  double *thE[2], *rE[2], *gE[2];  // [B][D] trajectory ends: 0 = backward (left), 1 = forward (right)
  double *thM, *thC;               // [B][D] current sample / candidate of the sub-trajectory being built
  double *ck_th, *ck_r;            // [max_depth][B][D] checkpoints (first point of every open sub-trajectory)
  double *outM, *outC;             // [B][8] ln-posterior parts at thM / thC
  int* ni;                         // [B][NI_COLS]
  double* nr;                      // [B][NR_COLS]
  int* n_active;                   // [1]
};

// block-wide sum of up to 4 values (256 threads), result valid in every thread
template <int NV>
__device__ __forceinline__ void nuts_block_sum(double (&v)[NV], double* scratch /*[NV][8]*/) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
  __syncthreads();  // scratch may still be read from a previous call
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) scratch[j * 8 + (threadIdx.x >> 5)] = v[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += scratch[j * 8 + w];
    v[j] = t;
  }
}

// first half of the next leapfrog step at the active end of a chain that goes on: half kick, drift, and the new
// position written straight into the chain's state row (fromVectorWith layout) for the evaluation kernels.
// Called by all threads of the chain's CTA after its scalars have been updated (and a __syncthreads()).
__device__ __forceinline__ void nuts_kick_drift(const NutsBuffers& nb, int b, const int* __restrict__ sidx,
                                                const double* __restrict__ inv_mass, double* __restrict__ states, int S, int D) {
  const int* ni = nb.ni + (size_t)b * NI_COLS;
  if (!ni[NI_ACTIVE]) return;
  const int dir = ni[NI_DIR];
  const double e = (dir ? 1.0 : -1.0) * nb.nr[(size_t)b * NR_COLS + NR_EPS];
  const size_t o = (size_t)b * D;
  double* __restrict__ th = nb.thE[dir] + o;
  double* __restrict__ rr = nb.rE[dir] + o;
  const double* __restrict__ gg = nb.gE[dir] + o;
  double* __restrict__ x = states + (size_t)ni[NI_SLOT] * S;
  // 4 independent elements per thread and trip: all loads of a trip are in flight before the first store
  for (int t0 = threadIdx.x; t0 < D; t0 += 4 * HMC_THREADS) {
    double r4[4], g4[4], q4[4], m4[4];
    int j4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * HMC_THREADS;
      if (t < D) { r4[u] = rr[t]; g4[u] = gg[t]; q4[u] = th[t]; m4[u] = inv_mass[t]; j4[u] = sidx[t]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * HMC_THREADS;
      if (t < D) {
        const double p = r4[u] + 0.5 * e * g4[u];
        const double q = q4[u] + e * m4[u] * p;
        rr[t] = p;
        th[t] = q;
        x[j4[u]] = q;
      }
    }
  }
}

// after the evaluation at theta0: momenta, slice variable, both trajectory ends = the start point
__global__ void __launch_bounds__(HMC_THREADS)
nuts_init_kernel(NutsBuffers nb, const double* __restrict__ theta0, const double* __restrict__ mom0 /*nullable*/,
                 const double* __restrict__ grad, const int* __restrict__ sidx, const double* __restrict__ inv_mass,
                 const double* __restrict__ eps, const double* __restrict__ out, const int* __restrict__ status,
                 double* __restrict__ states, uint64_t seed, uint32_t iteration, int S, int D, int B) {
  __shared__ double scratch[4 * 8];
  const int b = blockIdx.x;
  if (b >= B) return;
  const size_t o = (size_t)b * D;
  double v[1] = {0.0};
  for (int t = threadIdx.x; t < D; t += HMC_THREADS) {
    double p;
    if (mom0 != nullptr) {
      p = mom0[o + t];
    } else {  // r ~ N(0, M), M = diag(1 / inv_mass)
      double z0, z1;
      nuts_normal2(seed, (uint32_t)b, iteration, (uint32_t)(t >> 1), &z0, &z1);
      p = ((t & 1) ? z1 : z0) * rsqrt(inv_mass[t]);
    }
    const double th = theta0[o + t], g = grad[(size_t)b * S + sidx[t]];
    nb.thE[0][o + t] = th; nb.thE[1][o + t] = th; nb.thM[o + t] = th;
    nb.rE[0][o + t] = p; nb.rE[1][o + t] = p;
    nb.gE[0][o + t] = g; nb.gE[1][o + t] = g;
    v[0] = fma(p * p, inv_mass[t], v[0]);
  }
  nuts_block_sum<1>(v, scratch);
  if (threadIdx.x < 8) nb.outM[(size_t)b * 8 + threadIdx.x] = out[(size_t)b * 8 + threadIdx.x];
  if (threadIdx.x == 0) {
    int* ni = nb.ni + (size_t)b * NI_COLS;
    double* nr = nb.nr + (size_t)b * NR_COLS;
    const double h0neg = out[(size_t)b * 8 + 6] - 0.5 * v[0];
    const double u = nuts_uniform(seed, (uint32_t)b, iteration, 0u);
    const double ud = nuts_uniform(seed, (uint32_t)b, iteration, 1u);
    nr[NR_LOGU] = h0neg + log(u);
    nr[NR_H0NEG] = h0neg;
    nr[NR_ALPHA] = 0.0;
    nr[NR_EPS] = eps[b];
    const int ok = (h0neg == h0neg) && (h0neg > -1.7976931348623157e308) && (h0neg < 1.7976931348623157e308);
    ni[NI_ACTIVE] = ok;  // a chain that starts at a point of zero / undefined density stays where it is
    ni[NI_DEPTH] = 0; ni[NI_LEAF] = 0; ni[NI_DIR] = ud < 0.5 ? 0 : 1; ni[NI_N] = 1; ni[NI_NSUB] = 0; ni[NI_NALPHA] = 0;
    ni[NI_NLEAP] = 0; ni[NI_STATUS] = status[b]; ni[NI_DRAW] = 2; ni[NI_DIVERGED] = 0; ni[NI_TURNED] = 0;
    // still-active chains are compacted: each takes the next free row of the evaluation batch (the order is
    // arbitrary, the chains are independent, so the results do not depend on it)
    if (ok) ni[NI_SLOT] = atomicAdd(nb.n_active, 1);
  }
  __syncthreads();
  nuts_kick_drift(nb, b, sidx, inv_mass, states, S, D);
}

// second half kick + all tree bookkeeping of the new leaf + (if the chain goes on) the first half of its next
// leapfrog step; one CTA per chain.  Every thread keeps its NE elements (t = tid + 256 i) of the active end's
// position / momentum / gradient in registers from the first pass to the last, so that per tick the end is read
// once and written once (a chain that continues at the same end never writes its gradient at all).
// Requires D <= 256 NE.
template <int NE>
__global__ void __launch_bounds__(HMC_THREADS, NE <= 12 ? 2 : 1)
nuts_leaf_kernel(NutsBuffers nb, const double* __restrict__ grad, const int* __restrict__ sidx,
                 const double* __restrict__ inv_mass, const double* __restrict__ out, const int* __restrict__ status,
                 double* __restrict__ states, uint64_t seed, uint32_t iteration, int max_depth, int S, int D, int B) {
  __shared__ double scratch[4 * 8];
  __shared__ int sh[4];
  const int b = blockIdx.x;
  if (b >= B) return;
  int* ni = nb.ni + (size_t)b * NI_COLS;
  if (!ni[NI_ACTIVE]) return;
  double* nr = nb.nr + (size_t)b * NR_COLS;
  const int dir = ni[NI_DIR], depth = ni[NI_DEPTH], leaf = ni[NI_LEAF];
  const double vsgn = dir ? 1.0 : -1.0, e = vsgn * nr[NR_EPS];
  const size_t o = (size_t)b * D, BD = (size_t)B * D;
  double* __restrict__ th = nb.thE[dir] + o;
  double* __restrict__ rr = nb.rE[dir] + o;
  double* __restrict__ gg = nb.gE[dir] + o;
  const int tid = threadIdx.x;
  // ---- second half kick and kinetic energy, fused with the checkpoint work of this leaf: an even leaf opens a
  // balanced sub-trajectory (its position / momentum are parked at level idx_max), an odd leaf closes nsub of them
  // (first one checked here, deeper ones below); both are harmless if the leaf turns out to be divergent
  const int idx_max = __popc((unsigned)leaf >> 1);
  const bool even = (leaf & 1) == 0;
  double* __restrict__ cth0 = nb.ck_th + (size_t)idx_max * BD + o;
  double* __restrict__ cr0 = nb.ck_r + (size_t)idx_max * BD + o;
  const int slot = ni[NI_SLOT];  // row of this leaf in the compacted evaluation batch
  const double* __restrict__ grow = grad + (size_t)slot * S;
  out += (size_t)slot * 8;
  double P[NE], G[NE], Q[NE];
  double v1[3] = {0.0, 0.0, 0.0};
  {
    double A[NE], C[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int t = tid + i * HMC_THREADS;
      if (t < D) {
        G[i] = grow[sidx[t]]; P[i] = rr[t]; Q[i] = th[t];
        if (!even) { A[i] = cth0[t]; C[i] = cr0[t]; }
      }
    }
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int t = tid + i * HMC_THREADS;
      if (t < D) {
        const double m = inv_mass[t];
        P[i] = P[i] + 0.5 * e * G[i];
        v1[0] = fma(P[i] * P[i], m, v1[0]);
        if (even) {
          cth0[t] = Q[i];
          cr0[t] = P[i];
        } else {
          const double dth = (Q[i] - A[i]) * m;
          v1[1] = fma(dth, C[i], v1[1]);
          v1[2] = fma(dth, P[i], v1[2]);
        }
      }
    }
  }
  nuts_block_sum<3>(v1, scratch);
  // ---- leaf: validity under the slice, divergence, acceptance statistic, reservoir choice
  if (tid == 0) {
    const double hneg = out[6] - 0.5 * v1[0];
    const double logu = nr[NR_LOGU];
    const int valid = logu <= hneg;                       // false for NaN
    const int diverged = !(hneg > logu - NUTS_DELTA_MAX);  // true for NaN
    const double a = exp(hneg - nr[NR_H0NEG]);
    nr[NR_ALPHA] += (a == a) ? fmin(1.0, a) : 0.0;
    ni[NI_NALPHA] += 1;
    ni[NI_NLEAP] += 1;
    ni[NI_STATUS] |= status[slot];
    int take = 0;
    if (valid) {
      const int ns = ni[NI_NSUB] + 1;
      ni[NI_NSUB] = ns;
      const double u = nuts_uniform(seed, (uint32_t)b, iteration, (uint32_t)ni[NI_DRAW]);
      ni[NI_DRAW] += 1;
      take = u * (double)ns < 1.0;
    }
    if (diverged) ni[NI_DIVERGED] = 1;
    sh[0] = take;
    sh[1] = diverged;
  }
  __syncthreads();
  const int take = sh[0], diverged = sh[1];
  if (take) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int t = tid + i * HMC_THREADS;
      if (t < D) nb.thC[o + t] = Q[i];
    }
    if (tid < 8) nb.outC[(size_t)b * 8 + tid] = out[tid];
  }
  // ---- U-turn checks of the balanced sub-trajectories this leaf closes (checkpoint scheme)
  int turned = 0;
  if (!diverged && !even) {
    const int nsub = __ffs(~(unsigned)leaf) - 1;  // trailing one bits of the leaf index
    turned = (vsgn * v1[1] < 0.0) || (vsgn * v1[2] < 0.0);
    for (int k = idx_max - 1; k > idx_max - nsub && !turned; --k) {
      const double* __restrict__ cth = nb.ck_th + (size_t)k * BD + o;
      const double* __restrict__ cr = nb.ck_r + (size_t)k * BD + o;
      double d2[2] = {0.0, 0.0};
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int t = tid + i * HMC_THREADS;
        if (t < D) {
          const double dth = (Q[i] - cth[t]) * inv_mass[t];
          d2[0] = fma(dth, cr[t], d2[0]);
          d2[1] = fma(dth, P[i], d2[1]);
        }
      }
      nuts_block_sum<2>(d2, scratch);
      turned = (vsgn * d2[0] < 0.0) || (vsgn * d2[1] < 0.0);
    }
  }
  // ---- end of the sub-trajectory / of the whole trajectory
  const int s_sub = !diverged && !turned;
  const int sub_done = s_sub && (leaf + 1 == (1 << depth));
  if (sub_done) {
    // main tree: (theta+ - theta-) . M^-1 r- >= 0 and (theta+ - theta-) . M^-1 r+ >= 0; the active end is in
    // registers, the other one in memory
    double d2[2] = {0.0, 0.0};
    const double* __restrict__ tO = nb.thE[1 - dir] + o;
    const double* __restrict__ rO = nb.rE[1 - dir] + o;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int t = tid + i * HMC_THREADS;
      if (t < D) {
        const double dth = vsgn * (Q[i] - tO[t]) * inv_mass[t];  // theta_right - theta_left
        d2[0] = fma(dth, rO[t], d2[0]);
        d2[1] = fma(dth, P[i], d2[1]);
      }
    }
    nuts_block_sum<2>(d2, scratch);
    if (tid == 0) {
      const int n = ni[NI_N], ns = ni[NI_NSUB];
      const double u = nuts_uniform(seed, (uint32_t)b, iteration, (uint32_t)ni[NI_DRAW]);
      ni[NI_DRAW] += 1;
      sh[2] = ns > 0 && u * (double)n < (double)ns;  // accept the candidate with probability min(1, n'/n)
      ni[NI_N] = n + ns;
      const int turned_main = (d2[0] < 0.0) || (d2[1] < 0.0);
      const int nd = depth + 1;
      ni[NI_DEPTH] = nd;
      if (turned_main) ni[NI_TURNED] = 1;
      if (turned_main || nd >= max_depth) {
        ni[NI_ACTIVE] = 0;
      } else {
        const double ud = nuts_uniform(seed, (uint32_t)b, iteration, (uint32_t)ni[NI_DRAW]);
        ni[NI_DRAW] += 1;
        ni[NI_DIR] = ud < 0.5 ? 0 : 1;
        ni[NI_NSUB] = 0;
        ni[NI_LEAF] = 0;
        ni[NI_SLOT] = atomicAdd(nb.n_active, 1);
      }
    }
    __syncthreads();
    if (sh[2]) {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int t = tid + i * HMC_THREADS;
        if (t < D) nb.thM[o + t] = nb.thC[o + t];
      }
      if (tid < 8) nb.outM[(size_t)b * 8 + tid] = nb.outC[(size_t)b * 8 + tid];
    }
  } else {
    if (tid == 0) {
      if (!s_sub) {
        if (turned) ni[NI_TURNED] = 1;
        ni[NI_ACTIVE] = 0;  // the new half is discarded and the trajectory ends
      } else {
        ni[NI_LEAF] = leaf + 1;
        ni[NI_SLOT] = atomicAdd(nb.n_active, 1);
      }
    }
    __syncthreads();
  }
  // ---- write the end back; a chain that goes on at the same end takes the first half of its next step first
  const bool same_end = ni[NI_ACTIVE] && ni[NI_DIR] == dir;
  double* __restrict__ x = states + (size_t)ni[NI_SLOT] * S;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const int t = tid + i * HMC_THREADS;
    if (t < D) {
      if (same_end) {
        const double p = P[i] + 0.5 * e * G[i];
        const double q = Q[i] + e * inv_mass[t] * p;
        rr[t] = p;
        th[t] = q;
        x[sidx[t]] = q;
      } else {
        rr[t] = P[i];
        gg[t] = G[i];
      }
    }
  }
  if (!same_end) nuts_kick_drift(nb, b, sidx, inv_mass, states, S, D);  // other end (its data are in memory), or nothing
}

}  // namespace mcd
