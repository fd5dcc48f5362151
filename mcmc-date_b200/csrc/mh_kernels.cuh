// mh_kernels.cuh -- Metropolis-Hastings-Green proposals, heated chains and MC3 swaps on chains that live in HBM
// (SURVEY.md 8f rank 4).
//
// The reference's default sampler cycles through thousands of small proposals per iteration
// (app/Definitions.hs:145-279), each followed by a full prior + likelihood evaluation.  Here the chains' states stay
// resident on the device: one call applies one proposal to every chain (in place, with an undo log), the batched
// value-only evaluation scores the proposed states, and an accept kernel keeps or restores them.
//
// Every proposal of the reference's cycle is restated (first-party code of the reference unless noted):
//   kind                         reference                                                              state touched
//   MH_SLIDE_NODE                slideNodeAtUltrametric          Proposal/Ultrametric.hs:50-96          h_j
//   MH_SCALE_SUBTREE             scaleSubTreeAtUltrametric       Proposal/Ultrametric.hs:126-186        heights of sub tree j
//   MH_PULLEY                    pulleyUltrametric               Proposal/Ultrametric.hs:228-316        all heights below the root
//   MH_SLIDE_BRACE               slideBracedNodesUltrametric     Proposal/Brace.hs:37-90                heights of the braced nodes
//   MH_SCALE_BRANCH              scaleBranch (rate tree)         Proposal/Unconstrained.hs:45-85        r_i
//   MH_SCALE_RATE_SUBTREE        scaleTree on a sub tree         Proposal/Unconstrained.hs:87-175      rates of sub tree j (with stem)
//   MH_SCALE_NORM_TREE_CONTRA_M  scaleNormAndTreeContrarily      Proposal/Unconstrained.hs:232-284      m, all rates
//   MH_SCALE_NORM_TREE_CONTRA_H  (same, on timeHeight)           app/Definitions.hs:241-253             H, all rates
//   MH_SCALE_VAR_TREE            scaleVarianceAndTree            Proposal/Unconstrained.hs:286-371      v, all rates
//   MH_SCALE_VAR_TREE_AUTO       scaleVarianceAndTreeAutocorr.   Proposal/Unconstrained.hs:381-439      v, all rates
//   MH_SLIDE_NODE_CONTRA         slideNodesAtContrarily          Proposal/Contrary.hs:35-131            h_j, r_j, rates of the children
//   MH_SCALE_SUBTREE_CONTRA      scaleSubTreesAtContrarily       Proposal/Contrary.hs:269-395           heights + rates of sub tree j
//   MH_SLIDE_BRACE_CONTRA        slideBracedNodesContrarily      Proposal/Brace.hs:98-209               braced heights + adjacent rates
//   MH_SLIDE_ROOT_CONTRA         slideRootContrarily             Proposal/Contrary.hs:173-267           H, all heights, root-child rates
//   MH_SCALE_RATES_TREE_CONTRA   scaleRatesAndTreeContrarily     Proposal/Contrary.hs:420-486           lambda, rate mean m, all heights
//   MH_SCALE_SCALAR              scaleUnbiased (`mcmc` package)  app/Definitions.hs:259-262             one of lambda, mu, H, m, v
//   MH_SCALE_H_M_CONTRA          scaleContrarily (`mcmc`)        app/Definitions.hs:244                 H, m
// Truncated-normal moves use truncatedNormalSample (Proposal/Internal.hs:107-138) on the reference's own truncated normal
// (lib/Statistics/Distribution/TruncatedNormal.hs:61-131):
//   z(m) = Phi((b-m)/s') - Phi((a-m)/s'),  quantile(p) = erfinv(2 (p z + Phi(alpha)) - 1) sqrt(2) s' + m,
//   Hastings factor q = density_{x'}(x) / density_{x}(x') = z(x) / z(x').
// Multiplier moves draw u ~ Gamma(shape k/t, scale t/k) (mean 1) and use `genericContinuous` of the un-vendored `mcmc`
// package (rev 542c43f6, restated from its published source): q = pdf(1/u) / pdf(u), Jacobian as given by the caller.
// The acceptance ratio is the package's mhgRatio: r = [prior(y) lik(y)]^beta / [prior(x) lik(x)]^beta . q . |J| (. the
// ratio of jacobianRootBranch for proposals lifted with it); accept iff ln U < ln r.
// Uniform random numbers: Philox4x32-10 (hmc_kernels.cuh), counter (chain, iteration, draw, 2).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "hmc_kernels.cuh"
#include "posterior_kernels.cuh"

// compute-sanitizer is not available on the GPU pool: build with `make EXTRA=-DMCD_CHECK` to arm bounds checks of the
// list / range bookkeeping below (device asserts), then run the MH tests.
#ifdef MCD_CHECK
#include <assert.h>
#define MCD_ASSERT(c) assert(c)
#else
#define MCD_ASSERT(c) ((void)0)
#endif

namespace mcd {

enum {
  MH_SLIDE_NODE = 0, MH_SCALE_SUBTREE = 1, MH_PULLEY = 2, MH_SLIDE_BRACE = 3, MH_SCALE_BRANCH = 4, MH_SCALE_RATE_SUBTREE = 5,
  MH_SCALE_NORM_TREE_CONTRA_M = 6, MH_SCALE_NORM_TREE_CONTRA_H = 7, MH_SCALE_VAR_TREE = 8, MH_SCALE_VAR_TREE_AUTO = 9,
  MH_SLIDE_NODE_CONTRA = 10, MH_SCALE_SUBTREE_CONTRA = 11, MH_SLIDE_BRACE_CONTRA = 12, MH_SLIDE_ROOT_CONTRA = 13,
  MH_SCALE_RATES_TREE_CONTRA = 14, MH_SCALE_SCALAR = 15, MH_SCALE_H_M_CONTRA = 16, MH_N_KINDS = 17
};
enum { MH_ST_OK = 0, MH_ST_INVALID = 1 };  // invalid: the reference would call `error` (bounds crossed, bad parameters)
enum { MH_MAX_BRACE_NODES = 16, MH_MAX_OPS = 4 * MH_MAX_BRACE_NODES + 8 };
enum { OP_SET = 0, OP_MUL = 1, OP_DIV = 2, OP_ADD = 3, OP_AFFINE_POS = 4 };

struct MhOp {
  int off, cnt, mode, pad;
  double a, b;
};
struct MhTopo {
  int N, S, n_inner_nonroot, root_r, n_brace;
  int ldyc;               // row stride of the cached y (not a power of two: the chains' rows must not alias in L2 / HBM)
  const int* parent;      // leaf flag in bit 31
  const int* child1;      // second child (the first child of an inner node i is i + 1), -1 on leaves
  const int* sub_size;    // nodes of the sub tree (= its branch labels, `length`)
  const int* sub_inner;   // nInnerNodes of the sub tree
  const int* inner_list;  // inner nodes below the root, ascending
  const int *br_off, *br_node;
};
struct MhParams {
  int kind, node, use_root_jacobian, pad;
  double param, tune;  // standard deviation (truncated-normal moves) or shape k (multiplier moves); tuning parameter
  uint64_t seed;
  uint32_t iteration;
  int chain_offset;    // global index of this handle's first chain: the Philox counter uses the global chain index, so the
                       // draws do not depend on how the chains are split over handles / GPUs
};

__device__ __forceinline__ void mh_uniform2(uint64_t seed, uint32_t chain, uint32_t iteration, uint32_t draw, double* u0, double* u1) {
  uint32_t c[4] = {chain, iteration, draw, 2u};
  Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}(c);
  *u0 = ((double)(((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6)) + 0.5) * 1.1102230246251565e-16;
  *u1 = ((double)(((uint64_t)(c[2] >> 5) << 26) | (uint64_t)(c[3] >> 6)) + 0.5) * 1.1102230246251565e-16;
}
__device__ __forceinline__ double mh_uniform(uint64_t seed, uint32_t chain, uint32_t iteration, uint32_t draw) {
  double u0, u1;
  mh_uniform2(seed, chain, iteration, draw, &u0, &u1);
  return u0;
}
__device__ __forceinline__ double mh_phi2(double x) { return 0.5 * (1.0 + erf(x * 0.70710678118654752440)); }

// truncatedNormalSample: value and ln(qYX / qXY); false where truncatedNormalDistr / the bounds check would `error`
__device__ __forceinline__ bool mh_truncated_normal(double m, double s, double a, double b, double p, double* val, double* lnq) {
  if (!(s > 0.0) || !(a < b) || (a > m) || (b < m) || !(m == m)) return false;
  const double phiA = mh_phi2((a - m) / s), z = mh_phi2((b - m) / s) - phiA;
  const double u = erfinv(2.0 * (p * z + phiA) - 1.0) * 1.41421356237309504880 * s + m;
  if (a > u || b < u || !(u == u) || !(z > 0.0)) return false;
  const double z2 = mh_phi2((b - u) / s) - mh_phi2((a - u) / s);
  *val = u;
  *lnq = log(z) - log(z2);
  return true;
}
// u ~ Gamma(shape, scale): Marsaglia & Tsang (2000); attempt i uses draws 8 + 2 i (normal, Box-Muller on the two uniforms of
// one Philox block) and 9 + 2 i (uniform); shape < 1 is boosted with draw 7.
__device__ __forceinline__ bool mh_gamma(double shape, double scale, uint64_t seed, uint32_t chain, uint32_t iteration, double* out) {
  if (!(shape > 0.0) || !(scale > 0.0)) return false;
  double boost = 1.0, a = shape;
  if (a < 1.0) {
    boost = pow(mh_uniform(seed, chain, iteration, 7u), 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t i = 0; i < 64u; ++i) {
    double u0, u1;
    mh_uniform2(seed, chain, iteration, 8u + 2u * i, &u0, &u1);
    const double x = sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);
    const double t = 1.0 + c * x;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    const double uu = mh_uniform(seed, chain, iteration, 9u + 2u * i);
    if (log(uu) < 0.5 * x * x + d - d * v + d * log(v)) {
      *out = d * v * scale * boost;
      return true;
    }
  }
  return false;
}

// The scalar part of a proposal for chain b (one thread): draws, bounds, factors, ln(q |J|), and the move as a short list
// of range operations on the state row.  Returns the number of operations (0 and *ok = false where the reference would
// call `error`).  rate_sum: sum of the rates without the stem (MH_SCALE_VAR_TREE only).
__device__ __noinline__ int mh_build_ops(const MhTopo& T, const MhParams& P, const double* row, int b, double rate_sum, MhOp* ops,
                                         bool* ok_out, int* node_out, double* lq_out) {
  const int N = T.N;
  const int OH = 3, OR = 5 + N, OM = 3 + N, OV = 4 + N;  // offsets: heights, rates, rate mean, rate variance
  const double* h = row + OH;
  int nops = 0, node = P.node;
  bool ok = true;
  double lnq = 0.0, lnj = 0.0;
  auto push = [&](int off, int cnt, int mode, double a, double bb) {
    if (cnt <= 0) return;
    MCD_ASSERT(nops < MH_MAX_OPS && off >= 0 && off + cnt <= T.S);
    ops[nops].off = off; ops[nops].cnt = cnt; ops[nops].mode = mode; ops[nops].a = a; ops[nops].b = bb;
    ++nops;
  };
  const double s = P.param * P.tune;  // sd' = t * s (Internal.hs:117)
  const uint32_t gb = (uint32_t)(P.chain_offset + b);
  const double p = mh_uniform(P.seed, gb, P.iteration, 0u);
  const int kind = P.kind;
  const bool node_kind = kind == MH_SLIDE_NODE || kind == MH_SCALE_SUBTREE || kind == MH_SCALE_RATE_SUBTREE ||
                         kind == MH_SLIDE_NODE_CONTRA || kind == MH_SCALE_SUBTREE_CONTRA;
  if (node < 0) {
    const double un = mh_uniform(P.seed, gb, P.iteration, 2u);
    if (node_kind) {
      int pick = (int)(un * (double)T.n_inner_nonroot);
      if (pick >= T.n_inner_nonroot) pick = T.n_inner_nonroot - 1;
      node = T.inner_list[pick];
    } else if (kind == MH_SCALE_BRANCH) {
      int pick = (int)(un * (double)(N - 1));
      if (pick >= N - 1) pick = N - 2;
      node = 1 + pick;
    } else if (kind == MH_SLIDE_BRACE || kind == MH_SLIDE_BRACE_CONTRA) {
      int pick = (int)(un * (double)T.n_brace);
      if (pick >= T.n_brace) pick = T.n_brace - 1;
      node = pick;
    }
  }
  const int j = node;
  // multiplier u ~ Gamma(k / t, t / k) and its Hastings factor pdf(1/u) / pdf(u)  (genericContinuous)
  double u = 1.0;
  const bool mult_kind = kind == MH_SCALE_BRANCH || kind == MH_SCALE_RATE_SUBTREE || kind == MH_SCALE_NORM_TREE_CONTRA_M ||
                         kind == MH_SCALE_NORM_TREE_CONTRA_H || kind == MH_SCALE_VAR_TREE || kind == MH_SCALE_VAR_TREE_AUTO ||
                         kind == MH_SCALE_SCALAR || kind == MH_SCALE_H_M_CONTRA;
  if (mult_kind) {
    const double kk = P.param / P.tune, th = P.tune / P.param;
    ok = mh_gamma(kk, th, P.seed, gb, P.iteration, &u);
    if (ok) lnq = -2.0 * (kk - 1.0) * log(u) - (1.0 / u - u) / th;
  }
  if (ok) switch (kind) {
    case MH_SLIDE_NODE:
    case MH_SLIDE_NODE_CONTRA: {
      const int c0 = j + 1, c1 = T.child1[j];
      const double hj = h[j], hP = h[T.parent[j] & 0x7fffffff], h0 = h[c0], h1 = h[c1];
      double hn;
      ok = mh_truncated_normal(hj, s, fmax(h0, h1), hP, p, &hn, &lnq);  // hbdMaximumChildrenHeight .. hbdParentHeight
      if (!ok) break;
      push(OH + j, 1, OP_SET, 0.0, hn);
      if (kind == MH_SLIDE_NODE_CONTRA) {  // rates scale inversely to their branches' lengths
        const double xiS = (hP - hj) / (hP - hn), xi0 = (hj - h0) / (hn - h0), xi1 = (hj - h1) / (hn - h1);
        push(OR + j, 1, OP_MUL, xiS, 0.0);
        push(OR + c0, 1, OP_MUL, xi0, 0.0);
        push(OR + c1, 1, OP_MUL, xi1, 0.0);
        lnj = (log(xi0) + log(xi1)) + log(xiS);
      }
    } break;
    case MH_SCALE_SUBTREE:
    case MH_SCALE_SUBTREE_CONTRA: {
      const double hj = h[j], hP = h[T.parent[j] & 0x7fffffff];
      double hn;
      ok = mh_truncated_normal(hj, s, 0.0, hP, p, &hn, &lnq);
      if (!ok) break;
      const double xi = hn / hj;
      const int cnt = T.sub_size[j];
      push(OH + j, 1, OP_SET, 0.0, hn);  // the sub tree's root gets the sampled height itself (scaleUltrametricTreeF)
      push(OH + j + 1, cnt - 1, OP_MUL, xi, 0.0);
      if (kind == MH_SCALE_SUBTREE) {
        lnj = (double)(T.sub_inner[j] - 1) * log(xi);
      } else {
        const double xiR = 1.0 / xi, xiS = (hP - hj) / (hP - hn);
        push(OR + j + 1, cnt - 1, OP_MUL, xiR, 0.0);
        push(OR + j, 1, OP_MUL, xiS, 0.0);
        lnj = (double)(T.sub_inner[j] - cnt) * log(xi) + log(xiS);
      }
    } break;
    case MH_PULLEY: {
      const int l = 1, r = T.root_r;
      const double ht = h[0], hL = h[l], hR = h[r], brL = ht - hL, brR = ht - hR;
      if (!(brL > 0.0) || !(brR > 0.0)) { ok = false; break; }
      const double a = -fmin(brL, ht - brR), bb = fmin(brR, ht - brL);
      double uu;
      ok = mh_truncated_normal(0.0, s, a, bb, p, &uu, &lnq);
      if (!ok) break;
      const double hLn = hL - uu, hRn = hR + uu, xiL = hLn / hL, xiR = hRn / hR;
      push(OH + l, 1, OP_SET, 0.0, hLn);
      push(OH + l + 1, T.sub_size[l] - 1, OP_MUL, xiL, 0.0);
      push(OH + r, 1, OP_SET, 0.0, hRn);
      push(OH + r + 1, T.sub_size[r] - 1, OP_MUL, xiR, 0.0);
      lnj = (double)(T.sub_inner[l] - 1) * log(xiL) + (double)(T.sub_inner[r] - 1) * log(xiR);
    } break;
    case MH_SLIDE_BRACE:
    case MH_SLIDE_BRACE_CONTRA: {
      const int o0 = T.br_off[j], o1 = T.br_off[j + 1];
      double lo = -CUDART_INF, hi = CUDART_INF;
      for (int o = o0; o < o1; ++o) {
        const int x = T.br_node[o];
        const double hx = h[x];
        lo = fmax(lo, fmax(h[x + 1], h[T.child1[x]]) - hx);
        hi = fmin(hi, h[T.parent[x] & 0x7fffffff] - hx);
      }
      double dl;
      ok = mh_truncated_normal(0.0, s, lo, hi, p, &dl, &lnq);
      if (!ok) break;
      for (int o = o0; o < o1; ++o) push(OH + T.br_node[o], 1, OP_ADD, 0.0, dl);
      if (kind == MH_SLIDE_BRACE_CONTRA) {
        for (int o = o0; o < o1; ++o) {
          const int x = T.br_node[o], c0 = x + 1, c1 = T.child1[x];
          const double hx = h[x], hP = h[T.parent[x] & 0x7fffffff];
          const double xiS = (hP - hx) / (hP - hx - dl), xi0 = (hx - h[c0]) / (hx + dl - h[c0]), xi1 = (hx - h[c1]) / (hx + dl - h[c1]);
          push(OR + x, 1, OP_MUL, xiS, 0.0);
          push(OR + c0, 1, OP_MUL, xi0, 0.0);
          push(OR + c1, 1, OP_MUL, xi1, 0.0);
          lnj += (log(xi0) + log(xi1)) + log(xiS);
        }
      }
    } break;
    case MH_SCALE_BRANCH:
      push(OR + j, 1, OP_MUL, u, 0.0);
      lnj = -log(u);  // scaleUnbiased: Jacobian 1 / u
      break;
    case MH_SCALE_RATE_SUBTREE:
      push(OR + j, T.sub_size[j], OP_MUL, u, 0.0);  // stem included (scaleUnconstrainedTreeF)
      lnj = (double)(T.sub_size[j] - 2) * log(u);
      break;
    case MH_SCALE_NORM_TREE_CONTRA_M:
    case MH_SCALE_NORM_TREE_CONTRA_H:
      push(kind == MH_SCALE_NORM_TREE_CONTRA_M ? OM : 2, 1, OP_DIV, u, 0.0);
      push(OR + 1, N - 1, OP_MUL, u, 0.0);  // without the stem
      lnj = (double)(N - 1 - 3) * log(u);
      break;
    case MH_SCALE_VAR_TREE: {
      const double n = (double)(N - 1), n1 = 1.0 / n, mean = rate_sum / n;
      push(OV, 1, OP_MUL, u * u, 0.0);
      push(OR + 1, N - 1, OP_AFFINE_POS, u, mean);
      lnj = n * log(u - n1 * u + n1);
    } break;
    case MH_SCALE_VAR_TREE_AUTO:
      // r' = y_parent + u (r - r_parent) down the tree telescopes to m + u (r - m): closed form of scaleF
      push(OV, 1, OP_MUL, u * u, 0.0);
      push(OR + 1, N - 1, OP_AFFINE_POS, u, row[OM]);
      lnj = (double)(N - 1) * log(u);
      break;
    case MH_SLIDE_ROOT_CONTRA: {
      const int l = 1, r = T.root_r;
      const double H = row[2], hL = h[l], hR = h[r];
      if (fabs(h[0] - 1.0) > 1e-14) { ok = false; break; }
      double Hn;
      ok = mh_truncated_normal(H, s, H * fmax(hL, hR), CUDART_INF, p, &Hn, &lnq);
      if (!ok) break;
      const double uu = Hn / H, xiL = (1.0 - hL) / (uu - hL), xiR = (1.0 - hR) / (uu - hR);
      push(2, 1, OP_SET, 0.0, Hn);
      push(OH + 1, N - 1, OP_DIV, uu, 0.0);
      push(OR + l, 1, OP_MUL, xiL, 0.0);
      push(OR + r, 1, OP_MUL, xiR, 0.0);
      lnj = -(double)T.sub_inner[0] * log(uu) + (log(xiL) + log(xiR));
    } break;
    case MH_SCALE_RATES_TREE_CONTRA: {
      const int nn = T.sub_inner[0] - 1;
      if (nn < 1) { ok = false; break; }
      const double hc = fmax(h[1], h[T.root_r]);
      double hn;
      ok = mh_truncated_normal(hc, s, 0.0, h[0], p, &hn, &lnq);
      if (!ok) break;
      const double xi = hn / hc;
      push(OH + 1, N - 1, OP_MUL, xi, 0.0);
      push(0, 1, OP_DIV, xi, 0.0);   // timeBirthRate and rateMean: ratesTimeTreeL = tripleLens timeBirthRate rateMean timeTree
      push(OM, 1, OP_DIV, xi, 0.0);  // (app/Definitions.hs:236-237, 263); the PFunction's `mu` is the mean rate (Contrary.hs:435-436)
      lnj = (double)(nn - 1 - 2) * log(xi);
    } break;
    case MH_SCALE_SCALAR: {
      const int off = j == 0 ? 0 : j == 1 ? 1 : j == 2 ? 2 : j == 3 ? OM : OV;
      push(off, 1, OP_MUL, u, 0.0);
      lnj = -log(u);
    } break;
    case MH_SCALE_H_M_CONTRA:
      push(2, 1, OP_MUL, u, 0.0);
      push(OM, 1, OP_DIV, u, 0.0);
      lnj = -log(u * u);
      break;
    default: ok = false; break;
  }
  if (!ok) nops = 0;
  *ok_out = ok;
  *node_out = node;
  *lq_out = ok ? lnq + lnj : 0.0;
  return nops;
}

// Apply proposal P to every chain IN PLACE.  One CTA per chain: thread 0 draws and turns the move into a short list of
// range operations on the state row, then all threads save the old values to the undo log and apply them.
// undo[b][..]: old values, ranges packed in order; rng[b][o] = (offset, count); meta[b] = (ranges, validity, node, -);
// lq[b] = ln(q |J|).
__global__ void __launch_bounds__(256)
mh_propose_kernel(double* __restrict__ states, double* __restrict__ undo, int2* __restrict__ rng, int4* __restrict__ meta,
                  double* __restrict__ lq, const MhTopo T, const MhParams P, int undo_stride, int B) {
  __shared__ MhOp ops[MH_MAX_OPS];
  __shared__ int sh_nops;
  __shared__ double sh_warp[8], sh_sum;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= B) return;
  const int N = T.N, S = T.S;
  double* row = states + (size_t)b * S;
  const int OR = 5 + N;
  if (tid == 0) sh_sum = 0.0;
  const int nthr = blockDim.x;  // 64: the serial part of a proposal is latency, so many chains per SM beat many threads per chain
  if (P.kind == MH_SCALE_VAR_TREE) {  // sample mean of the rates without the stem (scaleVarianceAndTreeF)
    double s = 0.0;
    for (int i = 1 + tid; i < N; i += nthr) s += row[OR + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) sh_warp[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (nthr >> 5); ++w) t += sh_warp[w];
      sh_sum = t;
    }
    __syncthreads();
  }
  if (tid == 0) {
    bool ok;
    int node;
    double lqv;
    const int nops = mh_build_ops(T, P, row, b, sh_sum, ops, &ok, &node, &lqv);
    sh_nops = nops;
    meta[b] = make_int4(nops, ok ? MH_ST_OK : MH_ST_INVALID, node, 0);
    lq[b] = lqv;
  }
  __syncthreads();
  const int nops = sh_nops;
  double* ub = undo + (size_t)b * undo_stride;
  int pos = 0;
  for (int o = 0; o < nops; ++o) {
    const MhOp op = ops[o];
    for (int i = tid; i < op.cnt; i += nthr) {
      MCD_ASSERT(pos + i < undo_stride);
      const double old = row[op.off + i];
      ub[pos + i] = old;
      double y;
      switch (op.mode) {
        case OP_SET: y = op.b; break;
        case OP_MUL: y = old * op.a; break;
        case OP_DIV: y = old / op.a; break;
        case OP_ADD: y = old + op.b; break;
        default: y = (old - op.b) * op.a + op.b; if (!(y > 0.0)) y = CUDART_NAN; break;  // "force NaN when the new value is negative"
      }
      row[op.off + i] = y;
    }
    pos += op.cnt;
    __syncthreads();  // ranges may overlap (a braced node that is the child of another): strictly in order
  }
  for (int o = tid; o < nops; o += nthr) rng[(size_t)b * MH_MAX_OPS + o] = make_int2(ops[o].off, ops[o].cnt);
}

// ------------------------------------------------------------------------------------------------------------------
// Incremental evaluation of small moves.  The chains are resident, so the contraction result y = Sigma^-1 (d - mu) of
// every chain's CURRENT state can be kept in HBM.  A proposal that changes the residual by delta on a few branches then
// needs no contraction at all:
//     quad' = quad + 2 delta . y + delta^T Sigma^-1 delta       (|A| values of y, |A|^2 entries of Sigma^-1)
// and every prior part is a sum over nodes / branches in which only the touched terms change.  mh_delta_kernel turns
// the undo log of mh_propose_kernel into the ln-posterior parts of the proposed state in O(|A|^2) work per chain instead
// of 2 K^2 flops; on acceptance mh_accept_kernel updates y with |A| rows of Sigma^-1 (which stay in the 126 MB L2).
// Values differ from a fresh evaluation by accumulated rounding only (the host re-evaluates from scratch every
// `refresh` steps); a chain in the near-critical birth-death regime re-runs the literal D/E recursion (lane 0).
enum { DL_MAX_CHG = 96, DL_MAX_AB = 128, DL_SUBTREE_H = 32, DL_SUBTREE_R = 64 };
struct MhYUpdate {
  int mode;  // 0: no cached y; 1: rank-|A| update from the delta list; 2: copy the freshly contracted row
  int K, ldk, ldy, ldyc;  // ldy: stride of y_new (contraction output), ldyc: stride of y_cur
  double* y_cur;
  const double* y_new;
  const double* P;
  const int* dl_n;
  const int* dl_k;
  const double* dl_d;
};

template <int CLOCK>
__device__ __forceinline__ double mh_clock_term(double r, double t, double v, double lgk_v, double ln_v) {
  const double lnr = log(r);
  if (CLOCK == 0) {  // uncorrelatedGamma: k = 1/v, theta = v
    const double k = 1.0 / v;
    return lnr * (k - 1.0) - r * k - lgk_v - ln_v * k;
  } else if (CLOCK == 2) {  // white noise: k = t/v, theta = v/t
    const double k = t / v;
    return lnr * (k - 1.0) - r * k - lgamma(k) + log(k) * k;
  } else {  // logNormal' 1 w r, w = v | v t
    const double w = CLOCK == 1 ? v : v * t;
    const double bb = lnr + 0.5 * w;
    return -(MCD_LN_SQRT_2PI + lnr + 0.5 * (CLOCK == 1 ? ln_v : log(w))) - 0.5 / w * bb * bb;
  }
}

// y[0..ldk) += sum_a d[a] P[k[a]][0..ldk) by a group of G threads.  Rows of P (ldk a multiple of 16, zero padded) and y are
// 16-byte aligned: 128-bit accesses, four of them in flight per thread and row.  Measured on the benchmark shape this
// update is 2/3 of a small-move step (y: HBM read-modify-write, rows of P: L2) -- issuing the loads of four rows together
// was slower (redundant row traffic when fewer than four branches changed).
template <int G>
__device__ __forceinline__ void mh_update_y(double* y, const double* __restrict__ P, int ldk, const int* k, const double* d, int n,
                                            int lane) {
  constexpr int U = 4;
  const int n2 = ldk >> 1;
  double2* y2 = reinterpret_cast<double2*>(y);
  for (int c0 = lane; c0 < n2; c0 += G * U) {
    double2 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = c0 + u * G < n2 ? y2[c0 + u * G] : make_double2(0.0, 0.0);
    for (int a = 0; a < n; ++a) {
      const double2* row = reinterpret_cast<const double2*>(P + (size_t)k[a] * ldk);
      const double da = d[a];
      double2 p[U];
#pragma unroll
      for (int u = 0; u < U; ++u) p[u] = c0 + u * G < n2 ? __ldg(row + c0 + u * G) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u].x = fma(da, p[u].x, acc[u].x);
        acc[u].y = fma(da, p[u].y, acc[u].y);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c0 + u * G < n2) y2[c0 + u * G] = acc[u];
  }
}

// The evaluation proper, shared by mh_delta_kernel and the fused small-move kernel: one warp, the raw change list
// (offset, old value; first record of an offset = its old value) in shared memory, the state row already modified.
// Results: ln-posterior parts of the proposed state in o1[0..7] and its status (lane 0), the residual changes
// (abk, abd)[0..n_ab) for the update of y.
template <int CLOCK>
__device__ __forceinline__ void mh_delta_core(const DevModel& M, const MhTopo& T, const double* __restrict__ P, const double* row,
                                              const int* c_off, const double* c_old, int n_raw, unsigned* bm, int* ab, int* abk,
                                              double* abd, const double* __restrict__ y, const double* o0, int cur_st, int lane,
                                              double* o1, int* st_out, int* n_ab_out) {
  const int N = M.N, OH = 3, OR = 5 + N;
  const int nwords = (N + 31) >> 5;
  for (int w = lane; w < nwords; w += 32) bm[w] = 0u;
  __syncwarp();
  auto old_val = [&](int off) -> double {
    for (int e = 0; e < n_raw; ++e)
      if (c_off[e] == off) return c_old[e];
    return row[off];
  };
  auto changed = [&](int off) -> bool {
    for (int e = 0; e < n_raw; ++e)
      if (c_off[e] == off) return true;
    return false;
  };
  // 2. affected branches: a changed height moves the node's own branch and its children's, a changed rate its own
  for (int e = lane; e < n_raw; e += 32) {
    const int off = c_off[e];
    if (off > OH && off < OH + N) {
      const int x = off - OH;
      atomicOr(&bm[x >> 5], 1u << (x & 31));
      if (T.parent[x] >= 0) {  // inner node
        const int c0 = x + 1, c1 = T.child1[x];
        atomicOr(&bm[c0 >> 5], 1u << (c0 & 31));
        atomicOr(&bm[c1 >> 5], 1u << (c1 & 31));
      }
    } else if (off > OR && off < OR + N) {
      const int x = off - OR;
      atomicOr(&bm[x >> 5], 1u << (x & 31));
    }
  }
  __syncwarp();
  int n_ab = 0;  // ascending node order: the sums below do not depend on the order of the atomics
  for (int w0 = 0; w0 < nwords; w0 += 32) {
    const int w = w0 + lane;
    unsigned word = w < nwords ? bm[w] : 0u;
    const int cnt = __popc(word);
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int at = n_ab + pre - cnt;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      if (at < DL_MAX_AB) ab[at] = (w << 5) + bit;
      ++at;
    }
    n_ab += total;
  }
  MCD_ASSERT(n_ab <= DL_MAX_AB);
  if (n_ab > DL_MAX_AB) n_ab = DL_MAX_AB;
  __syncwarp();
  // 3. per-branch residual and clock-prior changes
  const double la = row[0], mu = row[1], H = row[2], m = row[3 + N], v = row[4 + N];
  const double sc = H * m;
  const double lgk_v = CLOCK == 0 ? lgamma(1.0 / v) : 0.0, ln_v = log(v);
  bool bad = false;
  double d_clock = 0.0, q1 = 0.0;
  for (int a = lane; a < n_ab; a += 32) {
    const int i = ab[a], p = T.parent[i] & 0x7fffffff;
    MCD_ASSERT(i >= 1 && i < N && p >= 0 && p < N);
    const double hi_n = row[OH + i], hp_n = row[OH + p], r_n = row[OR + i];
    const double hi_o = old_val(OH + i), hp_o = old_val(OH + p), r_o = old_val(OR + i);
    const double t_n = hp_n - hi_n, t_o = hp_o - hi_o;
    const bool ok = (t_n > 0.0) && (r_n > 0.0);
    bad = bad || !ok;
    const int k = branch_of(i, M.root_r);
    MCD_ASSERT(k >= 0 && k < M.K);
    const double d = (t_n * r_n) * sc - (t_o * r_o) * sc;
    abk[a] = k;
    abd[a] = d;
    q1 = fma(d, y[k], q1);
    if (ok) d_clock += mh_clock_term<CLOCK>(r_n, t_n, v, lgk_v, ln_v) - mh_clock_term<CLOCK>(r_o, t_o, v, lgk_v, ln_v);
  }
  __syncwarp();
  // 4. delta^T Sigma^-1 delta
  double q2 = 0.0;
  for (int idx = lane; idx < n_ab * n_ab; idx += 32) {
    const int a = idx / n_ab, bb = idx - a * n_ab;
    q2 = fma(abd[a] * abd[bb], P[(size_t)abk[a] * M.ldk + abk[bb]], q2);
  }
  // 5. birth-death and node priors of the nodes whose height changed
  const bool nearcrit = 1e-6 > fabs(la - mu);
  const bool bd_series = fabs(la - mu) * fmax(1.0, fabs(row[OH])) < 0.25;
  const double* h = row + OH;
  double d_bd = 0.0, d_A = 0.0;
  for (int e = lane; e < n_raw; e += 32) {
    const int off = c_off[e];
    if (!(off > OH && off < OH + N)) continue;
    bool first = true;
    for (int e2 = 0; e2 < e; ++e2) first = first && (c_off[e2] != off);
    if (!first) continue;
    const int x = off - OH;
    const double h_n = row[off], h_o = c_old[e];
    if (T.parent[x] >= 0 && !nearcrit) d_bd += ln_p1<false>(la, mu, h_n, bd_series).v - ln_p1<false>(la, mu, h_o, bd_series).v;
    for (int q = M.inc_off[x]; q < M.inc_off[x + 1]; ++q) {
      const int2 ent = M.inc_ent[q];
      if (ent.x == INC_CAL) {
        double dh, dH;
        int f = 0;
        d_A += calibration_term(M, ent.y, H, h_n, &dh, &dH, &f) - calibration_term(M, ent.y, H, h_o, &dh, &dH, &f);
      } else if (ent.x == INC_BRACE) {
        const int j0 = M.br_off[ent.y], j1 = M.br_off[ent.y + 1];
        bool owner = true;  // the changed node with the smallest index accounts for the brace
        for (int j = j0; j < j1; ++j) owner = owner && !(M.br_node[j] < x && changed(OH + M.br_node[j]));
        if (!owner) continue;
        const double sd = M.br_sd[ent.y];
        double vn = 0.0, vo = 0.0, sn = 0.0, so = 0.0;
        bool eqn = true, eqo = true;
        const double hn0 = h[M.br_node[j0]], ho0 = old_val(OH + M.br_node[j0]);
        for (int j = j0; j < j1; ++j) {
          const double a_n = h[M.br_node[j]], a_o = old_val(OH + M.br_node[j]);
          eqn = eqn && (a_n == hn0);
          eqo = eqo && (a_o == ho0);
          sn += a_n;
          so += a_o;
        }
        const double mn = sn / (double)(j1 - j0), mo = so / (double)(j1 - j0);
        for (int j = j0; j < j1; ++j) {
          const double dn = h[M.br_node[j]] - mn, d_o = old_val(OH + M.br_node[j]) - mo;
          vn += -(dn * dn) / (2.0 * sd * sd);
          vo += -(d_o * d_o) / (2.0 * sd * sd);
        }
        d_A += (eqn ? 0.0 : vn) - (eqo ? 0.0 : vo);
      } else {  // constraint: the younger-index changed node of the pair accounts for it
        const int ny = M.con_y[ent.y], no = M.con_o[ent.y], other = ny == x ? no : ny;
        if (other < x && changed(OH + other)) continue;
        const double s = M.con_s[ent.y];
        const double hYn = h[ny], hOn = h[no], hYo = old_val(OH + ny), hOo = old_val(OH + no);
        const double tn = (hYn < hOn) ? 0.0 : -((hYn - hOn) * (hYn - hOn)) / (2.0 * s * s);
        const double to = (hYo < hOo) ? 0.0 : -((hYo - hOo) * (hYo - hOo)) / (2.0 * s * s);
        d_A += tn - to;
      }
    }
  }
  // fixed shuffle trees: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d_clock += __shfl_xor_sync(0xffffffffu, d_clock, o);
    q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    d_bd += __shfl_xor_sync(0xffffffffu, d_bd, o);
    d_A += __shfl_xor_sync(0xffffffffu, d_A, o);
  }
  bad = __any_sync(0xffffffffu, bad);
  *n_ab_out = n_ab;
  if (lane == 0) {
    const double NINF = -CUDART_INF;
    int st = cur_st & ST_NEARCRIT;
    if (bad) {  // a non-positive branch or rate: probability zero (or NaN) -- rejected either way
      o1[0] = o0[0]; o1[1] = NINF; o1[2] = NINF; o1[3] = NINF; o1[4] = o0[4]; o1[5] = o0[5]; o1[6] = NINF; o1[7] = 0.0;
      st |= ST_ZERO;
    } else {
      double lnB = o0[1] + d_bd;
      if (nearcrit) {  // literal D/E recursion on the proposed state (BirthDeath.hs:90-114), as in the posterior kernel
        const double d = la - mu;
        double E = 0.0, bd = 0.0;
        for (int i = N - 1; i >= 1; --i) {
          const int pe = T.parent[i];
          const bool inner = pe >= 0;
          const double ti = h[pe & 0x7fffffff] - h[i];
          const double c = inner ? E : 0.0;
          const double yy = (mu - c * la) * ti, den = 1.0 + yy;
          const double D = (1.0 - d * ti) / den / den;
          E = (c + yy) / den;
          bd += log(D * (inner ? la : 1.0));
        }
        lnB = (0.0 - la) + (0.0 - mu) + bd;
      }
      const double lnA = o0[0] + d_A, lnC = o0[2] + d_clock;
      const double prior = lnA + lnB + lnC;
      const double lk = o0[4] + (-0.5) * (2.0 * q1 + q2);
      const double d0 = ((h[0] - h[1]) * row[OR + 1] + (h[0] - h[M.root_r]) * row[OR + M.root_r]) * sc;
      const double jac = log(1.0 / d0);
      const double post = prior + lk + jac;
      if (post == NINF) st |= ST_ZERO;
      if (post != post) st |= ST_NAN;
      o1[0] = lnA; o1[1] = lnB; o1[2] = lnC; o1[3] = prior; o1[4] = lk; o1[5] = jac; o1[6] = post; o1[7] = 0.0;
    }
    *st_out = st;
  }
}

// One warp per chain, eight chains per CTA.  Dynamic shared memory per warp: the bitmap of affected branches.
template <int CLOCK>
__global__ void __launch_bounds__(256)
mh_delta_kernel(const DevModel M, const MhTopo T, const double* __restrict__ P, const double* __restrict__ states,
                const double* __restrict__ undo, const int2* __restrict__ rng, const int4* __restrict__ meta,
                const double* __restrict__ y_cur, const double* __restrict__ cur_out, const int* __restrict__ cur_status,
                double* __restrict__ new_out, int* __restrict__ new_status, int* __restrict__ dl_n, int* __restrict__ dl_k,
                double* __restrict__ dl_d, int undo_stride, int B) {
  extern __shared__ unsigned dl_bitmaps[];
  __shared__ int s_off[8][DL_MAX_CHG];
  __shared__ double s_old[8][DL_MAX_CHG];
  __shared__ int s_ab[8][DL_MAX_AB], s_k[8][DL_MAX_AB];
  __shared__ double s_d[8][DL_MAX_AB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chain = blockIdx.x * 8 + warp;
  if (chain >= B) return;
  const int4 mt = meta[chain];
  if (mt.y != MH_ST_OK || mt.x == 0) {  // rejected by the accept kernel whatever new_out says
    if (lane == 0) dl_n[chain] = 0;
    return;
  }
  const int nwords = (M.N + 31) >> 5;
  unsigned* bm = dl_bitmaps + (size_t)warp * nwords;
  int* c_off = s_off[warp];
  double* c_old = s_old[warp];
  int* ab = s_ab[warp];
  int* abk = s_k[warp];
  double* abd = s_d[warp];
  const double* row = states + (size_t)chain * M.S;
  // 1. raw change list from the undo log (an entry touched twice appears twice: the FIRST record holds the old value)
  int n_raw = 0;
  {
    const double* ub = undo + (size_t)chain * undo_stride;
    for (int o = 0; o < mt.x; ++o) {
      const int2 r = rng[(size_t)chain * MH_MAX_OPS + o];
      for (int i = lane; i < r.y && n_raw + i < DL_MAX_CHG; i += 32) {
        c_off[n_raw + i] = r.x + i;
        c_old[n_raw + i] = ub[n_raw + i];
      }
      n_raw += r.y;
    }
    MCD_ASSERT(n_raw <= DL_MAX_CHG);
    if (n_raw > DL_MAX_CHG) n_raw = DL_MAX_CHG;  // the host only routes small moves here
  }
  double o1[8];
  int st = 0, n_ab = 0;
  mh_delta_core<CLOCK>(M, T, P, row, c_off, c_old, n_raw, bm, ab, abk, abd, y_cur + (size_t)chain * T.ldyc,
                       cur_out + (size_t)chain * 8, cur_status[chain], lane, o1, &st, &n_ab);
  for (int a = lane; a < n_ab; a += 32) {
    dl_k[(size_t)chain * DL_MAX_AB + a] = abk[a];
    dl_d[(size_t)chain * DL_MAX_AB + a] = abd[a];
  }
  if (lane == 0) {
    dl_n[chain] = n_ab;
    for (int j = 0; j < 8; ++j) new_out[(size_t)chain * 8 + j] = o1[j];
    new_status[chain] = st;
  }
}

// ln r = beta_p (ln prior(y) - ln prior(x)) + beta_l (ln lik(y) - ln lik(x)) + ln(q |J|) (+ the change of the root-branch
// Jacobian for proposals lifted with jacobianRootBranch, app/Definitions.hs:139-150); accept iff ln U < ln r.
// beta: heat of the chain's current temperature slot (MC3: prior and likelihood; stepping stone: likelihood only);
// slot == nullptr: cold chains.  Rejected chains get their values back from the undo log, ranges in reverse order.
// counters[0] += accepted, counters[1] += invalid (nullable).  One CTA per chain.
__global__ void __launch_bounds__(256)
mh_accept_kernel(double* __restrict__ states, const double* __restrict__ undo, const int2* __restrict__ rng,
                 const int4* __restrict__ meta, const double* __restrict__ lq, double* __restrict__ cur_out,
                 const double* __restrict__ new_out, int* __restrict__ cur_status, const int* __restrict__ new_status,
                 int* __restrict__ accepted, unsigned long long* __restrict__ counters, const int* __restrict__ slot,
                 const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int chain_offset,
                 int use_root_jacobian, uint64_t seed, uint32_t iteration, int S, int undo_stride, int B, const MhYUpdate Y) {
  __shared__ int sh_acc;
  __shared__ int2 sh_rng[MH_MAX_OPS];
  __shared__ int sh_k[DL_MAX_AB];
  __shared__ double sh_d[DL_MAX_AB];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= B) return;
  const int4 m = meta[b];
  if (tid < m.x) sh_rng[tid] = rng[(size_t)b * MH_MAX_OPS + tid];
  if (tid == 0) {
    int acc = 0;
    if (m.y == MH_ST_OK) {
      const double* o1 = new_out + (size_t)b * 8;
      const double* o0 = cur_out + (size_t)b * 8;
      double bp = 1.0, bl = 1.0;
      if (slot) {
        const int sl = slot[chain_offset + b];
        bp = ladder_prior[sl];
        bl = ladder_lik[sl];
      }
      double lr = bp * (o1[3] - o0[3]) + bl * (o1[4] - o0[4]) + lq[b];
      if (use_root_jacobian) lr += o1[5] - o0[5];
      const double u = mh_uniform(seed, (uint32_t)(chain_offset + b), iteration, 1u);
      acc = log(u) < lr;  // false for NaN and for -inf
    }
    sh_acc = acc;
    if (accepted) accepted[b] = m.y == MH_ST_INVALID ? -1 : acc;
    if (counters) {
      if (acc) atomicAdd(counters, 1ull);
      if (m.y == MH_ST_INVALID) atomicAdd(counters + 1, 1ull);
    }
  }
  __syncthreads();
  if (sh_acc) {
    if (tid < 8) cur_out[(size_t)b * 8 + tid] = new_out[(size_t)b * 8 + tid];
    if (tid == 8) cur_status[b] = new_status[b];
    // keep the cached y = Sigma^-1 dx of the accepted state (incremental evaluation, mh_delta_kernel)
    if (Y.mode == 2) {  // full evaluation: the contraction just produced it
      const double* yn = Y.y_new + (size_t)b * Y.ldy;
      double* yc = Y.y_cur + (size_t)b * Y.ldyc;
      for (int k = tid; k < Y.K; k += 256) yc[k] = yn[k];
    } else if (Y.mode == 1) {  // y += sum_a delta_a P[k_a][:]   (P symmetric: rows, coalesced, L2-resident)
      const int na = Y.dl_n[b];
      for (int a = tid; a < na; a += 256) {
        sh_k[a] = Y.dl_k[(size_t)b * DL_MAX_AB + a];
        sh_d[a] = Y.dl_d[(size_t)b * DL_MAX_AB + a];
      }
      __syncthreads();
      mh_update_y<256>(Y.y_cur + (size_t)b * Y.ldyc, Y.P, Y.ldk, sh_k, sh_d, na, tid);
    }
  } else if (m.x > 0) {
    double* row = states + (size_t)b * S;
    const double* ub = undo + (size_t)b * undo_stride;
    int pos = 0;
    for (int o = 0; o < m.x; ++o) pos += sh_rng[o].y;
    for (int o = m.x - 1; o >= 0; --o) {
      const int2 r = sh_rng[o];
      pos -= r.y;
      for (int i = tid; i < r.y; i += 256) row[r.x + i] = ub[pos + i];
      __syncthreads();
    }
  }
}

// Large sub-tree moves on a given node j (scale sub tree, its contrary form, scale rate sub tree): the residual and every
// prior term change inside the sub tree [j, j + size) only, and the rank-limited contraction has already produced
// y' = y + Sigma^-1[:, A] delta_A in y_new.  One CTA per chain scores the proposed state in O(size):
//   quad' = quad + delta . (y + y')   (= 2 delta.y + delta^T Sigma^-1 delta),
// clock terms of the sub tree's branches, ln p1 of its inner nodes, the node priors incident to its nodes (an entry shared
// by several moved nodes is accounted for by the one with the smallest index).  Old values come from the undo log by
// direct indexing (layout: see delta_split_kernel).  mode: 0 heights, 1 heights + rates (contrary), 2 rates.
template <int CLOCK>
__global__ void __launch_bounds__(256)
mh_range_delta_kernel(const DevModel M, const MhTopo T, const double* __restrict__ states, const double* __restrict__ undo,
                      int undo_stride, const int4* __restrict__ meta, const double* __restrict__ y_cur,
                      const double* __restrict__ y_new, const double* __restrict__ cur_out, const int* __restrict__ cur_status,
                      double* __restrict__ new_out, int* __restrict__ new_status, int mode, int j, int size, int B) {
  __shared__ double s_red[8][4];
  __shared__ int s_bad[8];
  const int chain = blockIdx.x, tid = threadIdx.x;
  if (chain >= B) return;
  if (meta[chain].x == 0) return;  // invalid proposal: the accept kernel leaves the chain where it is
  const int N = M.N, OH = 3, OR = 5 + N;
  const double* row = states + (size_t)chain * M.S;
  const double* h = row + OH;
  const double* r = row + OR;
  const double* ub = undo + (size_t)chain * undo_stride;
  const double* yc = y_cur + (size_t)chain * T.ldyc;
  const double* yn = y_new + (size_t)chain * M.ldy;
  const double la = row[0], mu = row[1], H = row[2], m = row[3 + N], v = row[4 + N];
  const double sc = H * m;
  const double lgk_v = CLOCK == 0 ? lgamma(1.0 / v) : 0.0, ln_v = log(v);
  const bool nearcrit = 1e-6 > fabs(la - mu);
  const bool bd_series = fabs(la - mu) * fmax(1.0, fabs(h[0])) < 0.25;
  const bool hmove = mode != 2;
  auto old_h = [&](int x) -> double { return hmove && x >= j && x < j + size ? ub[x - j] : h[x]; };
  double q = 0.0, d_clock = 0.0, d_bd = 0.0, d_A = 0.0;
  bool bad = false;
  for (int i = j + tid; i < j + size; i += blockDim.x) {
    const int pe = T.parent[i], p = pe & 0x7fffffff;
    MCD_ASSERT(i < N && p < i && (p >= j || i == j) && 2 * size - 1 < undo_stride);
    const double hi_n = h[i], hp_n = h[p], r_n = r[i];
    const double hi_o = old_h(i), hp_o = old_h(p);
    const double r_o = mode == 0 ? r_n : mode == 2 ? ub[i - j] : (i > j ? ub[size + (i - j - 1)] : ub[2 * size - 1]);
    const double t_n = hp_n - hi_n, t_o = hp_o - hi_o;
    const bool ok = (t_n > 0.0) && (r_n > 0.0);
    bad = bad || !ok;
    const int k = branch_of(i, M.root_r);
    const double d = (t_n * r_n) * sc - (t_o * r_o) * sc;
    q = fma(d, yc[k] + yn[k], q);
    if (ok) d_clock += mh_clock_term<CLOCK>(r_n, t_n, v, lgk_v, ln_v) - mh_clock_term<CLOCK>(r_o, t_o, v, lgk_v, ln_v);
    if (hmove) {
      if (pe >= 0 && !nearcrit) d_bd += ln_p1<false>(la, mu, hi_n, bd_series).v - ln_p1<false>(la, mu, hi_o, bd_series).v;
      for (int e = M.inc_off[i]; e < M.inc_off[i + 1]; ++e) {
        const int2 ent = M.inc_ent[e];
        if (ent.x == INC_CAL) {
          double dh, dH;
          int f = 0;
          d_A += calibration_term(M, ent.y, H, hi_n, &dh, &dH, &f) - calibration_term(M, ent.y, H, hi_o, &dh, &dH, &f);
        } else if (ent.x == INC_BRACE) {
          const int j0 = M.br_off[ent.y], j1 = M.br_off[ent.y + 1];
          bool owner = true;
          for (int e2 = j0; e2 < j1; ++e2) owner = owner && !(M.br_node[e2] < i && M.br_node[e2] >= j);
          if (!owner) continue;
          const double sd = M.br_sd[ent.y];
          double sn = 0.0, so = 0.0;
          bool eqn = true, eqo = true;
          const double hn0 = h[M.br_node[j0]], ho0 = old_h(M.br_node[j0]);
          for (int e2 = j0; e2 < j1; ++e2) {
            const double a_n = h[M.br_node[e2]], a_o = old_h(M.br_node[e2]);
            eqn = eqn && (a_n == hn0);
            eqo = eqo && (a_o == ho0);
            sn += a_n;
            so += a_o;
          }
          const double mn = sn / (double)(j1 - j0), mo = so / (double)(j1 - j0);
          double vn = 0.0, vo = 0.0;
          for (int e2 = j0; e2 < j1; ++e2) {
            const double dn = h[M.br_node[e2]] - mn, d_o = old_h(M.br_node[e2]) - mo;
            vn += -(dn * dn) / (2.0 * sd * sd);
            vo += -(d_o * d_o) / (2.0 * sd * sd);
          }
          d_A += (eqn ? 0.0 : vn) - (eqo ? 0.0 : vo);
        } else {
          const int ny = M.con_y[ent.y], no = M.con_o[ent.y], other = ny == i ? no : ny;
          if (other < i && other >= j) continue;  // the other node moved too and has the smaller index
          const double s = M.con_s[ent.y];
          const double hYn = h[ny], hOn = h[no], hYo = old_h(ny), hOo = old_h(no);
          const double tn = (hYn < hOn) ? 0.0 : -((hYn - hOn) * (hYn - hOn)) / (2.0 * s * s);
          const double to = (hYo < hOo) ? 0.0 : -((hYo - hOo) * (hYo - hOo)) / (2.0 * s * s);
          d_A += tn - to;
        }
      }
    }
  }
  // block sums: shuffle tree per warp, fixed warp order (deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q += __shfl_xor_sync(0xffffffffu, q, o);
    d_clock += __shfl_xor_sync(0xffffffffu, d_clock, o);
    d_bd += __shfl_xor_sync(0xffffffffu, d_bd, o);
    d_A += __shfl_xor_sync(0xffffffffu, d_A, o);
  }
  bad = __any_sync(0xffffffffu, bad);
  if ((tid & 31) == 0) {
    s_red[tid >> 5][0] = q; s_red[tid >> 5][1] = d_clock; s_red[tid >> 5][2] = d_bd; s_red[tid >> 5][3] = d_A;
    s_bad[tid >> 5] = bad ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {
    q = d_clock = d_bd = d_A = 0.0;
    int anybad = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      q += s_red[w][0]; d_clock += s_red[w][1]; d_bd += s_red[w][2]; d_A += s_red[w][3];
      anybad |= s_bad[w];
    }
    const double* o0 = cur_out + (size_t)chain * 8;
    double* o1 = new_out + (size_t)chain * 8;
    const double NINF = -CUDART_INF;
    int st = cur_status[chain] & ST_NEARCRIT;
    if (anybad) {
      o1[0] = o0[0]; o1[1] = NINF; o1[2] = NINF; o1[3] = NINF; o1[4] = o0[4]; o1[5] = o0[5]; o1[6] = NINF; o1[7] = 0.0;
      st |= ST_ZERO;
    } else {
      double lnB = o0[1] + d_bd;
      if (nearcrit) {  // literal D/E recursion on the proposed state (BirthDeath.hs:90-114)
        const double d = la - mu;
        double E = 0.0, bd = 0.0;
        for (int i = N - 1; i >= 1; --i) {
          const int pe = T.parent[i];
          const bool inner = pe >= 0;
          const double ti = h[pe & 0x7fffffff] - h[i];
          const double c = inner ? E : 0.0;
          const double yy = (mu - c * la) * ti, den = 1.0 + yy;
          const double D = (1.0 - d * ti) / den / den;
          E = (c + yy) / den;
          bd += log(D * (inner ? la : 1.0));
        }
        lnB = (0.0 - la) + (0.0 - mu) + bd;
      }
      const double lnA = o0[0] + d_A, lnC = o0[2] + d_clock;
      const double prior = lnA + lnB + lnC;
      const double lk = o0[4] + (-0.5) * q;
      const double d0 = ((h[0] - h[1]) * r[1] + (h[0] - h[M.root_r]) * r[M.root_r]) * sc;
      const double jac = log(1.0 / d0);
      const double post = prior + lk + jac;
      if (post == NINF) st |= ST_ZERO;
      if (post != post) st |= ST_NAN;
      o1[0] = lnA; o1[1] = lnB; o1[2] = lnC; o1[3] = prior; o1[4] = lk; o1[5] = jac; o1[6] = post; o1[7] = 0.0;
    }
    new_status[chain] = st;
  }
}

// Small moves in ONE launch: propose, evaluate incrementally, accept / restore and update the cached y, one warp per chain
// (eight chains per CTA).  Same draws, same arithmetic and therefore the same decisions as mh_propose_kernel ->
// mh_delta_kernel -> mh_accept_kernel, without the undo log's round trip through HBM and with a whole batch of 8192
// chains resident on the 148 SMs at once (the serial part of a proposal -- Philox, erf / erfinv, logs -- is latency, not
// throughput).  Dynamic shared memory: per warp the operation list and the bitmap of affected branches.
template <int CLOCK>
__global__ void __launch_bounds__(256, 4)
mh_fused_small_kernel(const DevModel M, const MhTopo T, const MhParams P, const double* __restrict__ Pm, double* states,
                      double* y_cur, double* cur_out, int* cur_status, int* __restrict__ accepted,
                      unsigned long long* __restrict__ counters, const int* __restrict__ slot,
                      const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int B) {
  extern __shared__ __align__(16) unsigned char fs_smem[];
  __shared__ int s_off[8][DL_MAX_CHG];
  __shared__ double s_old[8][DL_MAX_CHG];
  // per warp 2 KB: first the operation list (<= 64 operations for the moves routed here), then the affected branches
  __shared__ __align__(16) unsigned char s_u[8][DL_MAX_AB * 16];
  static_assert(sizeof(MhOp) * 4 * MH_MAX_BRACE_NODES <= DL_MAX_AB * 16, "operation list must fit the per-warp scratch");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chain = blockIdx.x * 8 + warp;
  if (chain >= B) return;
  const int nwords = (M.N + 31) >> 5;
  MhOp* ops = reinterpret_cast<MhOp*>(s_u[warp]);
  int* s_ab_w = reinterpret_cast<int*>(s_u[warp]);
  int* s_k_w = s_ab_w + DL_MAX_AB;
  double* s_d_w = reinterpret_cast<double*>(s_u[warp] + DL_MAX_AB * 8);
  unsigned* bm = reinterpret_cast<unsigned*>(fs_smem) + (size_t)warp * nwords;
  int* c_off = s_off[warp];
  double* c_old = s_old[warp];
  double* row = states + (size_t)chain * M.S;
  int nops = 0, okv = 0;
  double lqv = 0.0;
  if (lane == 0) {
    bool ok;
    int node;
    nops = mh_build_ops(T, P, row, chain, 0.0, ops, &ok, &node, &lqv);
    okv = ok ? 1 : 0;
  }
  nops = __shfl_sync(0xffffffffu, nops, 0);
  okv = __shfl_sync(0xffffffffu, okv, 0);
  __syncwarp();
  if (!okv || nops == 0) {  // the reference would call `error`: the chain stays where it is
    if (lane == 0) {
      if (accepted) accepted[chain] = okv ? 0 : -1;
      if (counters && !okv) atomicAdd(counters + 1, 1ull);
    }
    return;
  }
  // apply the operations in order, keeping (offset, old value) of every touched entry
  int n_raw = 0;
  for (int o = 0; o < nops; ++o) {
    const MhOp op = ops[o];
    for (int i = lane; i < op.cnt && n_raw + i < DL_MAX_CHG; i += 32) {
      const double old = row[op.off + i];
      c_off[n_raw + i] = op.off + i;
      c_old[n_raw + i] = old;
      double y;
      switch (op.mode) {
        case OP_SET: y = op.b; break;
        case OP_MUL: y = old * op.a; break;
        case OP_DIV: y = old / op.a; break;
        case OP_ADD: y = old + op.b; break;
        default: y = (old - op.b) * op.a + op.b; if (!(y > 0.0)) y = CUDART_NAN; break;
      }
      row[op.off + i] = y;
    }
    n_raw += op.cnt;
    __syncwarp();
  }
  MCD_ASSERT(n_raw <= DL_MAX_CHG && nops * (int)sizeof(MhOp) <= DL_MAX_AB * 16);
  if (n_raw > DL_MAX_CHG) n_raw = DL_MAX_CHG;  // the host only routes small moves here
  double o1[8];
  int st = 0, n_ab = 0;
  double* y = y_cur + (size_t)chain * T.ldyc;
  mh_delta_core<CLOCK>(M, T, Pm, row, c_off, c_old, n_raw, bm, s_ab_w, s_k_w, s_d_w, y, cur_out + (size_t)chain * 8,
                       cur_status[chain], lane, o1, &st, &n_ab);
  int acc = 0;
  if (lane == 0) {
    const double* o0 = cur_out + (size_t)chain * 8;
    double bp = 1.0, bl = 1.0;
    if (slot) {
      const int sl = slot[P.chain_offset + chain];
      bp = ladder_prior[sl];
      bl = ladder_lik[sl];
    }
    double lr = bp * (o1[3] - o0[3]) + bl * (o1[4] - o0[4]) + lqv;
    if (P.use_root_jacobian) lr += o1[5] - o0[5];
    const double u = mh_uniform(P.seed, (uint32_t)(P.chain_offset + chain), P.iteration, 1u);
    acc = log(u) < lr;
    if (accepted) accepted[chain] = acc;
    if (counters && acc) atomicAdd(counters, 1ull);
  }
  acc = __shfl_sync(0xffffffffu, acc, 0);
  if (acc) {
    if (lane == 0) {
      for (int j = 0; j < 8; ++j) cur_out[(size_t)chain * 8 + j] = o1[j];
      cur_status[chain] = st;
    }
    __syncwarp();
    mh_update_y<32>(y, Pm, M.ldk, s_k_w, s_d_w, n_ab, lane);
  } else {
    // restore: the FIRST record of an offset holds its old value (the operation list has been overwritten by now)
    for (int e = lane; e < n_raw; e += 32) {
      const int off = c_off[e];
      bool first = true;
      for (int e2 = 0; e2 < e; ++e2) first = first && (c_off[e2] != off);
      if (first) row[off] = c_old[e];
    }
  }
}

// Small trees (N <= 96 nodes, all reference data sets): a whole Metropolis-Hastings step in ONE launch -- these
// configurations are launch-latency bound (three launches of ~7 us each for a few hundred chains).  Structure of
// small_tree_fused_kernel (precision matrix in shared memory, one warp per chain: stage the row, residuals, y = P dx as a
// shared-memory mat-vec, the posterior passes), with the proposal in front of it and the decision behind it.  Same draws
// and arithmetic as mh_propose_kernel -> small_tree_fused_kernel -> mh_accept_kernel.
// Dynamic shared memory: reduction scratch | P [K][K] | per warp: state row, y, residuals | per warp: operation list,
// undo list (offset, old value) of up to 2N + 8 entries.
// One Metropolis-Hastings step of one chain by one warp (the body shared by the one-step and the whole-cycle kernels):
// propose in place, evaluate the proposed state from scratch, decide, restore on rejection.  Warp-level synchronisation only.
struct MhSmallWarp {
  double *sx, *sy, *sdx, *c_old;
  int* c_off;
  MhOp* ops;
  int n_undo;
  double *o_cur, *o_new;   // RESIDENT: the chain's current / proposed output row [8] in shared memory
  int* st;                 // RESIDENT: [0] proposed, [1] current status word
};
// RESIDENT (whole-cycle kernel): the chain's state row lives in W.sx and its output row / status word in W.o_cur / W.st[1] for the
// whole cycle; the step reads and writes shared memory only (a step is a serial chain of dependent operations: every global
// round trip -- row, proposed row, output rows -- costs ~1 us of its ~9).  Otherwise (one step per launch) the row is modified in
// global memory and staged, as mh_propose_kernel -> small_tree_fused_kernel -> mh_accept_kernel would.  Same arithmetic either way.
template <int CLOCK, bool RESIDENT = false>
__device__ __forceinline__ int mh_small_step(const DevModel& M, const MhTopo& T, const Topo& Tp, const MhParams& P, const double* sP,
                                              const MhSmallWarp& W, double* scratch, int* iscratch, double* states, double* cur_out,
                                              int* cur_status, double* new_out, int* new_status, int* __restrict__ accepted,
                                              unsigned long long* __restrict__ counters, const int* __restrict__ slot,
                                              const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int chain,
                                              int lane) {
  const int K = M.K, N = M.N, S = M.S, n_undo = W.n_undo;
  double *sx = W.sx, *sy = W.sy, *sdx = W.sdx, *c_old = W.c_old;
  int* c_off = W.c_off;
  MhOp* ops = W.ops;
  double* row = RESIDENT ? sx : states + (size_t)chain * S;
  double rate_sum = 0.0;
  if (P.kind == MH_SCALE_VAR_TREE) {
    for (int i = 1 + lane; i < N; i += 32) rate_sum += row[5 + N + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rate_sum += __shfl_xor_sync(0xffffffffu, rate_sum, o);
  }
  int nops = 0, okv = 0;
  double lqv = 0.0;
  if (lane == 0) {
    bool ok;
    int node;
    nops = mh_build_ops(T, P, row, chain, rate_sum, ops, &ok, &node, &lqv);
    okv = ok ? 1 : 0;
  }
  nops = __shfl_sync(0xffffffffu, nops, 0);
  okv = __shfl_sync(0xffffffffu, okv, 0);
  __syncwarp();
  if (!okv || nops == 0) {
    if (lane == 0) {
      if (!RESIDENT && accepted) accepted[chain] = okv ? 0 : -1;
      if (counters && !okv) atomicAdd(counters + 1, 1ull);
    }
    return okv ? 0 : -1;
  }
  int n_raw = 0;
  for (int o = 0; o < nops; ++o) {
    const MhOp op = ops[o];
    for (int i = lane; i < op.cnt && n_raw + i < n_undo; i += 32) {
      const double old = row[op.off + i];
      c_off[n_raw + i] = op.off + i;
      c_old[n_raw + i] = old;
      double y;
      switch (op.mode) {
        case OP_SET: y = op.b; break;
        case OP_MUL: y = old * op.a; break;
        case OP_DIV: y = old / op.a; break;
        case OP_ADD: y = old + op.b; break;
        default: y = (old - op.b) * op.a + op.b; if (!(y > 0.0)) y = CUDART_NAN; break;
      }
      row[op.off + i] = y;
    }
    n_raw += op.cnt;
    __syncwarp();
  }
  MCD_ASSERT(n_raw <= n_undo);
  if (n_raw > n_undo) n_raw = n_undo;
  // evaluate the proposed state: small_tree_fused_kernel's body
  if (!RESIDENT) {
    __threadfence_block();
    stage_chain<32>(M, chain, lane, sx, sy, states, nullptr);
  }
  if (M.lik == 0) {
    const double* h = sx + 3;
    const double* r = sx + 5 + N;
    const double sc = sx[2] * sx[3 + N];
    for (int i = 1 + lane; i < N; i += 32) {
      if (i == M.root_r) continue;
      double e = (h[M.parent[i] & ~LEAF_BIT] - h[i]) * r[i];
      if (i == 1) e = e + (h[0] - h[M.root_r]) * r[M.root_r];
      const int kk = i < M.root_r ? i - 1 : i - 2;
      sdx[kk] = e * sc - M.mu[kk];
    }
    __syncwarp();
    for (int kk = lane; kk < K; kk += 32) {
      const double* prow = sP + (size_t)kk * K;
      double a0 = 0.0, a1 = 0.0;
      int j = 0;
      for (; j + 2 <= K; j += 2) {
        a0 = fma(prow[j], sdx[j], a0);
        a1 = fma(prow[j + 1], sdx[j + 1], a1);
      }
      if (j < K) a0 = fma(prow[j], sdx[j], a0);
      sy[kk] = a0 + a1;
    }
    __syncwarp();
  }
  if (RESIDENT) process_chain<32, CLOCK, false>(M, Tp, 0, lane, sx, sy, scratch, iscratch, W.o_new, nullptr, W.st);
  else process_chain<32, CLOCK, false>(M, Tp, chain, lane, sx, sy, scratch, iscratch, new_out, nullptr, new_status);
  __syncwarp();
  int acc = 0;
  if (lane == 0) {  // lane 0 wrote the proposed output row / status word itself
    const double* o1 = RESIDENT ? W.o_new : new_out + (size_t)chain * 8;
    double* o0 = RESIDENT ? W.o_cur : cur_out + (size_t)chain * 8;
    double bp = 1.0, bl = 1.0;
    if (slot) {
      const int sl = slot[P.chain_offset + chain];
      bp = ladder_prior[sl];
      bl = ladder_lik[sl];
    }
    double lr = bp * (o1[3] - o0[3]) + bl * (o1[4] - o0[4]) + lqv;
    if (P.use_root_jacobian) lr += o1[5] - o0[5];
    const double u = mh_uniform(P.seed, (uint32_t)(P.chain_offset + chain), P.iteration, 1u);
    acc = log(u) < lr;
    if (!RESIDENT && accepted) accepted[chain] = acc;
    if (counters && acc) atomicAdd(counters, 1ull);
    if (acc) {
      for (int j = 0; j < 8; ++j) o0[j] = o1[j];
      if (RESIDENT) W.st[1] = W.st[0];
      else cur_status[chain] = new_status[chain];
    }
  }
  acc = __shfl_sync(0xffffffffu, acc, 0);
  if (!acc) {  // restore: the first record of an offset holds its old value
    for (int e = lane; e < n_raw; e += 32) {
      const int off = c_off[e];
      bool first = true;
      for (int e2 = 0; e2 < e; ++e2) first = first && (c_off[e2] != off);
      if (first) row[off] = c_old[e];
    }
  }
  return acc;
}

template <int CLOCK>
__global__ void __launch_bounds__(POST_THREADS, 2)
mh_small_tree_kernel(DevModel M, const MhTopo T, const MhParams P, const double* __restrict__ Pm /*[Mp][ldk] padded*/, double* states,
                     double* cur_out, int* cur_status, double* new_out, int* new_status, int* __restrict__ accepted,
                     unsigned long long* __restrict__ counters, const int* __restrict__ slot,
                     const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int B) {
  extern __shared__ __align__(16) unsigned char smem_p[];
  double* scratch = reinterpret_cast<double*>(smem_p);
  int* iscratch = reinterpret_cast<int*>(smem_p + 8 * NRED * 8);
  double* sP = reinterpret_cast<double*>(smem_p + POST_SMEM_FIXED);   // [K][K]
  const int K = M.K, N = M.N, S = M.S;
  double* stage = sP + (size_t)K * K;
  const int n_undo = 2 * N + 8;
  unsigned char* mh_base = reinterpret_cast<unsigned char*>(stage + (size_t)(POST_THREADS / 32) * (S + N + K));
  const size_t mh_per_warp = sizeof(MhOp) * MH_MAX_OPS + (size_t)n_undo * 16;
  const Topo Tp{M.parent, M.mu, M.var, M.inner, nullptr};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (M.lik == 0) {
    for (int e = threadIdx.x; e < K * K; e += POST_THREADS) sP[e] = Pm[(size_t)(e / K) * M.ldk + (e % K)];
  }
  __syncthreads();
  double* sx = stage + (size_t)warp * (S + N + K);
  double* sy = sx + S;
  double* sdx = sy + N;
  MhOp* ops = reinterpret_cast<MhOp*>(mh_base + (size_t)warp * mh_per_warp);
  double* c_old = reinterpret_cast<double*>(mh_base + (size_t)warp * mh_per_warp + sizeof(MhOp) * MH_MAX_OPS);
  int* c_off = reinterpret_cast<int*>(c_old + n_undo);
  for (int chain = blockIdx.x * (POST_THREADS / 32) + warp; chain < B; chain += gridDim.x * (POST_THREADS / 32)) {
    const MhSmallWarp W{sx, sy, sdx, c_old, c_off, ops, n_undo, nullptr, nullptr, nullptr};
    mh_small_step<CLOCK>(M, T, Tp, P, sP, W, scratch, iscratch, states, cur_out, cur_status, new_out, new_status, accepted, counters, slot,
                         ladder_prior, ladder_lik, chain, lane);
    __syncwarp();  // the warp's staging buffers are reused by its next chain
  }
}

// The whole proposal cycle of small trees in ONE launch: every warp keeps its chain and walks n_sweeps sweeps over the list
// (entry e `repeat` times, Philox iteration iteration0 + running step number), exactly the steps mcd_mh_cycle would enqueue one
// launch at a time -- same draws, same arithmetic, same results -- without ~250 launch latencies per sweep and with the precision
// matrix staged in shared memory once.  Chains never interact inside a cycle (the MC3 slots change between calls only).
constexpr int MH_CYCLE_WARP_EXTRA = 16 * 8 + 16;   // per warp: current / proposed output rows [8] + two status words
struct MhCycleEntry {
  int kind, node, use_root_jacobian, repeat;
  double param, tune;
};
template <int CLOCK>
__global__ void __launch_bounds__(POST_THREADS, 2)
mh_small_cycle_kernel(DevModel M, const MhTopo T, const MhCycleEntry* __restrict__ cycle, int n_entries, int n_sweeps, uint64_t seed,
                      uint32_t iteration0, int chain_offset, const double* __restrict__ Pm, double* states, double* cur_out,
                      int* cur_status, double* new_out, int* new_status, int* __restrict__ accepted,
                      unsigned long long* __restrict__ counters /*[n_entries][2]*/, const int* __restrict__ slot,
                      const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int B) {
  extern __shared__ __align__(16) unsigned char smem_p[];
  double* scratch = reinterpret_cast<double*>(smem_p);
  int* iscratch = reinterpret_cast<int*>(smem_p + 8 * NRED * 8);
  double* sP = reinterpret_cast<double*>(smem_p + POST_SMEM_FIXED);   // [K][K]
  const int K = M.K, N = M.N, S = M.S;
  double* stage = sP + (size_t)K * K;
  const int n_undo = 2 * N + 8;
  const int wpb = blockDim.x >> 5;   // warps (= chains in flight) per CTA: the per-warp areas are sized by it
  unsigned char* mh_base = reinterpret_cast<unsigned char*>(stage + (size_t)wpb * (S + N + K));
  const size_t mh_per_warp = sizeof(MhOp) * MH_MAX_OPS + (size_t)n_undo * 16 + MH_CYCLE_WARP_EXTRA;
  const Topo Tp{M.parent, M.mu, M.var, M.inner, nullptr};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (M.lik == 0) {
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) sP[e] = Pm[(size_t)(e / K) * M.ldk + (e % K)];
  }
  __syncthreads();
  MhSmallWarp W;
  W.sx = stage + (size_t)warp * (S + N + K);
  W.sy = W.sx + S;
  W.sdx = W.sy + N;
  W.ops = reinterpret_cast<MhOp*>(mh_base + (size_t)warp * mh_per_warp);
  W.c_old = reinterpret_cast<double*>(mh_base + (size_t)warp * mh_per_warp + sizeof(MhOp) * MH_MAX_OPS);
  W.c_off = reinterpret_cast<int*>(W.c_old + n_undo);
  W.n_undo = n_undo;
  W.o_cur = reinterpret_cast<double*>(mh_base + (size_t)warp * mh_per_warp + sizeof(MhOp) * MH_MAX_OPS + (size_t)n_undo * 16);
  W.o_new = W.o_cur + 8;
  W.st = reinterpret_cast<int*>(W.o_new + 8);
  for (int chain = blockIdx.x * wpb + warp; chain < B; chain += gridDim.x * wpb) {
    // the chain moves into shared memory for the whole cycle
    stage_chain<32>(M, chain, lane, W.sx, W.sy, states, nullptr);
    if (lane < 8) {
      W.o_cur[lane] = cur_out[(size_t)chain * 8 + lane];
      W.o_new[lane] = new_out[(size_t)chain * 8 + lane];
    }
    if (lane == 0) { W.st[1] = cur_status[chain]; W.st[0] = new_status[chain]; }
    __syncwarp();
    int last = 0;
    uint32_t it = iteration0;
    for (int sweep = 0; sweep < n_sweeps; ++sweep) {
      for (int e = 0; e < n_entries; ++e) {
        const MhCycleEntry ce = cycle[e];
        MhParams P;
        P.kind = ce.kind; P.node = ce.node; P.use_root_jacobian = ce.use_root_jacobian; P.pad = 0; P.param = ce.param; P.tune = ce.tune;
        P.seed = seed; P.chain_offset = chain_offset;
        for (int r = 0; r < ce.repeat; ++r) {
          P.iteration = it++;
          last = mh_small_step<CLOCK, true>(M, T, Tp, P, sP, W, scratch, iscratch, states, cur_out, cur_status, new_out, new_status,
                                            accepted, counters + 2 * e, slot, ladder_prior, ladder_lik, chain, lane);
          __syncwarp();  // the warp's staging buffers are reused by its next step
        }
      }
    }
    // ... and back: state row, output row, status word, the last step's decision and proposed outputs
    for (int i = lane; i < S; i += 32) states[(size_t)chain * S + i] = W.sx[i];
    if (lane < 8) {
      cur_out[(size_t)chain * 8 + lane] = W.o_cur[lane];
      new_out[(size_t)chain * 8 + lane] = W.o_new[lane];
    }
    if (lane == 0) {
      cur_status[chain] = W.st[1];
      new_status[chain] = W.st[0];
      if (accepted) accepted[chain] = last;
    }
    __syncwarp();
  }
}

// MC3 state swaps between neighbouring temperatures (the `mcmc` package's MC3 algorithm, called at app/Main.hs:476-479;
// restated from Altekar et al. 2004 / the package's published source): for the chains x_i, x_j holding slots p, p + 1 of a
// group, ln r = (beta_p - beta_{p+1}) (ln pi(x_j) - ln pi(x_i)) with prior and likelihood heated by their own ladders;
// accept iff ln U < ln r.  States never move: the chains exchange their temperature slots.  stats[c] = (ln prior,
// ln likelihood) of every chain of every rank (all-gathered by the host over NCCL when the groups span GPUs); every rank
// takes the same decisions from the same Philox draws (counter (group, iteration, draw, 3)).  One thread per group.
__global__ void mh_swap_kernel(const double* __restrict__ stats, int stats_stride, int* __restrict__ slot, int* __restrict__ chain_of_slot,
                               const double* __restrict__ ladder_prior, const double* __restrict__ ladder_lik, int n_groups,
                               int C, int pair, uint64_t seed, uint32_t iteration, int* __restrict__ accepted) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups || C < 2) return;
  uint32_t c[4] = {(uint32_t)g, iteration, 0u, 3u};
  Philox{(uint32_t)seed, (uint32_t)(seed >> 32)}(c);
  const double u0 = ((double)(((uint64_t)(c[0] >> 5) << 26) | (uint64_t)(c[1] >> 6)) + 0.5) * 1.1102230246251565e-16;
  const double u1 = ((double)(((uint64_t)(c[2] >> 5) << 26) | (uint64_t)(c[3] >> 6)) + 0.5) * 1.1102230246251565e-16;
  int p = pair;
  if (p < 0) {
    p = (int)(u1 * (double)(C - 1));
    if (p >= C - 1) p = C - 2;
  }
  const int i = chain_of_slot[g * C + p], j = chain_of_slot[g * C + p + 1];
  const double* si = stats + (size_t)i * stats_stride;
  const double* sj = stats + (size_t)j * stats_stride;
  const double lr = (ladder_prior[p] - ladder_prior[p + 1]) * (sj[0] - si[0]) + (ladder_lik[p] - ladder_lik[p + 1]) * (sj[1] - si[1]);
  const int acc = log(u0) < lr;
  if (acc) {
    slot[i] = p + 1;
    slot[j] = p;
    chain_of_slot[g * C + p] = j;
    chain_of_slot[g * C + p + 1] = i;
  }
  if (accepted) accepted[g] = acc;
}

}  // namespace mcd
