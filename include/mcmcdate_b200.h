/* mcmcdate_b200.h -- C ABI of the B200-native batched evaluator for McmcDate's hot path.
 *
 * The reference (dschrempf/mcmc-date) is pure Haskell with NO FFI today.  These entry
 * points are what a `foreign import ccall` shim would bind so that the functions the `mcmc` sampler
 * consumes keep their types (INTEGRATION.md shows the shim):
 *
 *   reference interface (file:line, relative to the reference repo)      -> replaced by
 *   ------------------------------------------------------------------------------------------
 *   priorFunction            app/Probability.hs:127-150   PriorFunction I        mcd_eval      (out[MCD_OUT_LNPRIOR])
 *   priorFunctionCalibrationsConstraintsBraces  :46-63    (monitor, Monitor.hs)  mcd_eval      (out[MCD_OUT_LNA])
 *   priorFunctionBirthDeath  :66-85                       (monitor)              mcd_eval      (out[MCD_OUT_LNB])
 *   priorFunctionRelaxedMolecularClock  :96-124           (monitor)              mcd_eval      (out[MCD_OUT_LNC])
 *   likelihoodFunction       app/Probability.hs:277-281   LikelihoodFunction I   mcd_eval      (out[MCD_OUT_LNLIK])
 *   jacobianRootBranch       app/Probability.hs:408-410   JacobianFunction I     mcd_eval      (out[MCD_OUT_LNJAC])
 *   htargetWith + `ad` grad  app/Hamiltonian.hs:72-92     HTarget IG             mcd_eval_grad (out[MCD_OUT_LNPOST], grad)
 *   getMask                  app/Hamiltonian.hs:33-47                            mcd_mask
 *   toVector / fromVectorWith app/Hamiltonian.hs:49-60    HStructure IG          mcd_to_vector / mcd_from_vector
 *   getBranches+sumFirstTwo  app/Tools.hs:36-48           (branch order)         mcd_branch_index
 *
 * Conventions
 *   - plain C types only; every call returns int: 0 = ok, < 0 = fatal (bad argument, CUDA error);
 *     the message is available from mcd_last_error().  Per-chain domain problems never fail a call:
 *     they set status[b] bits and yield -inf / NaN like the reference's `Log` arithmetic would.
 *   - a state is the canonical flattening `toList (x :: I)` (app/State.hs:70-100):
 *       [lambda, mu, H, h_0..h_{N-1}, m, v, r_0..r_{N-1}]        S = 5 + 2N doubles
 *     with node heights h and rates r in pre-order (root first).  Batches are chain-major [B][S].
 *   - the caller owns all in/out buffers; the library owns the handle and all device memory.
 *   - one handle = one GPU.  Calls on one handle are serialised internally (mutex); use `safe`
 *     foreign calls from Haskell.  No CPU fallback exists: without a CUDA device mcd_create fails.
 */
#ifndef MCMCDATE_B200_H
#define MCMCDATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcd_handle mcd_handle;

/* RelaxedMolecularClockModel, app/Probability.hs:88-93 */
enum { MCD_CLOCK_UNCORRELATED_GAMMA = 0, MCD_CLOCK_UNCORRELATED_LOGNORMAL = 1,
       MCD_CLOCK_UNCORRELATED_WHITENOISE = 2, MCD_CLOCK_AUTOCORRELATED_LOGNORMAL = 3 };
/* LikelihoodData, app/Probability.hs:210-235 */
enum { MCD_LIK_FULL = 0, MCD_LIK_UNIVARIATE = 1, MCD_LIK_NONE = 2, MCD_LIK_SPARSE = 3 };
/* per-chain status bits */
enum { MCD_ST_REF_ERROR = 1,    /* the reference would have called `error` (abort); value -inf */
       MCD_ST_ZERO = 2,         /* ln posterior = -inf (probability zero) */
       MCD_ST_NAN = 4,          /* ln posterior is NaN */
       MCD_ST_NEARCRIT = 8,     /* |lambda - mu| < 1e-6 (BirthDeath.hs:125-126) */
       MCD_ST_LEAF_HEIGHT = 16, /* a leaf height != 0: HeightTree invariant violated */
       MCD_ST_FP64_FALLBACK = 32 /* informational: this chain's residuals span too wide a range for the INT8 digit planes (at least half
                                  * of its standardised residuals are more than 512 x below the largest one); its contraction was
                                  * recomputed in plain FP64 */ };
/* columns of one output row (MCD_OUT_COLS doubles per chain) */
enum { MCD_OUT_LNA = 0,     /* calibrations * constraints * braces */
       MCD_OUT_LNB = 1,     /* exponential(lambda) * exponential(mu) * birth-death */
       MCD_OUT_LNC = 2,     /* exponential(m) * gamma(v) * relaxed clock */
       MCD_OUT_LNPRIOR = 3, /* product' [A, B, C] */
       MCD_OUT_LNLIK = 4, MCD_OUT_LNJAC = 5,
       MCD_OUT_LNPOST = 6,  /* prior * likelihood * jacobian (the HMC target) */
       MCD_OUT_COLS = 8 };

typedef struct mcd_model_desc {
  int32_t n_nodes;           /* N = 2n-1, strictly bifurcating, pre-order numbering, root = 0 */
  const int32_t* parent;     /* [N], parent[0] = -1 */
  int32_t clock_model;       /* MCD_CLOCK_* */
  int32_t likelihood;        /* MCD_LIK_* */
  const double* mean;        /* [K], K = N-2, branch order of mcd_branch_index */
  const double* precision;   /* FULL: [K*K] row-major Sigma^-1, symmetric up to rounding (its symmetric part is used: `prepare`
                              * writes an unsymmetrised LU inverse, app/Main.hs:230); UNIVARIATE: [K] variances */
  double logdet_sigma;       /* ln det Sigma (FULL) / sum ln variances (UNIVARIATE) */
  double ht;                 /* mean root height for the rate-mean prior (app/Main.hs:394), > 0 */
  int32_t n_cal;             /* calibrations (Calibration.hs:55-123) */
  const int32_t* cal_node;   /* [n_cal] pre-order node index */
  const double* cal_lo;      /* lower bound (absolute age); <= 0: none */
  const double* cal_lo_p;    /* probability mass at the lower bound */
  const double* cal_hi;      /* upper bound; +inf: none */
  const double* cal_hi_p;
  int32_t n_con;             /* constraints (Constraint.hs:61-74) */
  const int32_t* con_young;
  const int32_t* con_old;
  const double* con_p;
  int32_t n_brace;           /* braces (Brace.hs:54-59), CSR over nodes */
  const int32_t* brace_off;  /* [n_brace+1] */
  const int32_t* brace_node; /* [brace_off[n_brace]] */
  const double* brace_sd;    /* [n_brace], > 0 */
  int32_t device;            /* CUDA device ordinal */
  int32_t max_batch;         /* capacity hint; buffers grow on demand */
  const double* precision_chol; /* optional (FULL): lower-triangular Cholesky factor L of Sigma^-1 = L L^T,
                                   [K*K] row-major, used by the value-only path (half the flops); NULL: the
                                   library factorises on the GPU at the first value-only call (blocked FP64 Cholesky,
                                   ~K^3/3 flops: a few milliseconds at K = 2000; one extra K x K work matrix for its
                                   duration).  A matrix that is not positive definite keeps the symmetric product. */
  int32_t n_sparse;          /* SPARSE: the association list ((i, j), v) of the sparse Sigma^-1 exactly as the */
  const int32_t* sparse_row; /*   reference stores it (SparseS, app/Main.hs:75-81; mkSparse :95-97);        */
  const int32_t* sparse_col; /*   duplicates add up; `precision` is ignored, logdet_sigma = ln det of the   */
  const double* sparse_val;  /*   sparse covariance                                                         */
} mcd_model_desc;

/* lifecycle */
int mcd_create(const mcd_model_desc* desc, mcd_handle** out);
void mcd_destroy(mcd_handle* h);
const char* mcd_last_error(const mcd_handle* h); /* h may be NULL: error of the last failed mcd_create */

/* model queries (host side, no GPU work) */
int mcd_state_len(const mcd_handle* h);                  /* S = 5 + 2N */
int mcd_dim(const mcd_handle* h);                        /* K */
int mcd_branch_index(const mcd_handle* h, int32_t* out); /* [N] node -> k, root -1 */
int mcd_mask(const mcd_handle* h, uint8_t* out);         /* [S] 1 = free under HMC (getMask) */
int mcd_hmc_dim(const mcd_handle* h);                    /* D = number of free parameters */
int mcd_to_vector(const mcd_handle* h, const double* state, double* theta);  /* theta[D], reversed order */
int mcd_from_vector(const mcd_handle* h, const double* base_state, const double* theta, double* state_out);

/* evaluation with HOST buffers (copies in and out are part of the call; chunks are pipelined) */
int mcd_eval(mcd_handle* h, int32_t n_chains, const double* states /*[B][S]*/,
             double* out /*[B][MCD_OUT_COLS]*/, int32_t* status /*[B]*/);
int mcd_eval_grad(mcd_handle* h, int32_t n_chains, const double* states /*[B][S]*/,
                  double* out /*[B][MCD_OUT_COLS]*/, double* grad /*[B][S], masked entries 0*/,
                  int32_t* status /*[B]*/);

/* HMC form: positions are the masked, reversed-order vectors of toVector (app/Hamiltonian.hs:49-53),
 * theta[B][D]; fixed entries (root / leaf heights, rate stem, H without calibrations) come from one
 * shared base_state[S] (fromVectorWith); the gradient comes back packed the same way.  Moves D instead
 * of S doubles per chain over PCIe in each direction. */
int mcd_eval_grad_theta(mcd_handle* h, int32_t n_chains, const double* theta /*[B][D]*/,
                        const double* base_state /*[S]*/, double* out /*[B][MCD_OUT_COLS]*/,
                        double* grad_theta /*[B][D]*/, int32_t* status /*[B]*/);
/* Ordering between entry points.  All calls on one handle share its device work buffers.  The library orders them itself:
 * a call that runs on a single stream (the *_device forms on the caller's stream; mcd_leapfrog, mcd_nuts, mcd_chains_*,
 * mcd_mh_*, mcd_mc3_* on the library's) first waits, on the device, for whatever earlier pipelined host-buffer calls
 * (mcd_eval*, mcd_eval*_async) and the previous single-stream call still have in flight, and the pipelined calls wait for the
 * last single-stream call.  No mcd_wait / mcd_synchronize is needed between calls for correctness of the DEVICE work; host
 * OUTPUT buffers of asynchronous calls are of course valid only after their ticket has been waited for. */
/* The same without the final wait: returns a ticket (>= 0; < 0: error) after enqueueing; the outputs are valid after
 * mcd_wait(ticket) (or mcd_synchronize); the host buffers must stay alive and untouched until then.  Back-to-back calls
 * overlap the PCIe fill of one with the drain of the previous one (chunk k of every call runs in order on stream k % 4 on
 * its own slice of the staging buffers).  A double-buffered host loop: t1 = async(A); t2 = async(B); wait(t1); use A;
 * t3 = async(A); wait(t2); use B; ...  The last 8 tickets can be waited for. */
int64_t mcd_eval_grad_theta_async(mcd_handle* h, int32_t n_chains, const double* theta, const double* base_state, double* out,
                                  double* grad_theta, int32_t* status);
int mcd_wait(mcd_handle* h, int64_t ticket);
/* asynchronous forms of mcd_eval / mcd_eval_grad (same tickets) */
int64_t mcd_eval_async(mcd_handle* h, int32_t n_chains, const double* states, double* out, int32_t* status);
int64_t mcd_eval_grad_async(mcd_handle* h, int32_t n_chains, const double* states, double* out, double* grad, int32_t* status);

/* Device-resident leapfrog trajectory (first half of SURVEY 8f rank 1): n_steps leapfrog steps of the
 * Hamiltonian H = -ln post(theta) + 1/2 p^T M^-1 p (M diagonal) for every chain, positions / momenta /
 * gradients never leaving HBM between the two ends.  The integrator is the one behind the reference's
 * Hamiltonian proposal (app/Hamiltonian.hs:95-104 -> `mcmc`): half kick, drift, ..., half kick.
 * energy[b] = (H at the start, H at the end); out = ln-posterior parts at the end point;
 * status[b] = OR of the status words of every gradient evaluation of the trajectory. */
int mcd_leapfrog(mcd_handle* h, int32_t n_chains, int32_t n_steps, const double* theta0 /*[B][D]*/,
                 const double* momentum0 /*[B][D]*/, const double* base_state /*[S]*/,
                 const double* inv_mass /*[D]*/, const double* step_size /*[B]*/, double* theta_out /*[B][D]*/,
                 double* momentum_out /*[B][D]*/, double* out /*[B][MCD_OUT_COLS]*/, double* energy /*[B][2]*/,
                 int32_t* status /*[B]*/);

/* One No-U-Turn-sampler transition per chain, resident on the device (the reference's proposal is
 * `nuts defaultNParams ...` of the `mcmc` package, app/Hamiltonian.hs:95-104): Hoffman & Gelman (2014) Algorithm 3 --
 * slice variable, doubling in random directions, U-turn checks on every balanced sub-trajectory, divergence at an
 * energy error > 1000 -- for every chain at once, the chains ticking in lockstep (one leapfrog step per tick).
 * Random numbers are Philox4x32-10 with key `seed` and counter (chain, iteration, draw, stream): reproducible and
 * restated bit for bit in tests/nuts_ref.py.  momentum0 == NULL: momenta ~ N(0, M) are drawn on the device.
 * info[b] = (tree depth, leapfrog steps, diverged, valid points); accept_stat[b] = mean min(1, exp(-dH)) over the
 * leaves (the statistic dual averaging tunes the step size with). */
int mcd_nuts(mcd_handle* h, int32_t n_chains, const double* theta0 /*[B][D]*/, const double* base_state /*[S]*/,
             const double* inv_mass /*[D]*/, const double* step_size /*[B]*/, const double* momentum0 /*[B][D] or NULL*/,
             int32_t max_depth, uint64_t seed, uint32_t iteration, double* theta_out /*[B][D]*/,
             double* out /*[B][MCD_OUT_COLS]*/, double* accept_stat /*[B]*/, int32_t* info /*[B][4]*/, int32_t* status /*[B]*/);

/* Metropolis-Hastings-Green moves on chains that live in HBM.  mcd_chains_set uploads the states of
 * n chains and evaluates them; mcd_mh_step applies ONE proposal to every resident chain in place, evaluates the proposed
 * states (value-only path), and accepts or rejects per chain; mcd_mh_cycle runs sweeps over a whole list of proposals
 * without a host round trip; mcd_chains_get reads states / ln-posterior parts back.
 * Every proposal of the reference's cycle (app/Definitions.hs:145-279) is available; lib/ = lib/Mcmc/Tree/Proposal:
 *   kind                            reference                                           `node` argument      `param`
 *   MCD_MH_SLIDE_NODE               slideNodeAtUltrametric      lib/Ultrametric.hs:50-96     inner node | -1      sd
 *   MCD_MH_SCALE_SUBTREE            scaleSubTreeAtUltrametric   lib/Ultrametric.hs:126-186   inner node | -1      sd
 *   MCD_MH_PULLEY                   pulleyUltrametric           lib/Ultrametric.hs:228-316   ignored              sd
 *   MCD_MH_SLIDE_BRACE              slideBracedNodesUltrametric lib/Brace.hs:37-90           brace index | -1     sd
 *   MCD_MH_SCALE_BRANCH             scaleBranch (rate tree)     lib/Unconstrained.hs:45-85   node >= 1 | -1       shape k
 *   MCD_MH_SCALE_RATE_SUBTREE       scaleSubTreeAt (rate tree)  lib/Unconstrained.hs:87-175  inner node | -1      shape k
 *   MCD_MH_SCALE_NORM_TREE_CONTRA_M scaleNormAndTreeContrarily  lib/Unconstrained.hs:232-284 on (rateMean, rates)     k
 *   MCD_MH_SCALE_NORM_TREE_CONTRA_H   "                         app/Definitions.hs:241-253   on (timeHeight, rates)   k
 *   MCD_MH_SCALE_VAR_TREE           scaleVarianceAndTree        lib/Unconstrained.hs:286-371 ignored              shape k
 *   MCD_MH_SCALE_VAR_TREE_AUTO      scaleVarianceAndTreeAutocorrelated  lib/Unconstrained.hs:381-439 ignored      shape k
 *   MCD_MH_SLIDE_NODE_CONTRA        slideNodesAtContrarily      lib/Contrary.hs:35-131       inner node | -1      sd
 *   MCD_MH_SCALE_SUBTREE_CONTRA     scaleSubTreesAtContrarily   lib/Contrary.hs:269-395      inner node | -1      sd
 *   MCD_MH_SLIDE_BRACE_CONTRA       slideBracedNodesContrarily  lib/Brace.hs:98-209          brace index | -1     sd
 *   MCD_MH_SLIDE_ROOT_CONTRA        slideRootContrarily         lib/Contrary.hs:173-267      ignored              sd
 *   MCD_MH_SCALE_RATES_TREE_CONTRA  scaleRatesAndTreeContrarily lib/Contrary.hs:420-486      ignored              sd
 *   MCD_MH_SCALE_SCALAR             scaleUnbiased (`mcmc`)      app/Definitions.hs:259-262   0 lambda 1 mu 2 H 3 m 4 v   k
 *   MCD_MH_SCALE_H_M_CONTRA         scaleContrarily (`mcmc`)    app/Definitions.hs:244       ignored              shape k
 * node = -1: every chain draws its own node / brace uniformly among the eligible ones.  Inner node = inner node below
 * the root (the reference never builds these proposals for the root: HandleNode, app/Definitions.hs:132-138).
 * param, tune: standard deviation sd (truncated-normal moves, sd' = tune * sd, Internal.hs:117) or shape k of the
 * multiplier u ~ Gamma(k / tune, tune / k) (Unconstrained.hs:103, `mcmc` Scale proposals).
 * use_root_jacobian: include jacobianRootBranch in the ratio (proposals the reference lifts with liftProposalWith
 * jacobianRootBranch: the [R] ones of app/Definitions.hs).  accepted[b] (nullable) = 1 / 0, or -1 where the reference would
 * call `error` (truncatedNormalDistr: bounds crossed -- the chain's tree is invalid); such chains are left unchanged.
 * Heated chains (mcd_mc3_configure): ln r uses beta_prior * d ln prior + beta_lik * d ln lik.
 * Uniforms: Philox4x32-10, key = seed, counter = (chain, iteration, draw, 2); draws: 0 quantile, 1 acceptance, 2 node,
 * 7.. the gamma sampler (Marsaglia-Tsang).  Distinct proposals need distinct `iteration` values. */
enum { MCD_MH_SLIDE_NODE = 0, MCD_MH_SCALE_SUBTREE = 1, MCD_MH_PULLEY = 2, MCD_MH_SLIDE_BRACE = 3, MCD_MH_SCALE_BRANCH = 4,
       MCD_MH_SCALE_RATE_SUBTREE = 5, MCD_MH_SCALE_NORM_TREE_CONTRA_M = 6, MCD_MH_SCALE_NORM_TREE_CONTRA_H = 7,
       MCD_MH_SCALE_VAR_TREE = 8, MCD_MH_SCALE_VAR_TREE_AUTO = 9, MCD_MH_SLIDE_NODE_CONTRA = 10,
       MCD_MH_SCALE_SUBTREE_CONTRA = 11, MCD_MH_SLIDE_BRACE_CONTRA = 12, MCD_MH_SLIDE_ROOT_CONTRA = 13,
       MCD_MH_SCALE_RATES_TREE_CONTRA = 14, MCD_MH_SCALE_SCALAR = 15, MCD_MH_SCALE_H_M_CONTRA = 16 };
int mcd_chains_set(mcd_handle* h, int32_t n_chains, const double* states /*[B][S]*/);
int mcd_chains_get(mcd_handle* h, int32_t n_chains, double* states /*[B][S] or NULL*/, double* out /*[B][8] or NULL*/,
                   int32_t* status /*[B] or NULL*/);
int mcd_mh_step(mcd_handle* h, int32_t kind, int32_t node, double param, double tune, int32_t use_root_jacobian, uint64_t seed,
                uint32_t iteration, int32_t* accepted /*[B] or NULL*/);
/* The Hamiltonian proposal of the reference's cycle (`maybeHamiltonianProposal`, app/Definitions.hs:276-278) on the
 * RESIDENT chains: one mcd_nuts transition per chain, positions packed from and written back to the chains' state rows on
 * the device, then the chains' ln-posterior parts (and the cached contraction results) are re-evaluated.  Momenta are
 * drawn on the device.  The fixed entries (root / leaf heights, rate stem, H without calibrations) must agree across the
 * resident chains (they do for chains started from the reference's `initWith`).  Cold chains only. */
int mcd_chains_nuts(mcd_handle* h, const double* inv_mass /*[D]*/, const double* step_size /*[B]*/, int32_t max_depth,
                    uint64_t seed, uint32_t iteration, double* accept_stat /*[B]*/, int32_t* info /*[B][4]*/,
                    int32_t* status /*[B]*/);
/* One entry of a proposal cycle (the reference's `Cycle`, app/Definitions.hs:256-279): `repeat` = the proposal's weight. */
typedef struct mcd_mh_proposal {
  int32_t kind, node;
  double param, tune;
  int32_t use_root_jacobian, repeat;
} mcd_mh_proposal;
/* n_iterations sweeps over props[0..n_props) (each `repeat` times, in the given order), enqueued back to back; proposal
 * number s of the call uses Philox iteration iteration0 + s.  accepted / invalid [n_props] (nullable): chains x steps
 * accepted / rejected as invalid per list entry -- what the reference's auto tuner consumes.  *iteration_next (nullable):
 * first unused iteration value.  At most 4096 list entries per call.  Small trees (<= 96 nodes): the whole call is ONE kernel launch
 * (every chain stays in shared memory for all sweeps); same draws and results as one launch per step. */
int mcd_mh_cycle(mcd_handle* h, int32_t n_props, const mcd_mh_proposal* props, int32_t n_iterations, uint64_t seed,
                 uint32_t iteration0, uint64_t* accepted, uint64_t* invalid, uint32_t* iteration_next);

/* Incremental evaluation of small moves (default: on).  On large trees with a dense precision matrix the library keeps
 * y = Sigma^-1 (d - mu) of every resident chain's current state in HBM.  Proposals that touch a few branches only
 * (slide node, slide braced nodes, their contrary forms, scale branch, sub-tree moves on sub trees of <= 32 (heights) /
 * <= 64 (rates) nodes with a given node) are then scored from the change alone -- quad' = quad + 2 delta.y +
 * delta^T Sigma^-1 delta, prior terms of the touched nodes -- instead of 2 K^2 flops per chain; all other proposals are
 * evaluated from scratch.  The values differ from a fresh evaluation by accumulated rounding only; every
 * `refresh_every` incremental steps (default 512; 0 = keep) the resident chains are re-evaluated from scratch.  The mode
 * needs every chain's state to be valid at mcd_chains_set (else that set of chains runs with full evaluations).
 * mcd_mh_get_incremental: 1 if the resident chains are currently evaluated incrementally. */
int mcd_mh_set_incremental(mcd_handle* h, int32_t on, int32_t refresh_every);
int mcd_mh_get_incremental(const mcd_handle* h);

/* Heated chains: Metropolis-coupled MCMC (`mc3`, app/Main.hs:476-479) and the stepping-stone / thermodynamic-integration
 * points of `marginalLikelihood` (app/Main.hs:511-543); both live in the un-vendored `mcmc` package and are restated
 * from their published definitions.  n_global chains (over all ranks; this handle's resident chains are
 * [chain_offset, chain_offset + n_resident) of them) form groups of chains_per_group consecutive chains; chain c starts
 * at temperature slot c % chains_per_group with heats (ladder_prior[slot], ladder_lik[slot]).  MC3: both ladders equal
 * the beta_i; stepping stone: ladder_prior = 1, ladder_lik = beta_i, no swaps.  chains_per_group = 0: cold chains again
 * (chain_offset is still recorded).  chain_offset also enters the Philox counters of the proposals (global chain index),
 * so a run does not depend on how the chains are split over handles / GPUs.
 * mcd_mc3_swap proposes, in every group, the exchange of the chains at slots (pair, pair + 1) (pair = -1: drawn per
 * group): ln r = (beta_p - beta_{p+1}) (ln pi(x_{p+1}) - ln pi(x_p)); accepted exchanges swap the chains' SLOTS, states
 * never move.  d_stats_global: DEVICE pointer to the (ln prior, ln likelihood) pairs of all n_global chains [n_global][2]
 * (all-gathered over NCCL by the host from mcd_chains_out_device, columns MCD_OUT_LNPRIOR / MCD_OUT_LNLIK) or NULL when
 * this handle holds all chains.  Every rank takes identical decisions (Philox counter (group, iteration, 0, 3)).
 * accepted[g] (host, nullable) = 1 / 0 per group.  mcd_mc3_slots: current slot of every chain [n_global]. */
int mcd_mc3_configure(mcd_handle* h, int32_t n_global, int32_t chain_offset, int32_t chains_per_group,
                      const double* ladder_prior /*[chains_per_group]*/, const double* ladder_lik /*[chains_per_group]*/);
int mcd_mc3_swap(mcd_handle* h, int32_t pair, uint64_t seed, uint32_t iteration, const double* d_stats_global,
                 int32_t* accepted /*[n_global / chains_per_group] or NULL*/);
int mcd_mc3_slots(mcd_handle* h, int32_t* slots /*[n_global]*/);
void* mcd_chains_out_device(mcd_handle* h); /* device pointer: [n_resident][MCD_OUT_COLS] of the resident chains */
/* (ln prior, ln likelihood) of the resident chains into a caller's DEVICE buffer [n_resident][2] -- the send buffer of
 * the all-gather; returns after the copy has completed.  The caller synchronises its own collective before
 * mcd_mc3_swap reads the gathered table. */
int mcd_chains_stats_device(mcd_handle* h, double* d_stats);

/* The one exchange between GPUs: MC3 swap statistics.  One process per GPU, one handle per process; the chains of a
 * temperature group may be spread over the ranks.  NCCL is loaded at run time (dlopen "libnccl.so.2": the library itself has no
 * link-time dependency on it).
 *   mcd_comm_unique_id  rank 0 creates the 128-byte NCCL id; the host distributes it to the other ranks (any channel);
 *   mcd_comm_init       every rank joins with (world, rank, id); collective;
 *   mcd_allgather_stats (ln prior, ln likelihood) of this rank's resident chains -> [world * n_resident][2] on every rank, into a
 *                       caller's DEVICE buffer (the table mcd_mc3_swap takes); all ranks must hold the same number of resident
 *                       chains; ncclAllGather on the library's stream, returns after completion;
 *   mcd_comm_destroy    leaves the communicator (also done by mcd_destroy). */
int mcd_comm_unique_id(void* id128 /* [128] bytes out */);
int mcd_comm_init(mcd_handle* h, int32_t world, int32_t rank, const void* id128);
int mcd_allgather_stats(mcd_handle* h, double* d_stats_global /* device, [world * n_resident][2] */);
int mcd_comm_destroy(mcd_handle* h);

/* Arithmetic pipe of the precision-matrix contraction Y = DX . Sigma^-1 on large trees (the dominant kernel).
 *   MCD_CONTRACT_DMMA   FP64 tensor instructions (mma.sync m8n8k4.f64), plain FP64 GEMM rounding
 *   MCD_CONTRACT_I8_Sn  INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators) on n balanced base-256 digit
 *                       planes per operand row (error-free "Ozaki" splitting; integer dot products are exact, so
 *                       results are bit-reproducible).  S7 (default): 56 bits per operand, error of the order of
 *                       an FP64 GEMM's rounding; S6: 48 bits, ~2^-8 times coarser, relative to
 *                       |Sigma^-1 row| . |dx|  (DESIGN.md).
 * The environment variable MCD_CONTRACTION = dmma | i8s6 | i8s7 picks the default of new handles. */
enum { MCD_CONTRACT_DMMA = 0, MCD_CONTRACT_I8_S6 = 6, MCD_CONTRACT_I8_S7 = 7 };
int mcd_set_contraction(mcd_handle* h, int32_t mode);
int mcd_get_contraction(const mcd_handle* h);

/* evaluation with DEVICE buffers already resident in HBM (no copies); `stream` is a cudaStream_t
 * or NULL.  Asynchronous: returns after enqueueing. */
int mcd_eval_device(mcd_handle* h, int32_t n_chains, const double* d_states, double* d_out,
                    int32_t* d_status, void* stream);
int mcd_eval_grad_device(mcd_handle* h, int32_t n_chains, const double* d_states, double* d_out,
                         double* d_grad, int32_t* d_status, void* stream);

/* bookkeeping */
int64_t mcd_kernel_launches(const mcd_handle* h); /* kernels launched by this handle so far */
int mcd_synchronize(mcd_handle* h);
/* optional per-kernel timing with CUDA events on the launching stream (used by bench.py's roofline):
 * ms[0] = residual kernel, ms[1] = FP64 contraction, ms[2] = posterior kernel, summed over n_calls */
int mcd_set_kernel_timing(mcd_handle* h, int on);
int mcd_kernel_times(mcd_handle* h, double* ms /*[3]*/, int64_t* n_calls);
const char* mcd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCMCDATE_B200_H */
