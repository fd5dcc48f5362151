// oracle/oracle.hpp -- CPU restatement of McmcDate's prior / likelihood / Jacobian / HMC target.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is linked into, imported by or called
// from the product (mcmc-date_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, as the checker or as the timed CPU baseline.
//
// The reference (dschrempf/mcmc-date) is Haskell; no GHC toolchain exists in this image, so the
// reference itself cannot be compiled or run (DESIGN.md "Oracle").  Every function below restates
// one reference function and cites it (paths relative to /root/reference).  Scalar densities of
// the un-vendored third-party `mcmc` package (github dschrempf/mcmc, rev 542c43f6..., module
// Mcmc.Prior: exponential, gamma, normal, product', gammaMeanVarianceToShapeScale) are restated
// from their published definitions; their call sites in the reference are cited instead.
//
// PARITY PINNING.  Pinned by the reference's own known answers: the birth-death prior
// (lib/Mcmc/Tree/Prior/BirthDeath.hs:249-271, RevBayes-checked doc values) -- see
// tests/test_oracle_pins.py.  Everything else (clock priors, calibrations, constraints, braces,
// MVN likelihood, Jacobian, gradient): the reference holds NO golden vectors or unit tests ->
// "parity unpinned" by the reference for those functions; they are cross-checked against an
// independent mpmath restatement (oracle/mp_oracle.py) and closed forms instead.
//
// The evaluator is templated on the scalar type like the reference's `RealFloat a =>` code:
//   T = double       value path        (PriorFunction I, LikelihoodFunction I)
//   T = Dual         forward-mode AD   (stands in for the `ad` package used by the reference's NUTS,
//                                       app/Hamiltonian.hs:85-92) -- gradient ground truth
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------------------------
// model description (flattened; pre-order node numbering, root = 0)
// ----------------------------------------------------------------------------------------------
enum ClockModel { UGAMMA = 0, ULOGNORMAL = 1, UWHITENOISE = 2, ALOGNORMAL = 3 };  // app/Probability.hs:88-93
enum LikKind { LIK_FULL = 0, LIK_UNIVARIATE = 1, LIK_NONE = 2, LIK_SPARSE = 3 };   // app/Probability.hs:210-235
enum Status {
  ST_REF_ERROR = 1,   // the reference would have called Haskell `error` (process abort)
  ST_ZERO = 2,        // probability zero (ln = -inf) somewhere in prior*likelihood*jacobian
  ST_NAN = 4,         // NaN result
  ST_NEARCRIT = 8,    // |lambda-mu| < 1e-6: near-critical birth-death formulas were used
  ST_LEAF_HEIGHT = 16 // a leaf height is not exactly 0 (HeightTree invariant violated)
};

struct Model {
  int N = 0;                              // nodes (2n-1)
  std::vector<int> parent, child0, child1;  // pre-order; -1 = none
  std::vector<int> bidx;                  // node -> MVN dimension k(i); root -1
  int K = 0;                              // N-2
  std::vector<double> mu;                 // [K]
  std::vector<double> prec;               // [K*K] row-major Sigma^-1 (LIK_FULL) or [K] variances
  bool prec_symmetric = true;             // LIK_FULL: prec == prec^T exactly (`prepare` writes an unsymmetrised LU inverse)
  std::vector<int> sp_row, sp_col;        // LIK_SPARSE: association list ((i, j), v) as stored by the
  std::vector<double> sp_val;             //   reference (SparseS, app/Main.hs:75-81,95-97)
  double logdet = 0;                      // ln det Sigma   (or sum ln var)
  int clock = ULOGNORMAL, lik = LIK_FULL;
  double ht = 1.0;                        // mean root height (app/Main.hs:394)
  // calibrations (lib/Mcmc/Tree/Prior/Node/Calibration.hs:55-123)
  std::vector<int> cal_idx;
  std::vector<uint8_t> cal_has_lo, cal_has_hi;
  std::vector<double> cal_lo, cal_plo, cal_hi, cal_phi;
  // constraints (lib/Mcmc/Tree/Prior/Node/Constraint.hs:61-74)
  std::vector<int> con_y, con_o;
  std::vector<double> con_p;
  // braces (lib/Mcmc/Tree/Prior/Node/Brace.hs:54-59), CSR
  std::vector<int> br_off, br_idx;
  std::vector<double> br_sd;
  int S() const { return 5 + 2 * N; }
};

struct RefError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ----------------------------------------------------------------------------------------------
// forward-mode dual numbers
// ----------------------------------------------------------------------------------------------
inline double digamma(double x) {
  double r = 0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12 + f * (1.0 / 120 + f * (-1.0 / 252 + f * (1.0 / 240 +
             f * (-1.0 / 132 + f * (691.0 / 32760 + f * (-1.0 / 12)))))));
  return r + std::log(x) - 0.5 / x + t;
}

struct Dual {
  double v, d;
  Dual() : v(0), d(0) {}
  Dual(double v_) : v(v_), d(0) {}
  Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
inline Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline bool operator<=(Dual a, Dual b) { return a.v <= b.v; }
inline bool operator>=(Dual a, Dual b) { return a.v >= b.v; }
inline bool operator==(Dual a, Dual b) { return a.v == b.v; }
inline bool operator!=(Dual a, Dual b) { return a.v != b.v; }

inline double primal(double x) { return x; }
inline double primal(Dual x) { return x.v; }
inline double xlog(double x) { return std::log(x); }
inline Dual xlog(Dual x) { return {std::log(x.v), x.d / x.v}; }
inline double xexp(double x) { return std::exp(x); }
inline Dual xexp(Dual x) { double e = std::exp(x.v); return {e, e * x.d}; }
inline double xsqrt(double x) { return std::sqrt(x); }
inline Dual xsqrt(Dual x) { double s = std::sqrt(x.v); return {s, 0.5 * x.d / s}; }
inline double xabs(double x) { return std::fabs(x); }
inline Dual xabs(Dual x) { return x.v < 0 ? -x : x; }
inline double xlgamma(double x) { return std::lgamma(x); }
inline Dual xlgamma(Dual x) { return {std::lgamma(x.v), digamma(x.v) * x.d}; }

static const double LN_SQRT_2PI = 0.9189385332046727418;  // math-functions m_ln_sqrt_2_pi
static const double NEG_INF = -std::numeric_limits<double>::infinity();

// ----------------------------------------------------------------------------------------------
// Mcmc.Prior scalar densities (third-party `mcmc`; call sites: app/Probability.hs:72-75,105-113;
// RelaxedClock.hs:118-120,231-234; Calibration.hs:391; Constraint.hs:415; Brace.hs:227).
// All return ln(density); probability zero = -inf.
// ----------------------------------------------------------------------------------------------
template <class T>
T prior_exponential(T l, T x) {
  if (l <= T(0.0)) throw RefError("exponential: Rate is zero or negative.");
  if (x < T(0.0)) return T(NEG_INF);
  return xlog(l) - l * x;
}
template <class T>
T prior_gamma(T k, T th, T x) {
  if (k <= T(0.0)) throw RefError("gamma: Shape is zero or negative.");
  if (th <= T(0.0)) throw RefError("gamma: Scale is zero or negative.");
  if (x <= T(0.0)) return T(NEG_INF);
  return xlog(x) * (k - T(1.0)) - (x / th) - xlgamma(k) - xlog(th) * k;
}
template <class T>
T prior_normal(T m, T s, T x) {
  if (s <= T(0.0)) throw RefError("normal: Standard deviation is zero or negative.");
  T xm = x - m;
  return (-(xm * xm) / (T(2.0) * s * s)) - (T(LN_SQRT_2PI) + xlog(s));
}
template <class T>
void mean_var_to_shape_scale(T m, T v, T& k, T& th) {  // gammaMeanVarianceToShapeScale
  k = m * m / v;
  th = v / m;
}

// ----------------------------------------------------------------------------------------------
// per-chain state view: canonical order [lambda, mu, H, h[N], m, v, r[N]]  (app/State.hs:70-100,
// derived Foldable; tree folds are pre-order, lib/Mcmc/Tree/Types.hs:91-95,146-150)
// ----------------------------------------------------------------------------------------------
template <class T>
struct StateView {
  const T* x;
  int N;
  T la() const { return x[0]; }
  T mu() const { return x[1]; }
  T H() const { return x[2]; }
  T h(int i) const { return x[3 + i]; }
  T m() const { return x[3 + N]; }
  T v() const { return x[4 + N]; }
  T r(int i) const { return x[5 + N + i]; }
};

// heightTreeToLengthTree (lib/Mcmc/Tree/Types.hs:224-233): t_0 = h_0 - h_0, t_i = h_parent - h_i
template <class T>
std::vector<T> height_to_length(const Model& M, const StateView<T>& s) {
  std::vector<T> t(M.N);
  t[0] = s.h(0) - s.h(0);
  for (int i = 1; i < M.N; ++i) t[i] = s.h(M.parent[i]) - s.h(i);
  return t;
}

// getBranches + sumFirstTwo (app/Tools.hs:36-48): literal list construction.
// Returns the node order of the un-merged branch list: [l, r, rest of l (pre-order), rest of r].
inline std::vector<int> get_branches_order(const Model& M) {
  if (M.child0[0] < 0 || M.child1[0] < 0) throw RefError("getBranches: Root node is not bifurcating.");
  int l = M.child0[0], r = M.child1[0];
  std::vector<int> ls, rs;  // `branches l`: pre-order of the subtree
  for (int i = l; i < r; ++i) ls.push_back(i);
  for (int i = r; i < M.N; ++i) rs.push_back(i);
  std::vector<int> out;
  out.push_back(ls[0]);
  out.push_back(rs[0]);
  for (size_t i = 1; i < ls.size(); ++i) out.push_back(ls[i]);
  for (size_t i = 1; i < rs.size(); ++i) out.push_back(rs[i]);
  return out;
}
// node -> k after sumFirstTwo; literal restatement used to pin the closed form of SURVEY R2
inline std::vector<int> branch_index_literal(const Model& M) {
  std::vector<int> ord = get_branches_order(M);
  std::vector<int> k(M.N, -1);
  for (size_t j = 0; j < ord.size(); ++j) k[ord[j]] = j < 2 ? 0 : (int)j - 1;
  return k;
}

// distances (app/Probability.hs:201-207): zipWith (*) times rates -> sumFirstTwo -> map (*(tH*rMu))
template <class T>
std::vector<T> distances(const Model& M, const StateView<T>& s, const std::vector<T>& t) {
  std::vector<int> ord = get_branches_order(M);
  std::vector<T> prod(ord.size());
  for (size_t j = 0; j < ord.size(); ++j) prod[j] = t[ord[j]] * s.r(ord[j]);
  std::vector<T> d(M.K);
  d[0] = prod[0] + prod[1];
  for (int k = 1; k < M.K; ++k) d[k] = prod[k + 1];
  T sc = s.H() * s.m();
  for (int k = 0; k < M.K; ++k) d[k] = d[k] * sc;
  return d;
}

// ----------------------------------------------------------------------------------------------
// likelihood (app/Probability.hs:166-193 Double path; :286-326 generic path)
// ----------------------------------------------------------------------------------------------
// Double path: (dxs <# sigmaInv) <.> dxs   -- vector-matrix product then dot
inline double mvn_full_double(const Model& M, const std::vector<double>& d, std::vector<double>* y_out) {
  const int K = M.K;
  std::vector<double> dx(K), y(K, 0.0);
  for (int k = 0; k < K; ++k) dx[k] = d[k] - M.mu[k];
  for (int i = 0; i < K; ++i) {  // y = dx <# P = sum_i dx_i * P[i,:]
    const double a = dx[i];
    const double* row = &M.prec[(size_t)i * K];
    for (int j = 0; j < K; ++j) y[j] += a * row[j];
  }
  double quad = 0;
  for (int k = 0; k < K; ++k) quad += y[k] * dx[k];
  if (y_out) *y_out = y;
  double c = -(LN_SQRT_2PI * (double)K);
  return c + (-0.5) * (M.logdet + quad);
}
// generic path: reduceVMV (app/Probability.hs:286-298): foldl' (+) 0 [vl_i * m_ij * vr_j | i, j]
template <class T>
T mvn_full_generic(const Model& M, const std::vector<T>& d) {
  const int K = M.K;
  std::vector<T> dx(K);
  for (int k = 0; k < K; ++k) dx[k] = d[k] - T(M.mu[k]);
  T acc(0.0);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) acc = acc + dx[i] * T(M.prec[(size_t)i * K + j]) * dx[j];
  T c = T(-(LN_SQRT_2PI * (double)K));
  return c + T(-0.5) * (T(M.logdet) + acc);
}
// logDensitySparseMultivariateNormal (app/Probability.hs:178-184): dxs <.> (sigmaInvS !#> dxs)
template <class T>
T mvn_sparse(const Model& M, const std::vector<T>& d, std::vector<T>* y_out = nullptr) {
  const int K = M.K;
  std::vector<T> dx(K), y(K, T(0.0));
  for (int k = 0; k < K; ++k) dx[k] = d[k] - T(M.mu[k]);
  for (size_t e = 0; e < M.sp_val.size(); ++e) y[M.sp_row[e]] = y[M.sp_row[e]] + T(M.sp_val[e]) * dx[M.sp_col[e]];
  T quad(0.0);
  for (int k = 0; k < K; ++k) quad = quad + dx[k] * y[k];
  if (y_out) *y_out = y;
  T c = T(-(LN_SQRT_2PI * (double)K));
  return c + T(-0.5) * (T(M.logdet) + quad);
}
// logDensityUnivariateNormal (app/Probability.hs:186-193)
template <class T>
T mvn_univariate(const Model& M, const std::vector<T>& d) {
  T es(0.0);
  for (int k = 0; k < M.K; ++k) {
    T dx = d[k] - T(M.mu[k]);
    es = es + (dx * dx) / T(M.prec[k]);
  }
  T c = T(-(LN_SQRT_2PI * (double)M.K));
  return c + T(-0.5) * (T(M.logdet) + es);
}

// rootBranch / jacobianRootBranch (app/Probability.hs:393-410): Exp . log . recip . rootBranch
template <class T>
T ln_jacobian(const Model& M, const StateView<T>& s, const std::vector<T>& t) {
  int l = M.child0[0], r = M.child1[0];
  T rb = s.H() * s.m() * (t[l] * s.r(l) + t[r] * s.r(r));
  return xlog(T(1.0) / rb);
}

// ----------------------------------------------------------------------------------------------
// node priors (lib/Mcmc/Tree/Prior/Node/{Calibration,Constraint,Brace,Combined}.hs)
// ----------------------------------------------------------------------------------------------
static const double SQRT_2_OVER_PI = 0.7978845608028654;  // Calibration.hs:390

// calibrateSoftF (Calibration.hs:369-392) on an interval already transformed by
// transformCalibration (:426-430) / transformInterval (:94-105)
template <class T>
T calibrate_soft(const Model& M, int c, T H, T h) {
  T a(M.cal_lo[c]), b(M.cal_hi[c]);
  if (!(H == T(1.0))) {  // transformCalibration: `h == 1 = c`
    T x = T(1.0) / H;    // transformInterval (recip h)
    if (x <= T(0.0)) throw RefError("transformInterval: Multiplier is zero or negative.");
    a = x * a;
    b = x * b;
  }
  if (h < T(0.0)) return T(NEG_INF);
  T lower(0.0), upper(0.0);
  if (M.cal_has_lo[c] && h < a) {
    T s = T(SQRT_2_OVER_PI) * T(M.cal_plo[c]);
    lower = prior_normal(T(0.0), s, a - h) - prior_normal(T(0.0), s, T(0.0));
  }
  if (M.cal_has_hi[c] && h > b) {
    T s = T(SQRT_2_OVER_PI) * T(M.cal_phi[c]);
    upper = prior_normal(T(0.0), s, h - b) - prior_normal(T(0.0), s, T(0.0));
  }
  return lower + upper;
}
// constrainSoftF (Constraint.hs:403-416)
template <class T>
T constrain_soft(double p, T hY, T hO) {
  if (hY < hO) return T(0.0);
  T s = T(SQRT_2_OVER_PI) * T(p);
  return prior_normal(T(0.0), s, hY - hO) - prior_normal(T(0.0), s, T(0.0));
}
// braceSoftF (Brace.hs:218-231), allEqual (:195-199)
template <class T>
T brace_soft(double sd, const std::vector<T>& hs) {
  if (sd <= 0) throw RefError("braceSoftF: Standard deviation is zero or negative.");
  bool all_eq = true;
  for (size_t i = 1; i < hs.size(); ++i) all_eq = all_eq && (hs[i] == hs[0]);
  if (all_eq) return T(0.0);
  T sum(0.0);
  for (auto& h : hs) sum = sum + h;
  T mean = sum / T((double)hs.size());
  T d0 = prior_normal(T(0.0), T(sd), T(0.0));
  T acc(0.0);  // product = foldl (*) 1
  for (auto& h : hs) acc = acc + (prior_normal(T(0.0), T(sd), h - mean) - d0);
  return acc;
}
// calibrateConstrainBraceSoft (Combined.hs:70-85)
template <class T>
T prior_node(const Model& M, const StateView<T>& s) {
  T H = s.H();
  if (H <= T(0.0)) return T(NEG_INF);
  T cs(0.0), ks(0.0), bs(0.0);  // VB.product = foldl' (*) 1
  for (size_t c = 0; c < M.cal_idx.size(); ++c) cs = cs + calibrate_soft(M, (int)c, H, s.h(M.cal_idx[c]));
  for (size_t c = 0; c < M.con_y.size(); ++c)
    ks = ks + constrain_soft(M.con_p[c], s.h(M.con_y[c]), s.h(M.con_o[c]));
  for (size_t b = 0; b + 1 < M.br_off.size(); ++b) {
    std::vector<T> hs;
    for (int j = M.br_off[b]; j < M.br_off[b + 1]; ++j) hs.push_back(s.h(M.br_idx[j]));
    bs = bs + brace_soft(M.br_sd[b], hs);
  }
  return cs + ks + bs;
}

// ----------------------------------------------------------------------------------------------
// birth-death prior (lib/Mcmc/Tree/Prior/BirthDeath.hs)
// ----------------------------------------------------------------------------------------------
template <class T>
void compute_de(T la, T mu, T rho, T dt, T e0, T& pD, T& pE) {  // :53-79
  T d = la - mu;
  T x = xexp(-d * dt);
  T c = (T(1.0) - rho) + rho * e0;
  T y = (mu - c * la) * x;
  T nomD = d * d * x;
  T c1 = c - T(1.0);
  T nomE = mu * c1 + y;
  T denom = la * c1 + y;
  pD = nomD / denom / denom;
  pE = nomE / denom;
}
template <class T>
void compute_de_near_critical(T la, T mu, T rho, T dt, T e0, T& pD, T& pE) {  // :90-114
  T d = la - mu;
  T c = (T(1.0) - rho) + rho * e0;
  T y = (mu - c * la) * dt;
  T nomD = T(1.0) - d * dt;
  T nomE = c + y;
  T denom = T(1.0) + y;
  pD = nomD / denom / denom;
  pE = nomE / denom;
}
static const double EPS_NEAR_CRITICAL = 1e-6;  // :125-126

// birthDeathWith (:186-239); general tree given by children lists (unary nodes allowed, as in the
// reference); returns (ln D, E)
template <class T>
struct BDTree {
  const std::vector<int>*c0, *c1;
  const std::vector<T>* br;
};
template <class T>
void birth_death_with(bool nearcrit, T la, T mu, T rho, const BDTree<T>& tr, int node, T& lnD, T& E) {
  T br = (*tr.br)[node];
  int a = (*tr.c0)[node], b = (*tr.c1)[node];
  if (br <= T(0.0)) { lnD = T(NEG_INF); E = T(1.0); return; }
  T dT, eT;
  if (a >= 0 && b >= 0) {
    T dL, eL, dR, eR;
    birth_death_with(nearcrit, la, mu, rho, tr, a, dL, eL);
    birth_death_with(nearcrit, la, mu, rho, tr, b, dR, eR);
    if (nearcrit) compute_de_near_critical(la, mu, T(1.0), br, eL, dT, eT);
    else compute_de(la, mu, T(1.0), br, eL, dT, eT);
    lnD = xlog(dT * la) + dL + dR;
    E = eT;
  } else if (a >= 0) {
    T d, e;
    birth_death_with(nearcrit, la, mu, rho, tr, a, d, e);
    if (nearcrit) compute_de_near_critical(la, mu, T(1.0), br, e, dT, eT);
    else compute_de(la, mu, T(1.0), br, e, dT, eT);
    lnD = xlog(dT * rho) + d;
    E = eT;
  } else {
    if (nearcrit) compute_de_near_critical(la, mu, rho, br, T(0.0), dT, eT);
    else compute_de(la, mu, rho, br, T(0.0), dT, eT);
    lnD = xlog(dT * rho);
    E = eT;
  }
}
// birthDeath ConditionOnTimeOfOrigin (:158-172) on the subtree rooted at `node` (with its stem)
template <class T>
T birth_death_origin(T la, T mu, T rho, const BDTree<T>& tr, int node, bool* nearcrit_used = nullptr) {
  if (la < T(0.0)) throw RefError("birthDeath: Birth rate is negative.");
  if (mu < T(0.0)) throw RefError("birthDeath: Death rate is negative.");
  if (rho <= T(0.0)) throw RefError("birthDeath: Sampling rate is zero or negative.");
  if (rho > T(1.0)) throw RefError("birthDeath: Sampling rate is larger than 1.");
  bool nc = T(EPS_NEAR_CRITICAL) > xabs(la - mu);
  if (nearcrit_used) *nearcrit_used = nc;
  T lnD, E;
  birth_death_with(nc, la, mu, rho, tr, node, lnD, E);
  return lnD;
}
// birthDeath ConditionOnTimeOfMrca (:173-177)
template <class T>
T birth_death_mrca(T la, T mu, T rho, const BDTree<T>& tr, bool* nearcrit_used = nullptr) {
  int l = (*tr.c0)[0], r = (*tr.c1)[0];
  if (l < 0 || r < 0) throw RefError("birthDeath: Tree is not bifurcating.");
  return birth_death_origin(la, mu, rho, tr, l, nearcrit_used) + birth_death_origin(la, mu, rho, tr, r);
}

// ----------------------------------------------------------------------------------------------
// relaxed clock priors (lib/Mcmc/Tree/Prior/Branch/RelaxedClock.hs, lib/Mcmc/Tree/Prior/Branch.hs)
// ----------------------------------------------------------------------------------------------
template <class T>
T log_normal_prime(T m, T v, T x) {  // logNormal' (:141-150)
  if (v <= T(0.0)) throw RefError("logNormal': Variance is zero or negative.");
  if (x <= T(0.0)) return T(NEG_INF);
  T t = -(T(LN_SQRT_2PI) + xlog(x * xsqrt(v)));
  T a = T(1.0) / (T(2.0) * v);
  T b = xlog(x / m) + T(0.5) * v;
  T e = -(a * b * b);
  return t + e;
}
// per-branch density f(t_i, r_i) of the chosen model
template <class T>
T clock_branch(int model, T v, T t, T r) {
  const T one(1.0);
  switch (model) {
    case UGAMMA: {  // :110-126
      T k, th;
      mean_var_to_shape_scale(one, v, k, th);
      return prior_gamma(k, th, r);
    }
    case ULOGNORMAL:  // :160-172
      return log_normal_prime(one, v, r);
    case UWHITENOISE: {  // :209-241
      T vp = v / t, k, th;
      mean_var_to_shape_scale(one, vp, k, th);
      return prior_gamma(k, th, r);
    }
    default: {  // ALOGNORMAL :307-331 (no parent-rate term in the code)
      T vp = v * t;
      return log_normal_prime(one, vp, r);
    }
  }
}
// branchesWith WithStem (Branch.hs:24): foldl' (*) (f br) (map (branchesWith WithStem f) ts)
template <class T>
T branches_with_stem(const Model& M, int model, T v, const std::vector<T>& t, const StateView<T>& s, int node) {
  T acc = clock_branch(model, v, t[node], s.r(node));
  if (M.child0[node] >= 0) acc = acc + branches_with_stem(M, model, v, t, s, M.child0[node]);
  if (M.child1[node] >= 0) acc = acc + branches_with_stem(M, model, v, t, s, M.child1[node]);
  return acc;
}
// branchesWith WithoutStem (Branch.hs:25): foldl1' (*) over the root's children
template <class T>
T clock_model(const Model& M, const StateView<T>& s, const std::vector<T>& t) {
  T v = s.v();
  if (v <= T(0.0)) throw RefError("relaxed clock: Variance is zero or negative.");  // :117,:217,:315
  T acc = branches_with_stem(M, M.clock, v, t, s, M.child0[0]);
  acc = acc + branches_with_stem(M, M.clock, v, t, s, M.child1[0]);
  return acc;
}

// ----------------------------------------------------------------------------------------------
// product' (Mcmc.Prior): multiply left to right, return zero at the first zero factor; factors
// after it are never evaluated (laziness).  Implemented inline below with early returns.
// ----------------------------------------------------------------------------------------------
template <class T>
bool is_zero(T x) { return primal(x) == NEG_INF; }

template <class T>
struct Result {
  T lnA, lnB, lnC, lnPrior, lnLik, lnJac, lnPost;
  int status = 0;
};

// The three prior parts, each evaluated on its own (this is also how the reference's `prior`
// monitor logs them, app/Monitor.hs:27-56).
// A: priorFunctionCalibrationsConstraintsBraces (app/Probability.hs:46-63)
template <class T>
T eval_A(const Model& M, const StateView<T>& s) { return prior_node(M, s); }
// B: priorFunctionBirthDeath (:66-85) = product' [exponential 1 la, exponential 1 mu, birthDeath ...]
template <class T>
T eval_B(const Model& M, const StateView<T>& s, const std::vector<T>& t, bool* nearcrit) {
  const T Z(NEG_INF);
  T e1 = prior_exponential(T(1.0), s.la());
  if (is_zero(e1)) return Z;
  T e2 = prior_exponential(T(1.0), s.mu());
  if (is_zero(e2)) return Z;
  BDTree<T> tr{&M.child0, &M.child1, &t};
  T bd = birth_death_mrca(s.la(), s.mu(), T(1.0), tr, nearcrit);
  if (is_zero(bd)) return Z;
  return e1 + e2 + bd;
}
// C: priorFunctionRelaxedMolecularClock (:96-124) = product' [exponential ht m, gamma 1.5 (1/6) v, model]
template <class T>
T eval_C(const Model& M, const StateView<T>& s, const std::vector<T>& t) {
  const T Z(NEG_INF);
  T e = prior_exponential(T(M.ht), s.m());
  if (is_zero(e)) return Z;
  T g = prior_gamma(T(3.0 / 2.0), T(1.0 / 6.0), s.v());
  if (is_zero(g)) return Z;
  T c = clock_model(M, s, t);
  if (is_zero(c)) return Z;
  return e + g + c;
}

// priorFunction (app/Probability.hs:127-150) = product' [A, B, C]: factors after the first zero are
// never evaluated, so an `error` (RefError) hiding behind a zero does not fire.
template <class T>
void eval_prior(const Model& M, const StateView<T>& s, const std::vector<T>& t, Result<T>& R) {
  const T Z(NEG_INF);
  bool errA = false, errB = false, errC = false, nc = false;
  try { R.lnA = eval_A(M, s); } catch (const RefError&) { errA = true; R.lnA = Z; }
  try { R.lnB = eval_B(M, s, t, &nc); } catch (const RefError&) { errB = true; R.lnB = Z; }
  try { R.lnC = eval_C(M, s, t); } catch (const RefError&) { errC = true; R.lnC = Z; }
  if (nc) R.status |= ST_NEARCRIT;
  if (errA) { R.status |= ST_REF_ERROR; R.lnPrior = Z; return; }
  if (is_zero(R.lnA)) { R.lnPrior = Z; return; }
  if (errB) { R.status |= ST_REF_ERROR; R.lnPrior = Z; return; }
  if (is_zero(R.lnB)) { R.lnPrior = Z; return; }
  if (errC) { R.status |= ST_REF_ERROR; R.lnPrior = Z; return; }
  if (is_zero(R.lnC)) { R.lnPrior = Z; return; }
  R.lnPrior = R.lnA + R.lnB + R.lnC;
}

// Full evaluation of one state.  generic_lik selects the generic (HMC target) likelihood form.
template <class T>
Result<T> eval_state(const Model& M, const T* x, bool generic_lik) {
  Result<T> R;
  StateView<T> s{x, M.N};
  for (int i = 0; i < M.N; ++i)
    if (M.child0[i] < 0 && primal(s.h(i)) != 0.0) R.status |= ST_LEAF_HEIGHT;
  std::vector<T> t = height_to_length(M, s);
  eval_prior(M, s, t, R);
  if (M.lik == LIK_NONE) {
    R.lnLik = T(0.0);
  } else {
    std::vector<T> d = distances(M, s, t);
    if (M.lik == LIK_UNIVARIATE) R.lnLik = mvn_univariate(M, d);
    else if (M.lik == LIK_SPARSE) R.lnLik = mvn_sparse(M, d);
    else R.lnLik = mvn_full_generic(M, d);
    (void)generic_lik;
  }
  R.lnJac = ln_jacobian(M, s, t);
  R.lnPost = R.lnPrior + R.lnLik + R.lnJac;  // HTarget: prior * likelihood * jacobian (Hamiltonian.hs:85-92)
  double p = primal(R.lnPost);
  if (p == NEG_INF) R.status |= ST_ZERO;
  if (p != p) R.status |= ST_NAN;
  return R;
}
// double specialisation of the likelihood: BLAS-like form used by the reference's Double path
inline Result<double> eval_state_double(const Model& M, const double* x, std::vector<double>* y_out = nullptr) {
  Result<double> R;
  StateView<double> s{x, M.N};
  for (int i = 0; i < M.N; ++i)
    if (M.child0[i] < 0 && s.h(i) != 0.0) R.status |= ST_LEAF_HEIGHT;
  std::vector<double> t = height_to_length(M, s);
  eval_prior(M, s, t, R);
  if (M.lik == LIK_NONE) {
    R.lnLik = 0.0;
  } else {
    std::vector<double> d = distances(M, s, t);
    R.lnLik = M.lik == LIK_UNIVARIATE ? mvn_univariate(M, d)
              : M.lik == LIK_SPARSE   ? mvn_sparse(M, d)
                                      : mvn_full_double(M, d, y_out);
  }
  R.lnJac = ln_jacobian(M, s, t);
  R.lnPost = R.lnPrior + R.lnLik + R.lnJac;
  if (R.lnPost == NEG_INF) R.status |= ST_ZERO;
  if (R.lnPost != R.lnPost) R.status |= ST_NAN;
  return R;
}

// getMask (app/Hamiltonian.hs:33-47): canonical order; free = everything except the root height,
// leaf heights, the rate stem, and H unless calibrations are available.
inline std::vector<uint8_t> get_mask(const Model& M, bool calibrations_available) {
  const int N = M.N;
  std::vector<uint8_t> m(M.S(), 1);
  m[2] = calibrations_available ? 1 : 0;
  m[3 + 0] = 0;
  for (int i = 0; i < N; ++i)
    if (M.child0[i] < 0 && M.child1[i] < 0) m[3 + i] = 0;
  m[5 + N + 0] = 0;
  return m;
}
// toVector (app/Hamiltonian.hs:49-53): conses while folding left -> REVERSED canonical order
inline std::vector<double> to_vector(const std::vector<uint8_t>& mask, const double* x) {
  std::vector<double> ys;
  for (size_t i = 0; i < mask.size(); ++i)
    if (mask[i]) ys.insert(ys.begin(), x[i]);
  return ys;
}
// fromVectorWith (app/Hamiltonian.hs:55-60): reads theta from the end
inline void from_vector_with(const std::vector<uint8_t>& mask, const double* x, const double* theta, int D, double* out) {
  int i = D - 1;
  for (size_t j = 0; j < mask.size(); ++j) {
    if (mask[j]) out[j] = theta[i--];
    else out[j] = x[j];
  }
}

// ----------------------------------------------------------------------------------------------
// gradient ground truth: forward-mode duals through the restated reference code (one pass per free
// parameter), and directional derivatives (one pass) for large trees.
// ----------------------------------------------------------------------------------------------
inline double dir_derivative(const Model& M, const double* x, const double* dir, double* value = nullptr) {
  std::vector<Dual> xs(M.S());
  for (int i = 0; i < M.S(); ++i) xs[i] = Dual(x[i], dir[i]);
  Result<Dual> R = eval_state<Dual>(M, xs.data(), true);
  if (value) *value = R.lnPost.v;
  return R.lnPost.d;
}
inline void grad_dual(const Model& M, const double* x, const uint8_t* mask, double* grad) {
  std::vector<Dual> xs(M.S());
  for (int j = 0; j < M.S(); ++j) {
    grad[j] = 0.0;
    if (!mask[j]) continue;
    for (int i = 0; i < M.S(); ++i) xs[i] = Dual(x[i], i == j ? 1.0 : 0.0);
    grad[j] = eval_state<Dual>(M, xs.data(), true).lnPost.d;
  }
}

}  // namespace orc
