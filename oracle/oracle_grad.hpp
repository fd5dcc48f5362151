// oracle/oracle_grad.hpp -- hand-derived analytic gradient of ln(prior*likelihood*jacobian), CPU.
//
// *** TEST INFRASTRUCTURE ONLY (see oracle.hpp). ***  The reference has NO hand-written gradient:
// its NUTS proposal differentiates the generic target with the `ad` package
// (app/Hamiltonian.hs:85-104, app/Probability.hs:286-388).  This file is the "CPU port" of the
// same closed-form gradient the CUDA path evaluates ("Gradient that R18
// implies"); its ground truth is orc::grad_dual / orc::dir_derivative in oracle.hpp (forward-mode
// duals through the restated reference code), checked in tests/test_oracle_gradient.py.
// It is also what bench.py times as the CPU baseline ("port").
#pragma once
#include "oracle.hpp"

namespace orc {

// phi(z) = (1 - e^-z)/z and its derivative, stable near z = 0
inline double bd_phi(double z) { return z == 0.0 ? 1.0 : -std::expm1(-z) / z; }
inline double bd_dphi(double z) {
  if (std::fabs(z) < 0.3) {
    // sum_{n>=0} (-1)^(n+1) (n+1)/(n+2)! z^n
    double term_z = 1.0, fact = 2.0, s = 0.0;
    for (int n = 0; n <= 14; ++n) {
      s += ((n & 1) ? 1.0 : -1.0) * (n + 1) / fact * term_z;
      term_z *= z;
      fact *= (n + 3);
    }
    return s;
  }
  return (std::exp(-z) * (1.0 + z) - 1.0) / (z * z);
}
// ln p1(h) = -(la-mu) h - 2 ln(1 + mu h phi((la-mu) h)) and its partials (SURVEY.md section 9,
// "Verified identity", rewritten so that it stays finite at la == mu)
struct LnP1 { double v, dh, dla, dmu; };
inline LnP1 ln_p1(double la, double mu, double h) {
  double d = la - mu, z = d * h, x = std::exp(-z), phi = bd_phi(z), dphi = bd_dphi(z);
  double Q = 1.0 + mu * h * phi;
  LnP1 r;
  r.v = -z - 2.0 * std::log(Q);
  r.dh = -(la + mu * x) / Q;
  r.dla = -h - 2.0 * mu * h * h * dphi / Q;
  r.dmu = h - 2.0 * (h * phi - mu * h * h * dphi) / Q;
  return r;
}

// value (literal reference restatement) + analytic gradient in the canonical state layout
// [lambda, mu, H, h[N], m, v, r[N]]; masked entries (get_mask) are 0.
inline Result<double> eval_grad_double(const Model& M, const double* x, const uint8_t* mask, double* grad) {
  const int N = M.N, S = M.S();
  std::vector<double> y;
  Result<double> R = eval_state_double(M, x, &y);
  StateView<double> s{x, N};
  std::vector<double> t = height_to_length(M, s);
  for (int j = 0; j < S; ++j) grad[j] = 0.0;
  double* g_la = grad + 0; double* g_mu = grad + 1; double* g_H = grad + 2;
  double* g_h = grad + 3; double* g_m = grad + 3 + N; double* g_v = grad + 4 + N; double* g_r = grad + 5 + N;
  const double la = s.la(), mu = s.mu(), H = s.H(), m = s.m(), v = s.v();
  const double sc = H * m;
  const int l = M.child0[0], r = M.child1[0];
  std::vector<double> G(N, 0.0);  // d/dt_i

  // likelihood + jacobian:  w_k = dlnL/dd_k + dlnJ/dd_k
  {
    std::vector<double> w(M.K, 0.0);
    if (M.lik == LIK_FULL) {
      for (int k = 0; k < M.K; ++k) w[k] = -y[k];
      // y = dx <# P = P^T dx; the gradient of -1/2 dx^T P dx is -1/2 (P + P^T) dx.  `prepare` writes the LU inverse of the
      // covariance unsymmetrised (app/Main.hs:230), so P may differ from P^T in the last digits: add the other half then.
      if (!M.prec_symmetric) {
        std::vector<double> d = distances(M, s, t);
        for (int i = 0; i < M.K; ++i) {
          double z = 0;
          for (int j = 0; j < M.K; ++j) z += M.prec[(size_t)i * M.K + j] * (d[j] - M.mu[j]);
          w[i] = -0.5 * (y[i] + z);
        }
      }
    } else if (M.lik == LIK_UNIVARIATE) {
      std::vector<double> d = distances(M, s, t);
      for (int k = 0; k < M.K; ++k) w[k] = -(d[k] - M.mu[k]) / M.prec[k];
    } else if (M.lik == LIK_SPARSE) {  // d/d dx of -1/2 dx^T S dx = -1/2 (S + S^T) dx
      std::vector<double> d = distances(M, s, t);
      for (size_t e = 0; e < M.sp_val.size(); ++e) {
        const int i = M.sp_row[e], j = M.sp_col[e];
        w[i] -= 0.5 * M.sp_val[e] * (d[j] - M.mu[j]);
        w[j] -= 0.5 * M.sp_val[e] * (d[i] - M.mu[i]);
      }
    }
    double d0 = sc * (t[l] * s.r(l) + t[r] * s.r(r));
    w[0] -= 1.0 / d0;
    double sumWE = 0.0;
    for (int i = 1; i < N; ++i) {
      double wk = w[M.bidx[i]];
      g_r[i] += wk * sc * t[i];
      G[i] += wk * sc * s.r(i);
      sumWE += wk * t[i] * s.r(i);
    }
    *g_H += sumWE * m;
    *g_m += sumWE * H;
  }
  // clock prior C (app/Probability.hs:96-124)
  *g_m += -M.ht;
  *g_v += 0.5 / v - 6.0;
  for (int i = 1; i < N; ++i) {
    double ri = s.r(i), ti = t[i];
    if (M.clock == UGAMMA || M.clock == UWHITENOISE) {
      double k, th, dk_dv, dth_dv, dk_dt = 0, dth_dt = 0;
      if (M.clock == UGAMMA) { k = 1.0 / v; th = v; dk_dv = -1.0 / (v * v); dth_dv = 1.0; }
      else { k = ti / v; th = v / ti; dk_dv = -ti / (v * v); dth_dv = 1.0 / ti; dk_dt = 1.0 / v; dth_dt = -v / (ti * ti); }
      double f_k = std::log(ri) - digamma(k) - std::log(th);
      double f_th = ri / (th * th) - k / th;
      g_r[i] += (k - 1.0) / ri - 1.0 / th;
      *g_v += f_k * dk_dv + f_th * dth_dv;
      G[i] += f_k * dk_dt + f_th * dth_dt;
    } else {
      double w = M.clock == ULOGNORMAL ? v : v * ti;
      double u = std::log(ri) + 0.5 * w;
      double f_w = -0.5 / w + u * u / (2.0 * w * w) - u / (2.0 * w);
      g_r[i] += -1.0 / ri - u / (w * ri);
      if (M.clock == ULOGNORMAL) *g_v += f_w;
      else { *g_v += f_w * ti; G[i] += f_w * v; }
    }
  }
  // birth-death B
  if (std::fabs(la - mu) < EPS_NEAR_CRITICAL) {
    // near-critical regime: the reference switches to first-order formulas (BirthDeath.hs:90-126) whose
    // value differs from the exact one by O(|la-mu|); differentiate THOSE (reverse sweep over the
    // literal recursion; E is handed up from the LEFT child only, BirthDeath.hs:201-215)
    *g_la += -1.0; *g_mu += -1.0;
    const double d = la - mu;
    std::vector<double> E(N + 1, 0.0);
    for (int i = N - 1; i >= 1; --i) {  // children before parents; left child of i is i+1
      const bool inner = M.child0[i] >= 0;
      const double c = inner ? E[i + 1] : 0.0;
      const double yy = (mu - c * la) * t[i];
      E[i] = (c + yy) / (1.0 + yy);
    }
    double a = 0.0;  // adjoint of E_i
    for (int i = 1; i < N; ++i) {
      const bool inner = M.child0[i] >= 0;
      if (i == M.child0[0] || i == M.child1[0]) a = 0.0;  // E of the root's children is unused
      const double c = inner ? E[i + 1] : 0.0;
      const double yy = (mu - c * la) * t[i];
      const double gy = -2.0 / (1.0 + yy) + a * (1.0 - c) / ((1.0 + yy) * (1.0 + yy));
      G[i] += -d / (1.0 - d * t[i]) + gy * (mu - c * la);
      *g_la += -t[i] / (1.0 - d * t[i]) + gy * (-c * t[i]) + (inner ? 1.0 / la : 0.0);
      *g_mu += t[i] / (1.0 - d * t[i]) + gy * t[i];
      a = inner ? a / (1.0 + yy) + gy * (-la * t[i]) : 0.0;  // adjoint handed to the left child i+1
    }
  } else {
    // telescoped closed form of the D/E recursion
    int n_inner_nonroot = 0;
    *g_la += -1.0; *g_mu += -1.0;
    LnP1 p0 = ln_p1(la, mu, s.h(0));
    *g_la += 2.0 * p0.dla; *g_mu += 2.0 * p0.dmu;
    for (int i = 1; i < N; ++i) {
      if (M.child0[i] < 0) continue;
      ++n_inner_nonroot;
      LnP1 p = ln_p1(la, mu, s.h(i));
      g_h[i] += p.dh; *g_la += p.dla; *g_mu += p.dmu;
    }
    *g_la += n_inner_nonroot / la;
  }
  // node priors A
  for (size_t c = 0; c < M.cal_idx.size(); ++c) {
    int i = M.cal_idx[c];
    double h = s.h(i), a = M.cal_lo[c], b = M.cal_hi[c];
    bool scaled = !(H == 1.0);
    if (scaled) { a = (1.0 / H) * a; b = (1.0 / H) * b; }
    if (M.cal_has_lo[c] && h < a) {
      double sd = SQRT_2_OVER_PI * M.cal_plo[c];
      g_h[i] += (a - h) / (sd * sd);
      if (scaled) *g_H += (a - h) * M.cal_lo[c] / (sd * sd * H * H);
    }
    if (M.cal_has_hi[c] && h > b) {
      double sd = SQRT_2_OVER_PI * M.cal_phi[c];
      g_h[i] += -(h - b) / (sd * sd);
      if (scaled) *g_H += -(h - b) * M.cal_hi[c] / (sd * sd * H * H);
    }
  }
  for (size_t c = 0; c < M.con_y.size(); ++c) {
    double hY = s.h(M.con_y[c]), hO = s.h(M.con_o[c]);
    if (!(hY < hO)) {
      double sd = SQRT_2_OVER_PI * M.con_p[c];
      g_h[M.con_y[c]] += -(hY - hO) / (sd * sd);
      g_h[M.con_o[c]] += (hY - hO) / (sd * sd);
    }
  }
  for (size_t b = 0; b + 1 < M.br_off.size(); ++b) {
    int j0 = M.br_off[b], j1 = M.br_off[b + 1];
    bool all_eq = true; double sum = 0;
    for (int j = j0; j < j1; ++j) { all_eq = all_eq && s.h(M.br_idx[j]) == s.h(M.br_idx[j0]); sum += s.h(M.br_idx[j]); }
    if (all_eq) continue;
    double mean = sum / (j1 - j0), sd = M.br_sd[b];
    for (int j = j0; j < j1; ++j) g_h[M.br_idx[j]] += -(s.h(M.br_idx[j]) - mean) / (sd * sd);
  }
  // branch-length gradients back to node heights (children stencil)
  for (int j = 1; j < N; ++j) {
    if (M.child0[j] < 0) continue;
    g_h[j] += -G[j] + G[M.child0[j]] + G[M.child1[j]];
  }
  for (int j = 0; j < S; ++j)
    if (!mask[j]) grad[j] = 0.0;
  return R;
}

}  // namespace orc
