// oracle/oracle_capi.cpp -- extern "C" surface of the CPU oracle for ctypes (tests, smoke, bench).
// *** TEST INFRASTRUCTURE ONLY (see oracle.hpp). ***
#include <atomic>
#include <cstring>
#include <thread>

#include "oracle.hpp"
#include "oracle_grad.hpp"

using namespace orc;

// dynamic parallel-for over chains on host threads (std::thread; no OpenMP runtime needed)
template <class F>
static void parallel_for(int B, int nthreads, F f) {
  if (nthreads <= 1 || B <= 1) {
    for (int b = 0; b < B; ++b) f(b);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&]() {
      for (int b = next.fetch_add(1); b < B; b = next.fetch_add(1)) f(b);
    });
  for (auto& t : th) t.join();
}

extern "C" {

// Build a model.  Calibration c has a lower bound iff cal_lo[c] > 0 and an upper bound iff
// cal_hi[c] is finite (Interval, lib/Mcmc/Tree/Prior/Node/Calibration.hs:53-86).
void* orc_model_create(int N, const int* parent, const int* child0, const int* child1, int K, const double* mu,
                       const double* prec_or_var, double logdet, int clock, int lik, double ht, int ncal,
                       const int* cal_idx, const double* cal_lo, const double* cal_plo, const double* cal_hi,
                       const double* cal_phi, int ncon, const int* con_y, const int* con_o, const double* con_p,
                       int nbrace, const int* br_off, const int* br_idx, const double* br_sd, int nnz,
                       const int* sp_row, const int* sp_col, const double* sp_val) {
  Model* M = new Model();
  M->N = N;
  M->parent.assign(parent, parent + N);
  M->child0.assign(child0, child0 + N);
  M->child1.assign(child1, child1 + N);
  M->K = K;
  M->clock = clock;
  M->lik = lik;
  M->ht = ht;
  M->logdet = logdet;
  if (lik != LIK_NONE) M->mu.assign(mu, mu + K);
  if (lik == LIK_FULL || lik == LIK_UNIVARIATE) {
    size_t np = lik == LIK_FULL ? (size_t)K * K : (size_t)K;
    M->prec.assign(prec_or_var, prec_or_var + np);
    if (lik == LIK_FULL)
      for (int i = 0; i < K && M->prec_symmetric; ++i)
        for (int j = 0; j < i; ++j)
          if (M->prec[(size_t)i * K + j] != M->prec[(size_t)j * K + i]) { M->prec_symmetric = false; break; }
  }
  if (lik == LIK_SPARSE) {
    M->sp_row.assign(sp_row, sp_row + nnz);
    M->sp_col.assign(sp_col, sp_col + nnz);
    M->sp_val.assign(sp_val, sp_val + nnz);
  }
  for (int c = 0; c < ncal; ++c) {
    M->cal_idx.push_back(cal_idx[c]);
    M->cal_has_lo.push_back(cal_lo[c] > 0);
    M->cal_has_hi.push_back(std::isfinite(cal_hi[c]));
    M->cal_lo.push_back(cal_lo[c]);
    M->cal_plo.push_back(cal_plo[c]);
    M->cal_hi.push_back(cal_hi[c]);
    M->cal_phi.push_back(cal_phi[c]);
  }
  for (int c = 0; c < ncon; ++c) {
    M->con_y.push_back(con_y[c]);
    M->con_o.push_back(con_o[c]);
    M->con_p.push_back(con_p[c]);
  }
  if (nbrace > 0) {
    M->br_off.assign(br_off, br_off + nbrace + 1);
    M->br_idx.assign(br_idx, br_idx + br_off[nbrace]);
    M->br_sd.assign(br_sd, br_sd + nbrace);
  }
  try {
    M->bidx = branch_index_literal(*M);
  } catch (const RefError&) {
    delete M;
    return nullptr;
  }
  return M;
}
void orc_model_destroy(void* h) { delete static_cast<Model*>(h); }
int orc_state_len(void* h) { return static_cast<Model*>(h)->S(); }
void orc_branch_index(void* h, int* out) {
  Model* M = static_cast<Model*>(h);
  std::memcpy(out, M->bidx.data(), sizeof(int) * M->N);
}
void orc_mask(void* h, int calibrations_available, uint8_t* out) {
  Model* M = static_cast<Model*>(h);
  std::vector<uint8_t> m = get_mask(*M, calibrations_available != 0);
  std::memcpy(out, m.data(), m.size());
}
// HMC position vector <-> state (app/Hamiltonian.hs:49-60); returns D
int orc_to_vector(void* h, const uint8_t* mask, const double* x, double* theta) {
  Model* M = static_cast<Model*>(h);
  std::vector<uint8_t> m(mask, mask + M->S());
  std::vector<double> th = to_vector(m, x);
  std::memcpy(theta, th.data(), th.size() * 8);
  return (int)th.size();
}
void orc_from_vector(void* h, const uint8_t* mask, const double* x, const double* theta, int D, double* out) {
  Model* M = static_cast<Model*>(h);
  std::vector<uint8_t> m(mask, mask + M->S());
  from_vector_with(m, x, theta, D, out);
}

static void store(const Result<double>& R, double* out7, int* status) {
  out7[0] = R.lnA; out7[1] = R.lnB; out7[2] = R.lnC; out7[3] = R.lnPrior;
  out7[4] = R.lnLik; out7[5] = R.lnJac; out7[6] = R.lnPost;
  *status = R.status;
}

// states [B][S] chain-major; out [B][7] = lnA, lnB, lnC, lnPrior, lnLik, lnJac, lnPost
void orc_eval(void* h, int B, const double* states, double* out, int* status, int nthreads) {
  Model* M = static_cast<Model*>(h);
  const int S = M->S();
  parallel_for(B, nthreads, [&](int b) { store(eval_state_double(*M, states + (size_t)b * S), out + (size_t)b * 7, status + b); });
}
// value via the GENERIC target (reduceVMV order), as the reference's HMC target evaluates it
void orc_eval_generic(void* h, int B, const double* states, double* out, int* status) {
  Model* M = static_cast<Model*>(h);
  const int S = M->S();
  for (int b = 0; b < B; ++b) store(eval_state<double>(*M, states + (size_t)b * S, true), out + (size_t)b * 7, status + b);
}
// value + analytic gradient (CPU port; grad [B][S], masked entries 0)
void orc_eval_grad(void* h, int B, const double* states, const uint8_t* mask, double* out, double* grad, int* status,
                   int nthreads) {
  Model* M = static_cast<Model*>(h);
  const int S = M->S();
  parallel_for(B, nthreads, [&](int b) {
    store(eval_grad_double(*M, states + (size_t)b * S, mask, grad + (size_t)b * S), out + (size_t)b * 7, status + b);
  });
}
// gradient ground truth by forward-mode duals, one state
void orc_grad_dual(void* h, const double* state, const uint8_t* mask, double* grad) {
  grad_dual(*static_cast<Model*>(h), state, mask, grad);
}
double orc_dir_derivative(void* h, const double* state, const double* dir, double* value) {
  return dir_derivative(*static_cast<Model*>(h), state, dir, value);
}
// birth-death known-answer hook: ln birthDeath on an arbitrary tree given by children + branch
// lengths (pre-order), conditioning 0 = origin (whole tree with its stem), 1 = MRCA
double orc_birth_death(int N, const int* child0, const int* child1, const double* br, double la, double mu, double rho,
                       int condition_on_mrca) {
  std::vector<int> c0(child0, child0 + N), c1(child1, child1 + N);
  std::vector<double> t(br, br + N);
  BDTree<double> tr{&c0, &c1, &t};
  try {
    return condition_on_mrca ? birth_death_mrca(la, mu, rho, tr) : birth_death_origin(la, mu, rho, tr, 0);
  } catch (const RefError&) {
    return std::numeric_limits<double>::quiet_NaN();
  }
}
void orc_compute_de(double la, double mu, double rho, double dt, double e0, int nearcrit, double* out2) {
  if (nearcrit) compute_de_near_critical(la, mu, rho, dt, e0, out2[0], out2[1]);
  else compute_de(la, mu, rho, dt, e0, out2[0], out2[1]);
}
double orc_digamma(double x) { return digamma(x); }
int orc_max_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
