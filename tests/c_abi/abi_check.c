/* abi_check.c -- the C ABI of libmcmcdate_b200.so exercised WITHOUT Python: dlopen, mcd_create from plain arrays, mcd_eval_grad,
 * mcd_eval, mcd_mask / mcd_to_vector / mcd_from_vector, mcd_destroy, and a comparison with the oracle's values stored in a text
 * fixture (tests/golden/abi_case_12_leaves.txt, written by tests/golden/make_fixtures.py from the reference's 12-leaf data set).
 *
 *   gcc -O1 -o abi_check tests/c_abi/abi_check.c -ldl -lm
 *   ./abi_check <libmcmcdate_b200.so> --symbols            dlopen + dlsym of every entry point used here (no GPU needed)
 *   ./abi_check <libmcmcdate_b200.so> <fixture.txt>        full run on GPU 0; exit 0 iff everything agrees to 1e-10
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mcmcdate_b200.h"

#define LOAD(name)                                                                    \
  *(void**)(&p_##name) = dlsym(lib, #name);                                           \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static int (*p_mcd_create)(const mcd_model_desc*, mcd_handle**);
static void (*p_mcd_destroy)(mcd_handle*);
static const char* (*p_mcd_last_error)(const mcd_handle*);
static int (*p_mcd_state_len)(const mcd_handle*);
static int (*p_mcd_dim)(const mcd_handle*);
static int (*p_mcd_hmc_dim)(const mcd_handle*);
static int (*p_mcd_mask)(const mcd_handle*, uint8_t*);
static int (*p_mcd_to_vector)(const mcd_handle*, const double*, double*);
static int (*p_mcd_from_vector)(const mcd_handle*, const double*, const double*, double*);
static int (*p_mcd_eval)(mcd_handle*, int32_t, const double*, double*, int32_t*);
static int (*p_mcd_eval_grad)(mcd_handle*, int32_t, const double*, double*, double*, int32_t*);

static double* read_doubles(FILE* f, int n) {
  double* a = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    char tok[64];
    if (fscanf(f, "%63s", tok) != 1) { fprintf(stderr, "fixture truncated\n"); exit(3); }
    a[i] = strcmp(tok, "inf") == 0 ? INFINITY : strcmp(tok, "-inf") == 0 ? -INFINITY : strcmp(tok, "nan") == 0 ? NAN : strtod(tok, NULL);
  }
  return a;
}
static int32_t* read_ints(FILE* f, int n) {
  int32_t* a = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i)
    if (fscanf(f, "%d", &a[i]) != 1) { fprintf(stderr, "fixture truncated\n"); exit(3); }
  return a;
}
static double relerr(double a, double b) {
  if (a == b || (isnan(a) && isnan(b))) return 0.0;
  return fabs(a - b) / fmax(1.0, fabs(b));
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: abi_check <lib.so> --symbols | <fixture.txt>\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(mcd_create) LOAD(mcd_destroy) LOAD(mcd_last_error) LOAD(mcd_state_len) LOAD(mcd_dim) LOAD(mcd_hmc_dim) LOAD(mcd_mask)
  LOAD(mcd_to_vector) LOAD(mcd_from_vector) LOAD(mcd_eval) LOAD(mcd_eval_grad)
  if (strcmp(argv[2], "--symbols") == 0) { printf("symbols ok\n"); return 0; }

  FILE* f = fopen(argv[2], "r");
  if (!f) { perror(argv[2]); return 2; }
  int N, clock, lik, ncal, ncon, B;
  if (fscanf(f, "%d %d %d %d %d %d", &N, &clock, &lik, &ncal, &ncon, &B) != 6) return 3;
  const int K = N - 2, S = 5 + 2 * N;
  mcd_model_desc d;
  memset(&d, 0, sizeof d);
  d.n_nodes = N; d.clock_model = clock; d.likelihood = lik;
  d.parent = read_ints(f, N);
  d.mean = read_doubles(f, K);
  d.precision = read_doubles(f, K * K);
  double* sc = read_doubles(f, 2);
  d.logdet_sigma = sc[0]; d.ht = sc[1];
  d.n_cal = ncal; d.cal_node = read_ints(f, ncal);
  d.cal_lo = read_doubles(f, ncal); d.cal_lo_p = read_doubles(f, ncal); d.cal_hi = read_doubles(f, ncal); d.cal_hi_p = read_doubles(f, ncal);
  d.n_con = ncon; d.con_young = read_ints(f, ncon); d.con_old = read_ints(f, ncon); d.con_p = read_doubles(f, ncon);
  d.n_brace = 0; d.device = 0; d.max_batch = B;
  double* X = read_doubles(f, B * S);
  double* want_out = read_doubles(f, B * 7);
  double* want_grad = read_doubles(f, B * S);
  int32_t* want_status = read_ints(f, B);
  fclose(f);

  mcd_handle* h = NULL;
  if (p_mcd_create(&d, &h) != 0) { fprintf(stderr, "mcd_create: %s\n", p_mcd_last_error(NULL)); return 1; }
  if (p_mcd_state_len(h) != S || p_mcd_dim(h) != K) { fprintf(stderr, "shape mismatch\n"); return 1; }
  double* out = (double*)malloc(sizeof(double) * B * MCD_OUT_COLS);
  double* out2 = (double*)malloc(sizeof(double) * B * MCD_OUT_COLS);
  double* grad = (double*)malloc(sizeof(double) * B * S);
  int32_t* status = (int32_t*)malloc(sizeof(int32_t) * B);
  int32_t* status2 = (int32_t*)malloc(sizeof(int32_t) * B);
  if (p_mcd_eval_grad(h, B, X, out, grad, status) != 0 || p_mcd_eval(h, B, X, out2, status2) != 0) {
    fprintf(stderr, "mcd_eval*: %s\n", p_mcd_last_error(h));
    return 1;
  }
  double worst = 0.0, worst_g = 0.0;
  int bad = 0;
  for (int b = 0; b < B; ++b) {
    if (status[b] != want_status[b] || status2[b] != want_status[b]) { fprintf(stderr, "status of chain %d: %d / %d, expected %d\n", b, status[b], status2[b], want_status[b]); bad = 1; }
    for (int c = 0; c < 7; ++c) {
      worst = fmax(worst, relerr(out[b * MCD_OUT_COLS + c], want_out[b * 7 + c]));
      worst = fmax(worst, relerr(out2[b * MCD_OUT_COLS + c], want_out[b * 7 + c]));
    }
    if (!isfinite(want_out[b * 7 + 6])) continue;  /* the gradient is defined where ln posterior is finite */
    double gmax = 1.0;
    for (int j = 0; j < S; ++j) gmax = fmax(gmax, fabs(want_grad[b * S + j]));
    for (int j = 0; j < S; ++j) worst_g = fmax(worst_g, fabs(grad[b * S + j] - want_grad[b * S + j]) / gmax);
  }
  /* mask / vector packing round trip (getMask, toVector, fromVectorWith) */
  const int D = p_mcd_hmc_dim(h);
  uint8_t* mask = (uint8_t*)malloc((size_t)S);
  double* theta = (double*)malloc(sizeof(double) * (size_t)D);
  double* back = (double*)malloc(sizeof(double) * (size_t)S);
  p_mcd_mask(h, mask);
  p_mcd_to_vector(h, X, theta);
  p_mcd_from_vector(h, X, theta, back);
  int nfree = 0;
  for (int j = 0; j < S; ++j) { nfree += mask[j]; if (back[j] != X[j]) bad = 1; }
  if (nfree != D) bad = 1;
  printf("abi_check: %d chains, value relerr %.3e, gradient relerr %.3e, hmc dim %d, %s\n", B, worst, worst_g, D, bad ? "MISMATCH" : "ok");
  p_mcd_destroy(h);
  dlclose(lib);
  return (bad || !(worst < 1e-10) || !(worst_g < 1e-10)) ? 1 : 0;
}
