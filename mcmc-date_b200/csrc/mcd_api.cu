// mcd_api.cu -- C ABI (include/mcmcdate_b200.h) of the B200-native batched evaluator.
//
// Host-side mirror of the reference interface for this path (compiled code, like the reference):
// handle = what `getMcmcProps` closes over (app/Main.hs:370-457); mcd_eval / mcd_eval_grad =
// priorFunction / likelihoodFunction / jacobianRootBranch / HTarget evaluated for B states at once;
// mcd_mask / mcd_to_vector / mcd_from_vector = getMask / toVector / fromVectorWith
// (app/Hamiltonian.hs:33-60); mcd_branch_index = getBranches + sumFirstTwo (app/Tools.hs:36-48).
//
// There is no CPU fallback: every evaluation runs the three CUDA kernels
//   residual_kernel -> gemm_f64_dmma_kernel -> posterior_kernel                 (MCD_CONTRACT_DMMA)
//   residual_split_kernel -> gemm_i8_ozaki_kernel -> posterior_kernel           (MCD_CONTRACT_I8_*, default)
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <utility>
#include <string>
#include <vector>

#ifndef MCD_NO_NVTX
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler is attached
#endif

#include "../../include/mcmcdate_b200.h"
#include "gemm_f64.cuh"
#include "cholesky.cuh"
#include "gemm_i8_ozaki.cuh"
#include "hmc_kernels.cuh"
#include "mh_kernels.cuh"
#include "posterior_kernels.cuh"

using namespace mcd;

namespace {

std::string g_create_error;
constexpr int N_STREAMS = 4;
constexpr int MH_MAX_CYCLE = 4096;        // proposals per mcd_mh_cycle call
constexpr int SMALL_TREE_MAX_NODES = 96;  // warp-per-chain kernels up to this many nodes
#ifndef POST_MINB
#define POST_MINB 4
#endif

// NVTX range over one entry point (tracing): shows up as "mcd:<name>" in nsys / ncu --nvtx timelines
struct NvtxRange {
#ifndef MCD_NO_NVTX
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
#else
  explicit NvtxRange(const char*) {}
#endif
};

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct mcd_handle {
  int device = 0;
  // host copies
  int N = 0, K = 0, S = 0, D = 0, ldk = 0, Mp = 0, ldy = 0;
  std::vector<int32_t> parent, child1, bidx;
  std::vector<uint8_t> mask;
  // device model
  DevModel dm{};
  DevBuf d_parent, d_child1, d_inner, d_mu, d_var, d_P, d_U;
  DevBuf d_ck, d_ick;             // INT8 contraction: power-of-two equilibration c_k ~ 1 / sqrt(P_kk) and its reciprocal, [ld8]
  // MC3 swap statistics across GPUs (NCCL, loaded at run time)
  void* nccl_comm = nullptr;
  int comm_world = 0, comm_rank = -1;
  DevBuf d_stats_local;           // [n_resident][2] send buffer of the all-gather
  size_t d_stats_local_n = 0;
  DevBuf d_partials;              // [cap][POST_NPART] per-chain sums handed from posterior_kernel to posterior_assemble_kernel
  DevBuf d_wide, d_widecnt;       // [cap] chains whose residual range is too wide for the digit planes (FP64 fall-back); [cap][8] counts
  bool sparse = false;            // MCD_LIK_SPARSE on a large tree: CSR contraction instead of the dense one
  DevBuf d_sp_ptr, d_sp_col, d_sp_val;
  int chol_state = 0;             // 0 = not tried, 1 = U = L^T uploaded, -1 = not positive definite
  DevBuf d_cal_node, d_cal_lo, d_cal_hi, d_cal_slo, d_cal_shi, d_con_y, d_con_o, d_con_s, d_br_off, d_br_node, d_br_sd,
      d_inc_off, d_inc_ent;
  CUtensorMap tmP{}, tmX{}, tmU{};
  // work buffers (capacity `cap` chains, multiple of 128)
  int cap = 0;
  DevBuf d_dx, d_y;               // internal: residuals and P.dx
  DevBuf d_states, d_out, d_grad, d_status;  // staging for the host-buffer API
  DevBuf d_theta, d_gtheta, d_base, d_tidx, d_sidx;  // theta-packed API
  DevBuf d_mom, d_eps, d_invmass, d_energy, d_status_acc;  // device-resident leapfrog trajectories
  // chains resident in HBM for Metropolis-Hastings moves (mh_kernels.cuh)
  DevBuf d_chain, d_chain_out, d_chain_status, d_new_out, d_new_status, d_undo, d_rng, d_meta, d_lq, d_accepted, d_counters, d_cycle;
  std::vector<unsigned char> cycle_cache;   // bytes of the proposal list d_cycle holds
  DevBuf d_mh_child1, d_mh_size, d_mh_inner_cnt, d_mh_inner_list;
  int n_resident = 0, chain_cap = 0, n_inner_nonroot = 0, undo_stride = 0;
  std::vector<int32_t> br_off_h, br_node_h, sub_size_h;  // host copies of the brace table (argument checks of the brace proposals)
  // heated chains (MC3 / stepping stone): temperature ladders, slot of every chain of every rank, inverse table
  DevBuf d_slot, d_chain_of_slot, d_ladder_p, d_ladder_l, d_swap_acc;
  // incremental evaluation of small moves (mh_delta_kernel): cached y = Sigma^-1 dx of every resident chain + delta lists
  DevBuf d_chain_y, d_dl_n, d_dl_k, d_dl_d;
  bool inc_enabled = true;        // mcd_mh_set_incremental
  bool inc_ok = false;            // this resident set qualifies (large dense model, every chain's state valid)
  bool force_sym = false;         // value-only evaluations use the symmetric contraction (they must produce y, not L^T dx)
  int inc_steps = 0, refresh_every = 512, ldyc = 0;
  std::vector<double> base_host;       // base state currently on the device (theta-packed entry points)
  cudaEvent_t ticket_ev[8][N_STREAMS] = {};  // completion events of the last 8 asynchronous calls
  // ordering between the pipelined host-buffer calls (streams 0..3, shared work buffers) and everything else
  cudaEvent_t join_ev[N_STREAMS] = {};       // 'pipeline stream i has reached this point'
  cudaEvent_t serial_ev = nullptr;           // 'the last non-pipelined call has reached this point'
  cudaStream_t serial_stream = nullptr;
  bool pipeline_dirty = false, serial_dirty = false;
  int64_t next_ticket = 0;
  const double* y_override = nullptr;  // see enqueue(posterior_only)
  int y_override_ld = 0;
  int mc3_C = 0, mc3_n_global = 0, mc3_offset = 0;
  DevBuf d_nuts;                  // batched NUTS: trajectory ends, checkpoints, candidates, per-chain scalars
  size_t nuts_bytes = 0;
  int* nuts_flags = nullptr;      // pinned: active-chain counts of the last two ticks
  cudaEvent_t nuts_ev[2] = {nullptr, nullptr};
  // INT8 tensor-core contraction (gemm_i8_ozaki.cuh): digit planes of P (built once per plane count) and of
  // the chains' residuals (rebuilt by residual_split_kernel on every evaluation)
  int oz_S = 0;                   // 0: FP64 DMMA contraction; 6, 7: int8 digit planes
  int oz_P_S = 0, oz_X_S = 0;     // plane counts the buffers below were built for
  int ld8 = 0, Mp8 = 0;
  DevBuf d_pP, d_sP, d_pX, d_sX, d_pU, d_sU;   // planes + scales of Sigma^-1, of the residuals, of U = L^T
  int oz_U_S = 0;
  CUtensorMap tmA8{}, tmB8{}, tmU8{};
  CUtensorMap tmB8h{}, tmU8h{};   // the same planes with 32-row boxes: each CTA of a pair (cta_group::2 kernel) loads half a P-row tile
  cudaStream_t streams[N_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
  // device-path pipeline (eval_device on large dense models): chunks of chains run K1 / contraction / K3 on three streams so
  // that the HBM-side kernels of one chunk overlap the tensor-core contraction of its neighbours
  cudaStream_t pipe_st[3] = {nullptr, nullptr, nullptr};  // K1, contraction (high priority), K3
  std::vector<cudaEvent_t> pipe_ev;
  int pipe_chunks = 0;            // 0 / 1: off
  int pipe_k3_per_sm = 1;         // resident K3 CTAs per SM for all but the last chunk (0: unlimited)
  std::mutex mtx;
  std::string err;
  int64_t launches = 0;
  int n_sms = 148;
  // optional per-kernel timing (CUDA events on the launching stream), see mcd_set_kernel_timing
  bool timing = false;
  std::vector<cudaEvent_t> tev;  // 4 events per timed call: before K1, after K1, after GEMM, after K3
};

namespace {

int fail(mcd_handle* h, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return -1;
}
#define CU_TRY(h, call)                                                                       \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(h, std::string(#call) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

template <class T>
int upload(mcd_handle* h, DevBuf& b, const T* src, size_t n) {
  const size_t na = n ? n : 1;  // keep pointers valid for empty tables
  CU_TRY(h, cudaMalloc(&b.p, na * sizeof(T)));
  CU_TRY(h, cudaMemset(b.p, 0, na * sizeof(T)));
  if (src && n) CU_TRY(h, cudaMemcpy(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int ensure_i8(mcd_handle* h);

int ensure_capacity(mcd_handle* h, int n_chains, bool staging, bool grad) {
  int need = (n_chains + 2 * GEMM_BT - 1) / (2 * GEMM_BT) * (2 * GEMM_BT);   // whole CTA-pair tiles of 256 chains
  if (need > h->cap) {
    CU_TRY(h, cudaDeviceSynchronize());
    for (DevBuf* b : {&h->d_dx, &h->d_y, &h->d_states, &h->d_out, &h->d_grad, &h->d_status, &h->d_theta, &h->d_gtheta,
                      &h->d_mom, &h->d_eps, &h->d_energy, &h->d_status_acc, &h->d_pX, &h->d_sX, &h->d_wide, &h->d_widecnt, &h->d_partials}) {
      if (b->p) cudaFree(b->p);
      b->p = nullptr;
    }
    h->oz_X_S = 0;
    h->cap = need;
    CU_TRY(h, cudaMalloc(&h->d_partials.p, (size_t)need * POST_NPART * 8));
    {  // y = P dx (and scratch of the posterior kernel) exists for every likelihood kind
      const size_t ny = (size_t)need * h->ldy;
      CU_TRY(h, cudaMalloc(&h->d_y.p, ny * 8));
      CU_TRY(h, cudaMemset(h->d_y.p, 0, ny * 8));
    }
    if (h->dm.lik == MCD_LIK_FULL) {
      size_t ndx = (size_t)need * h->ldk;
      CU_TRY(h, cudaMalloc(&h->d_dx.p, ndx * 8));
      CU_TRY(h, cudaMemset(h->d_dx.p, 0, ndx * 8));  // k-padding columns and tail chains stay 0
      if (make_tile_map(&h->tmX, h->d_dx.as<double>(), need, h->ldk) != 0)
        return fail(h, "cuTensorMapEncodeTiled failed for the residual matrix");
    }
  }
  if (staging) {
    if (!h->d_states.p) CU_TRY(h, cudaMalloc(&h->d_states.p, (size_t)h->cap * h->S * 8));
    if (!h->d_out.p) CU_TRY(h, cudaMalloc(&h->d_out.p, (size_t)h->cap * MCD_OUT_COLS * 8));
    if (!h->d_status.p) CU_TRY(h, cudaMalloc(&h->d_status.p, (size_t)h->cap * 4));
    if (grad && !h->d_grad.p) CU_TRY(h, cudaMalloc(&h->d_grad.p, (size_t)h->cap * h->S * 8));
  }
  return ensure_i8(h);
}

// Value-only path (MH proposals): quad = |L^T dx|^2 with P = L L^T needs only the triangular half of the
// contraction's flops.  L is taken from the model description when supplied, else factorised here once ON THE DEVICE
// (cholesky.cuh, blocked right-looking, ~K^3/3 flops: milliseconds at K = 2000).  Not positive definite -> keep the symmetric
// product.
int ensure_cholesky(mcd_handle* h, const double* L_in) {
  if (h->chol_state != 0) return 0;
  if (!L_in && (h->sparse || !h->d_P.p || h->dm.lik != MCD_LIK_FULL)) { h->chol_state = -1; return 0; }  // sparse models: no dense factor
  const int K = h->K;
  const size_t nU = (size_t)h->Mp * h->ldk;
  if (!h->d_U.p) {
    CU_TRY(h, cudaMalloc(&h->d_U.p, nU * 8));
  }
  CU_TRY(h, cudaMemset(h->d_U.p, 0, nU * 8));
  if (L_in) {
    std::vector<double> U(nU, 0.0);  // U[m][k] = L[k][m], k >= m
    for (int k = 0; k < K; ++k)
      for (int m = 0; m <= k; ++m) U[(size_t)m * h->ldk + k] = L_in[(size_t)k * K + m];
    CU_TRY(h, cudaMemcpy(h->d_U.p, U.data(), nU * 8, cudaMemcpyHostToDevice));
  } else {
    double* W = nullptr;
    int* d_flag = nullptr;
    CU_TRY(h, cudaMalloc(&W, (size_t)K * h->ldk * 8));
    CU_TRY(h, cudaMalloc(&d_flag, 4));
    CU_TRY(h, cudaMemset(d_flag, 0, 4));
    CU_TRY(h, cudaMemcpy(W, h->d_P.p, (size_t)K * h->ldk * 8, cudaMemcpyDeviceToDevice));
    cudaError_t e = cholesky_device(W, h->ldk, K, d_flag, 0);
    if (e == cudaSuccess) {
      chol_transpose_kernel<<<dim3((K + 255) / 256, K), 256>>>(W, h->ldk, K, h->d_U.as<double>(), h->ldk);
      e = cudaGetLastError();
    }
    int flag = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
    cudaFree(W);
    cudaFree(d_flag);
    if (e != cudaSuccess) return fail(h, std::string("Cholesky factorisation: ") + cudaGetErrorString(e));
    h->launches += 3 * ((K + CH_NB - 1) / CH_NB) + 1;
    if (flag) { h->chol_state = -1; return 0; }   // not positive definite
  }
  if (make_tile_map(&h->tmU, h->d_U.as<double>(), h->Mp, h->ldk) != 0) return fail(h, "cuTensorMapEncodeTiled failed for the Cholesky factor");
  h->chol_state = 1;
  return 0;
}

// INT8 contraction: digit planes of P are built once per plane count; the chains' plane buffers follow the
// work-buffer capacity.  Called with the handle locked, after ensure_capacity.
template <int S>
int ensure_i8_planes(mcd_handle* h) {
  const int K = h->K;
  if (h->oz_P_S != S) {
    if (h->d_pP.p) { CU_TRY(h, cudaDeviceSynchronize()); cudaFree(h->d_pP.p); h->d_pP.p = nullptr; }
    if (h->d_sP.p) { cudaFree(h->d_sP.p); h->d_sP.p = nullptr; }
    const size_t stride = (size_t)h->Mp8 * h->ld8;
    CU_TRY(h, cudaMalloc(&h->d_pP.p, stride * S));
    CU_TRY(h, cudaMemset(h->d_pP.p, 0, stride * S));
    CU_TRY(h, cudaMalloc(&h->d_sP.p, (size_t)h->Mp8 * 8));
    CU_TRY(h, cudaMemset(h->d_sP.p, 0, (size_t)h->Mp8 * 8));
    // digit planes of the equilibrated matrix P' = C P C (C = diag(c_k), powers of two ~ 1 / sqrt(P_kk)): y = P x =
    // C (P' x') with x' = C x ... in the kernels' terms x'_k = x_k * ick[k] (ick = 1 / c_k), P'_mk = P_mk c_m c_k, and
    // y_m = ick[m] * sum_k P'_mk x'_k  -- the factor ick[m] goes into the published row scale
    oz_split_rows_kernel<S><<<(K + 7) / 8, 256>>>(h->d_P.as<double>(), h->ldk, K, K, h->d_pP.as<signed char>(), h->ld8,
                                                   stride, h->d_sP.as<double>(), 1.52587890625e-05 /* 2^-16 */,
                                                   h->d_ck.as<double>(), h->d_ck.as<double>(), h->d_ick.as<double>());
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaDeviceSynchronize());
    if (oz_make_plane_map(&h->tmB8, h->d_pP.as<signed char>(), (size_t)S * h->Mp8, h->ld8, OZ_N) != 0 ||
        oz_make_plane_map(&h->tmB8h, h->d_pP.as<signed char>(), (size_t)S * h->Mp8, h->ld8, OZ2_NH) != 0)
      return fail(h, "cuTensorMapEncodeTiled failed for the precision digit planes");
    CU_TRY(h, gemm_i8_ozaki_configure<S>());
    CU_TRY(h, cudaFuncSetAttribute(residual_split_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)((size_t)h->ld8 * 8)));
    CU_TRY(h, cudaFuncSetAttribute(delta_split_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)((size_t)h->ld8 * 8)));
    h->oz_P_S = S;
  }
  if (h->chol_state == 1 && h->oz_U_S != S) {  // value-only path: digit planes of the upper-triangular factor U = L^T
    if (h->d_pU.p) { CU_TRY(h, cudaDeviceSynchronize()); cudaFree(h->d_pU.p); h->d_pU.p = nullptr; }
    if (h->d_sU.p) { cudaFree(h->d_sU.p); h->d_sU.p = nullptr; }
    const size_t stride = (size_t)h->Mp8 * h->ld8;
    CU_TRY(h, cudaMalloc(&h->d_pU.p, stride * S));
    CU_TRY(h, cudaMemset(h->d_pU.p, 0, stride * S));
    CU_TRY(h, cudaMalloc(&h->d_sU.p, (size_t)h->Mp8 * 8));
    CU_TRY(h, cudaMemset(h->d_sU.p, 0, (size_t)h->Mp8 * 8));
    // z = U x = (U C)(C^-1 x): columns only, no row factor
    oz_split_rows_kernel<S><<<(K + 7) / 8, 256>>>(h->d_U.as<double>(), h->ldk, K, K, h->d_pU.as<signed char>(), h->ld8,
                                                   stride, h->d_sU.as<double>(), 1.52587890625e-05 /* 2^-16 */,
                                                   h->d_ck.as<double>(), nullptr, nullptr);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaDeviceSynchronize());
    if (oz_make_plane_map(&h->tmU8, h->d_pU.as<signed char>(), (size_t)S * h->Mp8, h->ld8, OZ_N) != 0 ||
        oz_make_plane_map(&h->tmU8h, h->d_pU.as<signed char>(), (size_t)S * h->Mp8, h->ld8, OZ2_NH) != 0)
      return fail(h, "cuTensorMapEncodeTiled failed for the Cholesky-factor digit planes");
    h->oz_U_S = S;
  }
  if (h->oz_X_S != S) {
    if (h->d_pX.p) { CU_TRY(h, cudaDeviceSynchronize()); cudaFree(h->d_pX.p); h->d_pX.p = nullptr; }
    if (h->d_sX.p) { cudaFree(h->d_sX.p); h->d_sX.p = nullptr; }
    const size_t stride = (size_t)h->cap * h->ld8;
    CU_TRY(h, cudaMalloc(&h->d_pX.p, stride * S));
    CU_TRY(h, cudaMemset(h->d_pX.p, 0, stride * S));
    CU_TRY(h, cudaMalloc(&h->d_sX.p, (size_t)h->cap * 8));
    CU_TRY(h, cudaMemset(h->d_sX.p, 0, (size_t)h->cap * 8));
    if (!h->d_wide.p) {
      CU_TRY(h, cudaMalloc(&h->d_wide.p, (size_t)h->cap * 4));
      CU_TRY(h, cudaMemset(h->d_wide.p, 0, (size_t)h->cap * 4));
      CU_TRY(h, cudaMalloc(&h->d_widecnt.p, (size_t)h->cap * 32));
      CU_TRY(h, cudaMemset(h->d_widecnt.p, 0, (size_t)h->cap * 32));
    }
    if (oz_make_plane_map(&h->tmA8, h->d_pX.as<signed char>(), (size_t)S * h->cap, h->ld8, OZ_M) != 0)
      return fail(h, "cuTensorMapEncodeTiled failed for the residual digit planes");
    h->oz_X_S = S;
  }
  return 0;
}
int ensure_i8(mcd_handle* h) {
  if (h->oz_S == 0 || h->dm.lik != MCD_LIK_FULL || h->sparse || h->N <= SMALL_TREE_MAX_NODES) return 0;
  return h->oz_S == 6 ? ensure_i8_planes<6>(h) : ensure_i8_planes<7>(h);
}
// the contraction launch: CTA pairs (256-chain tiles, cta_group::2) when selected and the chunk is aligned to them, else one CTA per tile
template <int S>
int oz_contract(mcd_handle* h, bool tri, int n, int c0, cudaStream_t st, int kb_lo = 0, int kb_hi = -1, const double* yadd = nullptr,
                int ldyadd = 0) {
  const double* sB = (tri ? h->d_sU : h->d_sP).as<double>();
  if (oz_pair_mode() && !yadd && c0 % (2 * OZ_M) == 0) {   // (the rank-limited MH update keeps the one-CTA kernel: its y buffer is padded to 128 rows)
    const int np2 = (n + 2 * OZ_M - 1) / (2 * OZ_M) * (2 * OZ_M);
    if (c0 + np2 <= h->cap) {
      CU_TRY(h, gemm_i8_ozaki_pair_launch<S>(h->tmA8, tri ? h->tmU8h : h->tmB8h, h->d_sX.as<double>(), sB, h->d_y.as<double>(), h->Mp8, np2,
                                             h->ld8, h->dm.ldy, h->cap, st, c0, h->n_sms, tri ? 1 : 0, kb_lo, kb_hi, yadd, ldyadd));
      return 0;
    }
  }
  const int np = (n + OZ_M - 1) / OZ_M * OZ_M;
  CU_TRY(h, gemm_i8_ozaki_launch<S>(h->tmA8, tri ? h->tmU8 : h->tmB8, h->d_sX.as<double>(), sB, h->d_y.as<double>(), h->Mp8, np, h->ld8,
                                    h->dm.ldy, h->cap, st, c0, h->n_sms, tri ? 1 : 0, kb_lo, kb_hi, yadd, ldyadd));
  return 0;
}

// K1 + contraction on the INT8 tensor pipe for chains [c0, c0 + n)
template <int S>
int enqueue_i8(mcd_handle* h, int c0, int n, const double* xs, cudaStream_t st, cudaEvent_t ev_mid, bool tri) {
  const DevModel& M = h->dm;
  const size_t stride = (size_t)h->cap * h->ld8;
  residual_split_kernel<S><<<n, 256, (size_t)h->ld8 * 8, st>>>(
      M.N, M.K, M.S, M.root_r, M.parent, M.mu, xs, h->d_pX.as<signed char>() + (size_t)c0 * h->ld8, h->ld8, stride,
      h->d_sX.as<double>() + c0, n, h->d_ick.as<double>(), h->d_widecnt.as<int>() + (size_t)c0 * 8);
  if (ev_mid) CU_TRY(h, cudaEventRecord(ev_mid, st));
  if (oz_contract<S>(h, tri, n, c0, st)) return -1;
  // chains K1 flagged (residual range too wide for 56-bit digits relative to the row maximum): their rows of y in plain FP64
  fp64_rows_kernel<<<n, 256, (size_t)M.K * 8, st>>>(M.N, M.K, M.S, M.root_r, M.parent, M.mu, xs,
                                                     (tri ? h->d_U : h->d_P).as<double>(), M.ldk, tri ? 1 : 0,
                                                     h->d_widecnt.as<int>() + (size_t)c0 * 8, h->ld8 - M.K, h->d_wide.as<int>() + c0,
                                                     h->d_y.as<double>() + (size_t)c0 * M.ldy, M.ldy, n);
  return 0;
}

// chains per pipelined chunk of the host-buffer APIs: >= 4 MiB of state per copy, multiple of 128;
// small chunks keep the PCIe fill/drain bubbles short (MCD_CHUNK overrides, for experiments)
int chunk_chains(int S) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MCD_CHUNK");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return (forced + 127) / 128 * 128;
  int chunk = 512;
  if ((size_t)chunk * S * 8 < (size_t)4 << 20) chunk = (int)(((size_t)4 << 20) / ((size_t)S * 8) / 128 + 1) * 128;
  return chunk;
}

// chunk schedule of the host-buffer pipeline: equal chunks (tapering the ends was measured: no gain)
std::vector<std::pair<int, int>> chunk_schedule(int n, int S) {
  const int chunk = chunk_chains(S);
  std::vector<std::pair<int, int>> out;
  for (int c0 = 0; c0 < n; c0 += chunk) out.push_back({c0, std::min(chunk, n - c0)});
  return out;
}

// enqueue the three kernels for chains [c0, c0 + n) of the given device buffers; c0 % 128 == 0
template <bool GRAD>
int enqueue(mcd_handle* h, int c0, int n, const double* d_states, double* d_out, double* d_grad, int32_t* d_status,
            cudaStream_t st, bool posterior_only = false) {
  DevModel M = h->dm;
  const bool tri = !GRAD && M.lik == MCD_LIK_FULL && h->chol_state == 1 && !h->sparse && !h->force_sym &&
                   (h->oz_S == 0 || h->oz_U_S == h->oz_S);
  M.quad_from_z = tri ? 1 : 0;
  const bool small = M.N <= SMALL_TREE_MAX_NODES;
  const int cpb = small ? POST_THREADS / 32 : 1;
  const int grid = (n + cpb - 1) / cpb;
  const double* xs = d_states + (size_t)c0 * M.S;
  double* dx = h->d_dx.as<double>() + (size_t)c0 * M.ldk;
  const double* y = h->d_y.as<double>() + (size_t)c0 * M.ldy;
  if (posterior_only && h->y_override) {  // score the states against another y buffer (the cached y of the resident chains)
    M.ldy = h->y_override_ld;
    y = h->y_override + (size_t)c0 * M.ldy;
  }
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (h->timing) {
    for (int i = 0; i < 4; ++i) {
      CU_TRY(h, cudaEventCreate(&ev[i]));
      h->tev.push_back(ev[i]);
    }
    CU_TRY(h, cudaEventRecord(ev[0], st));
  }
  if (small) {
    // one fused launch: tables and P in shared memory, one warp per chain
    const size_t fsmem = POST_SMEM_FIXED + ((size_t)M.K * M.K + (size_t)(POST_THREADS / 32) * (M.S + M.N + M.K)) * 8;
    const int fgrid = std::min(grid, 2 * h->n_sms);
    double* fo = d_out + (size_t)c0 * MCD_OUT_COLS;
    double* fg = GRAD ? d_grad + (size_t)c0 * M.S : nullptr;
    M.quad_from_z = 0;
#define MCD_LAUNCH_SMALL(CC) \
  small_tree_fused_kernel<CC, GRAD><<<fgrid, POST_THREADS, fsmem, st>>>(M, h->d_P.as<double>(), xs, fo, fg, d_status + c0, n)
    switch (M.clock) {
      case 0: MCD_LAUNCH_SMALL(0); break;
      case 1: MCD_LAUNCH_SMALL(1); break;
      case 2: MCD_LAUNCH_SMALL(2); break;
      default: MCD_LAUNCH_SMALL(3); break;
    }
#undef MCD_LAUNCH_SMALL
    h->launches += 1;
    if (h->timing) {
      CU_TRY(h, cudaEventRecord(ev[1], st));
      CU_TRY(h, cudaEventRecord(ev[2], st));
      CU_TRY(h, cudaEventRecord(ev[3], st));
    }
    CU_TRY(h, cudaGetLastError());
    return 0;
  }
  if (posterior_only) {
    // the caller has already put y = Sigma^-1 dx of these states into d_y (rank-limited update of the MH path)
    if (h->timing) CU_TRY(h, cudaEventRecord(ev[1], st));
  } else if (h->sparse) {
    sparse_contraction_kernel<<<n, POST_THREADS, (size_t)(M.S + M.K) * 8, st>>>(M, xs, h->d_y.as<double>() + (size_t)c0 * M.ldy, n);
    h->launches += 1;
    if (h->timing) CU_TRY(h, cudaEventRecord(ev[1], st));
  } else if (M.lik == MCD_LIK_FULL && h->oz_S != 0) {
    const int rc = h->oz_S == 6 ? enqueue_i8<6>(h, c0, n, xs, st, ev[1], tri) : enqueue_i8<7>(h, c0, n, xs, st, ev[1], tri);
    if (rc) return rc;
    h->launches += 3;
    M.wide = h->d_wide.as<int>() + c0;
  } else if (M.lik == MCD_LIK_FULL) {
    residual_kernel<256><<<grid, POST_THREADS, 0, st>>>(M, xs, dx, n);
    if (h->timing) CU_TRY(h, cudaEventRecord(ev[1], st));
    const int np = (n + GEMM_BT - 1) / GEMM_BT * GEMM_BT;
    CU_TRY(h, gemm_f64_dmma_launch(tri ? h->tmU : h->tmP, h->tmX, h->d_y.as<double>(), h->Mp, np, M.ldk, M.ldy, st, c0,
                                   tri ? 1 : 0));
    h->launches += 2;
  } else if (h->timing) {
    CU_TRY(h, cudaEventRecord(ev[1], st));
  }
  if (h->timing) CU_TRY(h, cudaEventRecord(ev[2], st));
  double* o = d_out + (size_t)c0 * MCD_OUT_COLS;
  double* g = GRAD ? d_grad + (size_t)c0 * M.S : nullptr;
  int32_t* s = d_status + c0;
  // shared memory: reduction scratch + per chain group the staged state row [S] and contraction result [K]
  const size_t smem = POST_SMEM_FIXED + (size_t)cpb * M.S * 8;
  double* pz = h->d_partials.as<double>() + (size_t)c0 * POST_NPART;
#define MCD_LAUNCH_POST(GG, CC, MB) \
  posterior_kernel<GG, CC, GRAD, MB><<<grid, POST_THREADS, smem, st>>>(M, xs, y, o, g, s, n, pz)
#define MCD_LAUNCH_POST_G(GG, MB)                                              \
  switch (M.clock) {                                                           \
    case 0: MCD_LAUNCH_POST(GG, 0, MB); break;                                 \
    case 1: MCD_LAUNCH_POST(GG, 1, MB); break;                                 \
    case 2: MCD_LAUNCH_POST(GG, 2, MB); break;                                 \
    default: MCD_LAUNCH_POST(GG, 3, MB); break;                                \
  }
  MCD_LAUNCH_POST_G(256, POST_MINB)
#undef MCD_LAUNCH_POST_G
#undef MCD_LAUNCH_POST
  // the scalar tail of every chain (product' semantics, status, scalar gradient entries) in parallel, one thread per chain
  posterior_assemble_kernel<GRAD><<<(n + 127) / 128, 128, 0, st>>>(M, xs, pz, o, g, s, n);
  h->launches += 2;
  if (h->timing) CU_TRY(h, cudaEventRecord(ev[3], st));
  CU_TRY(h, cudaGetLastError());
  return 0;
}

// ---- stream ordering.  All entry points share the handle's work buffers (digit planes, y, staging rows).  The pipelined
// host-buffer calls leave work in flight on streams 0..3; every other entry point works on ONE stream (streams[0] or the
// caller's).  begin_serial makes that stream wait for whatever the pipeline and the previous non-pipelined call still have in
// flight; end_serial publishes its own completion point; begin_pipelined makes the pipeline streams wait for that point.
// Without these, mixing e.g. mcd_eval_grad_theta_async with mcd_mh_cycle, or mcd_eval_device on two user streams, raced.
int begin_serial(mcd_handle* h, cudaStream_t st) {
  if (h->pipeline_dirty) {
    for (int i = 0; i < N_STREAMS; ++i) {
      if (h->streams[i] == st) continue;
      if (!h->join_ev[i]) CU_TRY(h, cudaEventCreateWithFlags(&h->join_ev[i], cudaEventDisableTiming));
      CU_TRY(h, cudaEventRecord(h->join_ev[i], h->streams[i]));
      CU_TRY(h, cudaStreamWaitEvent(st, h->join_ev[i], 0));
    }
    h->pipeline_dirty = false;
  }
  if (h->serial_dirty && h->serial_stream != st) CU_TRY(h, cudaStreamWaitEvent(st, h->serial_ev, 0));
  return 0;
}
int end_serial(mcd_handle* h, cudaStream_t st) {
  if (!h->serial_ev) CU_TRY(h, cudaEventCreateWithFlags(&h->serial_ev, cudaEventDisableTiming));
  CU_TRY(h, cudaEventRecord(h->serial_ev, st));
  h->serial_stream = st;
  h->serial_dirty = true;
  return 0;
}
int begin_pipelined(mcd_handle* h) {
  if (h->serial_dirty) {
    for (int i = 0; i < N_STREAMS; ++i)
      if (h->streams[i] != h->serial_stream) CU_TRY(h, cudaStreamWaitEvent(h->streams[i], h->serial_ev, 0));
    h->serial_dirty = false;
    // a later non-pipelined call on another stream is ordered after the pipeline, which is now ordered after this point
  }
  h->pipeline_dirty = true;
  return 0;
}
// scope guard for the single-stream entry points
struct SerialScope {
  mcd_handle* h;
  cudaStream_t st;
  int rc;
  SerialScope(mcd_handle* h_, cudaStream_t st_) : h(h_), st(st_), rc(begin_serial(h_, st_)) {}
  ~SerialScope() { if (rc == 0) end_serial(h, st); }
};

// Device-path pipeline for large dense models on the INT8 contraction: the batch is cut into `pipe_chunks` chunks (multiples
// of 128 chains); chunk j runs K1 on pipe_st[0], the contraction on pipe_st[1] (high priority: its persistent CTAs get the SMs
// first) and K3 on pipe_st[2].  K1 of chunk j + 1 and K3 of chunk j - 1 are resident on the SMs beside the contraction's one CTA
// per SM (its register budget leaves room for them), so the FP64 / HBM work hides behind the tensor-core work.
template <bool GRAD, int S>
int enqueue_pipelined(mcd_handle* h, int n, const double* d_states, double* d_out, double* d_grad, int32_t* d_status,
                      cudaStream_t user) {
  DevModel M = h->dm;
  M.quad_from_z = 0;
  const int nc = h->pipe_chunks;
  const int per = ((n + nc - 1) / nc + OZ_M - 1) / OZ_M * OZ_M;
  if (!h->pipe_st[0]) {
    int lo = 0, hi = 0;
    CU_TRY(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU_TRY(h, cudaStreamCreateWithPriority(&h->pipe_st[0], cudaStreamNonBlocking, lo));
    CU_TRY(h, cudaStreamCreateWithPriority(&h->pipe_st[1], cudaStreamNonBlocking, hi));
    CU_TRY(h, cudaStreamCreateWithPriority(&h->pipe_st[2], cudaStreamNonBlocking, lo));
    // every kernel of the pipeline asks for the maximum shared-memory carve-out: an SM never has to drain to change its
    // L1 / shared split before a CTA of another kernel can join
    if (!getenv("MCD_PIPE_NO_CARVEOUT")) {
      cudaFuncSetAttribute(residual_split_kernel<S>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(posterior_kernel<256, 0, GRAD, POST_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(posterior_kernel<256, 1, GRAD, POST_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(posterior_kernel<256, 2, GRAD, POST_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(posterior_kernel<256, 3, GRAD, POST_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
  }
  const size_t need = 2 + 3 * (size_t)nc;
  while (h->pipe_ev.size() < need) {
    cudaEvent_t e;
    CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->pipe_ev.push_back(e);
  }
  cudaEvent_t e_in = h->pipe_ev[0];
  CU_TRY(h, cudaEventRecord(e_in, user));
  for (int i = 0; i < 3; ++i) CU_TRY(h, cudaStreamWaitEvent(h->pipe_st[i], e_in, 0));
  const size_t stride = (size_t)h->cap * h->ld8;
  const size_t smem = POST_SMEM_FIXED + (size_t)M.S * 8;
  int j = 0;
  for (int c0 = 0; c0 < n; c0 += per, ++j) {
    const int m = std::min(per, n - c0);
    const double* xs = d_states + (size_t)c0 * M.S;
    cudaEvent_t e1 = h->pipe_ev[2 + 3 * j], e2 = h->pipe_ev[3 + 3 * j], e3 = h->pipe_ev[4 + 3 * j];
    residual_split_kernel<S><<<m, 256, (size_t)h->ld8 * 8, h->pipe_st[0]>>>(
        M.N, M.K, M.S, M.root_r, M.parent, M.mu, xs, h->d_pX.as<signed char>() + (size_t)c0 * h->ld8, h->ld8, stride,
        h->d_sX.as<double>() + c0, m, h->d_ick.as<double>(), h->d_widecnt.as<int>() + (size_t)c0 * 8);
    CU_TRY(h, cudaEventRecord(e1, h->pipe_st[0]));
    CU_TRY(h, cudaStreamWaitEvent(h->pipe_st[1], e1, 0));
    const int np = (m + OZ_M - 1) / OZ_M * OZ_M;
    CU_TRY(h, gemm_i8_ozaki_launch<S>(h->tmA8, h->tmB8, h->d_sX.as<double>(), h->d_sP.as<double>(), h->d_y.as<double>(), h->Mp8, np,
                                      h->ld8, M.ldy, h->cap, h->pipe_st[1], c0, h->n_sms, 0));
    fp64_rows_kernel<<<m, 256, (size_t)M.K * 8, h->pipe_st[1]>>>(M.N, M.K, M.S, M.root_r, M.parent, M.mu, xs, h->d_P.as<double>(), M.ldk, 0,
                                                                  h->d_widecnt.as<int>() + (size_t)c0 * 8, h->ld8 - M.K, h->d_wide.as<int>() + c0,
                                                                  h->d_y.as<double>() + (size_t)c0 * M.ldy, M.ldy, m);
    M.wide = h->d_wide.as<int>() + c0;
    CU_TRY(h, cudaEventRecord(e2, h->pipe_st[1]));
    CU_TRY(h, cudaStreamWaitEvent(h->pipe_st[2], e2, 0));
    const double* y = h->d_y.as<double>() + (size_t)c0 * M.ldy;
    double* o = d_out + (size_t)c0 * MCD_OUT_COLS;
    double* g = GRAD ? d_grad + (size_t)c0 * M.S : nullptr;
    int32_t* st_ = d_status + c0;
    // all but the last chunk: at most `pipe_k3_per_sm` CTAs per SM, so that the next chunk's contraction CTA fits beside them
    const bool last = c0 + per >= n;
    const int k3grid = last || h->pipe_k3_per_sm <= 0 ? m : std::min(m, h->pipe_k3_per_sm * h->n_sms);
    double* pz = h->d_partials.as<double>() + (size_t)c0 * POST_NPART;
#define MCD_LAUNCH_POST_P(CC) posterior_kernel<256, CC, GRAD, POST_MINB><<<k3grid, POST_THREADS, smem, h->pipe_st[2]>>>(M, xs, y, o, g, st_, m, pz)
    switch (M.clock) {
      case 0: MCD_LAUNCH_POST_P(0); break;
      case 1: MCD_LAUNCH_POST_P(1); break;
      case 2: MCD_LAUNCH_POST_P(2); break;
      default: MCD_LAUNCH_POST_P(3); break;
    }
#undef MCD_LAUNCH_POST_P
    posterior_assemble_kernel<GRAD><<<(m + 127) / 128, 128, 0, h->pipe_st[2]>>>(M, xs, pz, o, g, st_, m);
    CU_TRY(h, cudaEventRecord(e3, h->pipe_st[2]));
    CU_TRY(h, cudaStreamWaitEvent(user, e3, 0));
    h->launches += 5;
  }
  CU_TRY(h, cudaGetLastError());
  return 0;
}

template <bool GRAD>
int eval_device(mcd_handle* h, int n, const double* d_states, double* d_out, double* d_grad, int32_t* d_status,
                void* stream) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:eval_device");
  if (n <= 0) return 0;
  if (!d_states || !d_out || !d_status || (GRAD && !d_grad)) return fail(h, "null device buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  if (ensure_capacity(h, n, false, GRAD)) return -1;
  if (!GRAD && h->dm.lik == MCD_LIK_FULL && !getenv("MCD_NO_CHOLESKY") && (ensure_cholesky(h, nullptr) || ensure_i8(h))) return -1;
  SerialScope order(h, static_cast<cudaStream_t>(stream));
  if (order.rc) return -1;
  // pipelined form: value + gradient on a large dense model with the INT8 contraction, not while per-kernel timing is on
  if (GRAD && h->pipe_chunks > 1 && !h->timing && h->dm.lik == MCD_LIK_FULL && !h->sparse && h->oz_S != 0 &&
      h->N > SMALL_TREE_MAX_NODES && n >= 2 * OZ_M * h->pipe_chunks)
    return h->oz_S == 6 ? enqueue_pipelined<GRAD, 6>(h, n, d_states, d_out, d_grad, d_status, static_cast<cudaStream_t>(stream))
                        : enqueue_pipelined<GRAD, 7>(h, n, d_states, d_out, d_grad, d_status, static_cast<cudaStream_t>(stream));
  return enqueue<GRAD>(h, 0, n, d_states, d_out, d_grad, d_status, static_cast<cudaStream_t>(stream));
}

// completion events of an asynchronous host-buffer call on all pipeline streams -> ticket
int record_ticket(mcd_handle* h, int64_t* ticket_out) {
  CU_TRY(h, cudaSetDevice(h->device));
  const int64_t ticket = h->next_ticket++;
  for (int i = 0; i < N_STREAMS; ++i) {
    cudaEvent_t& ev = h->ticket_ev[ticket % 8][i];
    if (!ev) CU_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU_TRY(h, cudaEventRecord(ev, h->streams[i]));
  }
  if (ticket_out) *ticket_out = ticket;
  return 0;
}

// host buffers: chunks of chains are copied in, evaluated and copied out on rotating streams so
// that PCIe transfers in both directions overlap the kernels of neighbouring chunks
template <bool GRAD>
int eval_host(mcd_handle* h, int n, const double* states, double* out, double* grad, int32_t* status, bool async = false,
              int64_t* ticket_out = nullptr) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:eval_host");
  if (n <= 0) return async ? record_ticket(h, ticket_out) : 0;
  if (!states || !out || !status || (GRAD && !grad)) return fail(h, "null host buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  if (n > h->cap || !h->d_states.p || (GRAD && !h->d_grad.p)) CU_TRY(h, cudaDeviceSynchronize());  // (re)allocation ahead
  if (ensure_capacity(h, n, true, GRAD)) return -1;
  if (!GRAD && h->dm.lik == MCD_LIK_FULL && !getenv("MCD_NO_CHOLESKY") && (ensure_cholesky(h, nullptr) || ensure_i8(h))) return -1;
  const int S = h->S;
  int ci = 0;
  if (begin_pipelined(h)) return -1;
  for (const auto& cm : chunk_schedule(n, S)) {
    const int c0 = cm.first, m = cm.second;
    cudaStream_t st = h->streams[ci++ % N_STREAMS];
    double* d_x = h->d_states.as<double>();
    CU_TRY(h, cudaMemcpyAsync(d_x + (size_t)c0 * S, states + (size_t)c0 * S, (size_t)m * S * 8, cudaMemcpyHostToDevice, st));
    if (enqueue<GRAD>(h, c0, m, d_x, h->d_out.as<double>(), h->d_grad.as<double>(), h->d_status.as<int32_t>(), st)) return -1;
    CU_TRY(h, cudaMemcpyAsync(out + (size_t)c0 * MCD_OUT_COLS, h->d_out.as<double>() + (size_t)c0 * MCD_OUT_COLS,
                              (size_t)m * MCD_OUT_COLS * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(status + c0, h->d_status.as<int32_t>() + c0, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (GRAD)
      CU_TRY(h, cudaMemcpyAsync(grad + (size_t)c0 * S, h->d_grad.as<double>() + (size_t)c0 * S, (size_t)m * S * 8,
                                cudaMemcpyDeviceToHost, st));
  }
  if (async) return record_ticket(h, ticket_out);
  for (int i = 0; i < N_STREAMS; ++i) CU_TRY(h, cudaStreamSynchronize(h->streams[i]));
  h->pipeline_dirty = false;
  return 0;
}

// theta-packed host API: only the D free parameters cross PCIe in either direction
// async = true: returns after enqueueing; the outputs are valid after mcd_synchronize.  Consecutive calls then overlap
// their PCIe fill and drain (chunk k of every call runs on stream k % 4, in order, on its own slice of the staging buffers).
int eval_theta_host(mcd_handle* h, int n, const double* theta, const double* base, double* out, double* gtheta,
                    int32_t* status, bool async = false, int64_t* ticket_out = nullptr) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:eval_theta_host");
  if (n <= 0) return async ? record_ticket(h, ticket_out) : 0;
  if (!theta || !base || !out || !gtheta || !status) return fail(h, "null host buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  if (n > h->cap || !h->d_states.p || !h->d_grad.p) CU_TRY(h, cudaDeviceSynchronize());  // buffers are about to be (re)allocated
  if (ensure_capacity(h, n, true, true)) return -1;
  const int S = h->S, D = h->D;
  if (!h->d_theta.p) CU_TRY(h, cudaMalloc(&h->d_theta.p, (size_t)h->cap * D * 8));
  if (!h->d_gtheta.p) CU_TRY(h, cudaMalloc(&h->d_gtheta.p, (size_t)h->cap * D * 8));
  // the shared base state goes to the device only when it changes (work of an earlier asynchronous call may still read it)
  if (h->base_host.size() != (size_t)S || memcmp(h->base_host.data(), base, (size_t)S * 8) != 0) {
    CU_TRY(h, cudaDeviceSynchronize());
    CU_TRY(h, cudaMemcpy(h->d_base.p, base, (size_t)S * 8, cudaMemcpyHostToDevice));
    h->base_host.assign(base, base + S);
  }
  int ci = 0;
  if (begin_pipelined(h)) return -1;
  for (const auto& cm : chunk_schedule(n, S)) {
    const int c0 = cm.first, m = cm.second;
    cudaStream_t st = h->streams[ci++ % N_STREAMS];
    double* d_th = h->d_theta.as<double>() + (size_t)c0 * D;
    double* d_gt = h->d_gtheta.as<double>() + (size_t)c0 * D;
    double* d_x = h->d_states.as<double>();
    CU_TRY(h, cudaMemcpyAsync(d_th, theta + (size_t)c0 * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, st));
    unpack_theta_kernel<<<dim3((S + POST_THREADS - 1) / POST_THREADS, m), POST_THREADS, 0, st>>>(
        d_th, h->d_base.as<double>(), h->d_tidx.as<int>(), d_x + (size_t)c0 * S, S, D, m);
    if (enqueue<true>(h, c0, m, d_x, h->d_out.as<double>(), h->d_grad.as<double>(), h->d_status.as<int32_t>(), st)) return -1;
    pack_theta_kernel<<<dim3((D + POST_THREADS - 1) / POST_THREADS, m), POST_THREADS, 0, st>>>(
        h->d_grad.as<double>() + (size_t)c0 * S, h->d_sidx.as<int>(), d_gt, S, D, m);
    h->launches += 2;
    CU_TRY(h, cudaMemcpyAsync(out + (size_t)c0 * MCD_OUT_COLS, h->d_out.as<double>() + (size_t)c0 * MCD_OUT_COLS,
                              (size_t)m * MCD_OUT_COLS * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(status + c0, h->d_status.as<int32_t>() + c0, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(gtheta + (size_t)c0 * D, d_gt, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
  }
  if (!async) {
    for (int i = 0; i < N_STREAMS; ++i) CU_TRY(h, cudaStreamSynchronize(h->streams[i]));
    h->pipeline_dirty = false;
    return 0;
  }
  return record_ticket(h, ticket_out);
}
int wait_ticket(mcd_handle* h, int64_t ticket) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:wait_ticket");
  if (ticket < 0 || ticket >= h->next_ticket) return fail(h, "mcd_wait: unknown ticket");
  CU_TRY(h, cudaSetDevice(h->device));
  // a ticket older than the ring of 8: its slot now holds the events of a LATER call, recorded on the same in-order streams, so
  // their completion implies the old call's
  for (int i = 0; i < N_STREAMS; ++i)
    if (h->ticket_ev[ticket % 8][i]) CU_TRY(h, cudaEventSynchronize(h->ticket_ev[ticket % 8][i]));
  return 0;
}

// L leapfrog steps for n chains, everything resident on the device between the two ends
int leapfrog_host(mcd_handle* h, int n, int L, const double* theta0, const double* mom0, const double* base,
                  const double* inv_mass, const double* eps, double* theta_out, double* mom_out, double* out,
                  double* energy, int32_t* status) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:leapfrog_host");
  if (n <= 0) return 0;
  if (L < 1) return fail(h, "mcd_leapfrog: n_steps must be >= 1");
  if (!theta0 || !mom0 || !base || !inv_mass || !eps || !theta_out || !mom_out || !out || !energy || !status)
    return fail(h, "null host buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (ensure_capacity(h, n, true, true)) return -1;
  const int S = h->S, D = h->D;
  const size_t nd = (size_t)h->cap * D * 8;
  if (!h->d_theta.p) CU_TRY(h, cudaMalloc(&h->d_theta.p, nd));
  if (!h->d_gtheta.p) CU_TRY(h, cudaMalloc(&h->d_gtheta.p, nd));
  if (!h->d_mom.p) CU_TRY(h, cudaMalloc(&h->d_mom.p, nd));
  if (!h->d_eps.p) CU_TRY(h, cudaMalloc(&h->d_eps.p, (size_t)h->cap * 8));
  if (!h->d_energy.p) CU_TRY(h, cudaMalloc(&h->d_energy.p, (size_t)h->cap * 16));
  if (!h->d_status_acc.p) CU_TRY(h, cudaMalloc(&h->d_status_acc.p, (size_t)h->cap * 4));
  if (!h->d_invmass.p) CU_TRY(h, cudaMalloc(&h->d_invmass.p, (size_t)std::max(D, 1) * 8));
  cudaStream_t st = h->streams[0];
  double* th = h->d_theta.as<double>();
  double* gt = h->d_gtheta.as<double>();
  double* pm = h->d_mom.as<double>();
  double* xs = h->d_states.as<double>();
  double* o = h->d_out.as<double>();
  double* en = h->d_energy.as<double>();
  int32_t* stp = h->d_status.as<int32_t>();
  int32_t* sacc = h->d_status_acc.as<int32_t>();
  const size_t nb = (size_t)n * D * 8;
  CU_TRY(h, cudaMemcpyAsync(th, theta0, nb, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(pm, mom0, nb, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(h->d_eps.p, eps, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(h->d_invmass.p, inv_mass, (size_t)D * 8, cudaMemcpyHostToDevice, st));
  h->base_host.clear();
  CU_TRY(h, cudaMemcpyAsync(h->d_base.p, base, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemsetAsync(sacc, 0, (size_t)n * 4, st));
  const dim3 gS((S + POST_THREADS - 1) / POST_THREADS, n), gD((D + HMC_THREADS - 1) / HMC_THREADS, n);
  auto gradient = [&]() -> int {  // theta -> states -> (ln post, grad) -> packed gradient
    unpack_theta_kernel<<<gS, POST_THREADS, 0, st>>>(th, h->d_base.as<double>(), h->d_tidx.as<int>(), xs, S, D, n);
    if (enqueue<true>(h, 0, n, xs, o, h->d_grad.as<double>(), stp, st)) return -1;
    pack_theta_kernel<<<gD, POST_THREADS, 0, st>>>(h->d_grad.as<double>(), h->d_sidx.as<int>(), gt, S, D, n);
    h->launches += 2;
    return 0;
  };
  const double* im = h->d_invmass.as<double>();
  const double* ep = h->d_eps.as<double>();
  const int gE = (n + HMC_THREADS / 32 - 1) / (HMC_THREADS / 32);
  if (gradient()) return -1;
  hamiltonian_kernel<<<gE, HMC_THREADS, 0, st>>>(pm, im, o, en, 0, D, n);
  leapfrog_update_kernel<<<gD, HMC_THREADS, 0, st>>>(th, pm, gt, im, ep, 0.5, 1, stp, sacc, D, n);
  h->launches += 2;
  for (int l = 1; l <= L; ++l) {
    if (gradient()) return -1;
    if (l < L) leapfrog_update_kernel<<<gD, HMC_THREADS, 0, st>>>(th, pm, gt, im, ep, 1.0, 1, stp, sacc, D, n);
    else leapfrog_update_kernel<<<gD, HMC_THREADS, 0, st>>>(th, pm, gt, im, ep, 0.5, 0, stp, sacc, D, n);
    h->launches += 1;
  }
  hamiltonian_kernel<<<gE, HMC_THREADS, 0, st>>>(pm, im, o, en, 1, D, n);
  h->launches += 1;
  CU_TRY(h, cudaGetLastError());
  CU_TRY(h, cudaMemcpyAsync(theta_out, th, nb, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(mom_out, pm, nb, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out, o, (size_t)n * MCD_OUT_COLS * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(energy, en, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(status, sacc, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return 0;
}

// One NUTS transition for n chains (hmc_kernels.cuh): all chains tick in lockstep, one leapfrog step per tick, until
// every chain has made its U-turn (or reached max_depth / diverged); the host only reads the number of still
// active chains after each tick.
int mh_refresh(mcd_handle* h);
bool mh_incremental_capable(const mcd_handle* h);
// resident = true: the transition runs on the chains uploaded with mcd_chains_set (positions packed from / scattered back to
// their state rows on the device; theta0, base, theta_out, out are not used)
int nuts_host(mcd_handle* h, int n, const double* theta0, const double* base, const double* inv_mass, const double* eps,
              const double* mom0, int max_depth, uint64_t seed, uint32_t iteration, double* theta_out, double* out,
              double* accept_stat, int32_t* info, int32_t* status, bool resident = false) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:nuts_host");
  if (resident) n = h->n_resident;
  if (n <= 0) return resident ? fail(h, "mcd_chains_nuts: no resident chains (call mcd_chains_set first)") : 0;
  if (max_depth < 1 || max_depth > 16) return fail(h, "mcd_nuts: max_depth must be in 1..16");
  if (resident && h->mc3_C > 0) return fail(h, "mcd_chains_nuts: heated chains are not supported (cold chains only)");
  if (!inv_mass || !eps || !accept_stat || !info || !status || (!resident && (!theta0 || !base || !theta_out || !out)))
    return fail(h, "null host buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (ensure_capacity(h, n, true, true)) return -1;
  const int S = h->S, D = h->D;
  const size_t BD = (size_t)n * D;
  const size_t n_vec = 8 + 2 * (size_t)max_depth + (mom0 ? 1 : 0) + 1;   // + theta0 staging
  const size_t bytes = n_vec * BD * 8 + (size_t)n * (2 * 8 + NR_COLS) * 8 + (size_t)n * NI_COLS * 4 + 256 + (size_t)D * 8 + (size_t)n * 8;
  if (bytes > h->nuts_bytes) {
    CU_TRY(h, cudaDeviceSynchronize());
    if (h->d_nuts.p) cudaFree(h->d_nuts.p);
    h->d_nuts.p = nullptr;
    h->nuts_bytes = 0;
    CU_TRY(h, cudaMalloc(&h->d_nuts.p, bytes));
    h->nuts_bytes = bytes;
  }
  double* p = h->d_nuts.as<double>();
  auto take = [&](size_t count) { double* r = p; p += count; return r; };
  NutsBuffers nb;
  for (int e = 0; e < 2; ++e) { nb.thE[e] = take(BD); nb.rE[e] = take(BD); nb.gE[e] = take(BD); }
  nb.thM = take(BD); nb.thC = take(BD);
  nb.ck_th = take(BD * max_depth); nb.ck_r = take(BD * max_depth);
  double* d_theta0 = take(BD);
  double* d_mom0 = mom0 ? take(BD) : nullptr;
  nb.outM = take((size_t)n * 8); nb.outC = take((size_t)n * 8);
  nb.nr = take((size_t)n * NR_COLS);
  double* d_invm = take(D);
  double* d_eps = take(n);
  nb.n_active = reinterpret_cast<int*>(take(32));
  nb.ni = reinterpret_cast<int*>(p);
  cudaStream_t st = h->streams[0];
  double* xs = h->d_states.as<double>();
  double* o = h->d_out.as<double>();
  double* gr = h->d_grad.as<double>();
  int32_t* stp = h->d_status.as<int32_t>();
  const dim3 gD((D + POST_THREADS - 1) / POST_THREADS, n);
  if (resident) {  // positions = toVector of the resident states; the fixed entries come from the first chain's row
    pack_theta_kernel<<<gD, POST_THREADS, 0, st>>>(h->d_chain.as<double>(), h->d_sidx.as<int>(), d_theta0, S, D, n);
    h->launches += 1;
  } else {
    CU_TRY(h, cudaMemcpyAsync(d_theta0, theta0, BD * 8, cudaMemcpyHostToDevice, st));
  }
  if (mom0) CU_TRY(h, cudaMemcpyAsync(d_mom0, mom0, BD * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(d_invm, inv_mass, (size_t)D * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(d_eps, eps, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  h->base_host.clear();  // the theta-packed host calls re-upload their base state next time
  if (resident) CU_TRY(h, cudaMemcpyAsync(h->d_base.p, h->d_chain.p, (size_t)S * 8, cudaMemcpyDeviceToDevice, st));
  else CU_TRY(h, cudaMemcpyAsync(h->d_base.p, base, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemsetAsync(nb.n_active, 0, 4, st));
  const dim3 gS((S + POST_THREADS - 1) / POST_THREADS, n);
  unpack_theta_kernel<<<gS, POST_THREADS, 0, st>>>(d_theta0, h->d_base.as<double>(), h->d_tidx.as<int>(), xs, S, D, n);
  if (enqueue<true>(h, 0, n, xs, o, gr, stp, st)) return -1;
  nuts_init_kernel<<<n, HMC_THREADS, 0, st>>>(nb, d_theta0, d_mom0, gr, h->d_sidx.as<int>(), d_invm, d_eps, o, stp, xs, seed,
                                              iteration, S, D, n);
  h->launches += 2;
  CU_TRY(h, cudaGetLastError());
  // The number of still-active chains is read back with one tick of lag (pinned flags + events): tick k + 1 is already
  // queued when the host looks at tick k's count, so the GPU never waits for the host; at most one surplus tick runs at
  // the end (every kernel of it is a no-op for finished chains).
  if (!h->nuts_flags) {
    CU_TRY(h, cudaMallocHost(&h->nuts_flags, 2 * sizeof(int)));
    for (int i = 0; i < 2; ++i) CU_TRY(h, cudaEventCreateWithFlags(&h->nuts_ev[i], cudaEventDisableTiming));
  }
  volatile int* flags = h->nuts_flags;
  CU_TRY(h, cudaMemcpyAsync(h->nuts_flags, nb.n_active, 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  const long max_ticks = (1L << max_depth);
  bool pending = false;
  // The still-active chains are compacted into the first rows of the evaluation batch (each takes a row with an
  // atomic counter when it decides to go on), so a tick evaluates only as many states as were active one tick
  // earlier (the count the host knows; the real count can only be smaller, the surplus rows hold stale but valid
  // states).
  int n_eval = flags[0];
  if (n_eval > 0) {
    for (long tick = 0; tick < max_ticks; ++tick) {
      if (enqueue<true>(h, 0, n_eval, xs, o, gr, stp, st)) return -1;
      CU_TRY(h, cudaMemsetAsync(nb.n_active, 0, 4, st));
#define MCD_NUTS_LEAF(NE) \
  nuts_leaf_kernel<NE><<<n, HMC_THREADS, 0, st>>>(nb, gr, h->d_sidx.as<int>(), d_invm, o, stp, xs, seed, iteration, max_depth, S, D, n)
      if (D <= 1 * HMC_THREADS) MCD_NUTS_LEAF(1);
      else if (D <= 2 * HMC_THREADS) MCD_NUTS_LEAF(2);
      else if (D <= 4 * HMC_THREADS) MCD_NUTS_LEAF(4);
      else if (D <= 8 * HMC_THREADS) MCD_NUTS_LEAF(8);
      else if (D <= 12 * HMC_THREADS) MCD_NUTS_LEAF(12);
      else if (D <= 16 * HMC_THREADS) MCD_NUTS_LEAF(16);
      else if (D <= 24 * HMC_THREADS) MCD_NUTS_LEAF(24);
      else MCD_NUTS_LEAF(56);   /* D <= 14336 (trees of up to ~4700 leaves; larger ones exceed K3's staging anyway) */
#undef MCD_NUTS_LEAF
      h->launches += 1;
      CU_TRY(h, cudaGetLastError());
      CU_TRY(h, cudaMemcpyAsync(h->nuts_flags + (tick & 1), nb.n_active, 4, cudaMemcpyDeviceToHost, st));
      CU_TRY(h, cudaEventRecord(h->nuts_ev[tick & 1], st));
      if (pending) {  // look at the previous tick's count while this one runs
        CU_TRY(h, cudaEventSynchronize(h->nuts_ev[(tick - 1) & 1]));
        n_eval = flags[(tick - 1) & 1];
        if (n_eval == 0) break;
      }
      pending = true;
    }
  }
  CU_TRY(h, cudaStreamSynchronize(st));
  // results
  std::vector<int32_t> ni((size_t)n * NI_COLS);
  std::vector<double> nr((size_t)n * NR_COLS);
  if (resident) {  // the chosen points become the chains' states; their ln-posterior parts (and cached y) are re-evaluated
    scatter_theta_kernel<<<gD, POST_THREADS, 0, st>>>(nb.thM, h->d_sidx.as<int>(), h->d_chain.as<double>(), S, D, n);
    h->launches += 1;
    if (mh_incremental_capable(h) && h->inc_enabled && h->inc_ok) {
      if (mh_refresh(h)) return -1;
    } else if (enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_chain_out.as<double>(), nullptr,
                              h->d_chain_status.as<int32_t>(), st)) {
      return -1;
    }
  } else {
    CU_TRY(h, cudaMemcpyAsync(theta_out, nb.thM, BD * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaMemcpyAsync(out, nb.outM, (size_t)n * 8 * 8, cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(h, cudaMemcpyAsync(ni.data(), nb.ni, ni.size() * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(nr.data(), nb.nr, nr.size() * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  for (int b = 0; b < n; ++b) {
    const int32_t* r = &ni[(size_t)b * NI_COLS];
    info[4 * b + 0] = r[NI_DEPTH];
    info[4 * b + 1] = r[NI_NLEAP];
    info[4 * b + 2] = r[NI_DIVERGED];
    info[4 * b + 3] = r[NI_N];
    status[b] = r[NI_STATUS];
    accept_stat[b] = r[NI_NALPHA] > 0 ? nr[(size_t)b * NR_COLS + NR_ALPHA] / r[NI_NALPHA] : 0.0;
  }
  return 0;
}

// ---- chains resident in HBM + Metropolis-Hastings steps, heated chains, MC3 swaps (mh_kernels.cuh)
int mh_prepare_value_path(mcd_handle* h, int n) {
  if (ensure_capacity(h, n, false, false)) return -1;
  if (h->dm.lik == MCD_LIK_FULL && !getenv("MCD_NO_CHOLESKY") && (ensure_cholesky(h, nullptr) || ensure_i8(h))) return -1;
  return 0;
}
// incremental evaluation applies to the large-tree dense-precision pipeline (small trees are one fused launch anyway)
bool mh_incremental_capable(const mcd_handle* h) {
  return h->dm.lik == MCD_LIK_FULL && !h->sparse && h->N > SMALL_TREE_MAX_NODES;
}
// evaluate the resident chains from scratch (symmetric contraction) and cache y = Sigma^-1 dx of every chain
int mh_refresh(mcd_handle* h) {
  const int n = h->n_resident;
  cudaStream_t st = h->streams[0];
  h->force_sym = true;
  const int rc = enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_chain_out.as<double>(), nullptr, h->d_chain_status.as<int32_t>(), st);
  h->force_sym = false;
  if (rc) return -1;
  CU_TRY(h, cudaMemcpy2DAsync(h->d_chain_y.p, (size_t)h->ldyc * 8, h->d_y.p, (size_t)h->ldy * 8,
                              (size_t)std::min(h->ldy, h->ldyc) * 8, n, cudaMemcpyDeviceToDevice, st));
  h->inc_steps = 0;
  return 0;
}
bool mh_kind_incremental(const mcd_handle* h, int kind, int node) {
  switch (kind) {
    case MH_SLIDE_NODE: case MH_SLIDE_NODE_CONTRA: case MH_SCALE_BRANCH: case MH_SLIDE_BRACE: case MH_SLIDE_BRACE_CONTRA:
      return true;
    case MH_SCALE_SUBTREE: case MH_SCALE_SUBTREE_CONTRA:
      return node > 0 && h->sub_size_h[node] <= DL_SUBTREE_H;
    case MH_SCALE_RATE_SUBTREE:
      return node > 0 && h->sub_size_h[node] <= DL_SUBTREE_R;
    default: return false;  // moves that touch lambda, mu, H, m, v or most of the tree: evaluated from scratch
  }
}
int chains_set(mcd_handle* h, int n, const double* states) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:chains_set");
  if (n <= 0 || !states) return fail(h, "mcd_chains_set: need n > 0 and a state buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (mh_prepare_value_path(h, n)) return -1;
  const int S = h->S, N = h->N;
  h->undo_stride = S + MH_MAX_OPS;
  if (n > h->chain_cap) {
    CU_TRY(h, cudaDeviceSynchronize());
    for (DevBuf* b : {&h->d_chain, &h->d_chain_out, &h->d_chain_status, &h->d_new_out, &h->d_new_status, &h->d_undo, &h->d_meta,
                      &h->d_lq, &h->d_accepted, &h->d_rng, &h->d_chain_y, &h->d_dl_n, &h->d_dl_k, &h->d_dl_d}) {
      if (b->p) cudaFree(b->p);
      b->p = nullptr;
    }
    const int cap = (n + 127) / 128 * 128;
    CU_TRY(h, cudaMalloc(&h->d_chain.p, (size_t)cap * S * 8));
    CU_TRY(h, cudaMalloc(&h->d_chain_out.p, (size_t)cap * 8 * 8));
    CU_TRY(h, cudaMalloc(&h->d_new_out.p, (size_t)cap * 8 * 8));
    CU_TRY(h, cudaMalloc(&h->d_chain_status.p, (size_t)cap * 4));
    CU_TRY(h, cudaMalloc(&h->d_new_status.p, (size_t)cap * 4));
    CU_TRY(h, cudaMalloc(&h->d_undo.p, (size_t)cap * h->undo_stride * 8));
    CU_TRY(h, cudaMalloc(&h->d_rng.p, (size_t)cap * MH_MAX_OPS * sizeof(int2)));
    CU_TRY(h, cudaMalloc(&h->d_meta.p, (size_t)cap * sizeof(int4)));
    CU_TRY(h, cudaMalloc(&h->d_lq.p, (size_t)cap * 8));
    CU_TRY(h, cudaMalloc(&h->d_accepted.p, (size_t)cap * 4));
    if (mh_incremental_capable(h)) {
      h->ldyc = h->Mp8 + 16;  // >= the contraction's padded row count (its epilogue adds whole 64-column tiles of this buffer)
      CU_TRY(h, cudaMalloc(&h->d_chain_y.p, (size_t)cap * h->ldyc * 8));
      CU_TRY(h, cudaMemset(h->d_chain_y.p, 0, (size_t)cap * h->ldyc * 8));
      CU_TRY(h, cudaMalloc(&h->d_dl_n.p, (size_t)cap * 4));
      CU_TRY(h, cudaMalloc(&h->d_dl_k.p, (size_t)cap * DL_MAX_AB * 4));
      CU_TRY(h, cudaMalloc(&h->d_dl_d.p, (size_t)cap * DL_MAX_AB * 8));
    }
    h->chain_cap = cap;
  }
  if (!h->d_mh_child1.p) {  // topology tables of the proposals: second children, sub-tree sizes, inner-node counts
    std::vector<int> size(N, 1), inner(N, 0), list;
    std::vector<int> child0(N, -1);
    for (int i = N - 1; i >= 1; --i) {
      size[h->parent[i]] += size[i];
      child0[h->parent[i]] = i;  // descending i: the last one written is the first child
    }
    for (int i = N - 1; i >= 0; --i) {
      if (child0[i] >= 0) inner[i] += 1;
      if (i > 0) inner[h->parent[i]] += inner[i];
    }
    for (int i = 1; i < N; ++i)
      if (child0[i] >= 0) list.push_back(i);
    h->n_inner_nonroot = (int)list.size();
    h->sub_size_h.assign(size.begin(), size.end());
    if (upload(h, h->d_mh_child1, h->child1.data(), N) || upload(h, h->d_mh_size, size.data(), N) ||
        upload(h, h->d_mh_inner_cnt, inner.data(), N) || upload(h, h->d_mh_inner_list, list.data(), list.size()))
      return -1;
    CU_TRY(h, cudaMalloc(&h->d_counters.p, (size_t)2 * MH_MAX_CYCLE * sizeof(unsigned long long)));
  }
  cudaStream_t st = h->streams[0];
  CU_TRY(h, cudaMemcpyAsync(h->d_chain.p, states, (size_t)n * S * 8, cudaMemcpyHostToDevice, st));
  h->n_resident = n;
  h->inc_ok = false;
  if (mh_incremental_capable(h) && h->inc_enabled) {
    // symmetric contraction: y of every chain is kept; incremental moves need every chain's current state to be valid
    if (mh_refresh(h)) return -1;
    std::vector<int32_t> stv(n);
    CU_TRY(h, cudaMemcpyAsync(stv.data(), h->d_chain_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    bool ok = true;
    for (int b = 0; b < n; ++b) ok = ok && ((stv[b] & ~MCD_ST_NEARCRIT) == 0);
    h->inc_ok = ok;
    return 0;
  }
  if (enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_chain_out.as<double>(), nullptr, h->d_chain_status.as<int32_t>(), st))
    return -1;
  CU_TRY(h, cudaStreamSynchronize(st));
  return 0;
}
int chains_get(mcd_handle* h, int n, double* states, double* out, int32_t* status) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:chains_get");
  if (n <= 0 || n > h->n_resident) return fail(h, "mcd_chains_get: more chains requested than are resident");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  cudaStream_t st = h->streams[0];
  if (states) CU_TRY(h, cudaMemcpyAsync(states, h->d_chain.p, (size_t)n * h->S * 8, cudaMemcpyDeviceToHost, st));
  if (out) CU_TRY(h, cudaMemcpyAsync(out, h->d_chain_out.p, (size_t)n * 8 * 8, cudaMemcpyDeviceToHost, st));
  if (status) CU_TRY(h, cudaMemcpyAsync(status, h->d_chain_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return 0;
}
// argument checks of one proposal: what the reference's proposal constructors reject with `error`
int mh_check(mcd_handle* h, int kind, int node, double param, double tune) {
  const int N = h->N;
  if (kind < 0 || kind >= MH_N_KINDS) return fail(h, "mcd_mh: unknown proposal kind");
  if (!(param > 0.0) || !(tune > 0.0)) return fail(h, "mcd_mh: the standard deviation / shape and the tuning parameter must be positive");
  auto inner_nonroot = [&](int i) { return i > 0 && i < N && h->child1[i] >= 0; };
  switch (kind) {
    case MH_SLIDE_NODE: case MH_SCALE_SUBTREE: case MH_SCALE_RATE_SUBTREE: case MH_SLIDE_NODE_CONTRA: case MH_SCALE_SUBTREE_CONTRA:
      if (h->n_inner_nonroot == 0) return fail(h, "mcd_mh: the tree has no inner node below the root");
      if (node >= 0 && !inner_nonroot(node))
        return fail(h, "mcd_mh: the node must be an inner node below the root (slideNodeAtUltrametric: path leads to a leaf)");
      break;
    case MH_SCALE_BRANCH:
      if (node == 0 || node >= N) return fail(h, "mcd_mh: scaleBranch needs a node below the root");
      break;
    case MH_PULLEY:
      if (h->child1[1] < 0 || h->child1[h->child1[0]] < 0) return fail(h, "pulleyUltrametric: a sub tree of the root is a leaf");
      break;
    case MH_SLIDE_BRACE: case MH_SLIDE_BRACE_CONTRA: {
      const int nb = (int)h->br_off_h.size() - 1;
      if (nb <= 0) return fail(h, "mcd_mh: the model has no braces");
      if (node >= nb) return fail(h, "mcd_mh: brace index out of range");
      for (int b = (node < 0 ? 0 : node); b < (node < 0 ? nb : node + 1); ++b) {
        if (h->br_off_h[b + 1] - h->br_off_h[b] > MH_MAX_BRACE_NODES) return fail(h, "mcd_mh: braces of more than 16 nodes are not supported");
        for (int o = h->br_off_h[b]; o < h->br_off_h[b + 1]; ++o)
          if (!inner_nonroot(h->br_node_h[o])) return fail(h, "slideBracedNodesUltrametric: braced root node or leaf");
      }
    } break;
    case MH_SCALE_SCALAR:
      if (node < 0 || node > 4) return fail(h, "mcd_mh: scalar index must be 0 (lambda), 1 (mu), 2 (H), 3 (m) or 4 (v)");
      break;
    case MH_SCALE_RATES_TREE_CONTRA:
      if (h->n_inner_nonroot < 1) return fail(h, "scaleRatesAndTreeContrarilyPFunction: no internal nodes to scale");
      break;
    default: break;
  }
  return 0;
}
MhTopo mh_topo(mcd_handle* h) {
  MhTopo T;
  T.ldyc = h->ldyc; T.N = h->N; T.S = h->S; T.n_inner_nonroot = h->n_inner_nonroot; T.root_r = h->dm.root_r; T.n_brace = h->dm.n_brace;
  T.parent = h->dm.parent; T.child1 = h->d_mh_child1.as<int>(); T.sub_size = h->d_mh_size.as<int>();
  T.sub_inner = h->d_mh_inner_cnt.as<int>(); T.inner_list = h->d_mh_inner_list.as<int>();
  T.br_off = h->dm.br_off; T.br_node = h->dm.br_node;
  return T;
}
// enqueue propose -> evaluation of the proposed states -> accept on stream 0 (no synchronisation).  Small moves are
// evaluated incrementally from the cached y (mh_delta_kernel); everything else by the batched value-only evaluation.
int mh_enqueue(mcd_handle* h, int kind, int node, double param, double tune, int use_root_jacobian, uint64_t seed,
               uint32_t iteration, unsigned long long* d_counters) {
  const int n = h->n_resident, S = h->S;
  cudaStream_t st = h->streams[0];
  const bool inc_mode = h->inc_enabled && h->inc_ok;
  const bool inc = inc_mode && mh_kind_incremental(h, kind, node);
  if (inc && ++h->inc_steps > h->refresh_every && mh_refresh(h)) return -1;  // bound the accumulated rounding
  MhParams P;
  P.kind = kind; P.node = node; P.use_root_jacobian = use_root_jacobian; P.pad = 0; P.param = param; P.tune = tune;
  P.seed = seed; P.iteration = iteration; P.chain_offset = h->mc3_offset;
  const MhTopo T = mh_topo(h);
  const bool heated = h->mc3_C > 0;
  static const bool unfused = getenv("MCD_MH_UNFUSED") != nullptr;  // A/B switch: three launches instead of one
  if (h->N <= SMALL_TREE_MAX_NODES && !unfused) {  // small trees: the whole step in one launch
    DevModel M = h->dm;
    M.quad_from_z = 0;
    const size_t fsmem = POST_SMEM_FIXED + ((size_t)M.K * M.K + (size_t)(POST_THREADS / 32) * (M.S + M.N + M.K)) * 8 +
                         (size_t)(POST_THREADS / 32) * (sizeof(MhOp) * MH_MAX_OPS + (size_t)(2 * M.N + 8) * 16);
    const int fgrid = std::min((n + POST_THREADS / 32 - 1) / (POST_THREADS / 32), 2 * h->n_sms);
#define MCD_LAUNCH_MH_SMALL(CC)                                                                                              \
  mh_small_tree_kernel<CC><<<fgrid, POST_THREADS, fsmem, st>>>(M, T, P, h->d_P.as<double>(), h->d_chain.as<double>(),        \
      h->d_chain_out.as<double>(), h->d_chain_status.as<int32_t>(), h->d_new_out.as<double>(), h->d_new_status.as<int32_t>(), \
      h->d_accepted.as<int32_t>(), d_counters, heated ? h->d_slot.as<int>() : nullptr, h->d_ladder_p.as<double>(),            \
      h->d_ladder_l.as<double>(), n)
    switch (M.clock) {
      case 0: MCD_LAUNCH_MH_SMALL(0); break;
      case 1: MCD_LAUNCH_MH_SMALL(1); break;
      case 2: MCD_LAUNCH_MH_SMALL(2); break;
      default: MCD_LAUNCH_MH_SMALL(3); break;
    }
#undef MCD_LAUNCH_MH_SMALL
    h->launches += 1;
    CU_TRY(h, cudaGetLastError());
    return 0;
  }
  if (inc && !unfused) {
    const size_t smem = (size_t)8 * ((h->N + 31) / 32) * 4;
#define MCD_LAUNCH_FUSED(CC)                                                                                               \
  mh_fused_small_kernel<CC><<<(n + 7) / 8, 256, smem, st>>>(h->dm, T, P, h->d_P.as<double>(), h->d_chain.as<double>(),     \
      h->d_chain_y.as<double>(), h->d_chain_out.as<double>(), h->d_chain_status.as<int32_t>(), h->d_accepted.as<int32_t>(), \
      d_counters, heated ? h->d_slot.as<int>() : nullptr, h->d_ladder_p.as<double>(), h->d_ladder_l.as<double>(), n)
    switch (h->dm.clock) {
      case 0: MCD_LAUNCH_FUSED(0); break;
      case 1: MCD_LAUNCH_FUSED(1); break;
      case 2: MCD_LAUNCH_FUSED(2); break;
      default: MCD_LAUNCH_FUSED(3); break;
    }
#undef MCD_LAUNCH_FUSED
    h->launches += 1;
    CU_TRY(h, cudaGetLastError());
    return 0;
  }
  mh_propose_kernel<<<n, 64, 0, st>>>(h->d_chain.as<double>(), h->d_undo.as<double>(), h->d_rng.as<int2>(), h->d_meta.as<int4>(),
                                       h->d_lq.as<double>(), T, P, h->undo_stride, n);
  // Sub-tree moves on a given node that are too large for the per-chain incremental path still change the residual on
  // the sub tree's branches only: y' = y + Sigma^-1[:, A] delta_A is a contraction over the k-blocks covering A (all
  // chains share the range), on the tensor cores, instead of the full one; the posterior kernel then runs as usual.
  const bool range = inc_mode && h->oz_S != 0 && h->oz_P_S == h->oz_S && h->oz_X_S == h->oz_S && node > 1 && node != h->dm.root_r &&
                     (kind == MH_SCALE_SUBTREE || kind == MH_SCALE_SUBTREE_CONTRA || kind == MH_SCALE_RATE_SUBTREE) &&
                     !getenv("MCD_MH_NO_RANGE");
  if (range) {
    const int size = h->sub_size_h[node], k_lo = h->bidx[node];
    const int kb_lo = k_lo / OZ_KB, kb_hi = (k_lo + size - 1) / OZ_KB + 1;
    const int mode = kind == MH_SCALE_SUBTREE ? 0 : kind == MH_SCALE_SUBTREE_CONTRA ? 1 : 2;
    const size_t stride = (size_t)h->cap * h->ld8;
#define MCD_LAUNCH_RANGE(SS)                                                                                               \
  delta_split_kernel<SS><<<n, 256, (size_t)(kb_hi - kb_lo) * OZ_KB * 8, st>>>(                                            \
      h->N, h->S, h->dm.root_r, h->dm.parent, h->d_chain.as<double>(), h->d_undo.as<double>(), h->undo_stride,            \
      h->d_meta.as<int4>(), mode, node, size, k_lo, size, kb_lo, kb_hi, h->d_pX.as<signed char>(), h->ld8, stride,        \
      h->d_sX.as<double>(), n, h->d_ick.as<double>());                                                                    \
  if (oz_contract<SS>(h, false, n, 0, st, kb_lo, kb_hi, h->d_chain_y.as<double>(), h->ldyc)) return -1
    if (h->oz_S == 6) { MCD_LAUNCH_RANGE(6); } else { MCD_LAUNCH_RANGE(7); }
#undef MCD_LAUNCH_RANGE
    h->launches += 2;
    if (!getenv("MCD_MH_RANGE_FULL_POSTERIOR")) {  // score the move from the sub tree alone
#define MCD_LAUNCH_RDELTA(CC)                                                                                              \
  mh_range_delta_kernel<CC><<<n, 64, 0, st>>>(h->dm, T, h->d_chain.as<double>(), h->d_undo.as<double>(), h->undo_stride,  \
      h->d_meta.as<int4>(), h->d_chain_y.as<double>(), h->d_y.as<double>(), h->d_chain_out.as<double>(),                   \
      h->d_chain_status.as<int32_t>(), h->d_new_out.as<double>(), h->d_new_status.as<int32_t>(), mode, node, size, n)
      switch (h->dm.clock) {
        case 0: MCD_LAUNCH_RDELTA(0); break;
        case 1: MCD_LAUNCH_RDELTA(1); break;
        case 2: MCD_LAUNCH_RDELTA(2); break;
        default: MCD_LAUNCH_RDELTA(3); break;
      }
#undef MCD_LAUNCH_RDELTA
      h->launches += 1;
    } else {  // A/B switch: the full value-only posterior kernel on y'
      h->force_sym = true;  // d_y holds y = Sigma^-1 dx (not the Cholesky-form z)
      const int rc = enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_new_out.as<double>(), nullptr, h->d_new_status.as<int32_t>(), st, true);
      h->force_sym = false;
      if (rc) return -1;
    }
  }
  // Moves that leave every distance d_k = H m t_k r_k where it is -- lambda, mu, v; (H u, m / u); (m / u | H / u, rates u) --
  // leave y = Sigma^-1 (d - mu) where it is (up to one rounding of H m): the posterior kernel scores the proposed state
  // against the cached y, no contraction.
  const bool same_d = inc_mode && !range &&
                      ((kind == MH_SCALE_SCALAR && (node == 0 || node == 1 || node == 4)) || kind == MH_SCALE_H_M_CONTRA ||
                       kind == MH_SCALE_NORM_TREE_CONTRA_M || kind == MH_SCALE_NORM_TREE_CONTRA_H) &&
                      !getenv("MCD_MH_NO_SAME_D");
  if (same_d) {
    h->force_sym = true;
    h->y_override = h->d_chain_y.as<double>();
    h->y_override_ld = h->ldyc;
    const int rc = enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_new_out.as<double>(), nullptr, h->d_new_status.as<int32_t>(), st, true);
    h->y_override = nullptr;
    h->force_sym = false;
    if (rc) return -1;
    ++h->inc_steps;  // counts towards the periodic refresh (it is taken at the next small move)
  }
  MhYUpdate Y{};
  Y.mode = 0; Y.K = h->K; Y.ldk = h->ldk; Y.ldy = h->ldy; Y.ldyc = h->ldyc; Y.y_cur = h->d_chain_y.as<double>(); Y.y_new = h->d_y.as<double>();
  Y.P = h->d_P.as<double>(); Y.dl_n = h->d_dl_n.as<int>(); Y.dl_k = h->d_dl_k.as<int>(); Y.dl_d = h->d_dl_d.as<double>();
  if (inc) {
    const size_t smem = (size_t)8 * ((h->N + 31) / 32) * 4;
#define MCD_LAUNCH_DELTA(CC)                                                                                                 \
  mh_delta_kernel<CC><<<(n + 7) / 8, 256, smem, st>>>(h->dm, T, h->d_P.as<double>(), h->d_chain.as<double>(),              \
      h->d_undo.as<double>(), h->d_rng.as<int2>(), h->d_meta.as<int4>(), h->d_chain_y.as<double>(),                         \
      h->d_chain_out.as<double>(), h->d_chain_status.as<int32_t>(), h->d_new_out.as<double>(), h->d_new_status.as<int32_t>(), \
      h->d_dl_n.as<int>(), h->d_dl_k.as<int>(), h->d_dl_d.as<double>(), h->undo_stride, n)
    switch (h->dm.clock) {
      case 0: MCD_LAUNCH_DELTA(0); break;
      case 1: MCD_LAUNCH_DELTA(1); break;
      case 2: MCD_LAUNCH_DELTA(2); break;
      default: MCD_LAUNCH_DELTA(3); break;
    }
#undef MCD_LAUNCH_DELTA
    h->launches += 1;
    Y.mode = 1;
  } else if (range) {
    Y.mode = 2;
  } else if (same_d) {
    Y.mode = 0;  // y unchanged
  } else {
    h->force_sym = inc_mode;  // keep producing y while the incremental mode is on
    const int rc = enqueue<false>(h, 0, n, h->d_chain.as<double>(), h->d_new_out.as<double>(), nullptr, h->d_new_status.as<int32_t>(), st);
    h->force_sym = false;
    if (rc) return -1;
    Y.mode = inc_mode ? 2 : 0;
  }
  mh_accept_kernel<<<n, 256, 0, st>>>(h->d_chain.as<double>(), h->d_undo.as<double>(), h->d_rng.as<int2>(), h->d_meta.as<int4>(),
                                      h->d_lq.as<double>(), h->d_chain_out.as<double>(), h->d_new_out.as<double>(),
                                      h->d_chain_status.as<int32_t>(), h->d_new_status.as<int32_t>(), h->d_accepted.as<int32_t>(),
                                      d_counters, heated ? h->d_slot.as<int>() : nullptr, h->d_ladder_p.as<double>(),
                                      h->d_ladder_l.as<double>(), h->mc3_offset, use_root_jacobian, seed, iteration, S,
                                      h->undo_stride, n, Y);
  h->launches += 2;
  CU_TRY(h, cudaGetLastError());
  return 0;
}
int mh_step(mcd_handle* h, int kind, int node, double param, double tune, int use_root_jacobian, uint64_t seed, uint32_t iteration,
            int32_t* accepted) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mh_step");
  const int n = h->n_resident;
  if (n <= 0) return fail(h, "mcd_mh_step: no resident chains (call mcd_chains_set first)");
  if (mh_check(h, kind, node, param, tune)) return -1;
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (mh_prepare_value_path(h, n)) return -1;
  if (mh_enqueue(h, kind, node, param, tune, use_root_jacobian, seed, iteration, nullptr)) return -1;
  if (accepted) {
    cudaStream_t st = h->streams[0];
    CU_TRY(h, cudaMemcpyAsync(accepted, h->d_accepted.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
  }
  return 0;
}
// n_iterations sweeps over a list of proposals, everything enqueued back to back (no host round trip per proposal);
// the acceptance / invalid counts per proposal come back once at the end (what the reference's auto tuner consumes).
int mh_cycle(mcd_handle* h, int n_props, const mcd_mh_proposal* props, int n_iterations, uint64_t seed, uint32_t iteration0,
             uint64_t* accepted, uint64_t* invalid, uint32_t* iteration_next) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mh_cycle");
  const int n = h->n_resident;
  if (n <= 0) return fail(h, "mcd_mh_cycle: no resident chains (call mcd_chains_set first)");
  if (n_props <= 0 || n_props > MH_MAX_CYCLE || !props || n_iterations < 0) return fail(h, "mcd_mh_cycle: bad proposal list");
  for (int p = 0; p < n_props; ++p)
    if (props[p].repeat < 0 || mh_check(h, props[p].kind, props[p].node, props[p].param, props[p].tune)) return -1;
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (mh_prepare_value_path(h, n)) return -1;
  cudaStream_t st = h->streams[0];
  unsigned long long* cnt = h->d_counters.as<unsigned long long>();
  CU_TRY(h, cudaMemsetAsync(cnt, 0, (size_t)2 * n_props * sizeof(unsigned long long), st));
  uint32_t it = iteration0;
  static const bool per_step = getenv("MCD_MH_UNFUSED") != nullptr || getenv("MCD_MH_PER_STEP") != nullptr;  // A/B switches
  if (h->N <= SMALL_TREE_MAX_NODES && !per_step && n_iterations > 0) {
    // small trees: all sweeps in ONE launch (mh_small_cycle_kernel), one warp per chain
    std::vector<MhCycleEntry> tab((size_t)n_props);
    uint64_t steps = 0;
    for (int p = 0; p < n_props; ++p) {
      tab[p] = MhCycleEntry{props[p].kind, props[p].node, props[p].use_root_jacobian, props[p].repeat, props[p].param, props[p].tune};
      steps += (uint64_t)props[p].repeat;
    }
    steps *= (uint64_t)n_iterations;
    if (steps > 0xffffffffull - iteration0) return fail(h, "mcd_mh_cycle: the Philox iteration counter would wrap");
    if (!h->d_cycle.p) CU_TRY(h, cudaMalloc(&h->d_cycle.p, (size_t)MH_MAX_CYCLE * sizeof(MhCycleEntry)));
    // a sampler calls this with the same list iteration after iteration (it changes when the auto tuner has run): upload on change only
    const size_t tab_bytes = tab.size() * sizeof(MhCycleEntry);
    if (h->cycle_cache.size() != tab_bytes || memcmp(h->cycle_cache.data(), tab.data(), tab_bytes) != 0) {
      CU_TRY(h, cudaMemcpyAsync(h->d_cycle.p, tab.data(), tab_bytes, cudaMemcpyHostToDevice, st));
      CU_TRY(h, cudaStreamSynchronize(st));  // `tab` is pageable host memory
      h->cycle_cache.assign(reinterpret_cast<const unsigned char*>(tab.data()), reinterpret_cast<const unsigned char*>(tab.data()) + tab_bytes);
    }
    DevModel M = h->dm;
    M.quad_from_z = 0;
    const MhTopo T = mh_topo(h);
    const bool heated = h->mc3_C > 0;
    // few chains: four warps per CTA, one per SM sub-partition (the steps of a chain are a serial dependency chain, so a warp wants
    // a scheduler of its own; four chains walking the same code share the SM's instruction cache: 470 / 478 / 483 / 462
    // iterations/s with 1 / 2 / 4 / 8 warps per CTA on the 64-chain MC3 set); many chains: eight warps per CTA
    static const int wpb_forced = getenv("MCD_MH_WPB") ? atoi(getenv("MCD_MH_WPB")) : 0;   // experiment switch (1, 2, 4 or 8)
    const int wpb = (wpb_forced == 1 || wpb_forced == 2 || wpb_forced == 4 || wpb_forced == 8) ? wpb_forced
                    : n <= 2 * h->n_sms ? 4 : POST_THREADS / 32;
    const size_t fsmem = POST_SMEM_FIXED + ((size_t)M.K * M.K + (size_t)wpb * (M.S + M.N + M.K)) * 8 +
                         (size_t)wpb * (sizeof(MhOp) * MH_MAX_OPS + (size_t)(2 * M.N + 8) * 16 + MH_CYCLE_WARP_EXTRA);
    const int fgrid = std::min((n + wpb - 1) / wpb, 2 * h->n_sms);
#define MCD_LAUNCH_MH_CYCLE(CC)                                                                                              \
  mh_small_cycle_kernel<CC><<<fgrid, wpb * 32, fsmem, st>>>(M, T, h->d_cycle.as<MhCycleEntry>(), n_props, n_iterations, seed,  \
      iteration0, h->mc3_offset, h->d_P.as<double>(), h->d_chain.as<double>(), h->d_chain_out.as<double>(),                  \
      h->d_chain_status.as<int32_t>(), h->d_new_out.as<double>(), h->d_new_status.as<int32_t>(), h->d_accepted.as<int32_t>(),  \
      cnt, heated ? h->d_slot.as<int>() : nullptr, h->d_ladder_p.as<double>(), h->d_ladder_l.as<double>(), n)
    switch (M.clock) {
      case 0: MCD_LAUNCH_MH_CYCLE(0); break;
      case 1: MCD_LAUNCH_MH_CYCLE(1); break;
      case 2: MCD_LAUNCH_MH_CYCLE(2); break;
      default: MCD_LAUNCH_MH_CYCLE(3); break;
    }
#undef MCD_LAUNCH_MH_CYCLE
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    it += (uint32_t)steps;
  } else {
  for (int sweep = 0; sweep < n_iterations; ++sweep)
    for (int p = 0; p < n_props; ++p)
      for (int r = 0; r < props[p].repeat; ++r)
        if (mh_enqueue(h, props[p].kind, props[p].node, props[p].param, props[p].tune, props[p].use_root_jacobian, seed, it++,
                       cnt + 2 * p))
          return -1;
  }
  std::vector<unsigned long long> host((size_t)2 * n_props);
  CU_TRY(h, cudaMemcpyAsync(host.data(), cnt, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  for (int p = 0; p < n_props; ++p) {
    if (accepted) accepted[p] = host[2 * p];
    if (invalid) invalid[p] = host[2 * p + 1];
  }
  if (iteration_next) *iteration_next = it;
  return 0;
}
// temperature ladders + slot tables of the heated chains (replicated on every rank: n_global chains in groups of C)
int mc3_configure(mcd_handle* h, int n_global, int chain_offset, int C, const double* ladder_prior, const double* ladder_lik) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mc3_configure");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  if (C == 0) {  // back to cold chains (the global chain offset of the random streams stays)
    h->mc3_C = 0;
    if (chain_offset >= 0) h->mc3_offset = chain_offset;
    return 0;
  }
  if (C < 0 || n_global <= 0 || n_global % C != 0 || chain_offset < 0 || chain_offset > n_global || !ladder_prior || !ladder_lik)
    return fail(h, "mcd_mc3_configure: n_global must be a positive multiple of chains_per_group, ladders must be given");
  CU_TRY(h, cudaDeviceSynchronize());
  for (DevBuf* b : {&h->d_slot, &h->d_chain_of_slot, &h->d_ladder_p, &h->d_ladder_l, &h->d_swap_acc}) {
    if (b->p) cudaFree(b->p);
    b->p = nullptr;
  }
  std::vector<int> slot(n_global), cos(n_global);
  for (int c = 0; c < n_global; ++c) { slot[c] = c % C; cos[c] = c; }
  if (upload(h, h->d_slot, slot.data(), n_global) || upload(h, h->d_chain_of_slot, cos.data(), n_global) ||
      upload(h, h->d_ladder_p, ladder_prior, C) || upload(h, h->d_ladder_l, ladder_lik, C) ||
      upload(h, h->d_swap_acc, (const int*)nullptr, n_global / C))
    return -1;
  h->mc3_C = C; h->mc3_n_global = n_global; h->mc3_offset = chain_offset;
  return 0;
}
int mc3_swap(mcd_handle* h, int pair, uint64_t seed, uint32_t iteration, const double* d_stats_global, int32_t* accepted) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mc3_swap");
  if (h->mc3_C < 2) return fail(h, "mcd_mc3_swap: configure at least two temperatures first (mcd_mc3_configure)");
  if (pair >= h->mc3_C - 1) return fail(h, "mcd_mc3_swap: pair index out of range");
  if (!d_stats_global && (h->mc3_offset != 0 || h->mc3_n_global != h->n_resident))
    return fail(h, "mcd_mc3_swap: groups span ranks -- pass the all-gathered (ln prior, ln likelihood) table");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  cudaStream_t st = h->streams[0];
  const int G = h->mc3_n_global / h->mc3_C;
  const double* stats = d_stats_global ? d_stats_global : h->d_chain_out.as<double>() + MCD_OUT_LNPRIOR;
  mh_swap_kernel<<<(G + 127) / 128, 128, 0, st>>>(stats, d_stats_global ? 2 : MCD_OUT_COLS, h->d_slot.as<int>(),
                                                  h->d_chain_of_slot.as<int>(), h->d_ladder_p.as<double>(), h->d_ladder_l.as<double>(),
                                                  G, h->mc3_C, pair, seed, iteration, h->d_swap_acc.as<int>());
  h->launches += 1;
  CU_TRY(h, cudaGetLastError());
  if (accepted) {
    CU_TRY(h, cudaMemcpyAsync(accepted, h->d_swap_acc.p, (size_t)G * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
  }
  return 0;
}
int mc3_slots(mcd_handle* h, int32_t* slots) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mc3_slots");
  if (h->mc3_C <= 0 || !slots) return fail(h, "mcd_mc3_slots: no temperature ladder configured");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  CU_TRY(h, cudaMemcpyAsync(slots, h->d_slot.p, (size_t)h->mc3_n_global * 4, cudaMemcpyDeviceToHost, h->streams[0]));
  CU_TRY(h, cudaStreamSynchronize(h->streams[0]));
  return 0;
}

}  // namespace

// ---- NCCL, bound at run time: the evaluation needs no collective, only MC3's swap statistics cross GPUs (16 bytes per chain)
struct NcclId { char internal[128]; };   // ncclUniqueId
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId /* by value, as in nccl.h */, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      *(void**)(&api.GetUniqueId) = dlsym(api.lib, "ncclGetUniqueId");
      *(void**)(&api.CommInitRank) = dlsym(api.lib, "ncclCommInitRank");
      *(void**)(&api.AllGather) = dlsym(api.lib, "ncclAllGather");
      *(void**)(&api.CommDestroy) = dlsym(api.lib, "ncclCommDestroy");
      *(void**)(&api.GetErrorString) = dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    if (err) *err = "NCCL (libnccl.so.2) could not be loaded";
    return nullptr;
  }
  return &api;
}
static int comm_destroy(mcd_handle* h) {
  if (h->nccl_comm) {
    NcclApi* a = nccl_api(nullptr);
    if (a) a->CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr;
    h->comm_world = 0;
    h->comm_rank = -1;
  }
  return 0;
}

extern "C" {

const char* mcd_version(void) { return "mcmcdate_b200 0.1 (sm_100a)"; }

const char* mcd_last_error(const mcd_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mcd_create(const mcd_model_desc* d, mcd_handle** out) {
  if (!d || !out) return fail(nullptr, "mcd_create: null argument");
  *out = nullptr;
  const int N = d->n_nodes;
  if (N < 3 || N % 2 == 0 || !d->parent) return fail(nullptr, "mcd_create: need a bifurcating tree with >= 2 leaves");
  if (d->clock_model < 0 || d->clock_model > 3) return fail(nullptr, "mcd_create: unknown clock model");
  if (d->likelihood < 0 || d->likelihood > 3) return fail(nullptr, "mcd_create: unknown likelihood kind");
  if (d->likelihood == MCD_LIK_SPARSE) {
    if (d->n_sparse < 0 || (d->n_sparse > 0 && (!d->sparse_row || !d->sparse_col || !d->sparse_val)))
      return fail(nullptr, "mcd_create: sparse precision missing");
    for (int e = 0; e < d->n_sparse; ++e)
      if (d->sparse_row[e] < 0 || d->sparse_row[e] >= N - 2 || d->sparse_col[e] < 0 || d->sparse_col[e] >= N - 2)
        return fail(nullptr, "mcd_create: sparse precision index out of range");
  }
  if (!(d->ht > 0.0)) return fail(nullptr, "exponential: Rate is zero or negative.");  // exponential ht (Probability.hs:106)
  // topology checks: pre-order (parent < child), strictly bifurcating
  std::vector<int32_t> child0(N, -1), child1(N, -1);
  if (d->parent[0] != -1) return fail(nullptr, "mcd_create: parent[0] must be -1");
  for (int i = 1; i < N; ++i) {
    int p = d->parent[i];
    if (p < 0 || p >= i) return fail(nullptr, "mcd_create: nodes must be in pre-order (parent index < node index)");
    if (child0[p] < 0) child0[p] = i;
    else if (child1[p] < 0) child1[p] = i;
    else return fail(nullptr, "birthDeathWith: Tree is multifurcating.");
  }
  for (int i = 0; i < N; ++i) {
    if ((child0[i] < 0) != (child1[i] < 0)) return fail(nullptr, "mcd_create: unary nodes are not supported");
    if (child0[i] >= 0 && child0[i] != i + 1) return fail(nullptr, "mcd_create: nodes must be in pre-order");
  }
  if (child0[0] < 0) return fail(nullptr, "getBranches: Root node is not bifurcating.");
  for (int c = 0; c < d->n_cal; ++c) {
    if (d->cal_node[c] < 0 || d->cal_node[c] >= N) return fail(nullptr, "mcd_create: calibration node out of range");
    if (!(d->cal_lo_p[c] > 0 && d->cal_lo_p[c] < 1) && d->cal_lo[c] > 0) return fail(nullptr, "probabilityMass: out of (0,1)");
    if (!(d->cal_hi_p[c] > 0 && d->cal_hi_p[c] < 1) && std::isfinite(d->cal_hi[c])) return fail(nullptr, "probabilityMass: out of (0,1)");
  }
  for (int c = 0; c < d->n_con; ++c) {
    if (d->con_young[c] < 0 || d->con_young[c] >= N || d->con_old[c] < 0 || d->con_old[c] >= N)
      return fail(nullptr, "mcd_create: constraint node out of range");
    if (!(d->con_p[c] > 0 && d->con_p[c] < 1)) return fail(nullptr, "probabilityMass: out of (0,1)");
  }
  for (int b = 0; b < d->n_brace; ++b) {
    if (!(d->brace_sd[b] > 0)) return fail(nullptr, "braceSoftF: Standard deviation is zero or negative.");
    if (d->brace_off[b + 1] - d->brace_off[b] < 2) return fail(nullptr, "brace: need at least two nodes");
    for (int j = d->brace_off[b]; j < d->brace_off[b + 1]; ++j)
      if (d->brace_node[j] < 0 || d->brace_node[j] >= N) return fail(nullptr, "mcd_create: brace node out of range");
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, "mcd_create: no CUDA device (this library has no CPU fallback)");
  if (d->device < 0 || d->device >= ndev) return fail(nullptr, "mcd_create: bad device ordinal");
  mcd_handle* h = new mcd_handle();
  auto bail = [&](const char* what) {
    g_create_error = std::string(what) + (h->err.empty() ? "" : (": " + h->err));
    delete h;
    return -1;
  };
  h->device = d->device;
  if (cudaSetDevice(h->device) != cudaSuccess) return bail("cudaSetDevice failed");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major != 10) return bail("mcd_create: this build targets sm_100a (B200) only");
  h->n_sms = prop.multiProcessorCount;

  const int K = N - 2;
  h->N = N; h->K = K; h->S = 5 + 2 * N;
  h->ldk = (K + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
  h->Mp = (K + GEMM_PR - 1) / GEMM_PR * GEMM_PR;
  h->ldy = (N + GEMM_PR - 1) / GEMM_PR * GEMM_PR;  // >= Mp, and >= N: the near-critical sweep parks E[1..N-1] in the row
  h->ld8 = (K + OZ_KB - 1) / OZ_KB * OZ_KB;
  h->Mp8 = (K + OZ_N - 1) / OZ_N * OZ_N;
  {  // contraction pipe: INT8 tensor cores with 7 base-256 digit planes unless MCD_CONTRACTION says otherwise
    const char* e = getenv("MCD_CONTRACTION");
    h->oz_S = 7;
    if (e && !strcmp(e, "dmma")) h->oz_S = 0;
    else if (e && !strcmp(e, "i8s6")) h->oz_S = 6;
    else if (e && !strcmp(e, "i8s7")) h->oz_S = 7;
    else if (e && *e) return bail("mcd_create: MCD_CONTRACTION must be one of dmma, i8s6, i8s7");
    const char* pc = getenv("MCD_PIPE");   // chunks of the device-path pipeline (experiments; 0 / 1 = off)
    if (pc) h->pipe_chunks = std::max(0, std::min(16, atoi(pc)));
    const char* pk = getenv("MCD_PIPE_K3");
    if (pk) h->pipe_k3_per_sm = atoi(pk);
  }
  h->parent.assign(d->parent, d->parent + N);
  h->child1 = child1;
  const int root_r = child1[0];
  // branch order (closed form of getBranches + sumFirstTwo)
  h->bidx.resize(N);
  h->bidx[0] = -1;
  for (int i = 1; i < N; ++i) h->bidx[i] = (i == 1 || i == root_r) ? 0 : (i < root_r ? i - 1 : i - 2);
  // getMask (app/Hamiltonian.hs:33-47)
  h->mask.assign(h->S, 1);
  h->mask[2] = d->n_cal > 0 ? 1 : 0;
  h->mask[3] = 0;
  int n_inner_nonroot = 0;
  for (int i = 0; i < N; ++i) {
    if (child0[i] < 0) h->mask[3 + i] = 0;
    else if (i > 0) ++n_inner_nonroot;
  }
  h->mask[5 + N] = 0;
  h->D = 0;
  for (uint8_t m : h->mask) h->D += m;

  DevModel& M = h->dm;
  M.N = N; M.K = K; M.S = h->S; M.ldk = h->ldk; M.ldy = h->ldy;
  M.root_r = root_r; M.n_inner_nonroot = n_inner_nonroot;
  M.clock = d->clock_model; M.lik = d->likelihood; M.hmc_free_H = d->n_cal > 0; M.quad_from_z = 0;
  M.ht = d->ht; M.ln_ht = std::log(d->ht); M.logdet = d->logdet_sigma;
  M.lik_const = -(0.9189385332046727418 * (double)K);
  {
    std::vector<int32_t> penc(N);
    for (int i = 0; i < N; ++i) penc[i] = (i == 0 ? 0 : h->parent[i]) | (child0[i] < 0 ? LEAF_BIT : 0);
    if (upload(h, h->d_parent, penc.data(), N)) return bail("upload topology");
    M.parent = h->d_parent.as<int>();
  }
  // likelihood data
  std::vector<double> mu(h->ldk, 0.0);
  if (d->likelihood != MCD_LIK_NONE) {
    if (!d->mean || (!d->precision && d->likelihood != MCD_LIK_SPARSE)) return bail("mcd_create: mean / precision missing");
    std::memcpy(mu.data(), d->mean, K * 8);
  }
  if (upload(h, h->d_mu, mu.data(), h->ldk)) return bail("upload mean");
  M.mu = h->d_mu.as<double>();
  M.var = nullptr;
  if (d->likelihood == MCD_LIK_UNIVARIATE) {
    if (upload(h, h->d_var, d->precision, K)) return bail("upload variances");
    M.var = h->d_var.as<double>();
  } else if (d->likelihood == MCD_LIK_SPARSE) {
    // symmetrise: S_sym = (S + S^T)/2 has the same quadratic form, and -S_sym dx is the gradient
    std::vector<std::vector<std::pair<int, double>>> rows(K);
    for (int e = 0; e < d->n_sparse; ++e) {
      rows[d->sparse_row[e]].push_back({d->sparse_col[e], 0.5 * d->sparse_val[e]});
      rows[d->sparse_col[e]].push_back({d->sparse_row[e], 0.5 * d->sparse_val[e]});
    }
    std::vector<int> ptr(K + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < K; ++i) {
      std::sort(rows[i].begin(), rows[i].end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
      for (size_t e = 0; e < rows[i].size(); ++e) {
        if (!col.empty() && (int)col.size() > ptr[i] && col.back() == rows[i][e].first) val.back() += rows[i][e].second;
        else { col.push_back(rows[i][e].first); val.push_back(rows[i][e].second); }
      }
      ptr[i + 1] = (int)col.size();
    }
    if (N <= SMALL_TREE_MAX_NODES) {
      // small trees: densify, the fused single-launch kernel keeps the matrix in shared memory
      std::vector<double> P((size_t)h->Mp * h->ldk, 0.0);
      for (int i = 0; i < K; ++i)
        for (int e = ptr[i]; e < ptr[i + 1]; ++e) P[(size_t)i * h->ldk + col[e]] = val[e];
      if (upload(h, h->d_P, P.data(), P.size())) return bail("upload precision");
    } else {
      if (upload(h, h->d_sp_ptr, ptr.data(), ptr.size()) || upload(h, h->d_sp_col, col.data(), col.size()) ||
          upload(h, h->d_sp_val, val.data(), val.size()))
        return bail("upload sparse precision");
      M.sp_ptr = h->d_sp_ptr.as<int>(); M.sp_col = h->d_sp_col.as<int>(); M.sp_val = h->d_sp_val.as<double>();
      h->sparse = true;
      if (h->S + K > 27000) return bail("mcd_create: tree too large for the sparse contraction's shared memory");
      cudaFuncSetAttribute(sparse_contraction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    }
    M.lik = MCD_LIK_FULL;  // downstream kernels see y = S dx exactly like the dense case
  } else if (d->likelihood == MCD_LIK_FULL) {
    // P padded to [Mp][ldk].  The reference wraps the matrix as symmetric (L.Herm, app/Probability.hs:166) but `prepare`
    // writes the LU inverse of the covariance as it comes (L.invlndet, app/Main.hs:230), symmetric only up to rounding.
    // The kernels rely on symmetry (gradient = -P dx), so the symmetric part (P + P^T)/2 is what is stored: the same
    // quadratic form and exactly its gradient.  A matrix that is not symmetric beyond rounding is a caller error.
    std::vector<double> P((size_t)h->Mp * h->ldk, 0.0);
    double pmax = 0.0;
    for (size_t e = 0; e < (size_t)K * K; ++e) pmax = std::fmax(pmax, std::fabs(d->precision[e]));
    for (int i = 0; i < K; ++i) {
      for (int j = 0; j < K; ++j) {
        const double a = d->precision[(size_t)i * K + j], b = d->precision[(size_t)j * K + i];
        if (a != b && std::fabs(a - b) > 1e-6 * (std::fabs(a) + std::fabs(b)) && std::fabs(a - b) > 1e-9 * pmax)
          return bail("mcd_create: precision matrix is not symmetric");
        P[(size_t)i * h->ldk + j] = a == b ? a : 0.5 * (a + b);
      }
    }
    if (upload(h, h->d_P, P.data(), P.size())) return bail("upload precision");
    {  // power-of-two equilibration of the INT8 contraction: c_k = 2^round(-log2(P_kk) / 2), so that the digit planes of a
       // row of P' = C P C are relative to entries of order 1 and those of a chain's residuals to standardised residuals
       // (a precision matrix D P D with the diagonal D spread over many orders of magnitude splits like P itself)
      std::vector<double> ck(h->ld8, 1.0), ick(h->ld8, 1.0);
      bool ok = true;
      for (int k = 0; k < K; ++k) {
        const double pkk = P[(size_t)k * h->ldk + k];
        if (!(pkk > 0.0) || !std::isfinite(pkk)) { ok = false; break; }
      }
      if (ok)
        for (int k = 0; k < K; ++k) {
          const int e = (int)std::lround(-0.5 * std::log2(P[(size_t)k * h->ldk + k]));
          const int ec = std::max(-500, std::min(500, e));
          ck[k] = std::ldexp(1.0, ec);
          ick[k] = std::ldexp(1.0, -ec);
        }
      if (upload(h, h->d_ck, ck.data(), ck.size()) || upload(h, h->d_ick, ick.data(), ick.size())) return bail("upload equilibration");
    }
    if (d->precision_chol) {
      if (ensure_cholesky(h, d->precision_chol)) return bail("upload Cholesky factor");
    }  // else: factorised on the device at the first value-only evaluation (ensure_cholesky)
    if (make_tile_map(&h->tmP, h->d_P.as<double>(), h->Mp, h->ldk) != 0) return bail("cuTensorMapEncodeTiled failed for the precision matrix");
    if (gemm_f64_dmma_configure() != cudaSuccess) return bail("cudaFuncSetAttribute(gemm smem) failed");
  }
  // node prior tables
  const double SQRT_2_OVER_PI = 0.7978845608028654;
  std::vector<double> slo(d->n_cal), shi(d->n_cal), cs(d->n_con);
  for (int c = 0; c < d->n_cal; ++c) { slo[c] = SQRT_2_OVER_PI * d->cal_lo_p[c]; shi[c] = SQRT_2_OVER_PI * d->cal_hi_p[c]; }
  for (int c = 0; c < d->n_con; ++c) cs[c] = SQRT_2_OVER_PI * d->con_p[c];
  std::vector<std::vector<int2>> inc(N);
  for (int c = 0; c < d->n_cal; ++c) inc[d->cal_node[c]].push_back(make_int2(INC_CAL, c));
  for (int c = 0; c < d->n_con; ++c) {
    inc[d->con_young[c]].push_back(make_int2(INC_CON_YOUNG, c));
    inc[d->con_old[c]].push_back(make_int2(INC_CON_OLD, c));
  }
  for (int b = 0; b < d->n_brace; ++b)
    for (int j = d->brace_off[b]; j < d->brace_off[b + 1]; ++j) inc[d->brace_node[j]].push_back(make_int2(INC_BRACE, b));
  std::vector<int> inc_off(N + 1, 0);
  std::vector<int2> inc_ent;
  for (int i = 0; i < N; ++i) {
    inc_off[i] = (int)inc_ent.size();
    inc_ent.insert(inc_ent.end(), inc[i].begin(), inc[i].end());
  }
  inc_off[N] = (int)inc_ent.size();
  const int nbn = d->n_brace > 0 ? d->brace_off[d->n_brace] : 0;
  h->br_off_h.assign(1, 0);
  if (d->n_brace > 0) {
    h->br_off_h.assign(d->brace_off, d->brace_off + d->n_brace + 1);
    h->br_node_h.assign(d->brace_node, d->brace_node + nbn);
  }
  std::vector<int> br_off0(1, 0);
  if (upload(h, h->d_cal_node, d->cal_node, d->n_cal) || upload(h, h->d_cal_lo, d->cal_lo, d->n_cal) ||
      upload(h, h->d_cal_hi, d->cal_hi, d->n_cal) || upload(h, h->d_cal_slo, slo.data(), d->n_cal) ||
      upload(h, h->d_cal_shi, shi.data(), d->n_cal) || upload(h, h->d_con_y, d->con_young, d->n_con) ||
      upload(h, h->d_con_o, d->con_old, d->n_con) || upload(h, h->d_con_s, cs.data(), d->n_con) ||
      upload(h, h->d_br_off, d->n_brace > 0 ? d->brace_off : br_off0.data(), d->n_brace + 1) ||
      upload(h, h->d_br_node, d->brace_node, nbn) || upload(h, h->d_br_sd, d->brace_sd, d->n_brace) ||
      upload(h, h->d_inc_off, inc_off.data(), N + 1) || upload(h, h->d_inc_ent, inc_ent.data(), inc_ent.size()))
    return bail("upload prior tables");
  M.n_cal = d->n_cal; M.n_con = d->n_con; M.n_brace = d->n_brace;
  M.cal_node = h->d_cal_node.as<int>(); M.cal_lo = h->d_cal_lo.as<double>(); M.cal_hi = h->d_cal_hi.as<double>();
  M.cal_slo = h->d_cal_slo.as<double>(); M.cal_shi = h->d_cal_shi.as<double>();
  M.con_y = h->d_con_y.as<int>(); M.con_o = h->d_con_o.as<int>(); M.con_s = h->d_con_s.as<double>();
  M.br_off = h->d_br_off.as<int>(); M.br_node = h->d_br_node.as<int>(); M.br_sd = h->d_br_sd.as<double>();
  M.inc_off = h->d_inc_off.as<int>(); M.inc_ent = h->d_inc_ent.as<int2>();
  {
    std::vector<int4> inner;
    for (int i = 1; i < N; ++i)
      if (child0[i] >= 0) inner.push_back(make_int4(i, child1[i], inc_off[i], inc_off[i + 1] - inc_off[i]));
    if (upload(h, h->d_inner, inner.data(), inner.size())) return bail("upload inner-node list");
    M.inner = h->d_inner.as<int4>();
  }

  {  // theta <-> state index maps (toVector conses while folding left: reversed order)
    std::vector<int> tidx(h->S, -1), sidx(std::max(h->D, 1), 0);
    int t = h->D - 1;
    for (int j = 0; j < h->S; ++j)
      if (h->mask[j]) { tidx[j] = t; sidx[t] = j; --t; }
    if (upload(h, h->d_tidx, tidx.data(), tidx.size()) || upload(h, h->d_sidx, sidx.data(), sidx.size()) ||
        upload<double>(h, h->d_base, nullptr, h->S))
      return bail("upload theta maps");
  }
  for (int i = 0; i < N_STREAMS; ++i)
    if (cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
  // posterior kernels may need > 48 KiB dynamic smem on large trees
  {
    // large trees stage the state row [S] only (y is read from global memory); small ones P and per-warp rows
    const size_t post_smem = N <= SMALL_TREE_MAX_NODES
                                 ? POST_SMEM_FIXED + ((size_t)h->K * h->K + (size_t)(POST_THREADS / 32) * (h->S + N + h->K)) * 8
                                 : POST_SMEM_FIXED + (size_t)h->S * 8;
    if (post_smem > 220 * 1024)
      return bail("mcd_create: tree too large for the posterior kernel's shared-memory staging (more than ~14000 nodes)");
    const int lim = 220 * 1024;   // dynamic part; the kernels also hold ~2 KB of static shared memory (227 KB per CTA in total)
#define MCD_SET_SMEM(CC)                                                                                               \
  cudaFuncSetAttribute(posterior_kernel<256, CC, true, POST_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim); \
  cudaFuncSetAttribute(posterior_kernel<256, CC, false, POST_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    MCD_SET_SMEM(0) MCD_SET_SMEM(1) MCD_SET_SMEM(2) MCD_SET_SMEM(3)
#undef MCD_SET_SMEM
#define MCD_SET_SMEM(CC)                                                                                          \
  cudaFuncSetAttribute(small_tree_fused_kernel<CC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);     \
  cudaFuncSetAttribute(small_tree_fused_kernel<CC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    MCD_SET_SMEM(0) MCD_SET_SMEM(1) MCD_SET_SMEM(2) MCD_SET_SMEM(3)
#undef MCD_SET_SMEM
#define MCD_SET_SMEM(CC)                                                                               \
  cudaFuncSetAttribute(mh_small_tree_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);     \
  cudaFuncSetAttribute(mh_small_cycle_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    MCD_SET_SMEM(0) MCD_SET_SMEM(1) MCD_SET_SMEM(2) MCD_SET_SMEM(3)
#undef MCD_SET_SMEM
#define MCD_SET_SMEM(CC)                                                                                  \
  cudaFuncSetAttribute(mh_fused_small_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024); \
  cudaFuncSetAttribute(mh_fused_small_kernel<CC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    MCD_SET_SMEM(0) MCD_SET_SMEM(1) MCD_SET_SMEM(2) MCD_SET_SMEM(3)
#undef MCD_SET_SMEM
  }
  if (d->max_batch > 0 && ensure_capacity(h, d->max_batch, false, false)) return bail("allocating work buffers");
  if (cudaDeviceSynchronize() != cudaSuccess) return bail("device error during create");
  *out = h;
  return 0;
}

void mcd_destroy(mcd_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  comm_destroy(h);
  for (int i = 0; i < N_STREAMS; ++i)
    if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
  for (int i = 0; i < 3; ++i)
    if (h->pipe_st[i]) cudaStreamDestroy(h->pipe_st[i]);
  for (cudaEvent_t e : h->pipe_ev) cudaEventDestroy(e);
  if (h->nuts_flags) cudaFreeHost(h->nuts_flags);
  for (int i = 0; i < 2; ++i)
    if (h->nuts_ev[i]) cudaEventDestroy(h->nuts_ev[i]);
  for (auto& evs : h->ticket_ev)
    for (cudaEvent_t ev : evs)
      if (ev) cudaEventDestroy(ev);
  delete h;
}

int mcd_state_len(const mcd_handle* h) { return h ? h->S : -1; }
int mcd_dim(const mcd_handle* h) { return h ? h->K : -1; }
int mcd_hmc_dim(const mcd_handle* h) { return h ? h->D : -1; }
int mcd_branch_index(const mcd_handle* h, int32_t* out) {
  if (!h || !out) return -1;
  std::memcpy(out, h->bidx.data(), sizeof(int32_t) * h->N);
  return 0;
}
int mcd_mask(const mcd_handle* h, uint8_t* out) {
  if (!h || !out) return -1;
  std::memcpy(out, h->mask.data(), h->S);
  return 0;
}
// toVector (app/Hamiltonian.hs:49-53): free entries in REVERSED canonical order
int mcd_to_vector(const mcd_handle* h, const double* state, double* theta) {
  if (!h || !state || !theta) return -1;
  int i = h->D - 1;
  for (int j = 0; j < h->S; ++j)
    if (h->mask[j]) theta[i--] = state[j];
  return 0;
}
// fromVectorWith (app/Hamiltonian.hs:55-60)
int mcd_from_vector(const mcd_handle* h, const double* base, const double* theta, double* out) {
  if (!h || !base || !theta || !out) return -1;
  int i = h->D - 1;
  for (int j = 0; j < h->S; ++j) out[j] = h->mask[j] ? theta[i--] : base[j];
  return 0;
}

int mcd_eval(mcd_handle* h, int32_t n, const double* states, double* out, int32_t* status) {
  return eval_host<false>(h, n, states, out, nullptr, status);
}
int mcd_eval_grad(mcd_handle* h, int32_t n, const double* states, double* out, double* grad, int32_t* status) {
  return eval_host<true>(h, n, states, out, grad, status);
}
int mcd_eval_grad_theta(mcd_handle* h, int32_t n, const double* theta, const double* base_state, double* out,
                        double* grad_theta, int32_t* status) {
  return eval_theta_host(h, n, theta, base_state, out, grad_theta, status);
}
int64_t mcd_eval_grad_theta_async(mcd_handle* h, int32_t n, const double* theta, const double* base_state, double* out,
                                  double* grad_theta, int32_t* status) {
  int64_t ticket = -1;
  return eval_theta_host(h, n, theta, base_state, out, grad_theta, status, true, &ticket) == 0 ? ticket : -1;
}
int64_t mcd_eval_async(mcd_handle* h, int32_t n, const double* states, double* out, int32_t* status) {
  int64_t ticket = -1;
  return eval_host<false>(h, n, states, out, nullptr, status, true, &ticket) == 0 ? ticket : -1;
}
int64_t mcd_eval_grad_async(mcd_handle* h, int32_t n, const double* states, double* out, double* grad, int32_t* status) {
  int64_t ticket = -1;
  return eval_host<true>(h, n, states, out, grad, status, true, &ticket) == 0 ? ticket : -1;
}
int mcd_wait(mcd_handle* h, int64_t ticket) { return wait_ticket(h, ticket); }
int mcd_leapfrog(mcd_handle* h, int32_t n, int32_t n_steps, const double* theta0, const double* momentum0,
                 const double* base_state, const double* inv_mass, const double* step_size, double* theta_out,
                 double* momentum_out, double* out, double* energy, int32_t* status) {
  return leapfrog_host(h, n, n_steps, theta0, momentum0, base_state, inv_mass, step_size, theta_out, momentum_out, out,
                       energy, status);
}
int mcd_nuts(mcd_handle* h, int32_t n, const double* theta0, const double* base_state, const double* inv_mass,
             const double* step_size, const double* momentum0, int32_t max_depth, uint64_t seed, uint32_t iteration,
             double* theta_out, double* out, double* accept_stat, int32_t* info, int32_t* status) {
  return nuts_host(h, n, theta0, base_state, inv_mass, step_size, momentum0, max_depth, seed, iteration, theta_out, out,
                   accept_stat, info, status);
}
int mcd_chains_set(mcd_handle* h, int32_t n, const double* states) { return chains_set(h, n, states); }
int mcd_chains_nuts(mcd_handle* h, const double* inv_mass, const double* step_size, int32_t max_depth, uint64_t seed,
                    uint32_t iteration, double* accept_stat, int32_t* info, int32_t* status) {
  return nuts_host(h, 0, nullptr, nullptr, inv_mass, step_size, nullptr, max_depth, seed, iteration, nullptr, nullptr, accept_stat,
                   info, status, true);
}
int mcd_chains_get(mcd_handle* h, int32_t n, double* states, double* out, int32_t* status) {
  return chains_get(h, n, states, out, status);
}
int mcd_mh_cycle(mcd_handle* h, int32_t n_props, const mcd_mh_proposal* props, int32_t n_iterations, uint64_t seed,
                 uint32_t iteration0, uint64_t* accepted, uint64_t* invalid, uint32_t* iteration_next) {
  return mh_cycle(h, n_props, props, n_iterations, seed, iteration0, accepted, invalid, iteration_next);
}
int mcd_mc3_configure(mcd_handle* h, int32_t n_global, int32_t chain_offset, int32_t chains_per_group, const double* ladder_prior,
                      const double* ladder_lik) {
  return mc3_configure(h, n_global, chain_offset, chains_per_group, ladder_prior, ladder_lik);
}
int mcd_mc3_swap(mcd_handle* h, int32_t pair, uint64_t seed, uint32_t iteration, const double* d_stats_global, int32_t* accepted) {
  return mc3_swap(h, pair, seed, iteration, d_stats_global, accepted);
}
int mcd_mc3_slots(mcd_handle* h, int32_t* slots) { return mc3_slots(h, slots); }
void* mcd_chains_out_device(mcd_handle* h) { return h ? h->d_chain_out.p : nullptr; }
int mcd_mh_set_incremental(mcd_handle* h, int32_t on, int32_t refresh_every) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mcd_mh_set_incremental");
  if (refresh_every < 0) return fail(h, "mcd_mh_set_incremental: refresh_every must be >= 0 (0 keeps the current value)");
  if (on && !h->inc_enabled && h->n_resident > 0)
    return fail(h, "mcd_mh_set_incremental: switch the mode on before mcd_chains_set (the cached contraction results are built there)");
  h->inc_enabled = on != 0;
  if (refresh_every > 0) h->refresh_every = refresh_every;
  return 0;
}
int mcd_mh_get_incremental(const mcd_handle* h) { return h ? (h->inc_enabled && h->inc_ok ? 1 : 0) : -1; }
int mcd_chains_stats_device(mcd_handle* h, double* d_stats) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mcd_chains_stats_device");
  if (h->n_resident <= 0 || !d_stats) return fail(h, "mcd_chains_stats_device: no resident chains or null buffer");
  CU_TRY(h, cudaSetDevice(h->device));
  SerialScope order(h, h->streams[0]);  // after whatever the pipelined calls / other streams have in flight
  if (order.rc) return -1;
  CU_TRY(h, cudaMemcpy2DAsync(d_stats, 16, h->d_chain_out.as<double>() + MCD_OUT_LNPRIOR, MCD_OUT_COLS * 8, 16, h->n_resident,
                              cudaMemcpyDeviceToDevice, h->streams[0]));
  CU_TRY(h, cudaStreamSynchronize(h->streams[0]));
  return 0;
}
int mcd_comm_unique_id(void* id128) {
  std::string err;
  NcclApi* a = nccl_api(&err);
  if (!a || !id128) { g_create_error = a ? "mcd_comm_unique_id: null buffer" : err; return -1; }
  const int rc = a->GetUniqueId(id128);
  if (rc != 0) { g_create_error = std::string("ncclGetUniqueId: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"); return -1; }
  return 0;
}
int mcd_comm_init(mcd_handle* h, int32_t world, int32_t rank, const void* id128) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:comm_init");
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(h, "mcd_comm_init: bad world / rank / id");
  std::string err;
  NcclApi* a = nccl_api(&err);
  if (!a) return fail(h, err);
  CU_TRY(h, cudaSetDevice(h->device));
  comm_destroy(h);
  NcclId id;
  memcpy(id.internal, id128, 128);
  const int rc = a->CommInitRank(&h->nccl_comm, world, id, rank);
  if (rc != 0) { h->nccl_comm = nullptr; return fail(h, std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error")); }
  h->comm_world = world;
  h->comm_rank = rank;
  return 0;
}
int mcd_allgather_stats(mcd_handle* h, double* d_stats_global) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:allgather_stats");
  if (h->n_resident <= 0 || !d_stats_global) return fail(h, "mcd_allgather_stats: no resident chains or null buffer");
  if (!h->nccl_comm) return fail(h, "mcd_allgather_stats: no communicator (call mcd_comm_init first)");
  NcclApi* a = nccl_api(nullptr);
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->streams[0];
  SerialScope order(h, st);
  if (order.rc) return -1;
  const size_t n = (size_t)h->n_resident;
  if (!h->d_stats_local.p || h->d_stats_local_n < n) {
    if (h->d_stats_local.p) { CU_TRY(h, cudaStreamSynchronize(st)); cudaFree(h->d_stats_local.p); h->d_stats_local.p = nullptr; }
    CU_TRY(h, cudaMalloc(&h->d_stats_local.p, n * 16));
    h->d_stats_local_n = n;
  }
  CU_TRY(h, cudaMemcpy2DAsync(h->d_stats_local.p, 16, h->d_chain_out.as<double>() + MCD_OUT_LNPRIOR, MCD_OUT_COLS * 8, 16, n,
                              cudaMemcpyDeviceToDevice, st));
  const int rc = a->AllGather(h->d_stats_local.p, d_stats_global, n * 2, /* ncclFloat64 */ 8, h->nccl_comm, st);
  if (rc != 0) return fail(h, std::string("ncclAllGather: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"));
  CU_TRY(h, cudaStreamSynchronize(st));
  return 0;
}
int mcd_comm_destroy(mcd_handle* h) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  return comm_destroy(h);
}
int mcd_mh_step(mcd_handle* h, int32_t kind, int32_t node, double sd, double tune, int32_t use_root_jacobian, uint64_t seed,
                uint32_t iteration, int32_t* accepted) {
  return mh_step(h, kind, node, sd, tune, use_root_jacobian, seed, iteration, accepted);
}
int mcd_eval_device(mcd_handle* h, int32_t n, const double* d_states, double* d_out, int32_t* d_status, void* stream) {
  return eval_device<false>(h, n, d_states, d_out, nullptr, d_status, stream);
}
int mcd_eval_grad_device(mcd_handle* h, int32_t n, const double* d_states, double* d_out, double* d_grad,
                         int32_t* d_status, void* stream) {
  return eval_device<true>(h, n, d_states, d_out, d_grad, d_status, stream);
}
int mcd_set_contraction(mcd_handle* h, int32_t mode) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mcd_set_contraction");
  if (mode != MCD_CONTRACT_DMMA && mode != MCD_CONTRACT_I8_S6 && mode != MCD_CONTRACT_I8_S7)
    return fail(h, "mcd_set_contraction: unknown mode");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaDeviceSynchronize());
  h->oz_S = mode;
  return h->cap > 0 ? ensure_i8(h) : 0;
}
int mcd_get_contraction(const mcd_handle* h) { return h ? h->oz_S : -1; }
int64_t mcd_kernel_launches(const mcd_handle* h) { return h ? h->launches : -1; }
int mcd_set_kernel_timing(mcd_handle* h, int on) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mcd_set_kernel_timing");
  h->timing = on != 0;
  return 0;
}
// Sum of per-kernel durations (ms) over all calls since timing was enabled / last read:
// ms[0] = residual kernel, ms[1] = FP64 contraction, ms[2] = posterior kernel.  Synchronises.
int mcd_kernel_times(mcd_handle* h, double* ms, int64_t* n_calls) {
  if (!h || !ms || !n_calls) return -1;
  std::lock_guard<std::mutex> lock(h->mtx);
  NvtxRange nvtx_range("mcd:mcd_kernel_times");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaDeviceSynchronize());
  ms[0] = ms[1] = ms[2] = 0.0;
  *n_calls = (int64_t)h->tev.size() / 4;
  for (size_t i = 0; i + 3 < h->tev.size(); i += 4) {
    for (int j = 0; j < 3; ++j) {
      float t = 0.f;
      CU_TRY(h, cudaEventElapsedTime(&t, h->tev[i + j], h->tev[i + j + 1]));
      ms[j] += t;
    }
  }
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  h->tev.clear();
  return 0;
}
int mcd_synchronize(mcd_handle* h) {
  if (!h) return -1;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaDeviceSynchronize());
  return 0;
}

}  // extern "C"
