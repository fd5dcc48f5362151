// gemm_i8_ozaki.cuh -- the FP64 contraction  Y = DX . P  on the INT8 tensor pipe (tcgen05 + TMEM).
//
//   Y[b][m] = sum_k P[m][k] * DX[b][k]          b = chain, m/k = branch (MVN dimension)
//
// Same product as gemm_f64.cuh (the batched  dxs <# sigmaInv  of app/Probability.hs:169), evaluated with
// the error-free integer splitting of Ozaki et al.: every row of an operand is written as
//
//   x[k] = scale * sum_{s=0}^{S-1} q_s[k] * 256^-(s+1) + tail,     q_s[k] in [-128, 127]  (int8 "digit planes")
//
// scale = a power of two per row with |x| <= 0.498 scale; the digits are the balanced base-256 digits of the
// integer rint(x / scale * 2^(8S)), so the split is exact up to the rounding of that integer (|tail| <=
// 2^(-8S-1) scale).  The product of two rows is sum_{s,t} 256^-(s+t+2) <q_s, p_t>; the integer dot products are
// EXACT on the INT8 tensor cores (int32 accumulation: |q p| <= 2^14, K <= 2^14, at most S pairs per
// accumulator), pairs with s + t >= S are below the dropped tails and skipped, and all pairs with the same
// s + t share one TMEM accumulator.  S = 7 (56 bits per operand, 28 int8 products): error of the order of the
// rounding error of an FP64 GEMM; S = 6 (48 bits, 21 products): ~2^-8 times coarser.  Integer
// accumulation makes the result bit-reproducible and independent of tiling / summation order.
//
// Kernel: persistent CTAs, one tile = 128 chains (UMMA M, TMEM lanes) x 64 P rows (UMMA N), S accumulators of
// 64 TMEM columns (S*64 <= 512).  Warp-specialised: warp 0 = TMA producer (all digit planes of both operands
// for one 64-byte k-block per stage, SWIZZLE_64B), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer
// (kind::i8, M128 N64 K32, A operand kept in the collector across the products that share it), warps 2-9 =
// epilogue (tcgen05.ld, Horner in FP64, row/column scales).
#pragma once
#ifdef MCD_CHECK
#include <assert.h>
#endif
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gemm_f64.cuh"  // mbarrier / TMA wrappers, get_encode_tiled

namespace mcd {

constexpr int OZ_M = 128;        // chains per CTA tile
constexpr int OZ_N = 64;         // P rows per CTA tile
#ifndef MCD_OZ_KB
#define MCD_OZ_KB 64
#endif
constexpr int OZ_KB = MCD_OZ_KB; // reduction elements (bytes) per pipeline stage = one swizzle row (32 or 64)
constexpr int OZ_UK = 32;        // reduction depth of one tcgen05.mma kind::i8
#ifndef MCD_OZ_STAGES
#define MCD_OZ_STAGES (MCD_OZ_KB == 32 ? 4 : 2)
#endif
constexpr int OZ_STAGES = MCD_OZ_STAGES;
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_THREADS = 64 + 32 * OZ_EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int OZ_TMEM_COLS = 512;
constexpr int OZ_MAX_SLICES = 7;  // 8 S bits must fit one int64

// A chain in which at least half of the equilibrated residuals are more than this factor below the largest one is recomputed
// in FP64 (fp64_rows_kernel): the digits are relative to the row maximum, so with a typical residual R times below it the error
// of y relative to a typical component is bounded by K 2^-56 R -- 1.5e-11 at K = 2048 for R = 512.  (Relative to
// the chain's LARGEST component the split is always accurate to K 2^-55.)
#ifndef MCD_OZ_WIDE_RATIO
#define MCD_OZ_WIDE_RATIO 512.0
#endif
constexpr double OZ_WIDE_RATIO = MCD_OZ_WIDE_RATIO;
#ifndef MCD_K1_BATCH
#define MCD_K1_BATCH 2
#endif

template <int S>
__host__ __device__ constexpr int oz_stage_bytes() { return S * (OZ_M + OZ_N) * OZ_KB; }
template <int S>
__host__ __device__ constexpr size_t oz_smem_bytes() { return (size_t)OZ_STAGES * oz_stage_bytes<S>() + 1024 /*align*/ + 512 /*barriers*/; }

// ------------------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32, issued by ONE thread.  COLLECT selects what happens
// to the A operand in the tensor core's collector buffer: consecutive MMAs that share A read it from shared
// memory once (SASS: UTCIMMA ... .A_KEEP / .A_REUSE).
enum { OZ_A_DISCARD = 0, OZ_A_FILL = 1, OZ_A_USE = 2, OZ_A_LASTUSE = 3 };
template <int COLLECT>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
#define MCD_UMMA_I8(SUFFIX)                                                                       \
  asm volatile(                                                                                   \
      "{\n"                                                                                       \
      ".reg .pred p;\n"                                                                           \
      "setp.ne.b32 p, %4, 0;\n"                                                                   \
      "tcgen05.mma.cta_group::1.kind::i8" SUFFIX " [%0], %1, %2, %3, p;\n"                        \
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)                  \
      : "memory")
  if (COLLECT == OZ_A_FILL) MCD_UMMA_I8(".collector::a::fill");
  else if (COLLECT == OZ_A_USE) MCD_UMMA_I8(".collector::a::use");
  else if (COLLECT == OZ_A_LASTUSE) MCD_UMMA_I8(".collector::a::lastuse");
  else MCD_UMMA_I8("");
#undef MCD_UMMA_I8
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets lane (base lane + t)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// one lane of a converged warp (the compiler then predicates the single-thread tcgen05 / TMA instructions
// instead of wrapping each of them in an elect-and-retry loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(tm) : "memory");
}

// shared-memory matrix descriptor, K-major operand tile of OZ_KB-byte rows, swizzled over the row length:
//   start address >> 4 | LBO (unused for swizzled K-major) = 1 | SBO = 8 rows * OZ_KB bytes | version 1 |
//   layout 4 (SWIZZLE_64B) or 6 (SWIZZLE_32B)
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * OZ_KB) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(OZ_KB == 64 ? 4 : 6) << 61);
}
// exact int32 -> double without the conversion pipe: 2^52 + 2^31 + v as a bit pattern, minus the bias
__device__ __forceinline__ double oz_i2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}
// instruction descriptor: D = S32 (2 @4), A = B = signed int8 (1 @7, 1 @10), both K-major, N>>3 @17, M>>4 @24
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);

// ------------------------------------------------------------------------------ epilogue of one tile (all contraction kernels)
// One epilogue warp = 32 TMEM lanes (chains) x 32 columns (P rows) of every accumulator, in two chunks of 16 columns: tcgen05.ld of
// the S accumulators, exact INT32 -> FP64, Horner in 1/256, row x column scales, 16-byte stores (plus the cached y of the MH range
// update).  The accumulators are handed back to the MMA warp as soon as this warp's LAST tcgen05.ld has completed -- the arithmetic
// and the stores of that chunk only touch registers.  RANK0: arrive on the barrier of CTA rank 0 of the pair (cta_group::2 kernel).
template <int S, bool RANK0>
__device__ __forceinline__ void oz_epilogue_tile(uint32_t tlane, int chalf, int lane, double sa, const double* __restrict__ sb_tile,
                                                 double* __restrict__ yrow, const double* __restrict__ arow, uint64_t* acc_empty) {
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    const int c = 2 * chalf + cc;
    uint32_t v[S][16];
#pragma unroll
    for (int d = 0; d < S; ++d) tmem_ld_x16(tlane + (uint32_t)(d * OZ_N + c * 16), v[d]);
    tmem_ld_wait();
    if (cc == 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (RANK0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(smem_u32(acc_empty) & 0xFEFFFFFFu) : "memory");
        else mbar_arrive(acc_empty);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      double r0 = oz_i2d(v[S - 1][j]), r1 = oz_i2d(v[S - 1][j + 1]);
#pragma unroll
      for (int d = S - 2; d >= 0; --d) {
        r0 = fma(r0, 0.00390625, oz_i2d(v[d][j]));
        r1 = fma(r1, 0.00390625, oz_i2d(v[d][j + 1]));
      }
      const double2 sb = *reinterpret_cast<const double2*>(sb_tile + c * 16 + j);
      double2 o = make_double2(r0 * sa * sb.x, r1 * sa * sb.y);
      if (arow) {
        const double2 ya = *reinterpret_cast<const double2*>(arow + c * 16 + j);
        o.x += ya.x;
        o.y += ya.y;
      }
      *reinterpret_cast<double2*>(yrow + c * 16 + j) = o;
    }
  }
}

// ------------------------------------------------------------------------------ the contraction
// tmA: digit planes of the residuals, [S][Bp][ld8] int8 seen as a 2-D [S*Bp][ld8] tensor, box 128 x OZ_KB
// tmB: digit planes of P,             [S][Mp][ld8]                        [S*Mp][ld8],       box  64 x OZ_KB
// scaleA[b] = row scale of chain b (NaN marks a chain with non-finite residuals), scaleB[m] = row scale of
// P row m times 2^-16.
// Persistent: gridDim.x CTAs (one per SM) walk the tiles t = blockIdx.x, + gridDim.x, ...; tile t = (chain tile
// t / n_pr, P-row tile t % n_pr), so CTAs that run together share the chains' planes and sweep P, which stays
// L2-resident.  The shared-memory ring and its parities run on across tiles: the producer prefetches the next
// tile's first k-blocks while the epilogue warps drain TMEM.
#ifndef MCD_OZ_MAXNREG
#define MCD_OZ_MAXNREG 0
#endif
template <int S>
#if MCD_OZ_MAXNREG > 0
__global__ void __maxnreg__(MCD_OZ_MAXNREG)   // register cap: leaves room for CTAs of the HBM-side kernels on the same SM
#else
__global__ void __launch_bounds__(OZ_THREADS, 1)
#endif
gemm_i8_ozaki_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const double* __restrict__ scaleA, const double* __restrict__ scaleB, double* __restrict__ Y,
                     int nkb, int ldy, int Bp, int Mp, int bt_base, int n_pr, int n_tiles, int upper_tri, int kb_lo,
                     const double* __restrict__ yadd, int ldyadd) {
  static_assert(S >= 2 && S <= OZ_MAX_SLICES && S * OZ_N <= OZ_TMEM_COLS, "digit planes must fit TMEM");
  constexpr int STAGE = oz_stage_bytes<S>();
  constexpr int A_PLANE = OZ_M * OZ_KB, B_PLANE = OZ_N * OZ_KB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)OZ_STAGES * STAGE);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* acc_full = empty + OZ_STAGES;   // MMA warp -> epilogue: all accumulators of the tile are complete
  uint64_t* acc_empty = acc_full + 1;       // epilogue warps -> MMA warp: TMEM has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;

  // producer: fills stage g % STAGES with all digit planes of k-block kb of tile (bt0, pr0); g = running k-block count
  auto load_kblock = [&](int g, int kb, int bt0, int pr0) {
    const int st = g % OZ_STAGES;
    unsigned char* dst = smem + (size_t)st * STAGE;
    mbar_arrive_expect_tx(&full[st], STAGE);
#pragma unroll
    for (int s = 0; s < S; ++s) tma_load_2d(dst + s * A_PLANE, &tmA, kb * OZ_KB, s * Bp + bt0, &full[st]);
#pragma unroll
    for (int s = 0; s < S; ++s) tma_load_2d(dst + S * A_PLANE + s * B_PLANE, &tmB, kb * OZ_KB, s * Mp + pr0, &full[st]);
  };
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
#pragma unroll
      for (int s = 0; s < OZ_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, OZ_EPI_WARPS);
      mbar_fence_init();
    }
    __syncwarp();
  }
  if (warp == 1) tmem_alloc(tmem_slot, OZ_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int g = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * OZ_M;
      // upper-triangular operand (U = L^T of the Cholesky factor, value-only path): row m only has entries at
      // k >= m, so this tile's reduction starts at its own diagonal block
      // kb_lo .. nkb: the reduction may be restricted to the k-blocks where the chains' operand is non-zero (rank-limited
      // update Y = yadd + dX P of the MH path: dX is zero outside one sub tree's branches)
      for (int kb = upper_tri ? pr0 / OZ_KB : kb_lo; kb < nkb; ++kb, ++g) {
        if (g >= OZ_STAGES) mbar_wait(&empty[g % OZ_STAGES], ((g / OZ_STAGES) - 1) & 1);
        if (leader) load_kblock(g, kb, bt0, pr0);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    const bool leader = elect_one();
    int g = 0, it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      if (it > 0) {  // the epilogue warps must have read the previous tile out of TMEM
        mbar_wait(acc_empty, (it - 1) & 1);
        tc_fence_after();
      }
      const int kb0 = upper_tri ? ((t % n_pr) * OZ_N) / OZ_KB : kb_lo;
      for (int kb = kb0; kb < nkb; ++kb, ++g) {
        const int st = g % OZ_STAGES;
        mbar_wait(&full[st], (g / OZ_STAGES) & 1);
        tc_fence_after();
        if (leader) {
          // descriptor of the stage base; planes / k-steps are 16-byte-granular offsets added to the address field
          const uint64_t dbase = oz_smem_desc(smem_u32(smem + (size_t)st * STAGE));
#pragma unroll
          for (int ks = 0; ks < OZ_KB / OZ_UK; ++ks) {
#pragma unroll
            for (int a = 0; a < S; ++a) {
              const uint64_t da = dbase + (uint64_t)((a * A_PLANE + ks * OZ_UK) >> 4);
#pragma unroll
              for (int b = 0; b + a < S; ++b) {
                const uint64_t db = dbase + (uint64_t)((S * A_PLANE + b * B_PLANE + ks * OZ_UK) >> 4);
                // the first product into accumulator d = a + b is (a = 0, b = d) of the tile's first k-step
                const uint32_t acc = ((kb - kb0) | ks | a) != 0 ? 1u : 0u;
                const uint32_t td = tmem_base + (uint32_t)((a + b) * OZ_N);
                // plane a of the chains is shared by the S - a products of this inner loop: keep it in the collector
                if (S - a == 1) umma_i8<OZ_A_DISCARD>(td, da, db, OZ_IDESC, acc);
                else if (b == 0) umma_i8<OZ_A_FILL>(td, da, db, OZ_IDESC, acc);
                else if (b + a == S - 1) umma_i8<OZ_A_LASTUSE>(td, da, db, OZ_IDESC, acc);
                else umma_i8<OZ_A_USE>(td, da, db, OZ_IDESC, acc);
              }
            }
          }
          umma_commit(&empty[st]);  // frees the stage once these MMAs have read it
          if (kb == nkb - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps = 4 lane quarters x 2 column halves
    const int ew = warp - 2;
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes 32 (warp % 4) .. +31
    const int chalf = ew >> 2;     // columns [32 chalf, 32 chalf + 32) of every accumulator
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * OZ_M;
      const int b = bt0 + row;
      const double sa = scaleA[b];
      double* yrow = Y + (size_t)b * ldy + pr0;
      const double* arow = yadd ? yadd + (size_t)b * ldyadd + pr0 : nullptr;
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      oz_epilogue_tile<S, false>(tlane, chalf, lane, sa, scaleB + pr0, yrow, arow, acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, OZ_TMEM_COLS);
}

// ------------------------------------------------------------------------------ the contraction, plane-granular pipeline
// Same tiles, same MMAs, same epilogue as gemm_i8_ozaki_kernel -- bit-identical results (integer accumulation) -- but the
// shared-memory ring is recycled PLANE BY PLANE instead of stage by stage.  ncu on the stage-granular kernel: tensor pipe 78 %
// busy, every SM ingesting 38 B/clk of digit planes from L2; with two stages of one 84 KB k-block each, the refill of a stage
// could only start when ALL 56 MMAs of its k-block had retired and had to land within the 56 MMAs of the next one.  Here the
// MMAs of a k-block run with the chains' plane index a DESCENDING (a = S-1 needs only P's plane 0, a = S-2 planes 0..1, ...):
//   * plane a of the chains is released right after its 2 (S - a) MMAs (after 2, 6, 12, 20, 30, 42 of the 56 MMAs for a = 6..1),
//     so its refill for k-block g + 2 starts up to two k-blocks before it is needed;
//   * a k-block starts as soon as chains' plane S-1 and P's plane 0 (12 KB) have landed; the other planes may still be in flight.
// Barriers per stage: one "full" per plane (2 S), "empty" per chains' plane a >= 1 (S - 1) and one for chains' plane 0 + all of
// P's planes (released by the last MMAs of the k-block).
template <int S>
#if MCD_OZ_MAXNREG > 0
__global__ void __maxnreg__(MCD_OZ_MAXNREG)
#else
__global__ void __launch_bounds__(OZ_THREADS, 1)
#endif
gemm_i8_ozaki_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const double* __restrict__ scaleA, const double* __restrict__ scaleB, double* __restrict__ Y,
                        int nkb, int ldy, int Bp, int Mp, int bt_base, int n_pr, int n_tiles, int upper_tri, int kb_lo,
                        const double* __restrict__ yadd, int ldyadd) {
  static_assert(S >= 2 && S <= OZ_MAX_SLICES && S * OZ_N <= OZ_TMEM_COLS, "digit planes must fit TMEM");
  static_assert(OZ_STAGES * (3 * S) + 2 <= 62, "barrier block is 512 bytes");
  constexpr int STAGE = oz_stage_bytes<S>();
  constexpr int A_PLANE = OZ_M * OZ_KB, B_PLANE = OZ_N * OZ_KB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_A = reinterpret_cast<uint64_t*>(smem + (size_t)OZ_STAGES * STAGE);  // [stage][a]
  uint64_t* full_B = full_A + OZ_STAGES * S;                                         // [stage][b]
  uint64_t* empty_A = full_B + OZ_STAGES * S;   // [stage][a]; entry a = 0 stands for chains' plane 0 AND all planes of P
  uint64_t* acc_full = empty_A + OZ_STAGES * S;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int i = 0; i < OZ_STAGES * S; ++i) {
        mbar_init(&full_A[i], 1);
        mbar_init(&full_B[i], 1);
        mbar_init(&empty_A[i], 1);
      }
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, OZ_EPI_WARPS);
      mbar_fence_init();
    }
    __syncwarp();
  }
  if (warp == 1) tmem_alloc(tmem_slot, OZ_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane does everything)
    if (elect_one()) {
      int g = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * OZ_M;
        for (int kb = upper_tri ? pr0 / OZ_KB : kb_lo; kb < nkb; ++kb, ++g) {
          const int st = g % OZ_STAGES;
          const uint32_t prev = (uint32_t)((g / OZ_STAGES) - 1) & 1u;
          unsigned char* dst = smem + (size_t)st * STAGE;
          // chains' planes S-1 .. 1 in the order the MMA warp releases (and needs) them
#pragma unroll
          for (int a = S - 1; a >= 1; --a) {
            if (g >= OZ_STAGES) mbar_wait(&empty_A[st * S + a], prev);
            mbar_arrive_expect_tx(&full_A[st * S + a], A_PLANE);
            tma_load_2d(dst + a * A_PLANE, &tmA, kb * OZ_KB, a * Bp + bt0, &full_A[st * S + a]);
          }
          // P's planes (needed from the start of the k-block on, plane 0 first) and chains' plane 0 (needed last)
          if (g >= OZ_STAGES) mbar_wait(&empty_A[st * S + 0], prev);
#pragma unroll
          for (int b = 0; b < S; ++b) {
            mbar_arrive_expect_tx(&full_B[st * S + b], B_PLANE);
            tma_load_2d(dst + S * A_PLANE + b * B_PLANE, &tmB, kb * OZ_KB, b * Mp + pr0, &full_B[st * S + b]);
          }
          mbar_arrive_expect_tx(&full_A[st * S + 0], A_PLANE);
          tma_load_2d(dst, &tmA, kb * OZ_KB, bt0, &full_A[st * S + 0]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    if (elect_one()) {
      int g = 0, it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        if (it > 0) {  // the epilogue warps must have read the previous tile out of TMEM
          mbar_wait(acc_empty, (it - 1) & 1);
          tc_fence_after();
        }
        const int kb0 = upper_tri ? ((t % n_pr) * OZ_N) / OZ_KB : kb_lo;
        for (int kb = kb0; kb < nkb; ++kb, ++g) {
          const int st = g % OZ_STAGES;
          const uint32_t par = (uint32_t)(g / OZ_STAGES) & 1u;
          const uint64_t dbase = oz_smem_desc(smem_u32(smem + (size_t)st * STAGE));
#pragma unroll
          for (int a = S - 1; a >= 0; --a) {
            mbar_wait(&full_A[st * S + a], par);
            mbar_wait(&full_B[st * S + (S - 1 - a)], par);   // the plane of P this a needs for the first time
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < OZ_KB / OZ_UK; ++ks) {
              const uint64_t da = dbase + (uint64_t)((a * A_PLANE + ks * OZ_UK) >> 4);
#pragma unroll
              for (int b = 0; b + a < S; ++b) {
                const uint64_t db = dbase + (uint64_t)((S * A_PLANE + b * B_PLANE + ks * OZ_UK) >> 4);
                // the first product into accumulator d = a + b is (a = d, b = 0) of the tile's first k-step
                const uint32_t acc = ((kb - kb0) | ks | b) != 0 ? 1u : 0u;
                const uint32_t td = tmem_base + (uint32_t)((a + b) * OZ_N);
                if (S - a == 1) umma_i8<OZ_A_DISCARD>(td, da, db, OZ_IDESC, acc);
                else if (b == 0) umma_i8<OZ_A_FILL>(td, da, db, OZ_IDESC, acc);
                else if (b + a == S - 1) umma_i8<OZ_A_LASTUSE>(td, da, db, OZ_IDESC, acc);
                else umma_i8<OZ_A_USE>(td, da, db, OZ_IDESC, acc);
              }
            }
            umma_commit(&empty_A[st * S + a]);   // plane a of the chains (a = 0: and all of P's planes) may be refilled
          }
          if (kb == nkb - 1) umma_commit(acc_full);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (as in gemm_i8_ozaki_kernel)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * OZ_M;
      const int b = bt0 + row;
      const double sa = scaleA[b];
      double* yrow = Y + (size_t)b * ldy + pr0;
      const double* arow = yadd ? yadd + (size_t)b * ldyadd + pr0 : nullptr;
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      oz_epilogue_tile<S, false>(tlane, chalf, lane, sa, scaleB + pr0, yrow, arow, acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, OZ_TMEM_COLS);
}

// ------------------------------------------------------------------------------ the contraction on CTA pairs (cta_group::2)
// ncu on the one-CTA kernels: every tcgen05.mma M128 N64 K32 reads its whole B operand (2 KB) from shared memory, A once per
// S - a products (collector): 94 B/clk of operand reads + 47 B/clk of TMA fill against a 128 B/clk shared-memory pipe -- the
// tensor pipe tops out below 80 %.  Two SMs of a TPC as ONE tile of 256 chains x 64 P rows (UMMA M = 256): each CTA holds its own
// 128 chains' planes and HALF of P's rows (32), the pair's tensor cores read both halves, so the per-CTA operand traffic of B and
// its TMA fill halve (stage 70 KB instead of 84 KB: three stages fit).  CTA rank 0 issues the MMAs for the pair; both CTAs run a
// TMA producer (its completion bytes go to rank 0's "full" barrier) and the epilogue of their own 128 chains; tcgen05.commit is
// multicast to both CTAs' "empty" / "accumulators full" barriers; epilogue warps of both CTAs arrive on rank 0's "accumulators
// empty" barrier.  Same integer products as the one-CTA kernels: bit-identical results.
constexpr int OZ2_STAGES = 3;
constexpr int OZ2_NH = OZ_N / 2;   // P rows held by each CTA of the pair
template <int S>
__host__ __device__ constexpr int oz2_stage_bytes() { return S * (OZ_M + OZ2_NH) * OZ_KB; }
template <int S>
__host__ __device__ constexpr size_t oz2_smem_bytes() { return (size_t)OZ2_STAGES * oz2_stage_bytes<S>() + 1024 + 256; }
constexpr uint32_t OZ2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_N >> 3) << 17) | ((uint32_t)((2 * OZ_M) >> 4) << 24);
constexpr uint32_t OZ_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: rank 0's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_smem, uint32_t ncols) {  // whole warp, in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same offset in CTA rank 0
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
          "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar) & OZ_PEER_MASK) : "memory");
}
template <int COLLECT>
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
#define MCD_UMMA_I8_2(SUFFIX)                                                                     \
  asm volatile(                                                                                   \
      "{\n"                                                                                       \
      ".reg .pred p;\n"                                                                           \
      "setp.ne.b32 p, %4, 0;\n"                                                                   \
      "tcgen05.mma.cta_group::2.kind::i8" SUFFIX " [%0], %1, %2, %3, p;\n"                        \
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)                  \
      : "memory")
  if (COLLECT == OZ_A_FILL) MCD_UMMA_I8_2(".collector::a::fill");
  else if (COLLECT == OZ_A_USE) MCD_UMMA_I8_2(".collector::a::use");
  else if (COLLECT == OZ_A_LASTUSE) MCD_UMMA_I8_2(".collector::a::lastuse");
  else MCD_UMMA_I8_2("");
#undef MCD_UMMA_I8_2
}
// arrive (once every MMA issued so far has completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::
                   "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// tmA: chains' planes, box 128 rows; tmBh: P's planes, box 32 rows.  Grid = 2 x (number of pairs), persistent: pair p walks the
// tiles t = p, p + n_pairs, ...; tile t = (chain tile of 256 rows t / n_pr, P-row tile t % n_pr).  Chain rows beyond the batch (an odd
// number of 128-row tiles) are computed and discarded by the caller's padding (buffers hold a multiple of 256 rows).
template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
gemm_i8_ozaki_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                          const double* __restrict__ scaleA, const double* __restrict__ scaleB, double* __restrict__ Y,
                          int nkb, int ldy, int Bp, int Mp, int bt_base, int n_pr, int n_tiles, int upper_tri, int kb_lo,
                          const double* __restrict__ yadd, int ldyadd) {
  static_assert(S >= 2 && S <= OZ_MAX_SLICES && S * OZ_N <= OZ_TMEM_COLS, "digit planes must fit TMEM");
  constexpr int STAGE = oz2_stage_bytes<S>();
  constexpr int A_PLANE = OZ_M * OZ_KB, B_PLANE = OZ2_NH * OZ_KB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)OZ2_STAGES * STAGE);
  uint64_t* empty = full + OZ2_STAGES;
  uint64_t* acc_full = empty + OZ2_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmBh);
#pragma unroll
      for (int s = 0; s < OZ2_STAGES; ++s) {
        mbar_init(&full[s], 1);    // rank 0's producer (arrive.expect_tx for the bytes of BOTH CTAs)
        mbar_init(&empty[s], 1);   // the multicast commit
      }
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, 2 * OZ_EPI_WARPS);   // epilogue warps of both CTAs (used on rank 0 only)
      mbar_fence_init();
    }
    __syncwarp();
  }
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / TMA credit
  if (warp == 1) tmem_alloc2(tmem_slot, OZ_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int g = 0;
      for (int t = pair; t < n_tiles; t += n_pairs) {
        const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * (2 * OZ_M) + (int)rank * OZ_M;
        for (int kb = upper_tri ? pr0 / OZ_KB : kb_lo; kb < nkb; ++kb, ++g) {
          const int st = g % OZ2_STAGES;
          if (g >= OZ2_STAGES) mbar_wait(&empty[st], (uint32_t)((g / OZ2_STAGES) - 1) & 1u);
          unsigned char* dst = smem + (size_t)st * STAGE;
          if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * STAGE);
#pragma unroll
          for (int s = 0; s < S; ++s) tma_load_2d_pair(dst + s * A_PLANE, &tmA, kb * OZ_KB, s * Bp + bt0, &full[st]);
#pragma unroll
          for (int s = 0; s < S; ++s)
            tma_load_2d_pair(dst + S * A_PLANE + s * B_PLANE, &tmBh, kb * OZ_KB, s * Mp + pr0 + (int)rank * OZ2_NH, &full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one elected lane of CTA rank 0
    if (rank == 0 && elect_one()) {
      int g = 0, it = 0;
      for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
        if (it > 0) {
          mbar_wait(acc_empty, (it - 1) & 1);
          tc_fence_after();
        }
        const int kb0 = upper_tri ? ((t % n_pr) * OZ_N) / OZ_KB : kb_lo;
        for (int kb = kb0; kb < nkb; ++kb, ++g) {
          const int st = g % OZ2_STAGES;
          mbar_wait(&full[st], (uint32_t)(g / OZ2_STAGES) & 1u);
          tc_fence_after();
          const uint64_t dbase = oz_smem_desc(smem_u32(smem + (size_t)st * STAGE));
#pragma unroll
          for (int ks = 0; ks < OZ_KB / OZ_UK; ++ks) {
#pragma unroll
            for (int a = 0; a < S; ++a) {
              const uint64_t da = dbase + (uint64_t)((a * A_PLANE + ks * OZ_UK) >> 4);
#pragma unroll
              for (int b = 0; b + a < S; ++b) {
                const uint64_t db = dbase + (uint64_t)((S * A_PLANE + b * B_PLANE + ks * OZ_UK) >> 4);
                const uint32_t acc = ((kb - kb0) | ks | a) != 0 ? 1u : 0u;
                const uint32_t td = tmem_base + (uint32_t)((a + b) * OZ_N);
                if (S - a == 1) umma_i8_pair<OZ_A_DISCARD>(td, da, db, OZ2_IDESC, acc);
                else if (b == 0) umma_i8_pair<OZ_A_FILL>(td, da, db, OZ2_IDESC, acc);
                else if (b + a == S - 1) umma_i8_pair<OZ_A_LASTUSE>(td, da, db, OZ2_IDESC, acc);
                else umma_i8_pair<OZ_A_USE>(td, da, db, OZ2_IDESC, acc);
              }
            }
          }
          umma_commit_pair(&empty[st]);
          if (kb == nkb - 1) umma_commit_pair(acc_full);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue of this CTA's 128 chains
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
      const int pr0 = (t % n_pr) * OZ_N, bt0 = bt_base + (t / n_pr) * (2 * OZ_M) + (int)rank * OZ_M;
      const int b = bt0 + row;
      const double sa = scaleA[b];
      double* yrow = Y + (size_t)b * ldy + pr0;
      const double* arow = yadd ? yadd + (size_t)b * ldyadd + pr0 : nullptr;
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      oz_epilogue_tile<S, true>(tlane, chalf, lane, sa, scaleB + pr0, yrow, arow, acc_empty);
    }
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs are done with the pair's tensor memory
  if (warp == 1) tmem_dealloc2(tmem_base, OZ_TMEM_COLS);
}

// ------------------------------------------------------------------------------ digit planes
// scale = 2^(e+1) with 2^(e-1) <= max|x| < 2^e (2^(e+2) if max|x| > 0.996 * 2^e), so |x / scale| <= 0.498, the
// range S balanced base-256 digits cover.  A row holding a non-finite value gets scale = NaN (and zero
// digits): the whole output row becomes NaN.  Returns the scale; *mult = 2^(8S) / scale (0 for an all-zero or
// non-finite row).
template <int S>
__device__ __forceinline__ double oz_row_scale(double amax, bool finite, double* mult) {
  *mult = 0.0;
  // rows beyond 2^+-900 would overflow the power-of-two multipliers: treated as non-finite / as zero (their
  // contribution is below 1e-270 in absolute terms)
  if (!finite || amax > 8.452712498170644e+270) return __longlong_as_double(0x7ff8000000000000LL);
  if (amax < 1.1830521861667747e-271) return 0.0;
  int e;
  const double f = frexp(amax, &e);  // amax = f * 2^e, f in [0.5, 1)
  const int es = f > 0.996 ? e + 2 : e + 1;
  *mult = ldexp(1.0, 8 * S - es);
  return ldexp(1.0, es);
}
// The S digits of x as the low S bytes of a 64-bit word, least significant digit in byte 0:
//   I = rint(x * mult) = sum_j d_j 256^j, d_j in [-128, 127]  <=>  d_j = byte_j(I + sum_j 128 * 256^j) XOR 0x80
template <int S>
__device__ __forceinline__ unsigned long long oz_digit_bytes(double x, double mult) {
  constexpr unsigned long long C = 0x8080808080808080ULL >> (8 * (8 - S));
  return ((unsigned long long)__double2ll_rn(x * mult) + C) ^ C;
}

// colmul / rowmul (nullable): the row that is split is x[k] * colmul[k] * rowmul[row] -- the power-of-two equilibration of the
// precision matrix (P' = C P C, see oz_equilibration) -- and rowpost[row] (nullable) multiplies the published row scale.
template <int S>
__global__ void __launch_bounds__(256)
oz_split_rows_kernel(const double* __restrict__ X, int ldx, int rows, int K, signed char* __restrict__ planes, int ld8,
                     size_t plane_stride, double* __restrict__ scale, double post_scale,
                     const double* __restrict__ colmul = nullptr, const double* __restrict__ rowmul = nullptr,
                     const double* __restrict__ rowpost = nullptr) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const double* x = X + (size_t)row * ldx;
  const double rm = rowmul ? rowmul[row] : 1.0;
  double amax = 0.0;
  bool finite = true;
  for (int k = lane; k < K; k += 32) {
    const double a = fabs(x[k] * (colmul ? colmul[k] : 1.0) * rm);
    finite = finite && (a <= 1.7976931348623157e308);
    amax = fmax(amax, a);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    finite = __shfl_xor_sync(0xffffffffu, (int)finite, off) && finite;
  }
  double mult;
  const double sc = oz_row_scale<S>(amax, finite, &mult);
  if (lane == 0) scale[row] = sc * post_scale * (rowpost ? rowpost[row] : 1.0);
  signed char* out = planes + (size_t)row * ld8;
  for (int k = lane; k < K; k += 32) {
    const unsigned long long q = oz_digit_bytes<S>(finite ? x[k] * (colmul ? colmul[k] : 1.0) * rm : 0.0, mult);
#pragma unroll
    for (int s = 0; s < S; ++s) out[(size_t)s * plane_stride + k] = (signed char)(q >> (8 * (S - 1 - s)));
  }
}

// ------------------------------------------------------------------------------ K1 fused with the split
// One CTA per chain: residuals (residual_kernel's arithmetic: heightTreeToLengthTree, getBranches + sumFirstTwo,
// scaling, minus the mean) into shared memory, block-wide max, then every thread turns 4 consecutive residuals
// into one 32-bit word per digit plane and stores it (a warp writes 128 contiguous bytes per plane).  All
// digits of a residual come from ONE float-to-int64 conversion (oz_digit_bytes).  Reads 8S bytes per chain, writes S_planes * ld8 bytes (<= the 8K
// bytes of the FP64 residual row it replaces).  Dynamic shared memory: 8 ld8 bytes.
template <int S>
__global__ void __launch_bounds__(256)
residual_split_kernel(int N, int K, int SL, int root_r, const int* __restrict__ parent, const double* __restrict__ mu,
                      const double* __restrict__ states, signed char* __restrict__ planes, int ld8, size_t plane_stride,
                      double* __restrict__ scale, int B, const double* __restrict__ ick, int* __restrict__ widecnt) {
  extern __shared__ __align__(16) unsigned char smem_rs[];
  double* sdx = reinterpret_cast<double*>(smem_rs);  // [ld8]
  __shared__ double s_amax[8];
  __shared__ int s_bad[8];
  const int chain = blockIdx.x;
  if (chain >= B) return;
  const int tid = threadIdx.x;
  const double* x = states + (size_t)chain * SL;
  const double* h = x + 3;
  const double* r = x + 5 + N;
  const double sc = x[2] * x[3 + N];  // tH * rMu
  double amax = 0.0;
  int bad = 0;
  if (tid < ld8 - K) sdx[K + tid] = 0.0;  // k-padding (< 64 entries)
  {
    // Nodes in batches of NB per thread with the loads issued by dependency level: parent indices, heights and rates of the batch
    // first (independent), then the parent heights (the gather depends on the index), then the arithmetic.  ncu's source page
    // showed every node's subtraction waiting for its own gather; NB = 2 keeps 32 registers (8 CTAs per SM) and halves the
    // round trips: 0.107 -> 0.090 ms.  NB = 4 (40 registers) and NB = 8 (120 registers, 2 CTAs per SM) are slower.
    constexpr int NB = MCD_K1_BATCH;
    const double root_e = (h[0] - h[root_r]) * r[root_r];
    for (int base = 1; base < N; base += NB * 256) {
      int pi[NB];
      double hi[NB], ri[NB], hp[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int i = base + tid + j * 256;
        const bool on = i < N;
        pi[j] = on ? (parent[i] & 0x7fffffff) : 0;
        hi[j] = on ? h[i] : 0.0;
        ri[j] = on ? r[i] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) hp[j] = h[pi[j]];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int i = base + tid + j * 256;
        if (i >= N || i == root_r) continue;
        double e = (hp[j] - hi[j]) * ri[j];
        if (i == 1) e = e + root_e;
        const int k = i < root_r ? i - 1 : i - 2;
        const double d = (e * sc - mu[k]) * ick[k];   // (loading mu / 1/c with the first batch: 40 registers, no gain)
        sdx[k] = d;
        const double a = fabs(d);
        bad |= !(a <= 1.7976931348623157e308);
        amax = fmax(amax, a);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    bad |= __shfl_xor_sync(0xffffffffu, bad, off);
  }
  if ((tid & 31) == 0) { s_amax[tid >> 5] = amax; s_bad[tid >> 5] = bad; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) { amax = fmax(amax, s_amax[w]); bad |= s_bad[w]; }
  const bool finite = bad == 0;
  double mult;  // non-finite rows: zero digits, NaN scale
  const double scl = oz_row_scale<S>(amax, finite, &mult);
  if (tid == 0) scale[chain] = scl;
  // dynamic range of this chain's residuals: the digits are relative to the LARGEST one.  Every warp counts the residuals it
  // converts that are more than OZ_WIDE_RATIO times smaller than the largest; fp64_rows_kernel adds the counts up and
  // recomputes the chain in FP64 when they are the majority (no extra barrier here)
  const double small_below = finite ? amax * (1.0 / OZ_WIDE_RATIO) : -1.0;
  int n_small = 0;
  signed char* prow = planes + (size_t)chain * ld8;
  for (int q4 = tid; q4 < ld8 / 4; q4 += 256) {
    const double2 x01 = *reinterpret_cast<const double2*>(sdx + 4 * q4);
    const double2 x23 = *reinterpret_cast<const double2*>(sdx + 4 * q4 + 2);
    // (the zero padding beyond K counts as small here; fp64_rows_kernel subtracts it)
    n_small += (int)(fabs(x01.x) < small_below) + (int)(fabs(x01.y) < small_below) + (int)(fabs(x23.x) < small_below) +
               (int)(fabs(x23.y) < small_below);
    const unsigned long long j0 = oz_digit_bytes<S>(finite ? x01.x : 0.0, mult), j1 = oz_digit_bytes<S>(finite ? x01.y : 0.0, mult),
                             j2 = oz_digit_bytes<S>(finite ? x23.x : 0.0, mult), j3 = oz_digit_bytes<S>(finite ? x23.y : 0.0, mult);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int byte = S - 1 - s;  // plane s = digit S-1-s (most significant first)
      const uint32_t a0 = byte < 4 ? (uint32_t)j0 : (uint32_t)(j0 >> 32), a1 = byte < 4 ? (uint32_t)j1 : (uint32_t)(j1 >> 32),
                     a2 = byte < 4 ? (uint32_t)j2 : (uint32_t)(j2 >> 32), a3 = byte < 4 ? (uint32_t)j3 : (uint32_t)(j3 >> 32);
      const uint32_t sel = (uint32_t)(byte & 3) | ((uint32_t)(4 + (byte & 3)) << 4);
      const uint32_t lo = __byte_perm(a0, a1, sel), hi = __byte_perm(a2, a3, sel);   // {a0[b], a1[b]}, {a2[b], a3[b]}
      *reinterpret_cast<uint32_t*>(prow + (size_t)s * plane_stride + 4 * q4) = __byte_perm(lo, hi, 0x5410);
    }
  }
  if (widecnt) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_small += __shfl_xor_sync(0xffffffffu, n_small, off);
    if ((tid & 31) == 0) widecnt[chain * 8 + (tid >> 5)] = n_small;
  }
}

// MH path, sub-tree moves on a given node j: the residual changes on the branches of the sub tree only, k in
// [k_lo, k_lo + span) (contiguous in pre-order).  This kernel forms delta = dx(proposed state) - dx(current state) there
// from the state row (already modified in place) and the proposal's undo log, and writes its digit planes for the k-blocks
// [kb_lo, kb_hi) that cover the range (zero elsewhere in those blocks); the contraction then runs over these k-blocks only
// and adds the cached y of the current state.  Undo layout (mh_propose_kernel, ranges in order):
//   heights (scale sub tree / contrary): ub[i - j] = old h_i, i in [j, j + size);
//   contrary rates: ub[size + (i - j - 1)] = old r_i for i > j, ub[2 size - 1] = old r_j;
//   rate sub tree: ub[i - j] = old r_i.
// mode: 0 heights only, 1 heights + rates (contrary), 2 rates only.  One CTA per chain; dynamic smem 8 * span_pad bytes.
template <int S>
__global__ void __launch_bounds__(256)
delta_split_kernel(int N, int SL, int root_r, const int* __restrict__ parent, const double* __restrict__ states,
                   const double* __restrict__ undo, int undo_stride, const int4* __restrict__ meta, int mode, int j, int size,
                   int k_lo, int span, int kb_lo, int kb_hi, signed char* __restrict__ planes, int ld8, size_t plane_stride,
                   double* __restrict__ scale, int B, const double* __restrict__ ick) {
  extern __shared__ __align__(16) unsigned char smem_ds[];
  double* sd = reinterpret_cast<double*>(smem_ds);  // [(kb_hi - kb_lo) * OZ_KB]
  __shared__ double s_amax[8];
  __shared__ int s_bad[8];
  const int chain = blockIdx.x;
  if (chain >= B) return;
  const int tid = threadIdx.x;
  const double* x = states + (size_t)chain * SL;
  const double* h = x + 3;
  const double* r = x + 5 + N;
  const double* ub = undo + (size_t)chain * undo_stride;
  const double sc = x[2] * x[3 + N];
  const bool moved = meta[chain].x > 0;  // an invalid proposal left the chain unchanged: delta = 0
  const int k_base = kb_lo * OZ_KB, n_k = (kb_hi - kb_lo) * OZ_KB;
  double amax = 0.0;
  int bad = 0;
  for (int q = tid; q < n_k; q += 256) {
    const int k = k_base + q;
    double d = 0.0;
    if (moved && k >= k_lo && k < k_lo + span) {
      const int i = k + 1 < root_r ? k + 1 : k + 2;  // k > 0: node of branch k (left side i = k + 1, right side i = k + 2)
      const int p = parent[i] & 0x7fffffff;
#ifdef MCD_CHECK
      assert(i >= j && i < j + size && i < N && p < i && (p >= j || i == j) && 2 * size - 1 < undo_stride);
#endif
      const double hi_n = h[i], hp_n = h[p], r_n = r[i];
      double hi_o = hi_n, hp_o = hp_n, r_o = r_n;
      if (mode != 2) {
        hi_o = ub[i - j];
        if (p >= j) hp_o = ub[p - j];  // the parent of the sub tree's root keeps its height
      }
      if (mode == 1) r_o = i > j ? ub[size + (i - j - 1)] : ub[2 * size - 1];
      if (mode == 2) r_o = ub[i - j];
      d = (((hp_n - hi_n) * r_n) * sc - ((hp_o - hi_o) * r_o) * sc) * ick[k];  // equilibrated coordinates, as in K1
    }
    sd[q] = d;
    const double a = fabs(d);
    bad |= !(a <= 1.7976931348623157e308);
    amax = fmax(amax, a);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    bad |= __shfl_xor_sync(0xffffffffu, bad, off);
  }
  if ((tid & 31) == 0) { s_amax[tid >> 5] = amax; s_bad[tid >> 5] = bad; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) { amax = fmax(amax, s_amax[w]); bad |= s_bad[w]; }
  const bool finite = bad == 0;
  double mult;
  const double scl = oz_row_scale<S>(amax, finite, &mult);
  if (tid == 0) scale[chain] = scl;
  signed char* prow = planes + (size_t)chain * ld8 + k_base;
  for (int q4 = tid; q4 < n_k / 4; q4 += 256) {
    const double2 x01 = *reinterpret_cast<const double2*>(sd + 4 * q4);
    const double2 x23 = *reinterpret_cast<const double2*>(sd + 4 * q4 + 2);
    const unsigned long long j0 = oz_digit_bytes<S>(finite ? x01.x : 0.0, mult), j1 = oz_digit_bytes<S>(finite ? x01.y : 0.0, mult),
                             j2 = oz_digit_bytes<S>(finite ? x23.x : 0.0, mult), j3 = oz_digit_bytes<S>(finite ? x23.y : 0.0, mult);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int byte = S - 1 - s;
      const uint32_t a0 = byte < 4 ? (uint32_t)j0 : (uint32_t)(j0 >> 32), a1 = byte < 4 ? (uint32_t)j1 : (uint32_t)(j1 >> 32),
                     a2 = byte < 4 ? (uint32_t)j2 : (uint32_t)(j2 >> 32), a3 = byte < 4 ? (uint32_t)j3 : (uint32_t)(j3 >> 32);
      const uint32_t sel = (uint32_t)(byte & 3) | ((uint32_t)(4 + (byte & 3)) << 4);
      const uint32_t lo = __byte_perm(a0, a1, sel), hi = __byte_perm(a2, a3, sel);
      *reinterpret_cast<uint32_t*>(prow + (size_t)s * plane_stride + 4 * q4) = __byte_perm(lo, hi, 0x5410);
    }
  }
}

// ------------------------------------------------------------------------------ FP64 fall-back for flagged chains
// y[chain][m] = sum_k M[m][k] dx_k in plain FP64 for the chains whose residual range is too wide (majority of K1's per-warp
// counts, see residual_split_kernel; the decision is published in wide[chain] for the status word); every other CTA exits at once.
// M = P (symmetric product) or U = L^T (upper triangular, value-only path: k >= m).  One CTA per chain: residuals into shared
// memory (K1's arithmetic, not equilibrated), one warp per row with coalesced 256-byte reads of the row, fixed shuffle tree.
__global__ void __launch_bounds__(256)
fp64_rows_kernel(int N, int K, int SL, int root_r, const int* __restrict__ parent, const double* __restrict__ mu,
                 const double* __restrict__ states, const double* __restrict__ Mx, int ldm, int upper_tri,
                 const int* __restrict__ widecnt, int n_pad, int* __restrict__ wide, double* __restrict__ Y, int ldy, int B) {
  extern __shared__ __align__(16) unsigned char smem_fb[];
  double* sdx = reinterpret_cast<double*>(smem_fb);  // [K]
  const int chain = blockIdx.x;
  if (chain >= B) return;
  const int tid = threadIdx.x;
  const int4 c0 = *reinterpret_cast<const int4*>(widecnt + chain * 8), c1 = *reinterpret_cast<const int4*>(widecnt + chain * 8 + 4);
  const int n_small = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;   // includes the n_pad zero entries beyond K, if any counted
  const bool flagged = n_small > 0 && 2 * (n_small - n_pad) >= K;
  if (tid == 0) wide[chain] = flagged ? 1 : 0;
  if (!flagged) return;
  const double* x = states + (size_t)chain * SL;
  const double* h = x + 3;
  const double* r = x + 5 + N;
  const double sc = x[2] * x[3 + N];
  for (int i = 1 + tid; i < N; i += 256) {
    if (i == root_r) continue;
    double e = (h[parent[i] & 0x7fffffff] - h[i]) * r[i];
    if (i == 1) e = e + (h[0] - h[root_r]) * r[root_r];
    const int k = i < root_r ? i - 1 : i - 2;
    sdx[k] = e * sc - mu[k];
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  double* y = Y + (size_t)chain * ldy;
  for (int m = warp; m < K; m += 8) {
    const double* row = Mx + (size_t)m * ldm;
    double a0 = 0.0, a1 = 0.0;
    int k = (upper_tri ? (m & ~31) : 0) + lane;
    for (; k + 32 < K; k += 64) {
      a0 = fma(row[k], sdx[k], a0);
      a1 = fma(row[k + 32], sdx[k + 32], a1);
    }
    if (k < K) a0 = fma(row[k], sdx[k], a0);
    double a = a0 + a1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) y[m] = a;
  }
}

// ------------------------------------------------------------------------------ host side
// tensor map over digit planes stacked along rows: [total_rows][ld8] int8, box = box_rows x OZ_KB bytes, swizzled over OZ_KB
inline int oz_make_plane_map(CUtensorMap* tm, const signed char* base, size_t total_rows, int ld8, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)ld8, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld8};
  cuuint32_t box[2] = {OZ_KB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, OZ_KB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// MCD_OZ_V1 in the environment: the stage-granular pipeline of round 1 (A/B timing; results are bit-identical)
inline bool oz_use_v1() {
  static const bool v1 = getenv("MCD_OZ_V1") != nullptr;
  return v1;
}
template <int S>
inline cudaError_t gemm_i8_ozaki_configure() {
  cudaError_t e = cudaFuncSetAttribute(gemm_i8_ozaki_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem_bytes<S>());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_i8_ozaki_v2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem_bytes<S>());
}

inline int oz_pair_mode() {   // MCD_OZ_PAIR=1: the cta_group::2 kernel (experiment switch until it is the default)
  static const int m = getenv("MCD_OZ_PAIR") ? atoi(getenv("MCD_OZ_PAIR")) : 0;
  return m;
}
template <int S>
inline cudaError_t gemm_i8_ozaki_pair_launch(const CUtensorMap& tmA, const CUtensorMap& tmBh, const double* scaleA,
                                             const double* scaleB, double* Y, int Mp, int n_chains_padded256, int ld8, int ldy,
                                             int Bp_total, cudaStream_t st, int bt_base, int n_sms, int upper_tri, int kb_lo,
                                             int kb_hi, const double* yadd, int ldyadd) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_i8_ozaki_pair_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz2_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int n_pr = Mp / OZ_N, n_tiles = n_pr * (n_chains_padded256 / (2 * OZ_M));
  const int max_pairs = n_sms / 2;
  const int pairs = n_tiles < max_pairs ? n_tiles : max_pairs;
  const int nkb = kb_hi > 0 ? kb_hi : ld8 / OZ_KB;
  gemm_i8_ozaki_pair_kernel<S><<<2 * pairs, OZ_THREADS, oz2_smem_bytes<S>(), st>>>(tmA, tmBh, scaleA, scaleB, Y, nkb, ldy, Bp_total, Mp,
                                                                                    bt_base, n_pr, n_tiles, upper_tri, kb_lo, yadd, ldyadd);
  return cudaGetLastError();
}

// n_chains_padded chains (multiple of 128) starting at bt_base, Mp P rows (multiple of 64), ld8 = padded K
// (multiple of OZ_KB); one persistent CTA per SM
template <int S>
inline cudaError_t gemm_i8_ozaki_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const double* scaleA,
                                        const double* scaleB, double* Y, int Mp, int n_chains_padded, int ld8, int ldy,
                                        int Bp_total, cudaStream_t st, int bt_base = 0, int n_sms = 148, int upper_tri = 0,
                                        int kb_lo = 0, int kb_hi = -1, const double* yadd = nullptr, int ldyadd = 0) {
  const int n_pr = Mp / OZ_N, n_tiles = n_pr * (n_chains_padded / OZ_M);
  const int grid = n_tiles < n_sms ? n_tiles : n_sms;
  const int nkb = kb_hi > 0 ? kb_hi : ld8 / OZ_KB;  // [kb_lo, kb_hi): non-empty by the caller's contract
  if (oz_use_v1())
    gemm_i8_ozaki_kernel<S><<<grid, OZ_THREADS, oz_smem_bytes<S>(), st>>>(tmA, tmB, scaleA, scaleB, Y, nkb, ldy, Bp_total, Mp,
                                                                           bt_base, n_pr, n_tiles, upper_tri, kb_lo, yadd, ldyadd);
  else
    gemm_i8_ozaki_v2_kernel<S><<<grid, OZ_THREADS, oz_smem_bytes<S>(), st>>>(tmA, tmB, scaleA, scaleB, Y, nkb, ldy, Bp_total, Mp,
                                                                              bt_base, n_pr, n_tiles, upper_tri, kb_lo, yadd, ldyadd);
  return cudaGetLastError();
}

}  // namespace mcd
