// gemm_f64.cuh -- FP64 contraction of the precision matrix with the batched residuals.
//
//   Y[b][m] = sum_k P[m][k] * DX[b][k]          b = chain, m/k = branch (MVN dimension)
//
// This is the batched form of the reference's  (dxs <# sigmaInv)  at
// app/Probability.hs:169 (one BLAS gemv per likelihood call there; one GEMM per batch here).
// The same Y serves the value (quad = dx . y) and the gradient (-y), see DESIGN.md.
//
// Layout: every operand is "k-fastest" (chain-major rows), so both MMA operands are read
// exactly as they lie in HBM and no transposes are needed:
//   P  [Mp][ldk]  row-major; symmetric, zero padded to Mp rows (mult. of 128) and ldk cols
//   DX [Bp][ldk]  one residual vector per chain, zero padded to ldk (mult. of 16)
//   Y  [Bp][ldy]  ldy >= Mp
//
// Math: warp-level FP64 tensor instructions  mma.sync.aligned.m8n8k4.f64  (SASS DMMA.8x8x4 --
// the only FP64 tensor shape sm_100a executes natively; tcgen05 has no FP64 kind).  MMA-M (8)
// runs along chains, MMA-N (8) along P rows.  The reduction index inside one k16 block is
// permuted (k-step ks of lane-quad q uses k = 4q+ks) so that every fragment is fetched from
// shared memory with 128-bit loads.
//
// Data movement: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 6-stage shared-memory ring,
// completion on mbarriers; consumers never hit a CTA-wide barrier in the main loop.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcd {

constexpr int GEMM_BT = 128;       // chains per CTA tile
constexpr int GEMM_PR = 128;       // P rows per CTA tile
constexpr int GEMM_BK = 16;        // reduction depth per stage (16 doubles = one 128-byte row)
constexpr int GEMM_STAGES = 6;
constexpr int GEMM_PREFETCH = 4;   // tiles in flight ahead of the consumer (< STAGES: WAR slack)
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGE_BYTES = (GEMM_BT + GEMM_PR) * GEMM_BK * 8;   // 32 KiB
constexpr size_t GEMM_SMEM_BYTES = (size_t)GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*mbarriers*/;

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
          "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// D(8x8) += A(8x4) * B(4x8), FP64.  g = lane/4, q = lane%4:
//   a: row g, k q;   b: k q, col g;   c[0..1]: row g, cols 2q, 2q+1
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// One CTA computes a 128(chains) x 128(P rows) tile of Y.  Grid: x = P-row tile (fastest, so
// CTAs that run together share the DX tile and sweep P, which stays L2-resident), y = chain tile.
// 8 warps as 2 (chains, 64 each) x 4 (P rows, 32 each); each warp: 8x4 DMMA tiles, 64 accumulators
// per thread.  Thread 0 doubles as the TMA producer.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f64_dmma_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmX,
                     double* __restrict__ Y, int nk, int ldy, int bt_base, int upper_tri) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)GEMM_STAGES * GEMM_STAGE_BYTES);
  uint64_t* empty = full + GEMM_STAGES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int wbt = (warp & 1) * 64;
  const int wpr = (warp >> 1) * 32;
  const int pr0 = blockIdx.x * GEMM_PR;
  const int bt0 = bt_base + blockIdx.y * GEMM_BT;
  // upper-triangular operand (U = L^T of the Cholesky factor, value-only path): row m only has entries at
  // k >= m, so this tile's reduction starts at its own diagonal block
  const int kt0 = upper_tri ? pr0 / GEMM_BK : 0;
  nk -= kt0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GEMM_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int t) {  // producer: tile t -> stage t % STAGES
    const int s = t % GEMM_STAGES;
    if (t >= GEMM_STAGES) mbar_wait(&empty[s], ((t / GEMM_STAGES) - 1) & 1);
    unsigned char* dst = smem + (size_t)s * GEMM_STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], GEMM_STAGE_BYTES);
    tma_load_2d(dst, &tmX, (kt0 + t) * GEMM_BK, bt0, &full[s]);
    tma_load_2d(dst + GEMM_BT * GEMM_BK * 8, &tmP, (kt0 + t) * GEMM_BK, pr0, &full[s]);
  };
  if (tid == 0) {
    for (int t = 0; t < GEMM_PREFETCH && t < nk; ++t) issue(t);
  }

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // swizzled 16-byte chunk offsets of this lane's two chunks (k = 4q..4q+3); row & 7 == g
  const int off0 = ((2 * q) ^ g) << 4, off1 = ((2 * q + 1) ^ g) << 4;

  for (int kt = 0; kt < nk; ++kt) {
    if (tid == 0 && kt + GEMM_PREFETCH < nk) issue(kt + GEMM_PREFETCH);
    const int s = kt % GEMM_STAGES;
    mbar_wait(&full[s], (kt / GEMM_STAGES) & 1);
    const unsigned char* sA = smem + (size_t)s * GEMM_STAGE_BYTES;   // DX tile [128][16]
    const unsigned char* sB = sA + GEMM_BT * GEMM_BK * 8;            // P  tile [128][16]

    double bf[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned char* r = sB + (wpr + j * 8 + g) * 128;
      double2 v0 = *reinterpret_cast<const double2*>(r + off0);
      double2 v1 = *reinterpret_cast<const double2*>(r + off1);
      bf[j][0] = v0.x; bf[j][1] = v0.y; bf[j][2] = v1.x; bf[j][3] = v1.y;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double af[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned char* r = sA + (wbt + (half * 4 + i) * 8 + g) * 128;
        double2 v0 = *reinterpret_cast<const double2*>(r + off0);
        double2 v1 = *reinterpret_cast<const double2*>(r + off1);
        af[i][0] = v0.x; af[i][1] = v0.y; af[i][2] = v1.x; af[i][3] = v1.y;
      }
      // k-step outermost: 16 independent accumulators between dependent DMMAs
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[half * 4 + i][j], af[i][ks], bf[j][ks]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // epilogue: c[0..1] -> chain g, P rows 2q, 2q+1: 16-byte stores
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = bt0 + wbt + i * 8 + g;
      const int m = pr0 + wpr + j * 8 + 2 * q;
      *reinterpret_cast<double2*>(Y + (size_t)b * ldy + m) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// tensor map over a row-major [rows][ldk] FP64 matrix, box = 128 rows x 16 doubles, 128B swizzle
inline int make_tile_map(CUtensorMap* tm, const double* base, int rows, int ldk) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldk * 8};
  cuuint32_t box[2] = {GEMM_BK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

inline cudaError_t gemm_f64_dmma_configure() {
  return cudaFuncSetAttribute(gemm_f64_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)GEMM_SMEM_BYTES);
}

// Mp = P rows (mult of 128), ldk = padded K (mult of 16); chains [bt_base, bt_base + Bp), both mult of 128
inline cudaError_t gemm_f64_dmma_launch(const CUtensorMap& tmP, const CUtensorMap& tmX, double* Y, int Mp, int Bp,
                                        int ldk, int ldy, cudaStream_t st, int bt_base = 0, int upper_tri = 0) {
  dim3 grid(Mp / GEMM_PR, Bp / GEMM_BT);
  gemm_f64_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(tmP, tmX, Y, ldk / GEMM_BK, ldy, bt_base, upper_tri);
  return cudaGetLastError();
}

}  // namespace mcd
