// Dense FP64 kernels for sm_100a: DMMA (mma.sync m8n8k4 f64 -- the only FP64 tensor-core path on
// Blackwell; tcgen05 has no f64 kind) GEMM with a 4-stage cp.async pipeline, a register-resident
// 128x128 Cholesky/inverse leaf, and the recursive drivers built from them.
//
// Replaces, on the device: sytrf/sytri/sytrs (lapack.F90:165-363, called from model/type.F90:221-243)
// and gemm/gemv/symv (blas.F90, called from model/type.F90:235-290).
#include <cstdio>
#include <cooperative_groups.h>

#include "dense.cuh"
#include "erfgauss.cuh"
#include "diag16.cuh"

#include <algorithm>
#include <cstdlib>

namespace mchrg {

// ------------------------------------------------------------------------------------------------
// DMMA GEMM:  C = alpha * A * op(B) + beta * C
// CTA tile 128 x 128 x 16, 512 threads = 16 warps in a 4 (M) x 4 (N) grid, warp tile 32 x 32
// = 4 x 4 mma tiles of 8 x 8.  Shared memory holds k-major operand slabs padded so that the
// per-lane fragment loads (8 rows x 4 k) are bank-conflict free:
//   As[k][m], leading dimension 132 doubles (132 mod 16 == 4)
//   Bs[k][n] (transb: B given as n x k)      leading dimension 132
//   Bs[n][k] (no trans: B given as k x n)    leading dimension 20  (20 mod 16 == 4)
// ------------------------------------------------------------------------------------------------
namespace gemm {
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDAS = BM + 4;
constexpr int LDBN = BK + 4;
constexpr int A_STAGE = BK * LDAS;                                             // 2112 doubles
// CTA tile 128 x BNT (BNT = 128: 16 warps, one CTA per SM; BNT = 64: 8 warps, two CTAs per SM, so that the
// C read-modify-write and the pipeline fill of one CTA hide behind the DMMA loop of the other)
template <int BNT>
struct Cfg {
   static constexpr int LDBT = BNT + 4;
   static constexpr int B_STAGE = (BK * LDBT > BNT * LDBN) ? BK * LDBT : BNT * LDBN;
   static constexpr int STAGE = A_STAGE + B_STAGE;
   static constexpr int NTHREADS = 4 * BNT;  // 4 x (BNT / 32) warps of 32 x 32
   static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE * sizeof(double);  // 149504 B / 108544 B
};
}  // namespace gemm

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
   asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

template <bool TRANSB, int BNT>
__global__ void __launch_bounds__(gemm::Cfg<BNT>::NTHREADS, 128 / BNT)
dgemm_dmma_kernel(int64_t kdim, double alpha, const double* __restrict__ A, int64_t lda,
                  const double* __restrict__ B, int64_t ldb, double beta, double* C, int64_t ldc,
                  int lower_only, int4 cyc) {
   using namespace gemm;
   constexpr int LDBT = Cfg<BNT>::LDBT, STAGE = Cfg<BNT>::STAGE, NTHR = Cfg<BNT>::NTHREADS;
   constexpr int NSPLIT = BN / BNT;  // column slices of a 128 x 128 tile (one CTA each)
   extern __shared__ __align__(16) double smem[];

   const int bx = (lower_only ? (int)blockIdx.x / NSPLIT : (int)blockIdx.x);
   const int slice = lower_only ? (int)blockIdx.x % NSPLIT : 0;
   int tm, tn;
   if (lower_only == 2) {
      // block-cyclic trailing update of the multi-GPU factorisation: the owned block columns are cyc.w tiles wide,
      // start at tile column cyc.x and repeat every cyc.y tile columns; cyc.z = tile rows of the matrix.
      // Lower tiles (tm >= tn) only, numbered column by column.
      int b = bx, t0 = cyc.x;
      tm = tn = 0;
      for (bool found = false; !found; t0 += cyc.y) {
         for (int w = 0; w < cyc.w && t0 + w < cyc.z; ++w) {
            const int cnt = cyc.z - t0 - w;  // lower tiles of tile column t0 + w
            if (b < cnt) { tn = t0 + w; tm = tn + b; found = true; break; }
            b -= cnt;
         }
      }
   } else if (lower_only) {  // linear index over the lower block triangle
      const int b = bx;
      int r = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
      while ((int64_t)(r + 1) * (r + 2) / 2 <= b) ++r;
      while ((int64_t)r * (r + 1) / 2 > b) --r;
      tm = r;
      tn = b - r * (r + 1) / 2;
   } else {
      tm = blockIdx.x;
      tn = blockIdx.y;  // in units of BNT columns
   }
   const int64_t m0 = (int64_t)tm * BM, n0 = lower_only ? (int64_t)tn * BN + slice * BNT : (int64_t)tn * BNT;
   const int tid = threadIdx.x;
   const int lane = tid & 31, warp = tid >> 5;
   const int wm = warp & 3, wn = warp >> 2;
   const int g = lane >> 2, t = lane & 3;

   const double* Ag = A + m0;
   const double* Bg = TRANSB ? (B + n0) : (B + n0 * ldb);

   auto load_stage = [&](int stage, int64_t k0) {
      double* As = smem + (size_t)stage * STAGE;
      double* Bs = As + A_STAGE;
#pragma unroll
      for (int i = 0; i < BK * (BM / 2) / NTHR; ++i) {  // A: 16 k-columns x 64 16-byte chunks
         const int c = tid + i * NTHR;
         const int k = c >> 6, m2 = (c & 63) * 2;
         cp_async16(As + k * LDAS + m2, Ag + (k0 + k) * lda + m2);
      }
      if (TRANSB) {
#pragma unroll
         for (int i = 0; i < BK * (BNT / 2) / NTHR; ++i) {  // B: 16 k-columns x BNT / 2 chunks
            const int c = tid + i * NTHR;
            const int k = c / (BNT / 2), n2 = (c % (BNT / 2)) * 2;
            cp_async16(Bs + k * LDBT + n2, Bg + (k0 + k) * ldb + n2);
         }
      } else {
#pragma unroll
         for (int i = 0; i < BNT * (BK / 2) / NTHR; ++i) {  // B: BNT n-columns x 8 chunks along k
            const int c = tid + i * NTHR;
            const int n = c >> 3, k2 = (c & 7) * 2;
            cp_async16(Bs + n * LDBN + k2, Bg + (int64_t)n * ldb + k0 + k2);
         }
      }
   };

   double acc[4][4][2];
#pragma unroll
   for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

   // Two super-stages of 2 x BK = 32 k-columns: one barrier and one cp.async group per 32 k-columns.
   // While pair p is multiplied (8 k4-steps x 16 DMMA per warp, ~8k cycles per SM sub-partition) the
   // loads of pair p+1 are in flight.
   const int np2 = (int)(kdim / (2 * BK));
   load_stage(0, 0);
   load_stage(1, BK);
   cp_async_commit();
   for (int pr = 0; pr < np2; ++pr) {
      cp_async_wait<0>();
      __syncthreads();
      if (pr + 1 < np2) {
         const int sb = ((pr + 1) & 1) * 2;
         load_stage(sb, (int64_t)(pr + 1) * 2 * BK);
         load_stage(sb + 1, (int64_t)(pr + 1) * 2 * BK + BK);
         cp_async_commit();
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
         const double* As = smem + (size_t)((pr & 1) * 2 + half) * STAGE;
         const double* Bs = As + A_STAGE;
#pragma unroll
         for (int kk = 0; kk < BK / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = As[(kk * 4 + t) * LDAS + wm * 32 + mt * 8 + g];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
               b[nt] = TRANSB ? Bs[(kk * 4 + t) * LDBT + wn * 32 + nt * 8 + g]
                              : Bs[(wn * 32 + nt * 8 + g) * LDBN + kk * 4 + t];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
               for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
         }
      }
   }
   cp_async_wait<0>();

   // epilogue: lane holds C[row g][cols 2t, 2t+1] of each 8x8 tile
#pragma unroll
   for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
         const int64_t row = m0 + wm * 32 + mt * 8 + g;
         const int64_t col = n0 + wn * 32 + nt * 8 + 2 * t;
         double* c0 = C + row + col * ldc;
         double* c1 = c0 + ldc;
         double v0 = alpha * acc[mt][nt][0], v1 = alpha * acc[mt][nt][1];
         if (beta != 0.0) {
            v0 += beta * (*c0);
            v1 += beta * (*c1);
         }
         *c0 = v0;
         *c1 = v1;
      }
}

// MCHRG_B200_GEMM_BN = 64 | 128 (A/B timing); default chosen on B200
static int gemm_bn() { return env_ll("MCHRG_B200_GEMM_BN", 64) == 128 ? 128 : 64; }

static void gemm_setup(mchrg_b200_context* ctx) {
   ctx->ensure_smem(dgemm_dmma_kernel<true, 128>, gemm::Cfg<128>::SMEM_BYTES);
   ctx->ensure_smem(dgemm_dmma_kernel<false, 128>, gemm::Cfg<128>::SMEM_BYTES);
   ctx->ensure_smem(dgemm_dmma_kernel<true, 64>, gemm::Cfg<64>::SMEM_BYTES);
   ctx->ensure_smem(dgemm_dmma_kernel<false, 64>, gemm::Cfg<64>::SMEM_BYTES);
}

template <int BNT>
static void gemm_launch(mchrg_b200_context* ctx, bool transb, dim3 grid, int64_t k, double alpha, const double* A,
                        int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int mode, int4 cyc) {
   using cfg = gemm::Cfg<BNT>;
   if (transb)
      MCHRG_LAUNCH(ctx, (dgemm_dmma_kernel<true, BNT>), grid, cfg::NTHREADS, cfg::SMEM_BYTES, k, alpha, A, lda, B, ldb, beta,
                   C, ldc, mode, cyc);
   else
      MCHRG_LAUNCH(ctx, (dgemm_dmma_kernel<false, BNT>), grid, cfg::NTHREADS, cfg::SMEM_BYTES, k, alpha, A, lda, B, ldb, beta,
                   C, ldc, mode, cyc);
}

void dgemm_padded(mchrg_b200_context* ctx, bool transb, int64_t m, int64_t n, int64_t k, double alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                  int64_t ldc, bool lower_only) {
   if (m == 0 || n == 0) return;
   if (m % gemm::BM || n % gemm::BN || k % (2 * gemm::BK) || (lda & 1) || (ldb & 1))
      fail(MCHRG_B200_ERR_ARG, "dgemm_padded: unpadded operand (m=%lld n=%lld k=%lld lda=%lld ldb=%lld)",
           (long long)m, (long long)n, (long long)k, (long long)lda, (long long)ldb);
   gemm_setup(ctx);
   // in-place products (X <- X inv(L)^T: every CTA reads the whole row block of X) need one CTA per row block
   const int bnt = (A == C || B == C) ? 128 : gemm_bn();
   const unsigned nsplit = (unsigned)(gemm::BN / bnt);
   dim3 grid;
   if (lower_only) {
      const int64_t tiles = m / gemm::BM;
      grid = dim3((unsigned)(tiles * (tiles + 1) / 2) * nsplit, 1, 1);
   } else {
      grid = dim3((unsigned)(m / gemm::BM), (unsigned)(n / bnt), 1);
   }
   const int4 none = make_int4(0, 0, 0, 0);
   if (bnt == 64) gemm_launch<64>(ctx, transb, grid, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0, none);
   else gemm_launch<128>(ctx, transb, grid, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0, none);
}

// W(:, owned block columns) -= P P^T restricted to the lower tiles, P = W(:, c0 : c0 + k) (n x k panel, global
// rows): one launch for all block columns (two tiles wide) starting at tile column t0 and repeating every
// `stride` tile columns.  Multi-GPU factorisation only.
static void dsyrk_cyclic(mchrg_b200_context* ctx, int64_t n, int64_t k, const double* P, double* W, int64_t ldw, int t0,
                         int stride, int wt) {
   const int T = (int)(n / gemm::BM);
   int64_t tiles = 0;
   for (int t = t0; t < T; t += stride)
      for (int w = 0; w < wt && t + w < T; ++w) tiles += T - t - w;
   if (tiles == 0) return;
   gemm_setup(ctx);
   const int4 cyc = make_int4(t0, stride, T, wt);
   if (gemm_bn() == 64) gemm_launch<64>(ctx, true, dim3((unsigned)tiles * 2), k, -1.0, P, ldw, P, ldw, 1.0, W, ldw, 2, cyc);
   else gemm_launch<128>(ctx, true, dim3((unsigned)tiles), k, -1.0, P, ldw, P, ldw, 1.0, W, ldw, 2, cyc);
}

// ------------------------------------------------------------------------------------------------
// Small-tile DMMA product for the critical path of the distributed factorisation:
//    C (m x 128) = alpha * A (m x k) * B^T (B: 128 x k) + beta * C ,   m multiple of 16, k multiple of 8
// One CTA of 8 warps per 16 rows, one 16 x 16 tile per warp, fragments straight from global memory (L2): a
// 128 x 128 x 128 product runs on 8 SMs in ~2 us, where the 128 x 128 CTA tile of dgemm_dmma_kernel takes one SM
// ~20 us -- the two 128-wide products between consecutive leaf factorisations (row solve of the next diagonal
// block's rows, update of the next diagonal block) are latency, not throughput.
// C may alias A (the whole 16-row strip of A is read before anything is written).
// ------------------------------------------------------------------------------------------------
// rows [m0, m0 + 16) of C by the 256 threads of one CTA
__device__ __forceinline__ void small128_strip(int64_t m0, int k, double alpha, const double* A, int64_t lda, const double* B,
                                               int64_t ldb, double beta, double* C, int64_t ldc) {
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int g = lane >> 2, t = lane & 3;
   const int n0 = warp * 16;
   const double* a0p = A + m0 + g + (int64_t)t * lda;   // A(m0 + g [+ 8], kk + t)
   const double* b0p = B + n0 + g + (int64_t)t * ldb;   // B(n0 + g [+ 8], kk + t)
   double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 1
   for (int kk = 0; kk < k; kk += 32) {  // 8 k-steps of 4 per batch: 32 loads in flight per lane
      double a[8][2], b[8][2];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
         const bool on = kk + 4 * u < k;
         const int64_t ka = (int64_t)(kk + 4 * u) * lda, kb = (int64_t)(kk + 4 * u) * ldb;
         a[u][0] = on ? a0p[ka] : 0.0;
         a[u][1] = on ? a0p[ka + 8] : 0.0;
         b[u][0] = on ? b0p[kb] : 0.0;
         b[u][1] = on ? b0p[kb + 8] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
         dmma884(acc[0][0][0], acc[0][0][1], a[u][0], b[u][0]);
         dmma884(acc[0][1][0], acc[0][1][1], a[u][0], b[u][1]);
         dmma884(acc[1][0][0], acc[1][0][1], a[u][1], b[u][0]);
         dmma884(acc[1][1][0], acc[1][1][1], a[u][1], b[u][1]);
      }
   }
   double cold[2][2][2];
#pragma unroll
   for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
         const double* c = C + m0 + 8 * mt + g + (int64_t)(n0 + 8 * nt + 2 * t) * ldc;
         cold[mt][nt][0] = (beta != 0.0) ? c[0] : 0.0;
         cold[mt][nt][1] = (beta != 0.0) ? c[ldc] : 0.0;
      }
   __syncthreads();  // in-place products: every warp has read its strip of A
#pragma unroll
   for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
         double* c = C + m0 + 8 * mt + g + (int64_t)(n0 + 8 * nt + 2 * t) * ldc;
         c[0] = alpha * acc[mt][nt][0] + beta * cold[mt][nt][0];
         c[ldc] = alpha * acc[mt][nt][1] + beta * cold[mt][nt][1];
      }
}

__global__ void __launch_bounds__(256) dgemm_small128_kernel(int k, double alpha, const double* A, int64_t lda,
                                                             const double* B, int64_t ldb, double beta, double* C,
                                                             int64_t ldc) {
   small128_strip((int64_t)blockIdx.x * 16, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

static void dgemm_small128(mchrg_b200_context* ctx, int64_t m, int64_t k, double alpha, const double* A, int64_t lda,
                           const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
   if (m % 16 || k % 8) fail(MCHRG_B200_ERR_ARG, "dgemm_small128: m=%lld k=%lld", (long long)m, (long long)k);
   MCHRG_LAUNCH(ctx, dgemm_small128_kernel, (unsigned)(m / 16), 256, 0, (int)k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// ------------------------------------------------------------------------------------------------
// Cholesky leaf: factor one 128 x 128 diagonal block and invert its factor in ONE CTA (8 warps), everything in
// shared memory as 16 x 16 tiles (row stride 20: conflict-free FP64 mma fragment loads in both orientations).
//   Cholesky, right-looking over 8 tile columns:
//     D  diagonal tile: Cholesky + explicit inverse by one warp (d16::diag_factor_inv, a chain of 16 pivot steps)
//     S  tiles below:   L(i,k) = A(i,k) inv(L_kk)^T                    one DMMA tile product per warp
//     U  trailing tiles A(i,j) -= L(i,k) L(j,k)^T  on DMMA; warp 0 takes the next diagonal tile first and factors
//        it while the other warps finish the update (look-ahead), so a tile column costs ~ one D chain
//   inverse by recursive doubling over the tile grid (h = 1, 2, 4 tiles): X21 = -X22 (L21 X11), two DMMA stages
//   per level; X21 replaces the (dead) L21 tiles, the diagonal tiles of X come from D.
// aug != 0: the last two rows are right-hand sides of an augmented system (their own columns stay identity, so
// the forward substitution falls out of the factorisation; see dpotrf_dist).
// ~18 us (the register-cyclic rank-one version it replaces: 78 us = 128 barrier-separated column steps twice).
// ------------------------------------------------------------------------------------------------
namespace leaf {
constexpr int N = 128, T16 = 16, NTL = N / T16, TLD = 20, TSZ = T16 * TLD;
constexpr int THREADS = 256, NW = THREADS / 32;
constexpr int NTRI = NTL * (NTL + 1) / 2;
// doubles: L tiles (lower block triangle) | diagonal tiles of the inverse | D scratch.  113 KB, so that the leaf can
// share an SM with one 128 x 64 GEMM CTA (108 KB) of the update streams instead of waiting for an empty SM.
constexpr size_t SMEM_BYTES = ((size_t)(NTRI + NTL) * TSZ + 2 * d16::CBW + 2 * T16 + T16) * sizeof(double);
__device__ __forceinline__ int tix(int R, int C) { return R * (R + 1) / 2 + C; }

// acc += A B^T for 16 x 16 tiles (row stride TLD): D[m][n] = sum_c A[m][c] B[n][c]
__device__ __forceinline__ void mma_nt(double (&acc)[2][2][2], const double* A, const double* B, int g, int t) {
#pragma unroll
   for (int kk = 0; kk < 4; ++kk) {
      const double a0 = A[g * TLD + 4 * kk + t], a1 = A[(8 + g) * TLD + 4 * kk + t];
      const double b0 = B[g * TLD + 4 * kk + t], b1 = B[(8 + g) * TLD + 4 * kk + t];
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
   }
}
// acc += A B: D[m][n] = sum_k A[m][k] B[k][n]
__device__ __forceinline__ void mma_nn(double (&acc)[2][2][2], const double* A, const double* B, int g, int t) {
#pragma unroll
   for (int kk = 0; kk < 4; ++kk) {
      const double a0 = A[g * TLD + 4 * kk + t], a1 = A[(8 + g) * TLD + 4 * kk + t];
      const double b0 = B[(4 * kk + t) * TLD + g], b1 = B[(4 * kk + t) * TLD + 8 + g];
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
   }
}
// tile = sign * acc  /  tile -= acc   (lane holds rows 8 mt + g, columns 8 nt + 2 t, + 1)
__device__ __forceinline__ void put(double* tile, const double (&acc)[2][2][2], double sign, int g, int t) {
#pragma unroll
   for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
         *reinterpret_cast<double2*>(tile + (8 * mt + g) * TLD + 8 * nt + 2 * t) =
            make_double2(sign * acc[mt][nt][0], sign * acc[mt][nt][1]);
}
__device__ __forceinline__ void sub(double* tile, const double (&acc)[2][2][2], int g, int t) {
#pragma unroll
   for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
         double2* d = reinterpret_cast<double2*>(tile + (8 * mt + g) * TLD + 8 * nt + 2 * t);
         double2 v = *d;
         v.x -= acc[mt][nt][0];
         v.y -= acc[mt][nt][1];
         *d = v;
      }
}
}  // namespace leaf

__device__ __forceinline__ void potrf_leaf_body(double* __restrict__ W, int64_t ldw, double* __restrict__ dinv,
                                                int* __restrict__ info, int pivot_base, int aug, double* smem) {
   using namespace leaf;
   double* Lt = smem;               // tile (R, C), R >= C, at tix(R, C) * TSZ; later the off-diagonal tiles of inv(L)
   double* Xd = Lt + NTRI * TSZ;    // diagonal tiles of inv(L)
   double* cbuf = Xd + NTL * TSZ;
   double* rbuf = cbuf + 2 * d16::CBW;
   double* dvec = rbuf + 2 * T16;
   __shared__ int sfail;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, t = lane & 3;
   if (tid == 0) sfail = 0;
#ifdef LEAF_PROF
   long long tk[8];
   tk[0] = clock64();
#define LEAF_TICK(i) tk[i] = clock64()
#else
#define LEAF_TICK(i)
#endif
   {  // thread = (row r, column parity h): consecutive threads walk down a column (coalesced); the loads of a batch
      // are issued back to back (32 in flight per thread) before the first one is consumed
      const int r = tid & (N - 1), h = tid >> 7, R = r >> 4;
#pragma unroll 1
      for (int cb = 0; cb < N; cb += 64) {
         double v[32];
#pragma unroll
         for (int u = 0; u < 32; ++u) {
            const int c = cb + h + 2 * u;
            v[u] = (r >= c) ? W[r + (int64_t)c * ldw] : 0.0;
         }
#pragma unroll
         for (int u = 0; u < 32; ++u) {
            const int c = cb + h + 2 * u, C = c >> 4;
            if (R >= C) Lt[tix(R, C) * TSZ + (r & 15) * TLD + (c & 15)] = v[u];
         }
      }
   }
   __syncthreads();
   LEAF_TICK(1);
   if (warp == 0) {
      const int bad = d16::diag_factor_inv<TLD, false, TLD>(Lt, Xd, cbuf, rbuf, dvec, lane, T16, true);
      if (bad && lane == 0) sfail = bad;
   }
   __syncthreads();
   LEAF_TICK(2);
   for (int k = 0; k + 1 < NTL && !sfail; ++k) {
      const int nrem = NTL - 1 - k;
      if (warp < nrem) {  // S: one tile below the diagonal per warp, in place
         double* A = Lt + tix(k + 1 + warp, k) * TSZ;
         double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
         mma_nt(acc, A, Xd + k * TSZ, g, t);
         __syncwarp();
         put(A, acc, 1.0, g, t);
      }
      __syncthreads();
      // U: tiles (k+1+a, k+1+b), 0 <= b <= a < nrem, numbered u = a (a + 1) / 2 + b; u = 0 is the next diagonal tile
      const int ntile = nrem * (nrem + 1) / 2;
      if (warp == 0) {
         double* A = Lt + tix(k + 1, k + 1) * TSZ;
         const double* P = Lt + tix(k + 1, k) * TSZ;
         double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
         mma_nt(acc, P, P, g, t);
         sub(A, acc, g, t);
         __syncwarp();
         const int nfree = (aug && k + 2 == NTL) ? T16 - 2 : T16;
         const int bad = d16::diag_factor_inv<TLD, false, TLD>(A, Xd + (k + 1) * TSZ, cbuf, rbuf, dvec, lane, nfree, true);
         if (bad && lane == 0) sfail = T16 * (k + 1) + bad;
      } else {
         for (int u = warp; u < ntile; u += NW - 1) {
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= u) ++a;
            const int b = u - a * (a + 1) / 2;
            double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            mma_nt(acc, Lt + tix(k + 1 + a, k) * TSZ, Lt + tix(k + 1 + b, k) * TSZ, g, t);
            sub(Lt + tix(k + 1 + a, k + 1 + b) * TSZ, acc, g, t);
         }
      }
      __syncthreads();
   }
   if (sfail) {  // not positive definite: report the pivot; the caller re-assembles the matrix
      if (tid == 0) atomicCAS(info, 0, pivot_base + sfail);
      for (int e = tid; e < N * N; e += THREADS) dinv[e] = 0.0;
      return;
   }
   LEAF_TICK(3);
   // publish L (lower triangle only)
   for (int e = tid; e < N * N; e += THREADS) {
      const int r = e & (N - 1), c = e >> 7;
      if (r >= c) W[r + (int64_t)c * ldw] = Lt[tix(r >> 4, c >> 4) * TSZ + (r & 15) * TLD + (c & 15)];
   }
   LEAF_TICK(4);
   // inverse: blocks of h tiles are merged pairwise, X21 = -X22 (L21 X11); tile (R, C) of X: Xd when R == C.
   // The products T = L21 X11 go to the eight DIAGONAL tiles of Lt (dead once L is published: the diagonal tiles of
   // the inverse live in Xd); the last level (16 product tiles) runs as two passes over column halves -- the first
   // pass reads all of L21 and overwrites only its first two tile columns, which the second pass no longer needs.
   __syncthreads();  // the diagonal tiles of Lt are read by the publishing loop above
   auto xt = [&](int R, int C) -> const double* { return (R == C) ? Xd + R * TSZ : Lt + tix(R, C) * TSZ; };
   auto ts = [&](int slot) -> double* { return Lt + tix(slot, slot) * TSZ; };
   for (int h = 1; h < NTL; h *= 2) {
      const int jw = (h < 2) ? h : 2, npass = h / jw;  // tile columns of X21 per pass
      const int per = h * jw, nitems = (NTL / (2 * h)) * per;  // (pair p, tile row i, tile column j) per pass
      for (int pass = 0; pass < npass; ++pass) {
         const int jb = pass * jw;
         for (int item = warp; item < nitems; item += NW) {
            const int p = item / per, i = (item % per) / jw, j = jb + item % jw, b0 = 2 * h * p;
            double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            for (int k = j; k < h; ++k) mma_nn(acc, Lt + tix(b0 + h + i, b0 + k) * TSZ, xt(b0 + k, b0 + j), g, t);
            put(ts(item), acc, 1.0, g, t);
         }
         __syncthreads();
         for (int item = warp; item < nitems; item += NW) {
            const int p = item / per, i = (item % per) / jw, jl = item % jw, b0 = 2 * h * p;
            double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            for (int k = 0; k <= i; ++k) mma_nn(acc, xt(b0 + h + i, b0 + h + k), ts(p * per + k * jw + jl), g, t);
            put(Lt + tix(b0 + h + i, b0 + jb + jl) * TSZ, acc, -1.0, g, t);
         }
         __syncthreads();
      }
   }
   LEAF_TICK(5);
   for (int e = tid; e < N * N; e += THREADS) {
      const int r = e & (N - 1), c = e >> 7;
      dinv[e] = (r >= c) ? xt(r >> 4, c >> 4)[(r & 15) * TLD + (c & 15)] : 0.0;
   }
#ifdef LEAF_PROF
   LEAF_TICK(6);
   if (tid == 0)
      printf("[leaf] load %lld first D %lld cholesky loop %lld store L %lld inverse %lld store inv %lld cycles\n", tk[1] - tk[0],
             tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5]);
#endif
}

__global__ void __launch_bounds__(leaf::THREADS, 1)
potrf_leaf_kernel(double* W, int64_t ldw, double* dinv, int* info, int pivot_base, int aug) {
   extern __shared__ __align__(16) double smem[];
   potrf_leaf_body(W, ldw, dinv, info, pivot_base, aug, smem);
}

// One step of the critical path of the tracked (multi-GPU) factorisation as ONE launch of a cluster of 8 CTAs,
// D = W(c0.., c0..) the diagonal tile of the sub-panel:
//   phase 0 (pre != nullptr)  D -= P P^T with P = L(rows of D, an arrived sub-panel): the last update the tile waits for
//   leaf                      CTA 0 factors D and inverts the factor; the others wait at the cluster barrier
//   phase 1 (do_head)         the tile below: X = W(c0 + 128 .., c0 ..) <- X inv(L)^T        (16 rows per CTA)
//   phase 2 (do_next)         the next diagonal tile: W(c0 + 128 .., c0 + 128 ..) -= X X^T  (16 rows per CTA)
// Three dependent launches with their stream-scheduling gaps (~10 us each next to 55 + 3 + 3 us of work) become one;
// cluster barriers order the phases (release / acquire at cluster scope covers the global-memory hand-over).
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(leaf::THREADS, 1)
potrf_step_kernel(double* W, int64_t ldw, double* dinv, int* info, int pivot_base, int aug, const double* pre, int do_head,
                  int do_next) {
   extern __shared__ __align__(16) double smem[];
   namespace cg = cooperative_groups;
   cg::cluster_group cluster = cg::this_cluster();
   const int64_t r0 = 16 * (int64_t)cluster.block_rank();
   if (pre != nullptr) {
      small128_strip(r0, leaf::N, -1.0, pre, ldw, pre, ldw, 1.0, W, ldw);
      cluster.sync();
   }
   if (cluster.block_rank() == 0) potrf_leaf_body(W, ldw, dinv, info, pivot_base, aug, smem);
   if (!do_head) return;
   cluster.sync();
   double* X = W + leaf::N;
   small128_strip(r0, leaf::N, 1.0, X, ldw, dinv, leaf::N, 0.0, X, ldw);
   if (!do_next) return;
   cluster.sync();
   small128_strip(r0, leaf::N, -1.0, X, ldw, X, ldw, 1.0, W + leaf::N + leaf::N * ldw, ldw);
}

// ------------------------------------------------------------------------------------------------
// recursive drivers
// ------------------------------------------------------------------------------------------------
static int64_t split(int64_t n) { return (n / TILE / 2) * TILE; }

void dtrsm_right_lt(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx) {
   if (m == 0 || n == 0) return;
   if (n == TILE) {  // X <- X * inv(L_kk)^T
      dgemm_padded(ctx, true, m, TILE, TILE, 1.0, X, ldx, dinv, TILE, 0.0, X, ldx);
      return;
   }
   const int64_t n1 = split(n), n2 = n - n1;
   dtrsm_right_lt(ctx, m, n1, L, ldl, dinv, X, ldx);
   dgemm_padded(ctx, true, m, n2, n1, -1.0, X, ldx, L + n1, ldl, 1.0, X + n1 * ldx, ldx);
   dtrsm_right_lt(ctx, m, n2, L + n1 + n1 * ldl, ldl, dinv + (n1 / TILE) * TILE * TILE, X + n1 * ldx, ldx);
}

void dtrsm_right_ln(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx) {
   if (m == 0 || n == 0) return;
   if (n == TILE) {  // X <- X * inv(L_kk)
      dgemm_padded(ctx, false, m, TILE, TILE, 1.0, X, ldx, dinv, TILE, 0.0, X, ldx);
      return;
   }
   const int64_t n1 = split(n), n2 = n - n1;
   dtrsm_right_ln(ctx, m, n2, L + n1 + n1 * ldl, ldl, dinv + (n1 / TILE) * TILE * TILE, X + n1 * ldx, ldx);
   dgemm_padded(ctx, false, m, n1, n2, -1.0, X + n1 * ldx, ldx, L + n1, ldl, 1.0, X, ldx);
   dtrsm_right_ln(ctx, m, n1, L, ldl, dinv, X, ldx);
}

// aug_end: order of the whole matrix when its last two rows are right-hand sides (augmented system), else -1;
// the leaf that ends there keeps those rows' own columns at identity
static void potrf_rec(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info,
                      int64_t off, int64_t aug_end) {
   if (n == TILE) {
      ctx->ensure_smem(potrf_leaf_kernel, leaf::SMEM_BYTES);
      MCHRG_LAUNCH(ctx, potrf_leaf_kernel, 1, leaf::THREADS, leaf::SMEM_BYTES, W, ldw, dinv, info, (int)off,
                   (off + TILE == aug_end) ? 1 : 0);
      return;
   }
   const int64_t n1 = split(n), n2 = n - n1;
   double* W21 = W + n1;
   double* W22 = W + n1 + n1 * ldw;
   double* dinv2 = dinv + (n1 / TILE) * TILE * TILE;
   potrf_rec(ctx, n1, W, ldw, dinv, info, off, aug_end);
   dtrsm_right_lt(ctx, n2, n1, W, ldw, dinv, W21, ldw);
   dgemm_padded(ctx, true, n2, n2, n1, -1.0, W21, ldw, W21, ldw, 1.0, W22, ldw, /*lower_only=*/true);
   potrf_rec(ctx, n2, W22, ldw, dinv2, info, off + n1, aug_end);
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU right-looking Cholesky, 1 x R process grid, block-cyclic by column groups:
//   group g = columns [g DB, (g + 1) DB), owner(g) = g % R;  a group is factored in sub-panels of PW columns
//   (PW | DB).  Only the owner keeps a group up to date and only the owner's copy is read on entry (the assembly
//   skips the other groups, dpotrf_dist_panel); factored sub-panels are broadcast (NCCL / peer copies, comm.cuh)
//   into place, so every rank ends with the complete factor: the solves and the row-sharded dq/dr sweeps that
//   follow read all of L and need no further exchange.
// Why 1 x R and not P x Q: on NVSwitch a broadcast costs the same for 2 or 8 receivers, and a 2-D grid puts
// three collectives (diagonal block down the column, panel along the row, transposed panel down the columns)
// on the per-panel critical path where this layout has one; at N = 16k..32k on 8 GPUs that path, not the
// broadcast volume, bounds the factorisation (DESIGN.md "Multi-GPU").
// Per rank, three streams:
//   P (high priority) the critical path: for the owner, sub-panel factorisation (leaf + row solve) and the update
//                     of the rest of its group; for the owner of the NEXT group, the update of that group with
//                     every sub-panel as it arrives (so only the last sub-panel's update trails the broadcast)
//   C (communication) one grouped broadcast per sub-panel (panel + its inverted diagonal blocks)
//   U (bulk)          per completed group ONE rank-DB update of all later owned groups (strided-tile GEMM), the
//                     first of them -- the next one this rank will have to factor -- as its own launch
// ------------------------------------------------------------------------------------------------
__global__ void info_to_slot_kernel(const int* __restrict__ info, int rank, int nranks, double* __restrict__ slots) {
   for (int r = 0; r < nranks; ++r) slots[r] = (r == rank) ? (double)*info : 0.0;
}
// smallest nonzero entry of the ranks' pivot reports
__global__ void info_combine_kernel(const double* __restrict__ slots, int nranks, int* __restrict__ info) {
   int v = 0;
   for (int r = 0; r < nranks; ++r) {
      const int x = (int)slots[r];
      if (x != 0 && (v == 0 || x < v)) v = x;
   }
   *info = v;
}

// Below ~8k the serial panel chain plus one broadcast per panel outlasts what the shared trailing updates save
bool dist_rows_apply(const mchrg_b200_context* ctx, int64_t nat) {
   return ctx->comm != nullptr && ctx->comm->nranks > 1 && nat >= env_ll("MCHRG_B200_DIST_MIN", 8192);
}
bool dist_pbc_apply(const mchrg_b200_context* ctx, int64_t nat) {
   return ctx->comm != nullptr && ctx->comm->nranks > 1 && nat >= env_ll("MCHRG_B200_DIST_MIN_PBC", 1024) &&
          ctx->active == ctx->stream;
}
void row_share(const mchrg_b200_context* ctx, int nat, int* row0, int* row1) {
   const int64_t R = ctx->nranks(), r = ctx->rank();
   auto bound = [&](int64_t k) { return (int)std::min<int64_t>(nat, round_up((nat * k + R - 1) / R, 8)); };
   *row0 = bound(r);
   *row1 = (r + 1 == R) ? nat : bound(r + 1);
}
void comm_allreduce_sum(mchrg_b200_context* ctx, double* buf, size_t count) {
   cudaEvent_t e = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e, ctx->active));
   MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_comm, e, 0));
   ctx->comm->allreduce_sum(buf, count, ctx->stream_comm);
   cudaEvent_t e2 = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e2, ctx->stream_comm));
   MCHRG_CUDA(cudaStreamWaitEvent(ctx->active, e2, 0));
}
static bool dist_applies(const mchrg_b200_context* ctx, int64_t n) {
   return dist_rows_apply(ctx, n) && n >= 4 * TILE && ctx->active == ctx->stream;
}
static int64_t dist_group_width(const mchrg_b200_context* ctx, int64_t n) {
   // MCHRG_B200_DIST_GROUP: columns per ownership group (multiple of the sub-panel width)
   int64_t db = env_ll("MCHRG_B200_DIST_GROUP", 512) / TILE * TILE;
   // every rank should own at least two groups
   while (db > TILE && n / db < 2 * (int64_t)ctx->nranks()) db -= TILE;
   return std::max<int64_t>(db, TILE);
}
int64_t dpotrf_dist_panel(const mchrg_b200_context* ctx, int64_t n) { return dist_applies(ctx, n) ? dist_group_width(ctx, n) : 0; }

static void dpotrf_dist(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info, int64_t aug_end) {
   Comm* cm = ctx->comm;  // nullptr: one GPU, the same schedule without the broadcasts
   const int R = cm ? cm->nranks : 1, me = cm ? cm->rank : 0;
   const int64_t DB = dist_group_width(ctx, n), PW = TILE;  // sub-panels are one 128 x 128 leaf wide
   const int ngrp = (int)((n + DB - 1) / DB);
   const int nsub = (int)(n / PW), spg = (int)(DB / PW);  // sub-panels in all / per group
   cudaStream_t U = ctx->active, P = ctx->stream_hi, Q = ctx->stream_hi2, C = ctx->stream_comm;
   auto g0_of = [&](int g) { return (int64_t)g * DB; };
   auto gw_of = [&](int g) { return std::min<int64_t>(DB, n - g0_of(g)); };
   auto owner = [&](int g) { return g % R; };
   auto rec = [&](cudaStream_t s) {
      cudaEvent_t e = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(e, s));
      return e;
   };
   auto wait = [&](cudaStream_t s, cudaEvent_t e) {
      if (e) MCHRG_CUDA(cudaStreamWaitEvent(s, e, 0));
   };
   // MCHRG_B200_DIST_TRACE=1: timed events around the operations, summed per category and printed by every rank after
   // the factorisation (synchronises: measurement runs only)
   const bool trace = env_ll("MCHRG_B200_DIST_TRACE", 0) > 0;
   struct Span { int cat; cudaEvent_t a, b; };
   std::vector<Span> spans;
   auto tic = [&](cudaStream_t s, int cat) {
      if (!trace) return;
      Span sp{cat, nullptr, nullptr};
      MCHRG_CUDA(cudaEventCreate(&sp.a));
      MCHRG_CUDA(cudaEventCreate(&sp.b));
      MCHRG_CUDA(cudaEventRecord(sp.a, s));
      spans.push_back(sp);
   };
   auto toc = [&](cudaStream_t s) {
      if (trace) MCHRG_CUDA(cudaEventRecord(spans.back().b, s));
   };
   // ---- write tracker.  Per sub-panel column j three parts change hands between the streams:
   //        PD  the diagonal tile (j, j)      PE  the tile below it (j+1, j)      PB  the rows from (j+2) * 128 on
   //      Every operation that WRITES parts goes through `write`, which orders its stream after the last writer of
   //      each part and records itself as the new one; READS of factored panel data wait for the events in
   //      headE / tailB / arrived explicitly.
   enum { PD = 1, PE = 2, PB = 4, PALL = 7 };
   struct Touch { cudaEvent_t ev; cudaStream_t st; };
   cudaEvent_t e_begin = rec(U);  // everything queued on U so far (the assembly)
   std::vector<Touch> last((size_t)3 * nsub, Touch{e_begin, U});
   wait(C, e_begin);
   struct Range { int j0, j1, parts; };
   auto write_multi = [&](cudaStream_t X, std::initializer_list<Range> ranges, int cat, auto&& launch) {
      cudaEvent_t seen = nullptr;
      for (const Range& rg : ranges)
         for (int j = rg.j0; j < rg.j1; ++j)
            for (int pt = 0; pt < 3; ++pt)
               if ((rg.parts >> pt) & 1) {
                  const Touch& tc = last[(size_t)3 * j + pt];
                  if (tc.st != X && tc.ev != seen) {
                     wait(X, tc.ev);
                     seen = tc.ev;
                  }
               }
      StreamScope on(ctx, X);
      tic(X, cat);
      launch();
      toc(X);
      cudaEvent_t e = rec(X);
      for (const Range& rg : ranges)
         for (int j = rg.j0; j < rg.j1; ++j)
            for (int pt = 0; pt < 3; ++pt)
               if ((rg.parts >> pt) & 1) last[(size_t)3 * j + pt] = Touch{e, X};
      return e;
   };
   auto write = [&](cudaStream_t X, int j0, int j1, int parts, int cat, auto&& launch) {
      return write_multi(X, {Range{j0, j1, parts}}, cat, launch);
   };
   // events after which the factored sub-panel s may be read on this rank: its tile (s+1, s) / its rows below / all of it
   std::vector<cudaEvent_t> headE((size_t)nsub, nullptr), tailB((size_t)nsub, nullptr);
   auto need_panel = [&](cudaStream_t X, int s, bool head, bool tail) {
      if (head) wait(X, headE[s]);
      if (tail) wait(X, tailB[s]);
   };
   // Eager (non-bulk) updates of the columns [t0, t1) with the factored sub-panel s.  The diagonal tile of column t0
   // -- the next leaf -- is updated by the small-tile product on P; everything else goes to Q.
   // The update of the diagonal tile of column t0 -- the next leaf -- is: diag_mode 0 a small-tile product on P now;
   // 1 already done by the step kernel of sub-panel s (fused); 2 deferred into the step kernel of t0 (its phase 0).
   std::vector<const double*> pre_ptr((size_t)nsub + 1, nullptr);
   std::vector<int> pre_src((size_t)nsub + 1, -1);
   auto eager = [&](int s, int t0, int t1, int cat, int diag_mode) {
      const int64_t c0 = (int64_t)s * PW;
      const int64_t r0 = (int64_t)t0 * PW;               // first row / column of the target range
      const double* Lt0 = W + r0 + c0 * ldw;             // L(rows of tile t0, panel s)
      const bool head_is_E = (t0 == s + 1);              // that tile is the panel's tile (s+1, s), else part of its tail
      if (diag_mode == 0) {
         need_panel(P, s, head_is_E, !head_is_E);
         write(P, t0, t0 + 1, PD, cat, [&]() { dgemm_small128(ctx, PW, PW, -1.0, Lt0, ldw, Lt0, ldw, 1.0, W + r0 + r0 * ldw, ldw); });
      } else if (diag_mode == 2) {
         pre_ptr[t0] = Lt0;
         pre_src[t0] = s;
      }
      const int64_t rows2 = n - r0 - PW;                 // rows below the diagonal tile of column t0
      if (rows2 > 0) {
         need_panel(Q, s, head_is_E, true);
         write(Q, t0, t0 + 1, PE | PB, cat, [&]() {
            dgemm_padded(ctx, true, rows2, PW, PW, -1.0, Lt0 + PW, ldw, Lt0, ldw, 1.0, W + r0 + PW + r0 * ldw, ldw, false);
         });
         if (t0 + 1 < t1) {
            const int64_t cb = (int64_t)(t0 + 1) * PW, ce = std::min<int64_t>(n, (int64_t)t1 * PW);
            const double* A = W + cb + c0 * ldw;
            write(Q, t0 + 1, t1, PALL, cat, [&]() {
               dgemm_padded(ctx, true, n - cb, ce - cb, PW, -1.0, A, ldw, A, ldw, 1.0, W + cb + cb * ldw, ldw, false);
            });
         }
      }
   };
   // MCHRG_B200_DIST_FUSE=1: the three launches of a step as one cluster launch (potrf_step_kernel).  Measured SLOWER and
   // therefore off: a cluster of 8 CTAs has to be placed at once, and with the update streams keeping the SMs busy that
   // wait costs more than the two launch gaps it saves (one GPU, n = 4096: 4.05 vs 3.34 ms; 8 GPUs, n = 16384: 25.8 vs
   // 21.0 ms).  Kept for the parity tests of the cluster path and as the documented negative result.
   const bool fuse = env_ll("MCHRG_B200_DIST_FUSE", 0) != 0;
   for (int g = 0; g < ngrp; ++g) {
      const int64_t g0 = g0_of(g), gw = gw_of(g);
      const bool mine = owner(g) == me;
      const bool next_mine = g + 1 < ngrp && owner(g + 1) == me;
      const int s0 = g * spg, s1 = std::min(nsub, s0 + spg);  // sub-panels of this group
      const int h0 = s1, h1 = std::min(nsub, s1 + spg);       // sub-panels of the next group
      for (int s = s0; s < s1; ++s) {
         const int64_t c0 = (int64_t)s * PW, rem = n - c0 - PW;
         double* D = W + c0 + c0 * ldw;
         double* dinv_k = dinv + (c0 / TILE) * TILE * TILE;
         const bool next_diag_mine = rem > 0 && owner((s + 1) / spg) == me;
         if (mine) {
            // critical path on P: leaf, then only the 128 rows the next leaf depends on, then the next diagonal tile
            // -- one cluster launch (potrf_step_kernel; MCHRG_B200_DIST_FUSE=0: three launches); the rows below on Q
            cudaEvent_t e_leaf;
            if (fuse) {
               const double* pre = pre_ptr[s];
               if (pre) need_panel(P, pre_src[s], true, true);
               ctx->ensure_smem(potrf_step_kernel, leaf::SMEM_BYTES);
               e_leaf = write_multi(P, {Range{s, s + 1, rem > 0 ? (PD | PE) : PD}, Range{s + 1, s + (next_diag_mine ? 2 : 1), PD}}, 0, [&]() {
                  MCHRG_LAUNCH(ctx, potrf_step_kernel, 8, leaf::THREADS, leaf::SMEM_BYTES, D, ldw, dinv_k, info, (int)c0,
                               (c0 + TILE == aug_end) ? 1 : 0, pre, rem > 0 ? 1 : 0, next_diag_mine ? 1 : 0);
               });
               headE[s] = tailB[s] = e_leaf;
            } else {
               e_leaf = write(P, s, s + 1, PD, 0, [&]() { potrf_rec(ctx, PW, D, ldw, dinv_k, info, c0, aug_end); });
               headE[s] = tailB[s] = e_leaf;
               if (rem > 0)
                  headE[s] = write(P, s, s + 1, PE, 1, [&]() { dgemm_small128(ctx, PW, PW, 1.0, D + PW, ldw, dinv_k, TILE, 0.0, D + PW, ldw); });
            }
            if (rem > PW) {
               wait(Q, e_leaf);
               tailB[s] = write(Q, s, s + 1, PB, 1, [&]() { dtrsm_right_lt(ctx, rem - PW, PW, D, ldw, dinv_k, D + 2 * PW, ldw); });
            }
            wait(C, headE[s]);
            wait(C, tailB[s]);
         }
         // sub-panel from its diagonal element to the end of its last column (one contiguous range) + inverted block
         if (cm) {
            tic(C, 2);
            cm->group_begin();
            cm->bcast(D, (size_t)(((PW - 1) * ldw + (n - c0)) * sizeof(double)), owner(g), C);
            cm->bcast(dinv_k, (size_t)(TILE * TILE * sizeof(double)), owner(g), C);
            cm->group_end();
            toc(C);
            cudaEvent_t e_arr = rec(C);
            if (!mine) headE[s] = tailB[s] = e_arr;
         }
         if (rem == 0) break;
         // diagonal tile of the next sub-panel: fused into this rank's step kernels where it can be
         if (mine && s + 1 < s1) eager(s, s + 1, s1, 3, fuse ? 1 : 0);  // the rest of this group
         if (next_mine && h0 < h1)                                        // the group this rank factors next
            eager(s, h0, h1, 4, !fuse ? 0 : (s + 1 == h0 ? (mine ? 1 : 2) : 0));
      }
      // bulk: all owned groups behind the next one, with the whole group g at once (rank-gw update); the first of
      // them -- the next one this rank will have to factor -- as its own launch
      int j1 = g + 2;
      while (j1 < ngrp && owner(j1) != me) ++j1;
      if (j1 < ngrp) {
         for (int s = s0; s < s1; ++s) need_panel(U, s, false, true);  // rows far below: the tails
         const double* Pg = W + g0 * ldw;  // panel of the group, global rows
         const int64_t j0 = g0_of(j1);
         write(U, j1 * spg, std::min(nsub, (j1 + 1) * spg), PALL, 5, [&]() {
            const double* Aj = Pg + j0;
            dgemm_padded(ctx, true, n - j0, gw_of(j1), gw, -1.0, Aj, ldw, Aj, ldw, 1.0, W + j0 + j0 * ldw, ldw, false);
         });
         if (j1 + R < ngrp) {
            // (the tracker marks every sub-panel column from the first of them on; columns of foreign groups in between
            //  are never written here, the extra ordering is harmless)
            write(U, (j1 + R) * spg, nsub, PALL, 5, [&]() {
               dsyrk_cyclic(ctx, n, gw, Pg, W, ldw, (int)(g0_of(j1 + R) / TILE), (int)(R * DB / TILE), (int)(DB / TILE));
            });
         }
      }
   }
   // the factor is complete on this rank when the last broadcast has arrived and P / Q have drained
   wait(U, rec(P));
   wait(U, rec(Q));
   wait(U, rec(C));
   if (trace) {
      MCHRG_CUDA(cudaStreamSynchronize(U));
      static const char* names[6] = {"leaf", "row solve (head+tail)", "broadcast", "own-group update", "next-group update", "bulk update"};
      double sum[6] = {0, 0, 0, 0, 0, 0};
      int cnt[6] = {0, 0, 0, 0, 0, 0};
      for (auto& sp : spans) {
         float ms = 0.f;
         MCHRG_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
         sum[sp.cat] += ms;
         cnt[sp.cat]++;
      }
      fprintf(stderr, "[mchrg_b200 dist trace] rank %d of %d n=%lld DB=%lld:", me, R, (long long)n, (long long)DB);
      for (int c = 0; c < 6; ++c) fprintf(stderr, " %s %.2f ms/%d", names[c], sum[c], cnt[c]);
      fprintf(stderr, "\n");
      if (env_ll("MCHRG_B200_DIST_TRACE", 0) > 1 && me == (int)env_ll("MCHRG_B200_DIST_TRACE_RANK", 0)) {
         // timeline of this rank's operations: start / end in microseconds since its first one
         const size_t lo = (size_t)env_ll("MCHRG_B200_DIST_TRACE_FROM", 0), hi = lo + 400;
         for (size_t i = lo; i < spans.size() && i < hi; ++i) {
            float t0 = 0.f, t1 = 0.f;
            MCHRG_CUDA(cudaEventElapsedTime(&t0, spans.front().a, spans[i].a));
            MCHRG_CUDA(cudaEventElapsedTime(&t1, spans.front().a, spans[i].b));
            fprintf(stderr, "[timeline r%d] %4zu %-22s %9.1f -> %9.1f us\n", me, i, names[spans[i].cat], 1e3 * t0, 1e3 * t1);
         }
      }
      for (auto& sp : spans) {
         cudaEventDestroy(sp.a);
         cudaEventDestroy(sp.b);
      }
   }
   if (!cm) return;
   // every rank reports the same pivot failure
   double* slots = ctx->alloc<double>((size_t)R);
   MCHRG_LAUNCH(ctx, info_to_slot_kernel, 1, 1, 0, info, me, R, slots);
   wait(C, rec(U));
   cm->allreduce_sum(slots, (size_t)R, C);
   wait(U, rec(C));
   MCHRG_LAUNCH(ctx, info_combine_kernel, 1, 1, 0, slots, R, info);
}

void dpotrf_padded(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info, bool aug) {
   if (n % TILE || (ldw & 1)) fail(MCHRG_B200_ERR_ARG, "dpotrf_padded: n=%lld not a multiple of 128", (long long)n);
   const int64_t aug_end = aug ? n : -1;
   int64_t PANEL = env_ll("MCHRG_B200_PANEL", 256);  // tuning knob (multiple of 128); default chosen on B200 at n = 16384
   if (PANEL < TILE || PANEL % TILE) PANEL = 256;
   if (dist_applies(ctx, n)) {  // large system on several GPUs
      dpotrf_dist(ctx, n, W, ldw, dinv, info, aug_end);
      return;
   }
   // The schedule of the distributed factorisation on one GPU (critical path of leaf + two small-tile products on its
   // own stream) wins while the serial chain matters (measured on B200: n = 4096 3.34 vs 3.60 ms, 8192 10.9 vs 11.0 ms),
   // the two-stream look-ahead with its wider panels below wins at n = 16384 (56.9 vs 58.5 ms).
   // MCHRG_B200_POTRF_TRACKED = 0 | 1 forces one of them (A/B timing).
   const long long tracked = env_ll("MCHRG_B200_POTRF_TRACKED", -1);
   if ((tracked > 0 || (tracked < 0 && n <= 8192)) && n >= 8 * TILE && ctx->active == ctx->stream) {
      dpotrf_dist(ctx, n, W, ldw, dinv, info, aug_end);
      return;
   }
   if (n <= 2 * PANEL) {  // small: plain recursion on the current stream
      potrf_rec(ctx, n, W, ldw, dinv, info, 0, aug_end);
      return;
   }
   // Right-looking blocked Cholesky with one panel of look-ahead on two streams.
   //   P (high priority): diagonal block (recursive), panel TRSM, update of the NEXT block column
   //   U (main stream)  : the bulk trailing update (lower block triangle beyond the next block column)
   // leaf / skinny-GEMM latency of panel k+1 is hidden behind the bulk update of panel k.
   cudaStream_t U = ctx->active, P = ctx->stream_hi;
   cudaEvent_t e0 = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e0, U));
   MCHRG_CUDA(cudaStreamWaitEvent(P, e0, 0));
   cudaEvent_t ev_upd_prev = nullptr;
   // panel width: MCHRG_B200_PANEL fixes it; MCHRG_B200_PANEL_ADAPT=hi,lo widens the panels while the trailing
   // matrix is large (fewer read-modify-write passes over it) and narrows them in the tail (shorter serial chain)
   // (default measured at n = 16384 on B200: 512-wide panels while more than 7168 columns are left, 256 down to
   //  2048, 128 in the tail: 67.4 -> 62.2 ms)
   const char* adapt_env = std::getenv("MCHRG_B200_PANEL_ADAPT");
   const bool adapt = adapt_env != nullptr || std::getenv("MCHRG_B200_PANEL") == nullptr;
   long thr4 = 1L << 60, thr_hi = 7168, thr_lo = 2048;
   if (adapt_env) {
      long a = 0, b = 0, c = 0;
      const int got = sscanf(adapt_env, "%ld,%ld,%ld", &a, &b, &c);
      if (got == 3) { thr4 = a; thr_hi = b; thr_lo = c; }
      else if (got == 2) { thr_hi = a; thr_lo = b; }
   }
   auto pw_at = [&](int64_t c) -> int64_t {
      const int64_t left = n - c;
      if (!adapt) return std::min<int64_t>(PANEL, left);
      const int64_t w = left > thr4 ? 4 * PANEL : (left > thr_hi ? 2 * PANEL : (left > thr_lo ? PANEL : TILE));
      return std::min<int64_t>(w, left);
   };
   for (int64_t c0 = 0; c0 < n;) {
      const int64_t pw = pw_at(c0);
      const int64_t rem = n - c0 - pw;
      double* D = W + c0 + c0 * ldw;
      double* dinv_k = dinv + (c0 / TILE) * TILE * TILE;
      cudaEvent_t ev_panel = ctx->next_event();
      {
         StreamScope on_p(ctx, P);
         potrf_rec(ctx, pw, D, ldw, dinv_k, info, c0, aug_end);
         if (rem > 0) dtrsm_right_lt(ctx, rem, pw, D, ldw, dinv_k, D + pw, ldw);
         MCHRG_CUDA(cudaEventRecord(ev_panel, P));
      }
      if (rem == 0) break;
      const int64_t pw2 = pw_at(c0 + pw);
      double* A21 = D + pw;                   // (rem x pw) panel below the diagonal block
      double* C1 = W + (c0 + pw) + (c0 + pw) * ldw;  // next block column (rem x pw2)
      {
         StreamScope on_p(ctx, P);
         if (ev_upd_prev) MCHRG_CUDA(cudaStreamWaitEvent(P, ev_upd_prev, 0));
         dgemm_padded(ctx, true, rem, pw2, pw, -1.0, A21, ldw, A21, ldw, 1.0, C1, ldw, false);
      }
      const int64_t rem2 = rem - pw2;
      if (rem2 > 0) {
         StreamScope on_u(ctx, U);
         MCHRG_CUDA(cudaStreamWaitEvent(U, ev_panel, 0));
         double* A2 = A21 + pw2;              // rows below the next block column
         double* C2 = W + (c0 + pw + pw2) + (c0 + pw + pw2) * ldw;
         dgemm_padded(ctx, true, rem2, rem2, pw, -1.0, A2, ldw, A2, ldw, 1.0, C2, ldw, /*lower_only=*/true);
         cudaEvent_t ev_upd = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(ev_upd, U));
         ev_upd_prev = ev_upd;
      } else {
         ev_upd_prev = nullptr;
      }
      c0 += pw;
   }
   cudaEvent_t e1 = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e1, P));
   MCHRG_CUDA(cudaStreamWaitEvent(U, e1, 0));
}

// ------------------------------------------------------------------------------------------------
// (L L^T) Y = R for a few right-hand sides: blocked substitution with the inverted diagonal blocks
// ------------------------------------------------------------------------------------------------
constexpr int MAX_RHS = 4;

// r_k <- inv(L_kk) r_k  (forward)  or  inv(L_kk)^T r_k (backward); one CTA of 256 threads:
// thread (i, h) accumulates half of the 128-term dot product of row i with independent, unrolled loads
__global__ void __launch_bounds__(256) trsv_diag_kernel(const double* __restrict__ dinv_k, double* __restrict__ R,
                                                        int64_t ldr, int nrhs, int transpose) {
   __shared__ double v[MAX_RHS][TILE];
   __shared__ double part[MAX_RHS][TILE];
   const int i = threadIdx.x & (TILE - 1), h = threadIdx.x >> 7;
   if (h == 0)
      for (int c = 0; c < nrhs; ++c) v[c][i] = R[i + c * ldr];
   __syncthreads();
   double acc[MAX_RHS] = {0.0, 0.0, 0.0, 0.0};
   const int k0 = h * (TILE / 2);
#pragma unroll 16
   for (int kk = 0; kk < TILE / 2; ++kk) {
      const int k = k0 + kk;
      // the blocks are stored with an explicitly zero upper triangle, so no bounds on k are needed
      const double l = transpose ? dinv_k[k + i * TILE] : dinv_k[i + k * TILE];
#pragma unroll
      for (int c = 0; c < MAX_RHS; ++c) acc[c] = fma(l, v[c][k], acc[c]);
   }
   if (h == 1)
      for (int c = 0; c < nrhs; ++c) part[c][i] = acc[c];
   __syncthreads();
   if (h == 0)
      for (int c = 0; c < nrhs; ++c) R[i + c * ldr] = acc[c] + part[c][i];
}

// forward step k: R[i] -= sum_j L(i, j) w_j for the rows i below block k (w = solved block k), CTA = 128 rows x
// 2 column slices of 64 (two batches of 32 loads in flight per thread, coalesced over the rows); the first
// CTA owns exactly the rows of block k+1 and applies inv(L_{k+1,k+1}) to them right away, so that one launch per
// block step remains (no separate single-CTA kernel between the steps).
__global__ void __launch_bounds__(256) trsv_fwd_fused_kernel(const double* __restrict__ Lcol, int64_t ldl,
                                                             const double* __restrict__ w, double* __restrict__ R,
                                                             int64_t ldr, int64_t nrows, int nrhs,
                                                             const double* __restrict__ dinv_next) {
   __shared__ double ws[MAX_RHS][TILE];
   __shared__ double part[MAX_RHS][TILE];
   __shared__ double rn[MAX_RHS][TILE];
   for (int idx = threadIdx.x; idx < MAX_RHS * TILE; idx += blockDim.x) {
      const int c = idx / TILE, j = idx % TILE;
      ws[c][j] = (c < nrhs) ? w[j + c * ldr] : 0.0;
   }
   __syncthreads();
   const int rl = threadIdx.x & (TILE - 1), slice = threadIdx.x >> 7;
   const int64_t i = (int64_t)blockIdx.x * TILE + rl;
   const bool solve_next = (blockIdx.x == 0) && (dinv_next != nullptr);
   double acc[MAX_RHS] = {0.0, 0.0, 0.0, 0.0};
   if (i < nrows) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
         double l[32];
         const int j0 = 64 * slice + 32 * half;
#pragma unroll
         for (int j = 0; j < 32; ++j) l[j] = Lcol[i + (int64_t)(j0 + j) * ldl];
#pragma unroll
         for (int j = 0; j < 32; ++j)
#pragma unroll
            for (int c = 0; c < MAX_RHS; ++c) acc[c] = fma(l[j], ws[c][j0 + j], acc[c]);
      }
   }
   if (slice == 1)
#pragma unroll
      for (int c = 0; c < MAX_RHS; ++c) part[c][rl] = acc[c];
   __syncthreads();
   if (slice == 0 && i < nrows) {
      for (int c = 0; c < nrhs; ++c) {
         const double r = R[i + c * ldr] - (acc[c] + part[c][rl]);
         if (solve_next) rn[c][rl] = r;
         else R[i + c * ldr] = r;
      }
   }
   if (!solve_next) return;
   __syncthreads();
   // block k+1: w = inv(L_{k+1,k+1}) r  (blocks are stored with an explicitly zero upper triangle)
   double a2[MAX_RHS] = {0.0, 0.0, 0.0, 0.0};
   const int k0 = slice * (TILE / 2);
#pragma unroll 16
   for (int kk = 0; kk < TILE / 2; ++kk) {
      const double d = dinv_next[rl + (k0 + kk) * TILE];
#pragma unroll
      for (int c = 0; c < MAX_RHS; ++c) a2[c] = fma(d, rn[c][k0 + kk], a2[c]);
   }
   if (slice == 1)
      for (int c = 0; c < nrhs; ++c) part[c][rl] = a2[c];
   __syncthreads();
   if (slice == 0)
      for (int c = 0; c < nrhs; ++c) R[rl + c * ldr] = a2[c] + part[c][rl];
}

// backward update: R[j] -= sum_i L(k0 + i, j) w_i  for columns j left of the block (one warp per column)
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* __restrict__ Lrow, int64_t ldl,
                                                              const double* __restrict__ w, double* __restrict__ R,
                                                              int64_t ldr, int64_t ncols, int nrhs) {
   __shared__ double ws[MAX_RHS][TILE];
   for (int idx = threadIdx.x; idx < MAX_RHS * TILE; idx += blockDim.x) {
      const int c = idx / TILE, j = idx % TILE;
      ws[c][j] = (c < nrhs) ? w[j + c * ldr] : 0.0;
   }
   __syncthreads();
   const int lane = threadIdx.x & 31;
   const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (j >= ncols) return;
   double acc[MAX_RHS] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
   for (int ii = 0; ii < TILE / 32; ++ii) {
      const int i = lane + 32 * ii;
      const double l = Lrow[i + j * ldl];
#pragma unroll
      for (int c = 0; c < MAX_RHS; ++c) acc[c] = fma(l, ws[c][i], acc[c]);
   }
#pragma unroll
   for (int c = 0; c < MAX_RHS; ++c) {
      double v = acc[c];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && c < nrhs) R[j + c * ldr] -= v;
   }
}

void dpotrs_vec(mchrg_b200_context* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv,
                double* R, int64_t ldr, int nrhs, bool forward) {
   if (nrhs > MAX_RHS || n % TILE) fail(MCHRG_B200_ERR_ARG, "dpotrs_vec: bad arguments");
   const int64_t nb = n / TILE;
   if (forward) MCHRG_LAUNCH(ctx, trsv_diag_kernel, 1, 256, 0, dinv, R, ldr, nrhs, 0);
   for (int64_t k = 0; forward && k + 1 < nb; ++k) {  // L w = r: update the rows below block k, solve block k+1 on the way
      double* rk = R + k * TILE;
      const int64_t below = n - (k + 1) * TILE;
      MCHRG_LAUNCH(ctx, trsv_fwd_fused_kernel, (unsigned)(below / TILE), 256, 0, L + (k + 1) * TILE + k * TILE * ldl, ldl, rk,
                   rk + TILE, ldr, below, nrhs, dinv + (k + 1) * TILE * TILE);
   }
   for (int64_t k = nb - 1; k >= 0; --k) {  // L^T y = w
      double* rk = R + k * TILE;
      MCHRG_LAUNCH(ctx, trsv_diag_kernel, 1, 256, 0, dinv + k * TILE * TILE, rk, ldr, nrhs, 1);
      const int64_t left = k * TILE;
      if (left > 0)
         MCHRG_LAUNCH(ctx, trsv_update_bwd_kernel, (unsigned)((left + 7) / 8), 256, 0, L + k * TILE, ldl, rk, R, ldr,
                      left, nrhs);
   }
}

}  // namespace mchrg
