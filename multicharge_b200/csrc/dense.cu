// Dense FP64 kernels for sm_100a: DMMA (mma.sync m8n8k4 f64 -- the only FP64 tensor-core path on
// Blackwell; tcgen05 has no f64 kind) GEMM with a 4-stage cp.async pipeline, a register-resident
// 128x128 Cholesky/inverse leaf, and the recursive drivers built from them.
//
// Replaces, on the device: sytrf/sytri/sytrs (lapack.F90:165-363, called from model/type.F90:221-243)
// and gemm/gemv/symv (blas.F90, called from model/type.F90:235-290).
#include <cstdio>
#include "dense.cuh"
#include "erfgauss.cuh"

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
static int gemm_bn() {
   static const int v = std::getenv("MCHRG_B200_GEMM_BN") ? atoi(std::getenv("MCHRG_B200_GEMM_BN")) : 64;
   return v == 128 ? 128 : 64;
}

static void gemm_setup_once() {
   static bool done = false;
   if (done) return;
   MCHRG_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)gemm::Cfg<128>::SMEM_BYTES));
   MCHRG_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)gemm::Cfg<128>::SMEM_BYTES));
   MCHRG_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)gemm::Cfg<64>::SMEM_BYTES));
   MCHRG_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)gemm::Cfg<64>::SMEM_BYTES));
   done = true;
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
   gemm_setup_once();
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
   gemm_setup_once();
   const int4 cyc = make_int4(t0, stride, T, wt);
   if (gemm_bn() == 64) gemm_launch<64>(ctx, true, dim3((unsigned)tiles * 2), k, -1.0, P, ldw, P, ldw, 1.0, W, ldw, 2, cyc);
   else gemm_launch<128>(ctx, true, dim3((unsigned)tiles), k, -1.0, P, ldw, P, ldw, 1.0, W, ldw, 2, cyc);
}

// ------------------------------------------------------------------------------------------------
// Cholesky leaf: factor one 128 x 128 diagonal block and invert its factor, one CTA of 256
// threads.  Thread (tx, ty) = (tid & 15, tid >> 4) keeps the cyclic sub-matrix
// rows {tx + 16 r}, cols {ty + 16 c} (r, c = 0..7) in registers; every elimination step
// broadcasts one column (factorisation) or one row (Gauss-Jordan inversion) through shared memory,
// one barrier per step.
// ------------------------------------------------------------------------------------------------
namespace leaf {
constexpr int N = 128, THREADS = 256;
constexpr size_t SMEM_BYTES = (size_t(N) * N + 2 * N + N) * sizeof(double);  // Ls + bcast[2] + 1/diag
}  // namespace leaf

__global__ void __launch_bounds__(leaf::THREADS, 1)
potrf_leaf_kernel(double* __restrict__ W, int64_t ldw, double* __restrict__ dinv, int* __restrict__ info,
                  int pivot_base) {
   using namespace leaf;
   extern __shared__ __align__(16) double smem[];
   double* Ls = smem;                // Ls[col * N + row] = L(row, col)
   double* bc = smem + N * N;        // bc[2][N] broadcast buffers
   double* invd = bc + 2 * N;        // 1 / L(j, j)
   const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

   double a[8][8];
#pragma unroll
   for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
         const int row = tx + 16 * r, col = ty + 16 * c;
         a[r][c] = (row >= col) ? W[row + (int64_t)col * ldw] : W[col + (int64_t)row * ldw];
      }

   bool failed = false;
#pragma unroll
   for (int cj = 0; cj < 8; ++cj) {
      for (int jt = 0; jt < 16; ++jt) {
         const int j = 16 * cj + jt;
         double* col = bc + (j & 1) * N;
         if (ty == jt) {
#pragma unroll
            for (int r = 0; r < 8; ++r) col[tx + 16 * r] = a[r][cj];
         }
         __syncthreads();
         const double d = col[j];
         if (!(d > 0.0)) {  // also catches NaN; uniform across the CTA
            if (tid == 0) atomicCAS(info, 0, pivot_base + j + 1);
            failed = true;
            break;
         }
         const double inv = eg::fast_rsqrt(d);  // hardware seed + one third-order step (1.4e-16 rel.)
         if (ty == jt) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
               const int row = tx + 16 * r;
               if (row > j) a[r][cj] *= inv;
               else if (row == j) a[r][cj] = d * inv;
            }
         }
         // masked column factors: zero for rows / columns <= j, so the rank-one update needs no predicates
         double lr[8], lc[8];
#pragma unroll
         for (int r = 0; r < 8; ++r) lr[r] = (r >= cj && tx + 16 * r > j) ? -col[tx + 16 * r] * inv : 0.0;
#pragma unroll
         for (int c = 0; c < 8; ++c) lc[c] = (c >= cj && ty + 16 * c > j) ? col[ty + 16 * c] * inv : 0.0;
#pragma unroll
         for (int r = 0; r < 8; ++r) {
            if (r < cj) continue;  // rows tx + 16 r <= j for r < cj
#pragma unroll
            for (int c = 0; c < 8; ++c) {
               if (c < cj) continue;
               a[r][c] = fma(lr[r], lc[c], a[r][c]);
            }
         }
      }
      if (failed) break;
   }

   // publish L (lower triangle only) to global memory and shared memory
#pragma unroll
   for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
         const int row = tx + 16 * r, col = ty + 16 * c;
         if (row >= col) {
            W[row + (int64_t)col * ldw] = a[r][c];
            Ls[col * N + row] = a[r][c];
            if (row == col) invd[row] = 1.0 / a[r][c];
         }
      }
   __syncthreads();

   // Gauss-Jordan: X = inv(L), starting from the identity
   double x[8][8];
#pragma unroll
   for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) x[r][c] = (tx + 16 * r == ty + 16 * c) ? 1.0 : 0.0;

#pragma unroll
   for (int cj = 0; cj < 8; ++cj) {
      for (int jt = 0; jt < 16; ++jt) {
         const int j = 16 * cj + jt;
         double* rowb = bc + (j & 1) * N;
         if (tx == jt) {
            const double s = invd[j];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
               x[cj][c] *= s;
               rowb[ty + 16 * c] = x[cj][c];
            }
         }
         __syncthreads();
#pragma unroll
         for (int r = 0; r < 8; ++r) {
            if (r < cj) continue;
            const int row = tx + 16 * r;
            if (row > j) {
               const double l = Ls[j * N + row];
#pragma unroll
               for (int c = 0; c < 8; ++c) {
                  if (c > cj) continue;  // X(j, col) == 0 for col > j
                  x[r][c] = fma(-l, rowb[ty + 16 * c], x[r][c]);
               }
            }
         }
      }
   }
#pragma unroll
   for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
         const int row = tx + 16 * r, col = ty + 16 * c;
         dinv[row + col * N] = (row >= col && !failed) ? x[r][c] : 0.0;
      }
}

static void leaf_setup_once() {
   static bool done = false;
   if (done) return;
   MCHRG_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)leaf::SMEM_BYTES));
   done = true;
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

static void potrf_rec(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info,
                      int64_t off) {
   if (n == TILE) {
      MCHRG_LAUNCH(ctx, potrf_leaf_kernel, 1, leaf::THREADS, leaf::SMEM_BYTES, W, ldw, dinv, info, (int)off);
      return;
   }
   const int64_t n1 = split(n), n2 = n - n1;
   double* W21 = W + n1;
   double* W22 = W + n1 + n1 * ldw;
   double* dinv2 = dinv + (n1 / TILE) * TILE * TILE;
   potrf_rec(ctx, n1, W, ldw, dinv, info, off);
   dtrsm_right_lt(ctx, n2, n1, W, ldw, dinv, W21, ldw);
   dgemm_padded(ctx, true, n2, n2, n1, -1.0, W21, ldw, W21, ldw, 1.0, W22, ldw, /*lower_only=*/true);
   potrf_rec(ctx, n2, W22, ldw, dinv2, info, off + n1);
}

// Right-looking blocked Cholesky with one panel of look-ahead on two streams.
//   P (high priority): diagonal block (recursive), panel TRSM, update of the NEXT block column
//   U (main stream)  : the bulk trailing update (lower block triangle beyond the next block column)
// leaf / skinny-GEMM latency of panel k+1 is hidden behind the bulk update of panel k.
static int64_t panel_width() {  // tuning knob (multiple of 128); default chosen on B200 at n = 16384
   static int64_t w = 0;
   if (w == 0) {
      const char* e = std::getenv("MCHRG_B200_PANEL");
      w = e ? std::atoll(e) : 256;
      if (w < TILE || w % TILE) w = 256;
   }
   return w;
}

// smallest nonzero entry of the ranks' pivot reports
__global__ void info_combine_kernel(const int* __restrict__ xinfo, int nranks, int* __restrict__ info) {
   int v = 0;
   for (int r = 0; r < nranks; ++r)
      if (xinfo[r] != 0 && (v == 0 || xinfo[r] < v)) v = xinfo[r];
   *info = v;
}

static int64_t dist_panel_width() {  // MCHRG_B200_DIST_PANEL: columns per block column (multiple of 128), default 256
   static const int64_t w = std::getenv("MCHRG_B200_DIST_PANEL")
                               ? std::max<int64_t>(TILE, atoll(std::getenv("MCHRG_B200_DIST_PANEL")) / TILE * TILE)
                               : 2 * TILE;
   return w;
}

// Below ~8k the serial panel chain plus one broadcast per panel outlasts what the shared trailing updates save
// (measured on 2 B200: N = 3000 5.3 ms distributed vs 4.3 ms replicated; N = 16384 63 vs 68 ms, 4 GPUs 51.5 ms).
static bool dist_applies(const mchrg_b200_context* ctx, int64_t n) {
   static const int64_t nmin = std::getenv("MCHRG_B200_DIST_MIN") ? atoll(std::getenv("MCHRG_B200_DIST_MIN")) : 8192;
   return ctx->xch.on() && n >= std::max<int64_t>(nmin, 8 * TILE) && ctx->active == ctx->stream;
}

int64_t dpotrf_dist_panel(const mchrg_b200_context* ctx, int64_t n) { return dist_applies(ctx, n) ? dist_panel_width() : 0; }

// Multi-GPU right-looking Cholesky of a REPLICATED matrix: block columns of PANEL columns are
// owned block-cyclically (owner(b) = b % nranks).  Every rank keeps only its own block columns up to date (and
// needs only those assembled on entry: dpotrf_dist_panel tells the assembly which ones);
// the owner factors panel b (diagonal block + triangular solve of the rows below) and broadcasts it together
// with its inverted diagonal blocks, so that at the end EVERY rank holds the complete factor and the solves
// that follow stay replicated.  The broadcast of panel b+1 is started as soon as its owner has factored it
// (the owner updates that column first) and overlaps the trailing updates with panel b on all ranks.
static void dpotrf_dist(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info) {
   const auto& x = ctx->xch;
   const int64_t PANEL = dist_panel_width();
   const int nblk = (int)((n + PANEL - 1) / PANEL);
   auto c0_of = [&](int b) { return (int64_t)b * PANEL; };
   auto pw_of = [&](int b) { return std::min<int64_t>(PANEL, n - c0_of(b)); };
   auto owner = [&](int b) { return b % x.nranks; };
   auto factor_panel = [&](int b) {
      const int64_t c0 = c0_of(b), pw = pw_of(b), rem = n - c0 - pw;
      double* D = W + c0 + c0 * ldw;
      double* dinv_k = dinv + (c0 / TILE) * TILE * TILE;
      potrf_rec(ctx, pw, D, ldw, dinv_k, info, c0);
      if (rem > 0) dtrsm_right_lt(ctx, rem, pw, D, ldw, dinv_k, D + pw, ldw);
   };
   auto start_panel = [&](int b) {
      const int64_t c0 = c0_of(b), pw = pw_of(b);
      cudaEvent_t e = ctx->next_event();  // the exchange runs on the side stream, after what is queued here
      MCHRG_CUDA(cudaEventRecord(e, ctx->active));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_aux, e, 0));
      // block column from its diagonal element to the end of its last column (one contiguous range)
      x.start(x.user, W + c0 + c0 * ldw, (int64_t)(((pw - 1) * ldw + (n - c0)) * sizeof(double)), owner(b), 2 * b);
      x.start(x.user, dinv + (c0 / TILE) * TILE * TILE, (int64_t)((pw / TILE) * TILE * TILE * sizeof(double)), owner(b),
              2 * b + 1);
   };
   auto wait_panel = [&](int b) {
      x.wait(x.user, 2 * b);
      x.wait(x.user, 2 * b + 1);
   };
   auto update = [&](int b, int j) {  // block column j -= L(rows of j.., panel b) L(block j, panel b)^T
      const int64_t c0 = c0_of(b), cj = c0_of(j);
      const double* A21 = W + cj + c0 * ldw;
      dgemm_padded(ctx, true, n - cj, pw_of(j), pw_of(b), -1.0, A21, ldw, A21, ldw, 1.0, W + cj + cj * ldw, ldw, false);
   };
   if (owner(0) == x.rank) factor_panel(0);
   start_panel(0);
   for (int b = 0; b < nblk; ++b) {
      wait_panel(b);
      if (b + 1 < nblk) {
         if (owner(b + 1) == x.rank) {
            update(b, b + 1);
            factor_panel(b + 1);
         }
         start_panel(b + 1);
      }
      // all other owned block columns behind the look-ahead column in one launch
      int j0 = b + 2;
      while (j0 < nblk && owner(j0) != x.rank) ++j0;
      if (j0 < nblk)
         dsyrk_cyclic(ctx, n, pw_of(b), W + c0_of(b) * ldw, W, ldw, (int)(c0_of(j0) / TILE), (int)(x.nranks * PANEL / TILE),
                      (int)(PANEL / TILE));
   }
   // every rank reports the same pivot failure
   int* xinfo = ctx->alloc_zero<int>((size_t)x.nranks);
   MCHRG_CUDA(cudaMemcpyAsync(xinfo + x.rank, info, sizeof(int), cudaMemcpyDeviceToDevice, ctx->active));
   cudaEvent_t e = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e, ctx->active));
   MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_aux, e, 0));
   for (int r = 0; r < x.nranks; ++r) x.start(x.user, xinfo + r, (int64_t)sizeof(int), r, 2 * nblk + r);
   for (int r = 0; r < x.nranks; ++r) x.wait(x.user, 2 * nblk + r);
   MCHRG_LAUNCH(ctx, info_combine_kernel, 1, 1, 0, xinfo, x.nranks, info);
}

void dpotrf_padded(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info) {
   if (n % TILE || (ldw & 1)) fail(MCHRG_B200_ERR_ARG, "dpotrf_padded: n=%lld not a multiple of 128", (long long)n);
   leaf_setup_once();
   const int64_t PANEL = panel_width();
   if (dist_applies(ctx, n)) {  // large replicated system on several GPUs
      dpotrf_dist(ctx, n, W, ldw, dinv, info);
      return;
   }
   if (n <= 2 * PANEL) {  // small: plain recursion on the current stream
      potrf_rec(ctx, n, W, ldw, dinv, info, 0);
      return;
   }
   cudaStream_t U = ctx->active, P = ctx->stream_hi;
   cudaEvent_t e0 = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(e0, U));
   MCHRG_CUDA(cudaStreamWaitEvent(P, e0, 0));
   cudaEvent_t ev_upd_prev = nullptr;
   // panel width: MCHRG_B200_PANEL fixes it; MCHRG_B200_PANEL_ADAPT=hi,lo widens the panels while the trailing
   // matrix is large (fewer read-modify-write passes over it) and narrows them in the tail (shorter serial chain)
   // (default measured at n = 16384 on B200: 512-wide panels while more than 7168 columns are left, 256 down to
   //  2048, 128 in the tail: 67.4 -> 62.2 ms)
   static const char* adapt_env = std::getenv("MCHRG_B200_PANEL_ADAPT");
   static const bool adapt = adapt_env != nullptr || std::getenv("MCHRG_B200_PANEL") == nullptr;
   static long thr4 = 1L << 60, thr_hi = 7168, thr_lo = 2048;
   static bool parsed = false;
   if (adapt_env && !parsed) {
      long a = 0, b = 0, c = 0;
      const int got = sscanf(adapt_env, "%ld,%ld,%ld", &a, &b, &c);
      if (got == 3) { thr4 = a; thr_hi = b; thr_lo = c; }
      else if (got == 2) { thr_hi = a; thr_lo = b; }
      parsed = true;
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
         potrf_rec(ctx, pw, D, ldw, dinv_k, info, c0);
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
                double* R, int64_t ldr, int nrhs) {
   if (nrhs > MAX_RHS || n % TILE) fail(MCHRG_B200_ERR_ARG, "dpotrs_vec: bad arguments");
   const int64_t nb = n / TILE;
   MCHRG_LAUNCH(ctx, trsv_diag_kernel, 1, 256, 0, dinv, R, ldr, nrhs, 0);
   for (int64_t k = 0; k + 1 < nb; ++k) {  // L w = r: update the rows below block k, solve block k+1 on the way
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
