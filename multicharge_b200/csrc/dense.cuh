// Dense FP64 building blocks on the device (replacing the reference's lapack.F90 / blas.F90 calls):
// DMMA GEMM, recursive blocked Cholesky with inverted diagonal blocks, right-side triangular
// solves.  All matrices column-major; "padded" routines require dimensions that are multiples of
// the 128 tile and 16-byte aligned columns (even leading dimension).
#pragma once
#include "common.cuh"

namespace mchrg {

constexpr int TILE = 128;  // CTA tile edge == Cholesky leaf == diagonal-block size

// C(m x n) = alpha * A(m x k) * op(B) + beta * C ;  op(B) = B^T with B (n x k) when transb, else
// B (k x n).  m, n multiples of 128, k multiple of 16.  lower_only: only tiles on/below the block
// diagonal of C are computed (SYRK-style trailing update).  C may alias A only when n == 128
// (each CTA then owns complete rows).
void dgemm_padded(mchrg_b200_context* ctx, bool transb, int64_t m, int64_t n, int64_t k, double alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                  int64_t ldc, bool lower_only = false);

// In-place lower Cholesky of the n x n matrix W (n multiple of 128).  dinv receives the inverses
// of the 128x128 diagonal blocks of L (n/128 blocks of 128*128 doubles, column-major, lower).
// info (device int) is set to the 1-based index of the first non-positive pivot, else left 0.
void dpotrf_padded(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info);
// Width (columns) of the block columns dpotrf_padded would distribute block-cyclically over the ranks of the panel
// exchange (owner(b) = b % nranks) for an order-n matrix issued on the current stream; 0: not distributed.  A rank
// reads only the lower part of its OWN block columns on entry, so the assembly may skip the others.
int64_t dpotrf_dist_panel(const mchrg_b200_context* ctx, int64_t n);

// X (m x n) <- X * L^{-T} and X <- X * L^{-1} (right-side solves with the Cholesky factor);
// together X <- X * (L L^T)^{-1}.  m, n multiples of 128.
void dtrsm_right_lt(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx);
void dtrsm_right_ln(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx);

// Solve (L L^T) Y = R for nrhs (<= 4) right-hand sides stored as columns of R (n x nrhs, ld = ldr),
// in place.  Blocked forward + backward substitution using the inverted diagonal blocks.
void dpotrs_vec(mchrg_b200_context* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv,
                double* R, int64_t ldr, int nrhs);

}  // namespace mchrg
