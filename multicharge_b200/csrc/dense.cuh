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
// aug: rows n-2, n-1 hold right-hand sides [x | 1] of an augmented system (their own columns are kept at identity);
// after the factorisation they hold the forward-substituted vectors L^-1 x, L^-1 1 (rows of L).
// With a communicator on the context and n >= MCHRG_B200_DIST_MIN the factorisation is shared by the ranks
// (block-cyclic column groups, see dense.cu); every rank ends with the complete factor.
void dpotrf_padded(mchrg_b200_context* ctx, int64_t n, double* W, int64_t ldw, double* dinv, int* info,
                   bool aug = false);
// Width (columns) of the column groups dpotrf_padded would distribute block-cyclically over the ranks of the
// context's communicator (owner(g) = g % nranks) for an order-n matrix issued on the current stream; 0: not
// distributed.  A rank reads only the lower part of its OWN groups on entry, so the assembly may skip the others.
int64_t dpotrf_dist_panel(const mchrg_b200_context* ctx, int64_t n);
// One large system shared by the ranks of the context's communicator: true when a communicator is
// installed and the system has at least MCHRG_B200_DIST_MIN (default 8192) atoms.  Then the O(N^2) row passes
// (coordination number, (J q) / atrace / dadL, CN-gradient contraction) are row-sharded and joined by one
// all-reduce each, and the factorisation is block-cyclic.  row_share: this rank's rows (multiples of 8).
bool dist_rows_apply(const mchrg_b200_context* ctx, int64_t nat);
// periodic structures: the Ewald pair tiles (125 direct translations per image and pair) are shared between the ranks from
// MCHRG_B200_DIST_MIN_PBC atoms on (default 1024); the factorisation stays replicated below MCHRG_B200_DIST_MIN
bool dist_pbc_apply(const mchrg_b200_context* ctx, int64_t nat);
void row_share(const mchrg_b200_context* ctx, int nat, int* row0, int* row1);
// in-place sum over ranks, ordered after / before the context's current stream
void comm_allreduce_sum(mchrg_b200_context* ctx, double* buf, size_t count);

// X (m x n) <- X * L^{-T} and X <- X * L^{-1} (right-side solves with the Cholesky factor);
// together X <- X * (L L^T)^{-1}.  m, n multiples of 128.
void dtrsm_right_lt(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx);
void dtrsm_right_ln(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* L, int64_t ldl,
                    const double* dinv, double* X, int64_t ldx);

// Solve (L L^T) Y = R for nrhs (<= 4) right-hand sides stored as columns of R (n x nrhs, ld = ldr),
// in place.  Blocked forward + backward substitution using the inverted diagonal blocks.
// forward = false: R already holds L^-1 R (augmented factorisation), only L^T Y = R is solved.
void dpotrs_vec(mchrg_b200_context* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv,
                double* R, int64_t ldr, int nrhs, bool forward = true);

}  // namespace mchrg
