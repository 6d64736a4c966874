// C-ABI entry points (include/mchrg_b200.h) and the device orchestration of
// mchrg_model_type%solve (model/type.F90:151-292), get_charges (charge.f90:37-77) and the CLI
// sequence (app/main.f90:114-118).
//
// How solve() is restated for the GPU (EEQ2019):
//   reference (type.F90)                         here
//   -------------------------------------------  ------------------------------------------------
//   amat (N+1)^2 incl. constraint border         only the SPD Coulomb block J (padded to 128)
//   sytrf: Bunch-Kaufman LDL^T of bordered amat   blocked Cholesky J = L L^T (DMMA trailing updates)
//   sytrs / sytri+symv                            y = J^-1 x, z = J^-1 1, lambda = (1^T y - Q)/(1^T z),
//                                                 q = y - lambda z   (Schur complement of the border)
//   symv(0.5 J q - x), energy += q * (...)        (J q) re-evaluated from the pair formula (no J copy)
//   dadr, dxdr (3 N (N+1) each), gemv x4          gradient = q * atrace - sum_k thalf_k q_k dcndr(:,:,k)
//   dqdr = -(dadr - dxdr) * ainv(:, :N)  (gemm)   B = dadr - dxdr built once, X = B L^-T L^-1 (two
//                                                 recursive right-side TRSMs), dqdr = -X + (B z) z^T / s
#include <algorithm>
#include <cmath>

#include "dense.cuh"
#include "eeq.cuh"
#include "eeq_pbc.cuh"
#include "eeqbc.cuh"
#include "lattice.hpp"
#include "param_tables.inc"

using namespace mchrg;

namespace mchrg {

// ---------------------------------------------------------------------------------------------
// small kernels of the solve tail
// ---------------------------------------------------------------------------------------------

// R(:,0) = x (zero padded), R(:,1) = 1 for atoms, 0 for padding
__global__ void rhs_init_kernel(int nat, int64_t npad, const double* __restrict__ xvec, double* __restrict__ R) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npad) return;
   R[i] = (i < nat) ? xvec[i] : 0.0;
   R[i + npad] = (i < nat) ? 1.0 : 0.0;
}

// Augmented system: rows npad-2 / npad-1 of the work matrix carry the right-hand sides [x | 1], so the forward
// substitution falls out of the factorisation (the leaf keeps those rows' own columns at identity).
__global__ void aug_rows_kernel(int nat, int64_t npad, const double* __restrict__ xvec, double* __restrict__ W) {
   const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= npad) return;
   W[npad - 2 + j * npad] = (j < nat) ? xvec[j] : (j == npad - 2 ? 1.0 : 0.0);
   W[npad - 1 + j * npad] = (j < nat) ? 1.0 : (j == npad - 1 ? 1.0 : 0.0);
}
// R(:,0) = L^-1 x, R(:,1) = L^-1 1 from the last two rows of L; those rows (and the matching rows of the last
// inverted diagonal block) are cleared, which leaves L = blockdiag(L_J, I) for the dq/dr sweeps
__global__ void aug_extract_kernel(int64_t npad, double* __restrict__ W, double* __restrict__ dinv_last,
                                   double* __restrict__ R) {
   const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= npad) return;
   const bool body = j < npad - 2;
   R[j] = body ? W[npad - 2 + j * npad] : 0.0;
   R[j + npad] = body ? W[npad - 1 + j * npad] : 0.0;
   if (body) {
      W[npad - 2 + j * npad] = 0.0;
      W[npad - 1 + j * npad] = 0.0;
      const int64_t jl = j - (npad - TILE);
      if (jl >= 0) dinv_last[TILE - 2 + jl * TILE] = dinv_last[TILE - 1 + jl * TILE] = 0.0;
   }
}

// Schur complement of the total-charge constraint: one CTA.
// scal[0] = s = 1^T z, scal[1] = lambda; vrhs[0..nat-1] = q, vrhs[nat] = lambda
__global__ void __launch_bounds__(1024) schur_charges_kernel(int nat, int64_t npad, const double* __restrict__ R,
                                                             double charge, double* __restrict__ vrhs,
                                                             double* __restrict__ qvec, double* __restrict__ scal,
                                                             const double* __restrict__ shift) {
   __shared__ double sy[32], sz[32];
   __shared__ double lam;
   double ay = 0.0, az = 0.0;
   for (int i = threadIdx.x; i < nat; i += blockDim.x) {
      ay += R[i];
      az += R[i + npad];
   }
   for (int o = 16; o > 0; o >>= 1) {
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
      az += __shfl_xor_sync(0xffffffffu, az, o);
   }
   if ((threadIdx.x & 31) == 0) {
      sy[threadIdx.x >> 5] = ay;
      sz[threadIdx.x >> 5] = az;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      double ty = 0.0, tz = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
         ty += sy[w];
         tz += sz[w];
      }
      lam = (ty - charge) / tz;  // multiplier of the (possibly shifted) system
      scal[0] = tz;
      scal[1] = lam + shift[0] * charge;  // lambda of the reference's unshifted system
      vrhs[nat] = scal[1];
   }
   __syncthreads();
   const double l = lam;
   for (int i = threadIdx.x; i < nat; i += blockDim.x) {
      const double qi = R[i] - l * R[i + npad];
      vrhs[i] = qi;
      if (qvec) qvec[i] = qi;
   }
}

// energy(i) += q_i (0.5 (J q)_i - x_i)      (type.F90:259-261)
__global__ void energy_kernel(int nat, const double* __restrict__ q, const double* __restrict__ jq,
                              const double* __restrict__ xvec, double* __restrict__ energy) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nat) energy[i] += q[i] * (0.5 * jq[i] - xvec[i]);
}

// gradient(:, j) = q_j atrace(:, j) - gradx(:, j)    (type.F90:276-278 after eliminating dadr/dxdr)
__global__ void gradient_kernel(int jb, int nj, const double* __restrict__ q, const double* __restrict__ atrace,
                                const double* __restrict__ gradx, double* __restrict__ gradient) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;  // row within the block [jb, jb + nj)
   if (i < 3 * nj) gradient[i] = q[jb + i / 3] * atrace[3 * jb + i] - gradx[i];
}

// sigma(a,b) += sum_k q_k (0.5 dadL(a,b,k) - thalf_k dcndL(a,b,k))   (type.F90:279-280); 9 warps.
// dcndL == nullptr: the second sum arrives contracted in sig_cn (cn_gradient_device).
__global__ void __launch_bounds__(288) sigma_kernel(int nat, const double* __restrict__ q,
                                                    const double* __restrict__ dadL, const double* __restrict__ thalf,
                                                    const double* __restrict__ dcndL, const double* __restrict__ sig_cn,
                                                    double* __restrict__ sigma) {
   const int ab = threadIdx.x >> 5, lane = threadIdx.x & 31;
   double acc = 0.0;
   if (dcndL) {
      for (int k = lane; k < nat; k += 32)
         acc += q[k] * (0.5 * dadL[(size_t)9 * k + ab] - thalf[k] * dcndL[(size_t)9 * k + ab]);
   } else {
      for (int k = lane; k < nat; k += 32) acc += q[k] * (0.5 * dadL[(size_t)9 * k + ab]);
   }
   for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
   if (lane == 0) sigma[ab] += acc - (dcndL ? 0.0 : sig_cn[ab]);
}

// rows 3N..3N+8 of B: dadL(:,:,k) - dxdL(:,:,k) (type.F90:286), and their z-contraction
__global__ void __launch_bounds__(288) strain_rows_kernel(int nat, const double* __restrict__ dadL,
                                                          const double* __restrict__ thalf,
                                                          const double* __restrict__ dcndL,
                                                          const double* __restrict__ zvec, double* __restrict__ B,
                                                          int64_t ldb, double* __restrict__ bz, int64_t roff) {
   const int ab = threadIdx.x >> 5, lane = threadIdx.x & 31;
   double acc = 0.0;
   for (int k = lane; k < nat; k += 32) {
      const double v = dadL[(size_t)9 * k + ab] - thalf[k] * dcndL[(size_t)9 * k + ab];
      B[(size_t)k * ldb + roff + ab] = v;
      acc = fma(v, zvec[k], acc);
   }
   for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
   if (lane == 0) bz[roff + ab] = acc;
}

// dqdr(r, k) = -X(r, k) + bz_r z_k / s ; rows 3N.. go to dqdL     (type.F90:289-290 + Schur correction)
// nj = number of atoms in this row block (dqdr holds the rows of the block only, leading dimension 3 nj)
__global__ void __launch_bounds__(256) dq_epilogue_kernel(int nat, int nj, const double* __restrict__ X, int64_t ldx,
                                                          const double* __restrict__ bz,
                                                          const double* __restrict__ zvec,
                                                          const double* __restrict__ scal, double* __restrict__ dqdr,
                                                          double* __restrict__ dqdL) {
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t n3 = 3 * (int64_t)nj, nrow = n3 + 9;
   if (r >= nrow) return;
   const double u = bz[r] / scal[0];
   const int k0 = blockIdx.y * 16;
#pragma unroll 4
   for (int kk = 0; kk < 16; ++kk) {
      const int k = k0 + kk;
      if (k >= nat) break;
      const double v = fma(u, zvec[k], -X[r + (size_t)k * ldx]);
      if (r < n3) dqdr[r + (size_t)k * n3] = v;
      else if (dqdL) dqdL[(r - n3) + (size_t)k * 9] = v;
   }
}

// pack / unpack helpers for the exposed dense test entry points
__global__ void pad_copy_kernel(int64_t m, int64_t n, const double* __restrict__ src, int64_t lds, int64_t mp,
                                int64_t np, double* __restrict__ dst, int64_t ldd, int identity_pad) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t j = blockIdx.y;
   if (i >= mp || j >= np) return;
   double v = 0.0;
   if (i < m && j < n) v = src[i + j * lds];
   else if (identity_pad && i == j) v = 1.0;
   dst[i + j * ldd] = v;
}
__global__ void unpad_copy_kernel(int64_t m, int64_t n, const double* __restrict__ src, int64_t lds,
                                  double* __restrict__ dst, int64_t ldd, int lower_only) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t j = blockIdx.y;
   if (i >= m || j >= n) return;
   if (lower_only && i < j) return;
   dst[i + j * ldd] = src[i + j * lds];
}

// xvec_derivs (eeq.f90:152-180): dxdr(:,:,k) = thalf_k dcndr(:,:,k), dxdL likewise; last column zero
__global__ void xvec_derivs_kernel(int nat, const double* __restrict__ thalf, const double* __restrict__ dcndr,
                                   const double* __restrict__ dcndL, double* __restrict__ dxdr,
                                   double* __restrict__ dxdL) {
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t n3 = 3 * (int64_t)nat;
   const int64_t total = n3 * (nat + 1);
   if (idx < total) {
      const int64_t k = idx / n3;
      dxdr[idx] = (k < nat) ? thalf[k] * dcndr[idx] : 0.0;
   }
   if (idx < 9 * (int64_t)(nat + 1)) {
      const int64_t k = idx / 9;
      dxdL[idx] = (k < nat) ? thalf[k] * dcndL[idx] : 0.0;
   }
}

// amat(nat+1, nat+1) <- J (padded work matrix) + constraint border (eeq.f90:303-305)
__global__ void border_copy_kernel(int nat, const double* __restrict__ W, int64_t ldw, double* __restrict__ amat) {
   const int64_t nd = nat + 1;
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t c = blockIdx.y;
   if (r >= nd || c >= nd) return;
   double v;
   if (r < nat && c < nat) v = W[r + c * ldw];
   else v = (r == nat && c == nat) ? 0.0 : 1.0;
   amat[r + c * nd] = v;
}

// dadr(3, nat, nat+1) <- off-diagonal blocks of the padded B; diagonal blocks and last column zero
__global__ void dadr_copy_kernel(int nat, const double* __restrict__ B, int64_t ldb, double* __restrict__ dadr) {
   const int64_t n3 = 3 * (int64_t)nat;
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t k = blockIdx.y;
   if (r >= n3 || k > nat) return;
   dadr[r + k * n3] = (k < nat && r / 3 != k) ? B[r + k * ldb] : 0.0;
}

// ---- rank-one shift of the Coulomb block:  J' = J + mu 1 1^T ---------------------------------
// The constrained solution q (and the projected inverse J^-1 - z z^T / s that dq/dr needs) is invariant
// under this shift; only the multiplier moves: lambda = lambda' + mu Q.  Periodic Ewald matrices are
// indefinite along the constraint vector 1 (the omitted G = 0 term), which the reference's LDL^T
// tolerates and a Cholesky factorisation does not.
__global__ void __launch_bounds__(256) matsum_kernel(int nat, const double* __restrict__ W, int64_t ld,
                                                     double* __restrict__ total) {
   __shared__ double part[8];
   const int col = blockIdx.x;
   double acc = 0.0;
   for (int r = threadIdx.x; r < nat; r += blockDim.x) acc += W[r + (int64_t)col * ld];
   for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
   if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
   __syncthreads();
   if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += part[w];
      atomicAdd(total, s);
   }
}
// sh[0] = mu, sh[1] = 1^T J 1 (input).  first: mu from the Rayleigh quotient rho = 1^T J 1 / N along 1,
// so that rho + mu N = 1 + |rho| when rho < 0; later calls (retry after a failed pivot) multiply by 8.
__global__ void shift_update_kernel(int nat, double* __restrict__ sh) {
   const double rho = sh[1] / (double)nat;
   if (sh[0] == 0.0) sh[0] = (2.0 * fmax(0.0, -rho) + 1.0) / (double)nat;
   else sh[0] *= 8.0;
}
__global__ void shift_apply_kernel(int nat, double* __restrict__ W, int64_t ld, const double* __restrict__ sh) {
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   const int c = blockIdx.y;
   if (r < nat && c < nat) W[r + (int64_t)c * ld] += sh[0];
}

// ---- indefinite fallback (see factor_solve): bordered matrix K of order nk (padded to npk with the identity) ----
__global__ void bordered_kernel(int nat, int64_t npad, const double* __restrict__ W, int64_t npk, double* __restrict__ K) {
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t c = blockIdx.y;
   if (r >= npk || c >= npk) return;
   double v;
   if (r < nat && c < nat) v = W[r + c * npad];
   else if (r <= nat && c <= nat) v = (r == nat && c == nat) ? 0.0 : 1.0;  // constraint row / column
   else v = (r == c) ? 1.0 : 0.0;
   K[r + c * npk] = v;
}
// b = [x; Q; 0 ...] in column 0 of R (leading dimension npk), column 1 zero
__global__ void bordered_rhs_kernel(int nat, int64_t npk, const double* __restrict__ xvec, double charge, double* __restrict__ R) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npk) return;
   R[i] = (i < nat) ? xvec[i] : (i == nat ? charge : 0.0);
   R[i + npk] = 0.0;
}
// r = b - t   /   y += d   /   copy
__global__ void axpby_kernel(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                             double* __restrict__ out) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = a * x[i] + b * y[i];
}
__global__ void fallback_charges_kernel(int nat, const double* __restrict__ y, double* __restrict__ vrhs, double* __restrict__ qvec,
                                        double* __restrict__ scal) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i <= nat) vrhs[i] = y[i];
   if (i < nat && qvec) qvec[i] = y[i];
   if (i == 0) {
      scal[0] = 1.0;
      scal[1] = y[nat];
   }
}

__global__ void fill_kernel(int64_t n, double v, double* __restrict__ p) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
struct SolveIO {
   const double *cn = nullptr, *qloc = nullptr, *dcndr = nullptr, *dcndL = nullptr, *dqlocdr = nullptr,
                *dqlocdL = nullptr;
   double *energy = nullptr, *gradient = nullptr, *sigma = nullptr, *qvec = nullptr, *dqdr = nullptr,
          *dqdL = nullptr;
   // row block [jb, je) of the derivative outputs (gradient rows, dq/dr rows); je < 0: all atoms.
   // With a proper sub-block, gradient is (3, nj) and dqdr (3, nj, nat).
   int jb = 0, je = -1;
   // EEQ2019 gradient without dq/dr: CN derivatives contracted on the fly instead of dcndr / dcndL
   const CnLazy* lazy = nullptr;
};

static AtomPar gather_eeq(const mchrg_b200_model* model, const DevMol& mol) {
   mchrg_b200_context* ctx = model->ctx;
   AtomPar ap;
   auto g = [&](int slot) {
      double* d = ctx->alloc<double>(mol.nat);
      gather_species(ctx, mol.nat, mol.id, model->dev(slot), d);
      return d;
   };
   ap.chi = g(0);
   ap.rad = g(1);
   ap.eta = g(2);
   ap.kcnchi = g(3);
   ap.rcov = g(4);
   return ap;
}

static DevMol stage_mol(mchrg_b200_context* ctx, const mchrg_b200_structure* mol) {
   if (mol == nullptr || mol->nat <= 0 || mol->id == nullptr || mol->xyz == nullptr)
      fail(MCHRG_B200_ERR_ARG, "invalid structure (nat=%d)", mol ? mol->nat : -1);
   DevMol d;
   d.nat = mol->nat;
   d.id = ctx->in(mol->id, (size_t)mol->nat);
   d.xyz = ctx->in(mol->xyz, (size_t)3 * mol->nat);
   d.charge = mol->charge;
   d.periodic = mol->periodic[0] || mol->periodic[1] || mol->periodic[2];
   for (int k = 0; k < 9; ++k) d.lattice[k] = mol->lattice[k];
   return d;
}

// Factorisation + constrained solve shared by both models.
//   assemble():  (re)build the Coulomb block into W (order npad, identity padding)
//   pre_shift:   apply the rank-one shift before the first attempt (periodic Ewald matrices)
// On success: W = L, dinv = inverted diagonal blocks, R(:,1) = z = J'^-1 1, vrhs = [q, lambda], scal = [s, lambda].
struct Factor {
   double *W = nullptr, *dinv = nullptr, *R = nullptr, *vrhs = nullptr, *scal = nullptr, *shift = nullptr;
   int info = 0;
   bool indefinite = false;  // solved through the normal equations of the bordered matrix: no factor of J (no dq/dr)
   const double* z(int64_t npad) const { return R + npad; }
};

// Augmented right-hand-side rows (MCHRG_B200_AUG=0 switches them off): used when they cost no extra tile (nat + 2
// fits the padded order anyway) or when the system is shared by several ranks (one replicated substitution sweep
// less per rank).
static bool aug_enabled(const mchrg_b200_context* ctx, int nat) {
   if (env_ll("MCHRG_B200_AUG", 1) == 0) return false;
   return round_up(nat + 2, TILE) == round_up(nat, TILE) || dist_rows_apply(ctx, nat);
}

template <class Assemble, class AfterAssemble>
static Factor factor_solve(mchrg_b200_context* ctx, int nat, int64_t npad, Assemble&& assemble, AfterAssemble&& after_first,
                           bool pre_shift, const double* xvec, double charge, double* qvec_out, bool aug = false) {
   if (aug && npad < nat + 2) fail(MCHRG_B200_ERR_ARG, "factor_solve: augmented order too small");
   Factor f;
   f.W = ctx->alloc<double>((size_t)npad * npad);
   f.dinv = ctx->alloc<double>((size_t)npad * TILE);
   int* info = ctx->alloc_zero<int>(1);
   f.shift = ctx->alloc_zero<double>(2);  // [0] mu of J' = J + mu 1 1^T, [1] 1^T J 1
   auto apply_shift = [&]() {
      MCHRG_CUDA(cudaMemsetAsync(f.shift + 1, 0, sizeof(double), ctx->stream));
      MCHRG_LAUNCH(ctx, matsum_kernel, (unsigned)nat, 256, 0, nat, f.W, npad, f.shift + 1);
      MCHRG_LAUNCH(ctx, shift_update_kernel, 1, 1, 0, nat, f.shift);
      MCHRG_LAUNCH(ctx, shift_apply_kernel, dim3((unsigned)((nat + 255) / 256), (unsigned)nat), 256, 0, nat, f.W, npad,
                   f.shift);
   };
   assemble(f.W, /*first=*/true);
   after_first(f.W);  // e.g. keep a copy of the unshifted matrix for the energy
   if (pre_shift) apply_shift();
   int* h_info = static_cast<int*>(ctx->pinned_buf(sizeof(int)));
   for (int attempt = 0;; ++attempt) {
      if (aug) MCHRG_LAUNCH(ctx, aug_rows_kernel, (unsigned)((npad + 255) / 256), 256, 0, nat, npad, xvec, f.W);
      dpotrf_padded(ctx, npad, f.W, npad, f.dinv, info, aug);
      MCHRG_CUDA(cudaMemcpyAsync(h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));
      f.info = *h_info;
      if (env_ll("MCHRG_B200_DIST_TRACE", 0) > 0)
         fprintf(stderr, "[mchrg_b200 factor_solve] rank %d attempt %d npad %lld info %d\n", ctx->rank(), attempt, (long long)npad, f.info);
      if (f.info == 0) break;
      if (attempt == 3) {
         // J is indefinite on the charge-conserving subspace as well (e.g. two atoms far inside each other's charge
         // width).  The reference's Bunch-Kaufman LDL^T (lapack.F90:165-201) factors any nonsingular bordered matrix
         // K = [J 1; 1^T 0]; here K^2 = K K^T is positive definite, so  K^-1 = (K K^T)^-1 K : DMMA product, the same
         // blocked Cholesky, and two steps of iterative refinement against K itself (the squared condition number
         // costs digits only in the first solve).  Charges, energy and gradient; dq/dr is not offered on this route.
         if (std::getenv("MCHRG_B200_NO_INDEFINITE")) return f;
         const int pivot = f.info;
         const int64_t npk = round_up(nat + 1, TILE);
         assemble(f.W, /*first=*/false);
         double* K = ctx->alloc<double>((size_t)npk * npk);
         double* M = ctx->alloc<double>((size_t)npk * npk);
         dim3 gk((unsigned)((npk + 255) / 256), (unsigned)npk);
         MCHRG_LAUNCH(ctx, bordered_kernel, gk, 256, 0, nat, npad, f.W, npk, K);
         dgemm_padded(ctx, true, npk, npk, npk, 1.0, K, npk, K, npk, 0.0, M, npk, /*lower_only=*/true);
         double* dinvM = ctx->alloc<double>((size_t)npk * TILE);
         MCHRG_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
         {  // (never shared between ranks: every rank solves the small bordered problem itself)
            Comm* keep = ctx->comm;
            ctx->comm = nullptr;
            try {
               dpotrf_padded(ctx, npk, M, npk, dinvM, info);
            } catch (...) {
               ctx->comm = keep;
               throw;
            }
            ctx->comm = keep;
         }
         MCHRG_CUDA(cudaMemcpyAsync(h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
         MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));
         if (*h_info != 0) {  // singular to working precision: report the pivot of the Coulomb block like before
            f.info = pivot;
            return f;
         }
         const unsigned gb = (unsigned)((npk + 255) / 256);
         double* b = ctx->alloc<double>((size_t)npk * 2);
         double* y = ctx->alloc_zero<double>((size_t)npk);
         double* r = ctx->alloc<double>((size_t)npk * 2);
         double* t = ctx->alloc<double>((size_t)npk);
         MCHRG_LAUNCH(ctx, bordered_rhs_kernel, gb, 256, 0, nat, npk, xvec, charge, b);
         MCHRG_CUDA(cudaMemcpyAsync(r, b, (size_t)npk * 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
         for (int it = 0; it < 3; ++it) {  // y += (K K^T)^-1 K r ;  r = b - K y
            MCHRG_CUDA(cudaMemsetAsync(r + npk, 0, (size_t)npk * sizeof(double), ctx->stream));
            pbc_symv(ctx, (int)npk, K, npk, r, t);
            MCHRG_CUDA(cudaMemcpyAsync(r, t, (size_t)npk * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            dpotrs_vec(ctx, npk, M, npk, dinvM, r, npk, 2);
            MCHRG_LAUNCH(ctx, axpby_kernel, gb, 256, 0, npk, 1.0, y, 1.0, r, y);
            pbc_symv(ctx, (int)npk, K, npk, y, t);
            MCHRG_LAUNCH(ctx, axpby_kernel, gb, 256, 0, npk, 1.0, b, -1.0, t, r);
         }
         f.vrhs = ctx->alloc<double>(nat + 1);
         f.scal = ctx->alloc<double>(2);
         f.R = ctx->alloc_zero<double>((size_t)npad * 2);
         MCHRG_LAUNCH(ctx, fallback_charges_kernel, (unsigned)((nat + 256) / 256), 256, 0, nat, y, f.vrhs, qvec_out, f.scal);
         f.info = 0;
         f.indefinite = true;
         return f;
      }
      MCHRG_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
      assemble(f.W, /*first=*/false);
      apply_shift();
   }
   // y = J^-1 x, z = J^-1 1, Schur complement of the constraint
   f.R = ctx->alloc<double>((size_t)npad * 2);
   if (aug)
      MCHRG_LAUNCH(ctx, aug_extract_kernel, (unsigned)((npad + 255) / 256), 256, 0, npad, f.W,
                   f.dinv + (npad / TILE - 1) * TILE * TILE, f.R);
   else
      MCHRG_LAUNCH(ctx, rhs_init_kernel, (unsigned)((npad + 255) / 256), 256, 0, nat, npad, xvec, f.R);
   dpotrs_vec(ctx, npad, f.W, npad, f.dinv, f.R, npad, 2, /*forward=*/!aug);
   f.vrhs = ctx->alloc<double>(nat + 1);
   f.scal = ctx->alloc<double>(2);
   MCHRG_LAUNCH(ctx, schur_charges_kernel, 1, 1024, 0, nat, npad, f.R, charge, f.vrhs, qvec_out, f.scal, f.shift);
   return f;
}

// dq/dr, dq/dL from B (type.F90:283-291): X = B L^-T L^-1, dqdr = -X + (B z) z^T / s
static void dq_tail(mchrg_b200_context* ctx, int nat, int nj, int64_t npad, int64_t mp, const Factor& f, double* Bm,
                    const double* bz, double* dqdr, double* dqdL) {
   dtrsm_right_lt(ctx, mp, npad, f.W, npad, f.dinv, Bm, mp);
   dtrsm_right_ln(ctx, mp, npad, f.W, npad, f.dinv, Bm, mp);
   dim3 grid((unsigned)((3 * (int64_t)nj + 9 + 255) / 256), (unsigned)((nat + 15) / 16));
   MCHRG_LAUNCH(ctx, dq_epilogue_kernel, grid, 256, 0, nat, nj, Bm, mp, bz, f.z(npad), f.scal, dqdr, dqdL);
}

// releases the periodic setup on every exit path
struct PbcScope {
   PbcHost host;
   PbcWork work;
   bool on = false;
   ~PbcScope() {
      if (on) pbc_end(work);
   }
};

// the EEQ2019 solve on device arrays; returns LAPACK-style info (0 ok, > 0 failed pivot)
static int solve_eeq_device(const mchrg_b200_model* model, const DevMol& mol, const AtomPar& ap, const SolveIO& io) {
   mchrg_b200_context* ctx = model->ctx;
   const int nat = mol.nat;
   const bool pbc = mol.periodic;
   // augmented right-hand-side rows: non-periodic path only (the periodic kernels pad to round_up(nat) themselves)
   const bool aug = !pbc && aug_enabled(ctx, nat);
   const int64_t npad = aug ? round_up(nat + 2, TILE) : round_up(nat, TILE);
   const bool dcn = io.dcndr && io.dcndL;
   const bool cpq = io.dqdr && io.dqdL && dcn;
   const bool lazy = io.lazy && !cpq;
   const bool grad = io.gradient && io.sigma && (dcn || lazy);
   const int jb = io.jb, je = io.je < 0 ? nat : io.je, nj = je - jb;
   if (jb < 0 || je > nat || nj <= 0) fail(MCHRG_B200_ERR_ARG, "invalid row block [%d, %d)", jb, je);
   // one large system on several GPUs: the O(N^2) row passes are row-sharded and joined by one all-reduce; periodic
   // structures share the Ewald pair tiles instead
   const bool shard = pbc ? dist_pbc_apply(ctx, nat) : dist_rows_apply(ctx, nat);
   int row0 = 0, row1 = nat;
   if (shard) row_share(ctx, nat, &row0, &row1);

   double* xvec = ctx->alloc<double>(nat + 1);
   double* thalf = ctx->alloc<double>(nat);
   eeq_xvec(ctx, mol, ap.chi, ap.kcnchi, io.cn, xvec, thalf);

   PbcScope ps;
   double* Jc = nullptr;  // periodic: copy of J for the energy (type.F90:257-259)
   if (pbc) {
      pbc_begin(ctx, mol, ps.host, ps.work);
      ps.on = true;
   }
   auto assemble = [&](double* W, bool first) {
      // first attempt on several GPUs: every rank assembles only the column groups it owns in the factorisation;
      // the shifted retries sum over the whole matrix and need all of it
      if (pbc) pbc_amat(ctx, mol, ap.rad, ap.eta, ps.work, W, shard);
      else eeq_amat_0d(ctx, mol, ap.rad, ap.eta, W, npad, /*border=*/false, npad, first ? dpotrf_dist_panel(ctx, npad) : 0);
   };
   auto keep_copy = [&](double* W) {
      if (pbc && io.energy) {
         Jc = ctx->alloc<double>((size_t)npad * npad);
         MCHRG_CUDA(cudaMemcpyAsync(Jc, W, (size_t)npad * npad * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      }
   };
   // Ewald matrices are indefinite along 1 (omitted G = 0 term): shift up front
   Factor f = factor_solve(ctx, nat, npad, assemble, keep_copy, pbc, xvec, mol.charge, io.qvec, aug);
   if (f.info != 0) {
      ctx->copybacks.clear();
      return f.info;
   }
   if (f.indefinite && cpq)
      fail(MCHRG_B200_ERR_MODEL, "dq/dr is not available for an indefinite Coulomb block (solved through the bordered system)");
   const double* vrhs = f.vrhs;
   const double* zvec = f.z(npad);

   if (io.energy || grad || cpq) {
      // [jq | atrace | dadL | gradx (all atoms) | sig_cn]: one buffer, so that row shards are joined by ONE all-reduce
      const size_t n1 = (size_t)nat;
      double* pack = shard ? ctx->alloc_zero<double>(16 * n1 + 16) : nullptr;
      double* jq = io.energy ? (shard ? pack : ctx->alloc<double>(nat)) : nullptr;
      double* atrace = (grad || cpq) ? (shard ? pack + n1 : ctx->alloc<double>((size_t)3 * nat)) : nullptr;
      double* dadL = (grad || cpq) ? (shard ? pack + 4 * n1 : ctx->alloc<double>((size_t)9 * nat)) : nullptr;
      double* Bm = nullptr;
      double* bz = nullptr;
      int64_t mp = 0;
      if (cpq) {
         mp = round_up(3 * (int64_t)nj + 9, TILE);
         Bm = ctx->alloc_zero<double>((size_t)mp * npad);
         bz = ctx->alloc_zero<double>((size_t)mp);
      }
      if (pbc) {
         // sharded: (J q) by rank 0 alone, pair tiles of the derivatives by all ranks -- partial sums in `pack`
         if (io.energy && (!shard || ctx->rank() == 0)) pbc_symv(ctx, nat, Jc, npad, vrhs, jq);
         if (grad || cpq) pbc_derivs(ctx, mol, ap.rad, vrhs, ps.work, Bm, mp, mp, atrace, dadL, shard, jb, je);
      } else {
         eeq_pair_rows_0d(ctx, mol, ap.rad, ap.eta, vrhs, jq, atrace, dadL, row0, row1);
      }
      double* gradx = nullptr;
      double* sig_cn = nullptr;
      if (grad && lazy) {
         if (shard) {
            gradx = pack + 13 * n1;  // all atoms; the caller's rows are picked below
            sig_cn = pack + 16 * n1;
            cn_gradient_device(ctx, mol, *io.lazy, thalf, vrhs, 0, nat, gradx, sig_cn, row0, row1);
            gradx += 3 * (size_t)jb;
         } else {
            gradx = ctx->alloc_zero<double>((size_t)3 * nj);
            sig_cn = ctx->alloc<double>(9);
            cn_gradient_device(ctx, mol, *io.lazy, thalf, vrhs, jb, je, gradx, sig_cn);
         }
      }
      if (shard) comm_allreduce_sum(ctx, pack, (lazy && grad) ? 16 * n1 + 9 : 13 * n1);
      if (io.energy) MCHRG_LAUNCH(ctx, energy_kernel, (nat + 255) / 256, 256, 0, nat, vrhs, jq, xvec, io.energy);

      if (grad || cpq) {
         if (!lazy) {
            gradx = grad ? ctx->alloc_zero<double>((size_t)3 * nj) : nullptr;
            eeq_build_b_0d(ctx, mol, ap.rad, vrhs, atrace, thalf, io.dcndr, Bm, mp, gradx, zvec, bz, pbc && cpq, jb, je);
         }
         if (grad) {
            MCHRG_LAUNCH(ctx, gradient_kernel, (3 * nj + 255) / 256, 256, 0, jb, nj, vrhs, atrace, gradx, io.gradient);
            MCHRG_LAUNCH(ctx, sigma_kernel, 1, 288, 0, nat, vrhs, dadL, thalf, lazy ? nullptr : io.dcndL, sig_cn, io.sigma);
         }
         if (cpq) {
            MCHRG_LAUNCH(ctx, strain_rows_kernel, 1, 288, 0, nat, dadL, thalf, io.dcndL, zvec, Bm, mp, bz,
                         3 * (int64_t)nj);
            dq_tail(ctx, nat, nj, npad, mp, f, Bm, bz, io.dqdr, io.dqdL);
         }
      }
   }
   ctx->finish();
   return 0;
}

// ---------------------------------------------------------------------------------------------
// EEQ-BC (model/eeqbc.f90), molecular and periodic
// ---------------------------------------------------------------------------------------------
static BcAtoms gather_eeqbc(const mchrg_b200_model* model, const DevMol& mol) {
   mchrg_b200_context* ctx = model->ctx;
   auto g = [&](int slot) {
      double* d = ctx->alloc<double>(mol.nat);
      gather_species(ctx, mol.nat, mol.id, model->dev(slot), d);
      return (const double*)d;
   };
   BcAtoms at;
   at.chi = g(0); at.rad = g(1); at.eta = g(2); at.kcnchi = g(3);
   at.kqchi = g(5); at.kqeta = g(6); at.cap = g(7); at.avg_cn = g(8);
   return at;
}

// gradient = 0.5 ga - gx (type.F90:277-278)
__global__ void bc_gradient_kernel(int n3, const double* __restrict__ ga, const double* __restrict__ gx,
                                   double* __restrict__ gradient) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n3) gradient[i] = 0.5 * ga[i] - gx[i];
}
// sigma += sum_k q_k (0.5 dadL_k - dxdL_k); strain rows of B = dadL_k - dxdL_k and their z contraction
__global__ void __launch_bounds__(288) bc_sigma_kernel(int nat, const double* __restrict__ q,
                                                       const double* __restrict__ dadL, const double* __restrict__ dxdL,
                                                       double* __restrict__ sigma, const double* __restrict__ zvec,
                                                       double* __restrict__ B, int64_t ldb, double* __restrict__ bz,
                                                       int64_t roff) {
   const int ab = threadIdx.x >> 5, lane = threadIdx.x & 31;
   double acc = 0.0, accz = 0.0;
   for (int k = lane; k < nat; k += 32) {
      const double da = dadL[(size_t)9 * k + ab], dx = dxdL[(size_t)9 * k + ab];
      acc += q[k] * (0.5 * da - dx);
      if (B) {
         B[(size_t)k * ldb + roff + ab] = da - dx;
         accz = fma(da - dx, zvec[k], accz);
      }
   }
   for (int o = 16; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      accz += __shfl_xor_sync(0xffffffffu, accz, o);
   }
   if (lane == 0) {
      if (sigma) sigma[ab] += acc;
      if (B) bz[roff + ab] = accz;
   }
}
__global__ void set_last_kernel(int nat, double v, double* __restrict__ x) { x[nat] = v; }

static int solve_eeqbc_device(const mchrg_b200_model* model, const DevMol& mol, const SolveIO& io) {
   mchrg_b200_context* ctx = model->ctx;
   if (!io.qloc) fail(MCHRG_B200_ERR_MODEL, "qloc required for eeqbc");  // error stop, eeqbc.f90:202
   const int nat = mol.nat;
   const int64_t npad = round_up(nat, TILE);
   const int jb = io.jb, je = io.je < 0 ? nat : io.je, nj = je - jb;
   if (jb < 0 || je > nat || nj <= 0) fail(MCHRG_B200_ERR_ARG, "invalid row block [%d, %d)", jb, je);
   if (mol.periodic && nj != nat) fail(MCHRG_B200_ERR_MODEL, "row blocks are not available for periodic structures");
   // one large system on several GPUs: the O(N^2) row passes are row-sharded and joined by one all-reduce
   const bool shard = !mol.periodic && dist_rows_apply(ctx, nat);
   int row0 = 0, row1 = nat;
   if (shard) row_share(ctx, nat, &row0, &row1);
   // eeqbc.f90:195 (update) needs the local-charge derivatives for every derivative output
   const bool dcn = io.dcndr && io.dcndL && io.dqlocdr && io.dqlocdL;
   const bool grad = io.gradient && io.sigma && dcn;
   const bool cpq = io.dqdr && io.dqdL && dcn;

   BcAtoms at = gather_eeqbc(model, mol);
   BcWork w;
   bc_prepare(ctx, mol, model, at, io.cn, io.qloc, grad || cpq, w);
   PbcScope pbc;
   double *cself = nullptr, *Wpbc = nullptr;
   if (mol.periodic) {  // get_cmat_3d + get_amat_3d in one pass over the unique pairs
      pbc_begin(ctx, mol, pbc.host, pbc.work);
      pbc.on = true;
      cself = ctx->alloc<double>(nat);
      Wpbc = ctx->alloc<double>((size_t)npad * npad);
      bc_pbc_matrices(ctx, mol, w.atoms, model->dev_rvdw(), model->nid, model->kbc, pbc.work, w.Cm, Wpbc, cself);
   }
   // xvec = C xtmp (eeqbc.f90:275-282); xtmp lives in w.atoms[5 i + 3]
   double* xtmp = ctx->alloc<double>(nat + 1);
   MCHRG_CUDA(cudaMemcpy2DAsync(xtmp, sizeof(double), w.atoms + 3, 5 * sizeof(double), sizeof(double), nat + 1,
                                cudaMemcpyDeviceToDevice, ctx->stream));
   double* xvec = ctx->alloc<double>(nat + 1);
   pbc_symv(ctx, nat, w.Cm, npad, xtmp, xvec);
   if (mol.periodic) bc_pbc_xvec_fix(ctx, nat, w.atoms, cself, xvec);
   MCHRG_LAUNCH(ctx, set_last_kernel, 1, 1, 0, nat, mol.charge, xvec);

   double* Jc = nullptr;
   auto assemble = [&](double* W, bool first) {
      if (mol.periodic)
         MCHRG_CUDA(cudaMemcpyAsync(W, Wpbc, (size_t)npad * npad * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      else  // several GPUs: only the column groups this rank owns in the factorisation (the shifted retries need all)
         bc_amat(ctx, mol, w, W, first ? dpotrf_dist_panel(ctx, npad) : 0);
   };
   auto keep_copy = [&](double* W) {
      if (io.energy && mol.periodic) {  // molecular: (A q) is re-evaluated from the pair formula (bc_jq_rows)
         Jc = ctx->alloc<double>((size_t)npad * npad);
         MCHRG_CUDA(cudaMemcpyAsync(Jc, W, (size_t)npad * npad * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      }
   };
   Factor f = factor_solve(ctx, nat, npad, assemble, keep_copy, false, xvec, mol.charge, io.qvec);
   if (f.info != 0) {
      ctx->copybacks.clear();
      return f.info;
   }
   if (f.indefinite && cpq)
      fail(MCHRG_B200_ERR_MODEL, "dq/dr is not available for an indefinite Coulomb block (solved through the bordered system)");
   const double* q = f.vrhs;
   // molecular: [jq | rows (24 N) | dadL (9 N) | dxdL (9 N)] in one buffer, so that row shards are joined by ONE all-reduce
   const size_t n1 = (size_t)nat;
   double* pack = (!mol.periodic && (io.energy || grad || cpq)) ? ctx->alloc_zero<double>(43 * n1) : nullptr;
   if (io.energy && mol.periodic) {
      double* jq = ctx->alloc<double>(nat);
      pbc_symv(ctx, nat, Jc, npad, q, jq);
      MCHRG_LAUNCH(ctx, energy_kernel, (nat + 255) / 256, 256, 0, nat, q, jq, xvec, io.energy);
   }
   if ((grad || cpq) && mol.periodic) {
      // get_damat_3d / get_dcmat_3d / periodic get_xvec_derivs: everything is assembled by bc_pbc_derivs (eeq_pbc.cu)
      double* dadL = ctx->alloc_zero<double>((size_t)9 * nat);
      double* dxdL = ctx->alloc_zero<double>((size_t)9 * nat);
      double* cq = ctx->alloc<double>(nat);
      pbc_symv(ctx, nat, w.Cm, npad, q, cq);
      double *Bm = nullptr, *bz = nullptr;
      int64_t mp = 0;
      if (cpq) {
         mp = round_up(3 * (int64_t)nat + 9, TILE);
         Bm = ctx->alloc<double>((size_t)mp * npad);
         bz = ctx->alloc_zero<double>((size_t)mp);
         bc_dxdr_gemm(ctx, mol, at, w, io.dcndr, io.dqlocdr, -1.0, Bm, mp);  // Bm = -dtmpdr C
      }
      double* ga = ctx->alloc_zero<double>((size_t)3 * nat);
      double* gx = ctx->alloc_zero<double>((size_t)3 * nat);
      bc_pbc_derivs(ctx, mol, w.atoms, model->dev_rvdw(), model->nid, model->kbc, at.kcnchi, at.kqchi, at.kqeta, q, cq, w.Cm,
                    cself, io.dcndr, io.dcndL, io.dqlocdr, io.dqlocdL, f.z(npad), pbc.work, Bm, mp, ga, gx, dadL, dxdL, bz);
      if (grad) MCHRG_LAUNCH(ctx, bc_gradient_kernel, (3 * nat + 255) / 256, 256, 0, 3 * nat, ga, gx, io.gradient);
      MCHRG_LAUNCH(ctx, bc_sigma_kernel, 1, 288, 0, nat, q, dadL, dxdL, grad ? io.sigma : nullptr, f.z(npad), Bm, mp, bz,
                   3 * (int64_t)nat);
      if (cpq) dq_tail(ctx, nat, nat, npad, mp, f, Bm, bz, io.dqdr, io.dqdL);
   } else if (!mol.periodic && (io.energy || grad || cpq)) {
      double *jq = pack, *rows = pack + n1, *dadL = pack + 25 * n1, *dxdL = pack + 34 * n1;
      if (io.energy) bc_jq_rows(ctx, mol, w, q, jq, row0, row1);
      if (grad || cpq) {
         bc_rows(ctx, mol, model, w, q, io.dcndr, io.dcndL, rows, row0, row1);
         bc_strain(ctx, mol, at, w, q, rows, io.dcndL, io.dqlocdL, dadL, dxdL, row0, row1);
      }
      if (shard) comm_allreduce_sum(ctx, pack, (grad || cpq) ? 43 * n1 : n1);
      if (io.energy) MCHRG_LAUNCH(ctx, energy_kernel, (nat + 255) / 256, 256, 0, nat, q, jq, xvec, io.energy);
      if (grad || cpq) {
         double* cq = ctx->alloc<double>(nat);
         pbc_symv(ctx, nat, w.Cm, npad, q, cq);
         double *Bm = nullptr, *bz = nullptr;
         int64_t mp = 0;
         if (cpq) {
            mp = round_up(3 * (int64_t)nj + 9, TILE);
            Bm = ctx->alloc<double>((size_t)mp * npad);
            bz = ctx->alloc_zero<double>((size_t)mp);
            bc_dxdr_gemm(ctx, mol, at, w, io.dcndr, io.dqlocdr, -1.0, Bm, mp, jb, je);  // Bm = -dtmpdr C, rows of the block
         }
         // several GPUs, gradient of ALL atoms wanted on every rank (no dq/dr): the pair pass behind the gradient rows
         // is shared like the other row passes, one all-reduce of the 3 N values joins it
         const bool gshard = shard && grad && !cpq && nj == nat;
         const int gb = gshard ? row0 : jb, ge = gshard ? row1 : je, ng = ge - gb;
         double* ga = (grad && ng > 0) ? ctx->alloc_zero<double>((size_t)3 * ng) : nullptr;
         double* gx = (grad && ng > 0) ? ctx->alloc_zero<double>((size_t)3 * ng) : nullptr;
         if (ng > 0)
            bc_build_b(ctx, mol, model, at, w, q, cq, rows, io.dcndr, io.dqlocdr, 1.0, 1.0, true, Bm, mp, ga, gx, f.z(npad), bz,
                       gb, ge);
         if (gshard) {
            double* gall = ctx->alloc_zero<double>((size_t)3 * nat);
            if (ng > 0) MCHRG_LAUNCH(ctx, bc_gradient_kernel, (3 * ng + 255) / 256, 256, 0, 3 * ng, ga, gx, gall + 3 * (size_t)gb);
            comm_allreduce_sum(ctx, gall, (size_t)3 * nat);
            MCHRG_CUDA(cudaMemcpyAsync(io.gradient, gall, (size_t)3 * nat * sizeof(double), cudaMemcpyDeviceToDevice, ctx->active));
         } else if (grad) {
            MCHRG_LAUNCH(ctx, bc_gradient_kernel, (3 * nj + 255) / 256, 256, 0, 3 * nj, ga, gx, io.gradient);
         }
         MCHRG_LAUNCH(ctx, bc_sigma_kernel, 1, 288, 0, nat, q, dadL, dxdL, grad ? io.sigma : nullptr, f.z(npad), Bm, mp, bz,
                      3 * (int64_t)nj);
         if (cpq) dq_tail(ctx, nat, nj, npad, mp, f, Bm, bz, io.dqdr, io.dqdL);
      }
   }
   ctx->finish();
   return 0;
}

static int solve_dispatch(const mchrg_b200_model* model, const DevMol& mol, const SolveIO& io) {
   if (io.cn == nullptr) fail(MCHRG_B200_ERR_ARG, "cn is required");
   if (model->kind == 1) {
      AtomPar ap = gather_eeq(model, mol);
      return solve_eeq_device(model, mol, ap, io);
   }
   return solve_eeqbc_device(model, mol, io);
}

// qloc += charge / nat (type.F90:320)
__global__ void add_const_kernel(int n, double v, double* __restrict__ x) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) x[i] += v;
}

// translations for the coordination number: get_lattice_points(periodic, lattice, cutoff) (charge.f90:70)
static const double* cn_translations(mchrg_b200_context* ctx, const mchrg_b200_structure* mol, double cutoff, int* ntr) {
   std::vector<double> t = lat::points_cutoff(mol->periodic, mol->lattice, cutoff);
   *ntr = (int)(t.size() / 3);
   double* d = ctx->alloc<double>(t.size());
   MCHRG_CUDA(cudaMemcpyAsync(d, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
   MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));  // t goes out of scope
   return d;
}

}  // namespace mchrg

// =============================================================================================
extern "C" {

static int new_model_common(mchrg_b200_context* ctx, mchrg_b200_model* m, mchrg_b200_model** out) {
   return guarded(ctx, [&]() {
      const int nid = m->nid;
      std::vector<double> host((size_t)10 * nid + (size_t)nid * nid, 0.0);
      auto put = [&](int slot, const std::vector<double>& v) {
         for (size_t i = 0; i < v.size() && i < (size_t)nid; ++i) host[(size_t)slot * nid + i] = v[i];
      };
      put(0, m->chi); put(1, m->rad); put(2, m->eta); put(3, m->kcnchi); put(4, m->rcov);
      put(5, m->kqchi); put(6, m->kqeta); put(7, m->cap); put(8, m->avg_cn); put(9, m->en);
      for (size_t i = 0; i < m->rvdw.size(); ++i) host[(size_t)10 * nid + i] = m->rvdw[i];
      if (cudaMalloc(&m->d_par, host.size() * sizeof(double)) != cudaSuccess) {
         cudaGetLastError();
         fail(MCHRG_B200_ERR_ALLOC, "model parameter allocation failed");
      }
      MCHRG_CUDA(cudaMemcpy(m->d_par, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
      *out = m;
      return 0;
   });
}

int mchrg_b200_new_eeq_model(mchrg_b200_context* ctx, int32_t nid, const double* chi, const double* rad,
                             const double* eta, const double* kcnchi, const double* rcov, double cutoff,
                             double cn_exp, double cn_max, mchrg_b200_model** model) {
   if (!ctx || !model || nid <= 0 || !chi || !rad || !eta || !kcnchi || !rcov) {
      set_last_error("new_eeq_model: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   auto* m = new mchrg_b200_model();
   m->ctx = ctx;
   m->kind = 1;
   m->nid = nid;
   m->chi.assign(chi, chi + nid);
   m->rad.assign(rad, rad + nid);
   m->eta.assign(eta, eta + nid);
   m->kcnchi.assign(kcnchi, kcnchi + nid);
   m->rcov.assign(rcov, rcov + nid);
   m->ncoord.cutoff = cutoff;
   m->ncoord.kcn = cn_exp;
   m->ncoord.cut = cn_max;
   m->ncoord.norm_exp = 1.0;
   int rc = new_model_common(ctx, m, model);
   if (rc != 0) delete m;
   return rc;
}

int mchrg_b200_new_eeq2019_model(mchrg_b200_context* ctx, int32_t nid, const int32_t* num,
                                 mchrg_b200_model** model) {
   if (!ctx || !model || nid <= 0 || !num) {
      set_last_error("new_eeq2019_model: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   const double aatoau = 1.0 / 0.529177210903;
   std::vector<double> chi(nid), rad(nid), eta(nid), kcnchi(nid), rcov(nid);
   for (int i = 0; i < nid; ++i) {
      const int z = num[i];
      if (z < 1 || z > 103) {  // the reference getters return -1 for unknown elements (eeq2019.f90:174-178)
         chi[i] = rad[i] = eta[i] = kcnchi[i] = -1.0;
         rcov[i] = 0.0;
         continue;
      }
      chi[i] = mchrg_eeq_chi[z - 1];
      eta[i] = mchrg_eeq_eta[z - 1];
      kcnchi[i] = mchrg_eeq_kcnchi[z - 1];
      rad[i] = mchrg_eeq_rad[z - 1];
      rcov[i] = mchrg_pa2009_covrad_aa[z - 1] * aatoau * 4.0 / 3.0;  // get_covalent_rad, D3 convention
   }
   return mchrg_b200_new_eeq_model(ctx, nid, chi.data(), rad.data(), eta.data(), kcnchi.data(), rcov.data(), 25.0,
                                   7.5, 8.0, model);
}

int mchrg_b200_new_eeqbc_model(mchrg_b200_context* ctx, int32_t nid, const double* chi, const double* rad,
                               const double* eta, const double* kcnchi, const double* kqchi, const double* kqeta,
                               double kcnrad, const double* cap, const double* avg_cn, const double* rvdw, double kbc,
                               double cutoff, double cn_exp, const double* rcov, const double* en, double norm_exp,
                               mchrg_b200_model** model) {
   if (!ctx || !model || nid <= 0 || !chi || !rad || !eta || !kcnchi || !kqchi || !kqeta || !cap || !avg_cn || !rvdw ||
       !rcov || !en) {
      set_last_error("new_eeqbc_model: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   auto* m = new mchrg_b200_model();
   m->ctx = ctx;
   m->kind = 2;
   m->nid = nid;
   m->chi.assign(chi, chi + nid);
   m->rad.assign(rad, rad + nid);
   m->eta.assign(eta, eta + nid);
   m->kcnchi.assign(kcnchi, kcnchi + nid);
   m->kqchi.assign(kqchi, kqchi + nid);
   m->kqeta.assign(kqeta, kqeta + nid);
   m->cap.assign(cap, cap + nid);
   m->avg_cn.assign(avg_cn, avg_cn + nid);
   m->rvdw.assign(rvdw, rvdw + (size_t)nid * nid);
   m->rcov.assign(rcov, rcov + nid);
   m->en.assign(en, en + nid);
   m->kcnrad = kcnrad;
   m->kbc = kbc;
   m->norm_exp = norm_exp;
   m->ncoord.cutoff = m->ncoord_en.cutoff = cutoff;
   m->ncoord.kcn = m->ncoord_en.kcn = cn_exp;
   m->ncoord.norm_exp = m->ncoord_en.norm_exp = norm_exp;
   m->ncoord.cut = m->ncoord_en.cut = -1.0;  // cn_max is absent in new_eeqbc2025_model (param.f90:118-122)
   int rc = new_model_common(ctx, m, model);
   if (rc != 0) delete m;
   return rc;
}

int mchrg_b200_new_eeqbc2025_model(mchrg_b200_context* ctx, int32_t nid, const int32_t* num, const double* rvdw,
                                   const double* en_pauling, mchrg_b200_model** model) {
   if (!ctx || !model || nid <= 0 || !num || !rvdw || !en_pauling) {
      set_last_error("new_eeqbc2025_model: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   std::vector<double> chi(nid), rad(nid), eta(nid), kcnchi(nid), kqchi(nid), kqeta(nid), cap(nid), avg_cn(nid),
      rcov(nid), en(nid);
   for (int i = 0; i < nid; ++i) {
      const int z = num[i];
      if (z < 1 || z > 103) {
         set_last_error("new_eeqbc2025_model: atomic number out of range");
         return MCHRG_B200_ERR_ARG;
      }
      chi[i] = mchrg_eeqbc_chi[z - 1];
      eta[i] = mchrg_eeqbc_eta[z - 1];
      rad[i] = mchrg_eeqbc_rad[z - 1];
      kcnchi[i] = mchrg_eeqbc_kcnchi[z - 1];
      kqchi[i] = mchrg_eeqbc_kqchi[z - 1];
      kqeta[i] = mchrg_eeqbc_kqeta[z - 1];
      cap[i] = mchrg_eeqbc_cap[z - 1];
      rcov[i] = mchrg_eeqbc_cov_radii[z - 1];
      avg_cn[i] = mchrg_eeqbc_avg_cn[z - 1];
      // Pauling EN normalised to fluorine, actinide overrides of param.f90:105-113
      double e = en_pauling[i];
      if (z >= 90) e = 1.30;
      if (z == 87) e = 0.80;
      if (z == 89) e = 1.00;
      if (z == 90 || z == 91 || z == 92 || z == 95) e = 1.10;
      if (z == 93 || z == 94 || z == 97 || z == 103) e = 1.20;
      en[i] = e / 3.98;
   }
   return mchrg_b200_new_eeqbc_model(ctx, nid, chi.data(), rad.data(), eta.data(), kcnchi.data(), kqchi.data(),
                                     kqeta.data(), 0.14, cap.data(), avg_cn.data(), rvdw, 0.60, 25.0, 2.0, rcov.data(),
                                     en.data(), 0.75, model);
}

void mchrg_b200_model_free(mchrg_b200_model* model) {
   if (!model) return;
   if (model->ctx) cudaSetDevice(model->ctx->device);
   if (model->d_par) cudaFree(model->d_par);
   delete model;
}

int mchrg_b200_model_kind(const mchrg_b200_model* model) { return model ? model->kind : 0; }

// ---------------------------------------------------------------------------------------------
int mchrg_b200_get_coordination_number(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* cn,
                                       double* dcndr, double* dcndL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!cn || ((dcndr == nullptr) != (dcndL == nullptr))) fail(MCHRG_B200_ERR_ARG, "get_coordination_number: bad outputs");
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      int ntr = 0;
      const double* trans = cn_translations(ctx, mol, model->ncoord.cutoff, &ntr);
      double* rcov_a = ctx->alloc<double>(n);
      gather_species(ctx, d.nat, d.id, model->dev(4), rcov_a);
      ncoord_device(ctx, d, rcov_a, nullptr, model->ncoord, ntr, trans, ctx->out(cn, n), ctx->out(dcndr, 3 * n * n),
                    ctx->out(dcndL, 9 * n));
      return 0;
   });
}

int mchrg_b200_local_charge(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* qloc,
                            double* dqlocdr, double* dqlocdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!qloc) fail(MCHRG_B200_ERR_ARG, "local_charge: qloc is required");
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      double* dq = ctx->out(qloc, n);
      if (model->kind == 1) {  // no ncoord_en: qloc = charge / nat, derivatives zero (type.F90:308-320)
         MCHRG_LAUNCH(ctx, fill_kernel, (unsigned)((n + 255) / 256), 256, 0, (int64_t)n, d.charge / (double)n, dq);
         if (dqlocdr && dqlocdL) {
            MCHRG_CUDA(cudaMemsetAsync(ctx->out(dqlocdr, 3 * n * n), 0, 3 * n * n * sizeof(double), ctx->stream));
            MCHRG_CUDA(cudaMemsetAsync(ctx->out(dqlocdL, 9 * n), 0, 9 * n * sizeof(double), ctx->stream));
         }
         return 0;
      }
      // EEQ-BC: electronegativity-weighted coordination number (cn_count%erf_en) + charge / nat
      if ((dqlocdr == nullptr) != (dqlocdL == nullptr)) fail(MCHRG_B200_ERR_ARG, "local_charge: bad outputs");
      int ntr = 0;
      const double* trans = cn_translations(ctx, mol, model->ncoord_en.cutoff, &ntr);
      double* rcov_a = ctx->alloc<double>(n);
      double* en_a = ctx->alloc<double>(n);
      gather_species(ctx, d.nat, d.id, model->dev(4), rcov_a);
      gather_species(ctx, d.nat, d.id, model->dev(9), en_a);
      ncoord_device(ctx, d, rcov_a, en_a, model->ncoord_en, ntr, trans, dq, ctx->out(dqlocdr, 3 * n * n),
                    ctx->out(dqlocdL, 9 * n));
      MCHRG_LAUNCH(ctx, add_const_kernel, (unsigned)((n + 255) / 256), 256, 0, d.nat, d.charge / (double)n, dq);
      return 0;
   });
}

int mchrg_b200_get_xvec(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                        const double* qloc, double* xvec) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!cn || !xvec) fail(MCHRG_B200_ERR_ARG, "get_xvec: cn and xvec are required");
      DevMol d = stage_mol(ctx, mol);
      if (model->kind == 2) {
         if (!qloc) fail(MCHRG_B200_ERR_MODEL, "qloc required for eeqbc");
         const size_t n = d.nat;
         BcAtoms at = gather_eeqbc(model, d);
         BcWork w;
         bc_prepare(ctx, d, model, at, ctx->in(cn, n), ctx->in(qloc, n), false, w);
         PbcScope ps;
         double* cself = nullptr;
         if (d.periodic) {
            pbc_begin(ctx, d, ps.host, ps.work);
            ps.on = true;
            cself = ctx->alloc<double>(n);
            double* W = ctx->alloc<double>((size_t)w.npad * w.npad);
            bc_pbc_matrices(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, ps.work, w.Cm, W, cself);
         }
         double* xtmp = ctx->alloc<double>(n + 1);
         MCHRG_CUDA(cudaMemcpy2DAsync(xtmp, sizeof(double), w.atoms + 3, 5 * sizeof(double), sizeof(double), n + 1,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
         double* xv = ctx->out(xvec, n + 1);
         pbc_symv(ctx, d.nat, w.Cm, w.npad, xtmp, xv);
         if (d.periodic) bc_pbc_xvec_fix(ctx, d.nat, w.atoms, cself, xv);
         MCHRG_LAUNCH(ctx, set_last_kernel, 1, 1, 0, d.nat, d.charge, xv);
         if (d.periodic) ctx->finish();  // host setup arrays of `ps` are consumed by async copies
         return 0;
      }
      AtomPar ap = gather_eeq(model, d);
      eeq_xvec(ctx, d, ap.chi, ap.kcnchi, ctx->in(cn, (size_t)d.nat), ctx->out(xvec, (size_t)d.nat + 1), nullptr);
      return 0;
   });
}

int mchrg_b200_get_xvec_derivs(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                               const double* qloc, const double* dcndr, const double* dcndL, const double* dqlocdr,
                               const double* dqlocdL, double* dxdr, double* dxdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!cn || !dcndr || !dcndL || !dxdr || !dxdL) fail(MCHRG_B200_ERR_ARG, "get_xvec_derivs: missing argument");
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      if (model->kind == 2) {
         if (!qloc || !dqlocdr || !dqlocdL) fail(MCHRG_B200_ERR_MODEL, "qloc and its derivatives required for eeqbc");
         BcAtoms at = gather_eeqbc(model, d);
         BcWork w;
         const double* dcn = ctx->in(dcndr, 3 * n * n);
         const double* dcl = ctx->in(dcndL, 9 * n);
         const double* dqr = ctx->in(dqlocdr, 3 * n * n);
         const double* dql = ctx->in(dqlocdL, 9 * n);
         bc_prepare(ctx, d, model, at, ctx->in(cn, n), ctx->in(qloc, n), true, w);
         const int64_t mp = round_up(3 * (int64_t)n + 9, TILE);
         double* zeros = ctx->alloc_zero<double>(n + 1);  // q = 0: only the q-independent (dxdr) parts survive
         if (d.periodic) {  // periodic branch of get_xvec_derivs (eeqbc.f90:364-418)
            PbcScope ps;
            pbc_begin(ctx, d, ps.host, ps.work);
            ps.on = true;
            double* cself = ctx->alloc<double>(n);
            double* Wtmp = ctx->alloc<double>((size_t)w.npad * w.npad);
            bc_pbc_matrices(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, ps.work, w.Cm, Wtmp, cself);
            double* Bm = ctx->alloc<double>((size_t)mp * w.npad);
            bc_dxdr_gemm(ctx, d, at, w, dcn, dqr, 1.0, Bm, mp);  // Bm = +dtmpdr C
            double* ga = ctx->alloc_zero<double>(3 * n);
            double* gx = ctx->alloc_zero<double>(3 * n);
            double* dadL_tmp = ctx->alloc_zero<double>(9 * n);
            double* d_dxdL = ctx->out(dxdL, 9 * (n + 1));
            MCHRG_CUDA(cudaMemsetAsync(d_dxdL, 0, 9 * (n + 1) * sizeof(double), ctx->stream));
            bc_pbc_derivs(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, at.kcnchi, at.kqchi, at.kqeta, zeros, zeros,
                          w.Cm, cself, dcn, dcl, dqr, dql, nullptr, ps.work, Bm, mp, ga, gx, dadL_tmp, d_dxdL, nullptr, 0.0,
                          -1.0, false, nullptr);
            double* d_dxdr = ctx->out(dxdr, 3 * n * (n + 1));
            MCHRG_LAUNCH(ctx, unpad_copy_kernel, dim3((unsigned)((3 * n + 255) / 256), (unsigned)n), 256, 0, (int64_t)(3 * n),
                         (int64_t)n, Bm, mp, d_dxdr, (int64_t)(3 * n), 0);
            MCHRG_CUDA(cudaMemsetAsync(d_dxdr + 3 * n * n, 0, 3 * n * sizeof(double), ctx->stream));
            ctx->finish();  // host setup arrays of `ps` are consumed by async copies
            return 0;
         }
         double* rows = ctx->alloc<double>(24 * n);
         bc_rows(ctx, d, model, w, zeros, dcn, dcl, rows);
         double* dadL_tmp = ctx->alloc<double>(9 * n);
         double* d_dxdL = ctx->out(dxdL, 9 * (n + 1));
         bc_strain(ctx, d, at, w, zeros, rows, dcl, dql, dadL_tmp, d_dxdL);
         MCHRG_CUDA(cudaMemsetAsync(d_dxdL + 9 * n, 0, 9 * sizeof(double), ctx->stream));
         double* Bm = ctx->alloc<double>((size_t)mp * w.npad);
         bc_dxdr_gemm(ctx, d, at, w, dcn, dqr, 1.0, Bm, mp);  // Bm = +dtmpdr C
         bc_build_b(ctx, d, model, at, w, zeros, nullptr, rows, dcn, dqr, 0.0, -1.0, false, Bm, mp, nullptr, nullptr,
                    nullptr, nullptr);
         double* d_dxdr = ctx->out(dxdr, 3 * n * (n + 1));
         MCHRG_LAUNCH(ctx, unpad_copy_kernel, dim3((unsigned)((3 * n + 255) / 256), (unsigned)n), 256, 0, (int64_t)(3 * n),
                      (int64_t)n, Bm, mp, d_dxdr, (int64_t)(3 * n), 0);
         MCHRG_CUDA(cudaMemsetAsync(d_dxdr + 3 * n * n, 0, 3 * n * sizeof(double), ctx->stream));
         return 0;
      }
      AtomPar ap = gather_eeq(model, d);
      double* xv = ctx->alloc<double>(n + 1);
      double* thalf = ctx->alloc<double>(n);
      eeq_xvec(ctx, d, ap.chi, ap.kcnchi, ctx->in(cn, n), xv, thalf);
      const int64_t total = 3 * (int64_t)n * (n + 1);
      MCHRG_LAUNCH(ctx, xvec_derivs_kernel, (unsigned)((total + 255) / 256), 256, 0, d.nat, thalf,
                   ctx->in(dcndr, 3 * n * n), ctx->in(dcndL, 9 * n), ctx->out(dxdr, 3 * n * (n + 1)),
                   ctx->out(dxdL, 9 * (n + 1)));
      return 0;
   });
}

int mchrg_b200_get_coulomb_matrix(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                                  const double* qloc, double* amat) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!amat) fail(MCHRG_B200_ERR_ARG, "get_coulomb_matrix: amat is required");
      DevMol d = stage_mol(ctx, mol);
      const size_t nd = (size_t)d.nat + 1;
      if (model->kind == 2) {
         if (!cn || !qloc) fail(MCHRG_B200_ERR_MODEL, "cn and qloc required for eeqbc");
         BcAtoms at = gather_eeqbc(model, d);
         BcWork w;
         bc_prepare(ctx, d, model, at, ctx->in(cn, (size_t)d.nat), ctx->in(qloc, (size_t)d.nat), false, w);
         double* W = ctx->alloc<double>((size_t)w.npad * w.npad);
         PbcScope ps;
         if (d.periodic) {
            pbc_begin(ctx, d, ps.host, ps.work);
            ps.on = true;
            double* cself = ctx->alloc<double>((size_t)d.nat);
            bc_pbc_matrices(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, ps.work, w.Cm, W, cself);
         } else {
            bc_amat(ctx, d, w, W);
         }
         MCHRG_LAUNCH(ctx, border_copy_kernel, dim3((unsigned)((nd + 255) / 256), (unsigned)nd), 256, 0, d.nat, W, w.npad,
                      ctx->out(amat, nd * nd));
         if (d.periodic) ctx->finish();  // host setup arrays of `ps` are consumed by async copies
         return 0;
      }
      AtomPar ap = gather_eeq(model, d);
      double* out = ctx->out(amat, nd * nd);
      if (d.periodic) {
         PbcScope ps;
         pbc_begin(ctx, d, ps.host, ps.work);
         ps.on = true;
         const int64_t npad = ps.work.npad;
         double* W = ctx->alloc<double>((size_t)npad * npad);
         pbc_amat(ctx, d, ap.rad, ap.eta, ps.work, W);
         MCHRG_LAUNCH(ctx, border_copy_kernel, dim3((unsigned)((nd + 255) / 256), (unsigned)nd), 256, 0, d.nat, W, npad, out);
         ctx->finish();  // host setup arrays of `ps` are consumed by async copies
      } else {
         eeq_amat_0d(ctx, d, ap.rad, ap.eta, out, (int64_t)nd, true, 0);
      }
      return 0;
   });
}

int mchrg_b200_get_coulomb_derivs(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                                  const double* qloc, const double* dcndr, const double* dcndL,
                                  const double* dqlocdr, const double* dqlocdL, const double* qvec, double* dadr,
                                  double* dadL, double* atrace) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (!qvec || !dadr || !dadL || !atrace) fail(MCHRG_B200_ERR_ARG, "get_coulomb_derivs: missing argument");
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      if (model->kind == 2) {
         if (!cn || !qloc || !dcndr || !dcndL || !dqlocdr || !dqlocdL)
            fail(MCHRG_B200_ERR_MODEL, "cn, qloc and their derivatives required for eeqbc");
         BcAtoms at = gather_eeqbc(model, d);
         BcWork w;
         const double* dcn = ctx->in(dcndr, 3 * n * n);
         const double* dcl = ctx->in(dcndL, 9 * n);
         const double* dqr = ctx->in(dqlocdr, 3 * n * n);
         const double* dql = ctx->in(dqlocdL, 9 * n);
         const double* q = ctx->in(qvec, n);
         bc_prepare(ctx, d, model, at, ctx->in(cn, n), ctx->in(qloc, n), true, w);
         if (d.periodic) {  // get_damat_3d (eeqbc.f90:789-924)
            PbcScope ps;
            pbc_begin(ctx, d, ps.host, ps.work);
            ps.on = true;
            const int64_t mp = round_up(3 * (int64_t)n + 9, TILE);
            double* cself = ctx->alloc<double>(n);
            double* Wtmp = ctx->alloc<double>((size_t)w.npad * w.npad);
            bc_pbc_matrices(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, ps.work, w.Cm, Wtmp, cself);
            double* cq = ctx->alloc<double>(n);
            pbc_symv(ctx, d.nat, w.Cm, w.npad, q, cq);
            double* Bm = ctx->alloc_zero<double>((size_t)mp * w.npad);
            double* ga = ctx->alloc_zero<double>(3 * n);
            double* gx = ctx->alloc_zero<double>(3 * n);
            double* dxdL_tmp = ctx->alloc_zero<double>(9 * n);
            double* d_dadL = ctx->out(dadL, 9 * (n + 1));
            MCHRG_CUDA(cudaMemsetAsync(d_dadL, 0, 9 * (n + 1) * sizeof(double), ctx->stream));
            double* d_atr = ctx->out(atrace, 3 * n);
            MCHRG_CUDA(cudaMemsetAsync(d_atr, 0, 3 * n * sizeof(double), ctx->stream));
            bc_pbc_derivs(ctx, d, w.atoms, model->dev_rvdw(), model->nid, model->kbc, at.kcnchi, at.kqchi, at.kqeta, q, cq, w.Cm,
                          cself, dcn, dcl, dqr, dql, nullptr, ps.work, Bm, mp, ga, gx, d_dadL, dxdL_tmp, nullptr, 1.0, 0.0,
                          false, d_atr);
            double* d_dadr = ctx->out(dadr, 3 * n * (n + 1));
            MCHRG_LAUNCH(ctx, unpad_copy_kernel, dim3((unsigned)((3 * n + 255) / 256), (unsigned)n), 256, 0, (int64_t)(3 * n),
                         (int64_t)n, Bm, mp, d_dadr, (int64_t)(3 * n), 0);
            MCHRG_CUDA(cudaMemsetAsync(d_dadr + 3 * n * n, 0, 3 * n * sizeof(double), ctx->stream));
            ctx->finish();  // host setup arrays of `ps` are consumed by async copies
            return 0;
         }
         double* rows = ctx->alloc<double>(24 * n);
         bc_rows(ctx, d, model, w, q, dcn, dcl, rows);
         double* d_dadL = ctx->out(dadL, 9 * (n + 1));
         double* dxdL_tmp = ctx->alloc<double>(9 * n);
         bc_strain(ctx, d, at, w, q, rows, dcl, dql, d_dadL, dxdL_tmp);
         MCHRG_CUDA(cudaMemsetAsync(d_dadL + 9 * n, 0, 9 * sizeof(double), ctx->stream));
         double* d_dadr = ctx->out(dadr, 3 * n * (n + 1));
         bc_build_b(ctx, d, model, at, w, q, nullptr, rows, dcn, dqr, 1.0, 0.0, false, d_dadr, 3 * (int64_t)n, nullptr,
                    nullptr, nullptr, nullptr);
         MCHRG_CUDA(cudaMemsetAsync(d_dadr + 3 * n * n, 0, 3 * n * sizeof(double), ctx->stream));
         double* d_atr = ctx->out(atrace, 3 * n);
         MCHRG_CUDA(cudaMemcpy2DAsync(d_atr, 3 * sizeof(double), rows, 24 * sizeof(double), 3 * sizeof(double), n,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
         return 0;
      }
      AtomPar ap = gather_eeq(model, d);
      const double* q = ctx->in(qvec, n);
      double* d_dadr = ctx->out(dadr, 3 * n * (n + 1));
      double* d_dadL = ctx->out(dadL, 9 * (n + 1));
      double* d_atr = ctx->out(atrace, 3 * n);
      // last column (the constraint) is zero in both arrays (eeq.f90:388-389)
      MCHRG_CUDA(cudaMemsetAsync(d_dadr + 3 * n * n, 0, 3 * n * sizeof(double), ctx->stream));
      MCHRG_CUDA(cudaMemsetAsync(d_dadL + 9 * n, 0, 9 * sizeof(double), ctx->stream));
      if (d.periodic) {
         PbcScope ps;
         pbc_begin(ctx, d, ps.host, ps.work);
         ps.on = true;
         const int64_t npad = ps.work.npad, mp = round_up(3 * (int64_t)n + 9, TILE);
         double* Bm = ctx->alloc_zero<double>((size_t)mp * npad);
         pbc_derivs(ctx, d, ap.rad, q, ps.work, Bm, mp, mp, d_atr, d_dadL);
         MCHRG_LAUNCH(ctx, dadr_copy_kernel, dim3((unsigned)((3 * n + 255) / 256), (unsigned)(n + 1)), 256, 0, d.nat, Bm, mp,
                      d_dadr);
         ctx->finish();
      } else {
         eeq_pair_rows_0d(ctx, d, ap.rad, ap.eta, q, nullptr, d_atr, d_dadL);
         eeq_build_b_0d(ctx, d, ap.rad, q, nullptr, nullptr, nullptr, d_dadr, 3 * (int64_t)n, nullptr, nullptr, nullptr);
      }
      return 0;
   });
}

// ---------------------------------------------------------------------------------------------
int mchrg_b200_solve(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                     const double* qloc, const double* dcndr, const double* dcndL, const double* dqlocdr,
                     const double* dqlocdL, double* energy, double* gradient, double* sigma, double* qvec,
                     double* dqdr, double* dqdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      SolveIO io;
      io.cn = ctx->in(cn, n);
      io.qloc = ctx->in(qloc, n);
      io.dcndr = ctx->in(dcndr, 3 * n * n);
      io.dcndL = ctx->in(dcndL, 9 * n);
      io.dqlocdr = ctx->in(dqlocdr, 3 * n * n);
      io.dqlocdL = ctx->in(dqlocdL, 9 * n);
      // presence rules of type.F90:199-201 (EEQ-BC: eeqbc.f90:195 needs the local-charge derivatives as well): an
      // output that will not be produced is not staged, so a host array stays untouched like the reference's
      // intent(inout) arguments
      const bool dcn = dcndr && dcndL && (model->kind == 1 || (dqlocdr && dqlocdL));
      const bool grad = gradient && sigma && dcn, cpq = dqdr && dqdL && dcn;
      io.energy = ctx->out(energy, n, /*preload=*/true);
      io.gradient = grad ? ctx->out(gradient, 3 * n) : nullptr;
      io.sigma = grad ? ctx->out(sigma, 9, /*preload=*/true) : nullptr;
      io.qvec = ctx->out(qvec, n);
      io.dqdr = cpq ? ctx->out(dqdr, 3 * n * n) : nullptr;
      io.dqdL = cpq ? ctx->out(dqdL, 9 * n) : nullptr;
      return solve_dispatch(model, d, io);
   });
}

// MCHRG_B200_CN_DERIVS=dense keeps the (3, N, N) CN derivative array on the gradient-only path (A/B timing, tests)
static bool lazy_cn_enabled() {
   const char* e = std::getenv("MCHRG_B200_CN_DERIVS");
   return !(e && e[0] == 'd');
}

static bool small_fused_enabled() {
   const char* e = std::getenv("MCHRG_B200_SMALL_FUSED");
   return !(e && e[0] == '0');
}

static int singlepoint_impl(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* cn, double* qloc,
                            double* energy, double* gradient, double* sigma, double* qvec, double* dqdr,
                            double* dqdL, bool accumulate, int jb = 0, int je = -1) {
   mchrg_b200_context* ctx = model->ctx;
   // Charges only, of a small molecule (what get_charges is called for most): ONE launch of the fused
   // one-CTA-per-molecule kernel (batch.cu) instead of the ~40 launches of the blocked single-system path.
   // MCHRG_B200_SMALL_FUSED=0 keeps the general path (A/B timing).
   if (mol && mol->id && mol->xyz && qvec && !cn && !qloc && !energy && !gradient && !sigma && !dqdr && !dqdL && jb == 0 &&
       je < 0 && model->kind == 1 && mol->nat > 0 && mol->nat <= BATCH_MAX_ATOMS &&
       !(mol->periodic[0] || mol->periodic[1] || mol->periodic[2]) && small_fused_enabled()) {
      const int64_t offsets[2] = {0, mol->nat};
      const double charge = mol->charge;
      int32_t status = 0;
      const int rc = mchrg_b200_singlepoint_batch(model, 1, offsets, mol->id, mol->xyz, &charge, qvec, nullptr, nullptr, &status);
      if (rc != 0 || status <= 0) return rc != 0 ? rc : status;
      // failed pivot: the general path below retries with the rank-one shift, so that both routes agree
   }
   return guarded(ctx, [&]() {
      DevMol d = stage_mol(ctx, mol);
      const size_t n = d.nat;
      const bool grad = (gradient && sigma) || dqdr;
      int ntr = 0;
      const double* trans = cn_translations(ctx, mol, model->ncoord.cutoff, &ntr);
      double* rcov_a = ctx->alloc<double>(n);
      gather_species(ctx, d.nat, d.id, model->dev(4), rcov_a);
      double* d_cn = cn ? ctx->out(cn, n) : ctx->alloc<double>(n);
      // EEQ2019 without dq/dr: the CN derivatives are contracted pair by pair after the solve (CnLazy)
      const bool lazy = grad && !dqdr && model->kind == 1 && lazy_cn_enabled();
      double* d_dcndr = (grad && !lazy) ? ctx->alloc<double>(3 * n * n) : nullptr;
      double* d_dcndL = (grad && !lazy) ? ctx->alloc<double>(9 * n) : nullptr;
      CnLazy lz;
      ncoord_device(ctx, d, rcov_a, nullptr, model->ncoord, ntr, trans, d_cn, d_dcndr, d_dcndL, &lz.cn_raw);
      lz.par = model->ncoord;
      lz.rcov_a = rcov_a;
      lz.trans = trans;
      lz.ntr = ntr;
      double* d_qloc = qloc ? ctx->out(qloc, n) : ctx->alloc<double>(n);
      double *d_dqlr = nullptr, *d_dqlL = nullptr;
      if (model->kind == 1) {
         MCHRG_LAUNCH(ctx, fill_kernel, (unsigned)((n + 255) / 256), 256, 0, (int64_t)n, d.charge / (double)n, d_qloc);
      } else {
         double* en_a = ctx->alloc<double>(n);
         gather_species(ctx, d.nat, d.id, model->dev(9), en_a);
         if (grad) {
            d_dqlr = ctx->alloc<double>(3 * n * n);
            d_dqlL = ctx->alloc<double>(9 * n);
         }
         ncoord_device(ctx, d, rcov_a, en_a, model->ncoord_en, ntr, trans, d_qloc, d_dqlr, d_dqlL);
         MCHRG_LAUNCH(ctx, add_const_kernel, (unsigned)((n + 255) / 256), 256, 0, d.nat, d.charge / (double)n, d_qloc);
      }
      SolveIO io;
      io.cn = d_cn;
      io.qloc = d_qloc;
      io.dcndr = d_dcndr;
      io.dcndL = d_dcndL;
      io.dqlocdr = d_dqlr;
      io.dqlocdL = d_dqlL;
      const size_t nj = (je < 0) ? n : (size_t)std::max(0, je - jb);
      io.jb = jb;
      io.je = je;
      io.lazy = lazy ? &lz : nullptr;
      io.energy = ctx->out(energy, n, accumulate);
      io.gradient = (gradient && sigma) ? ctx->out(gradient, 3 * nj) : nullptr;  // gradient and sigma only together
      io.sigma = (gradient && sigma) ? ctx->out(sigma, 9, accumulate) : nullptr;
      io.qvec = ctx->out(qvec, n);
      io.dqdr = ctx->out(dqdr, 3 * nj * n);
      io.dqdL = ctx->out(dqdL, 9 * n);
      if (io.dqdr && !io.dqdL) io.dqdL = ctx->alloc<double>(9 * n);  // dq/dL rides along with dq/dr
      return solve_dispatch(model, d, io);
   });
}

int mchrg_b200_singlepoint_rows(const mchrg_b200_model* model, const mchrg_b200_structure* mol, int32_t row_begin,
                                int32_t row_end, double* cn, double* energy, double* gradient_rows, double* sigma,
                                double* qvec, double* dqdr_rows, double* dqdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   if (!mol || row_begin < 0 || row_end > mol->nat || row_end <= row_begin) {
      set_last_error("singlepoint_rows: invalid row block");
      return MCHRG_B200_ERR_ARG;
   }
   return singlepoint_impl(model, mol, cn, nullptr, energy, gradient_rows, sigma, qvec, dqdr_rows, dqdL, true, row_begin,
                           row_end);
}

int mchrg_b200_get_charges(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* qvec,
                           double* dqdr, double* dqdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   if (!qvec) {
      set_last_error("get_charges: qvec is required");
      return MCHRG_B200_ERR_ARG;
   }
   return singlepoint_impl(model, mol, nullptr, nullptr, nullptr, nullptr, nullptr, qvec, dqdr, dqdL, true);
}

int mchrg_b200_singlepoint(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* cn, double* qloc,
                           double* energy, double* gradient, double* sigma, double* qvec, double* dqdr,
                           double* dqdL) {
   if (!model) return MCHRG_B200_ERR_ARG;
   return singlepoint_impl(model, mol, cn, qloc, energy, gradient, sigma, qvec, dqdr, dqdL, true);
}

// ---------------------------------------------------------------------------------------------
// dense test entry points
// ---------------------------------------------------------------------------------------------
int mchrg_b200_dpotrf(mchrg_b200_context* ctx, int64_t n, double* a, int64_t lda) {
   return guarded(ctx, [&]() {
      if (n <= 0 || !a || lda < n) fail(MCHRG_B200_ERR_ARG, "dpotrf: invalid argument");
      const int64_t np = round_up(n, TILE);
      const double* src = ctx->in(a, (size_t)lda * n);
      double* W = ctx->alloc<double>((size_t)np * np);
      double* dinv = ctx->alloc<double>((size_t)np * TILE);
      int* info = ctx->alloc_zero<int>(1);
      dim3 g((unsigned)((np + 255) / 256), (unsigned)np);
      MCHRG_LAUNCH(ctx, pad_copy_kernel, g, 256, 0, n, n, src, lda, np, np, W, np, 1);
      dpotrf_padded(ctx, np, W, np, dinv, info);
      double* dst = ctx->out(a, (size_t)lda * n, /*preload=*/true);
      dim3 g2((unsigned)((n + 255) / 256), (unsigned)n);
      MCHRG_LAUNCH(ctx, unpad_copy_kernel, g2, 256, 0, n, n, W, np, dst, lda, 1);
      int* h_info = static_cast<int*>(ctx->pinned_buf(sizeof(int)));
      MCHRG_CUDA(cudaMemcpyAsync(h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      ctx->finish();
      return *h_info;
   });
}

int mchrg_b200_dgemm(mchrg_b200_context* ctx, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* a, int64_t lda, const double* b, int64_t ldb, double beta, double* c,
                     int64_t ldc) {
   return guarded(ctx, [&]() {
      const bool tb = (transb == 't' || transb == 'T');
      if (m <= 0 || n <= 0 || k <= 0 || !a || !b || !c) fail(MCHRG_B200_ERR_ARG, "dgemm: invalid argument");
      const int64_t mp = round_up(m, TILE), np = round_up(n, TILE), kp = round_up(k, 32);
      const double* sa = ctx->in(a, (size_t)lda * k);
      const double* sb = ctx->in(b, (size_t)ldb * (tb ? k : n));
      double* dc = ctx->out(c, (size_t)ldc * n, /*preload=*/true);
      double* A = ctx->alloc<double>((size_t)mp * kp);
      double* B = ctx->alloc<double>(tb ? (size_t)np * kp : (size_t)kp * np);
      double* Cp = ctx->alloc<double>((size_t)mp * np);
      MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((mp + 255) / 256), (unsigned)kp), 256, 0, m, k, sa, lda, mp, kp,
                   A, mp, 0);
      if (tb)
         MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((np + 255) / 256), (unsigned)kp), 256, 0, n, k, sb, ldb, np,
                      kp, B, np, 0);
      else
         MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((kp + 255) / 256), (unsigned)np), 256, 0, k, n, sb, ldb, kp,
                      np, B, kp, 0);
      MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((mp + 255) / 256), (unsigned)np), 256, 0, m, n, dc, ldc, mp, np,
                   Cp, mp, 0);
      dgemm_padded(ctx, tb, mp, np, kp, alpha, A, mp, B, tb ? np : kp, beta, Cp, mp, false);
      MCHRG_LAUNCH(ctx, unpad_copy_kernel, dim3((unsigned)((m + 255) / 256), (unsigned)n), 256, 0, m, n, Cp, mp, dc, ldc,
                   0);
      return 0;
   });
}

int mchrg_b200_dpotrs_right(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* l, int64_t ldl, double* x,
                            int64_t ldx) {
   return guarded(ctx, [&]() {
      if (m <= 0 || n <= 0 || !l || !x) fail(MCHRG_B200_ERR_ARG, "dpotrs_right: invalid argument");
      const int64_t mp = round_up(m, TILE), np = round_up(n, TILE);
      const double* sl = ctx->in(l, (size_t)ldl * n);
      double* dx = ctx->out(x, (size_t)ldx * n, /*preload=*/true);
      double* W = ctx->alloc<double>((size_t)np * np);
      double* dinv = ctx->alloc<double>((size_t)np * TILE);
      double* X = ctx->alloc<double>((size_t)mp * np);
      int* info = ctx->alloc_zero<int>(1);
      // re-factor L L^T to obtain the inverted diagonal blocks: W = L L^T is formed by the GEMM itself
      double* Lp = ctx->alloc<double>((size_t)np * np);
      MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((np + 255) / 256), (unsigned)np), 256, 0, n, n, sl, ldl, np, np,
                   Lp, np, 1);
      // zero the strict upper triangle of Lp (caller may pass a full matrix)
      dgemm_padded(ctx, true, np, np, np, 1.0, Lp, np, Lp, np, 0.0, W, np, false);
      dpotrf_padded(ctx, np, W, np, dinv, info);
      MCHRG_LAUNCH(ctx, pad_copy_kernel, dim3((unsigned)((mp + 255) / 256), (unsigned)np), 256, 0, m, n, dx, ldx, mp, np,
                   X, mp, 0);
      dtrsm_right_lt(ctx, mp, np, W, np, dinv, X, mp);
      dtrsm_right_ln(ctx, mp, np, W, np, dinv, X, mp);
      MCHRG_LAUNCH(ctx, unpad_copy_kernel, dim3((unsigned)((m + 255) / 256), (unsigned)n), 256, 0, m, n, X, mp, dx, ldx,
                   0);
      return 0;
   });
}

}  // extern "C"
