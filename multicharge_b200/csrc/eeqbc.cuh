// EEQ-BC (model/eeqbc.f90), molecular (0-D) device pieces -- see eeqbc.cu; the periodic ones are in eeq_pbc.cu.
#pragma once
#include "model.cuh"

namespace mchrg {

// per-atom gathered model parameters
struct BcAtoms {
   const double *chi, *rad, *eta, *kcnchi, *kqchi, *kqeta, *cap, *avg_cn;
};

// eeqbc_cache (eeqbc.f90:39-54) on the device
struct BcWork {
   int64_t npad = 0;
   double* atoms = nullptr;      // [5 * (nat + 1)]: rad_eff, drad/dcn, h, xtmp, cap
   double* Cm = nullptr;         // capacitance matrix, order npad, zero padded
   double* dcdr_diag = nullptr;  // dcdr(:, a, a)
   double* dcdL = nullptr;       // dcdL(:, :, a)
};

// update(): capacitance matrix (+ derivative row sums when grad)
void bc_prepare(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcAtoms& at,
                const double* cn, const double* qloc, bool grad, BcWork& w);
// get_amat_0d into the padded work matrix W (order w.npad)
// own_panel > 0 (multi-GPU factorisation, dpotrf_dist_panel): only the column groups this rank owns
void bc_amat(mchrg_b200_context* ctx, const DevMol& mol, const BcWork& w, double* W, int64_t own_panel = 0);
// jq[a] = (A q)_a for the rows [row0, row1) (row1 < 0: all), re-evaluated from the pair formula
void bc_jq_rows(mchrg_b200_context* ctx, const DevMol& mol, const BcWork& w, const double* q, double* jq, int row0 = 0,
                int row1 = -1);
// row sums: rows[24 * a] = atrace(3), dadL pair part(9), sum_b xtmp_b D_ab (3), sum_b xtmp_b D(be) v(al) (9)
void bc_rows(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcWork& w,
             const double* q, const double* dcndr, const double* dcndL, double* rows, int row0 = 0, int row1 = -1);
// complete dadL(:,:,a) and dxdL(:,:,a)
void bc_strain(mchrg_b200_context* ctx, const DevMol& mol, const BcAtoms& at, const BcWork& w, const double* q,
               const double* rows, const double* dcndL, const double* dqlocdL, double* dadL, double* dxdL, int row0 = 0,
               int row1 = -1);
// Bm = alpha * dtmpdr * C   (the dense part of get_xvec_derivs, eeqbc.f90:361)
void bc_dxdr_gemm(mchrg_b200_context* ctx, const DevMol& mol, const BcAtoms& at, const BcWork& w, const double* dcndr,
                  const double* dqlocdr, double alpha, double* Bm, int64_t mp, int jb = 0, int je = -1);
// B = sa dadr - sx dxdr (see eeqbc.cu), with the q / z contractions
void bc_build_b(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcAtoms& at,
                const BcWork& w, const double* q, const double* cq, const double* rows, const double* dcndr,
                const double* dqlocdr, double sa, double sx, bool add_atrace, double* Bout, int64_t ldb, double* ga,
                double* gx, const double* zvec, double* bz, int jb = 0, int je = -1);
// [row0, row1) / [jb, je) above: row shards of one large system over several GPUs (the O(N^2) row passes are joined
// by one all-reduce) and the atom block whose rows of B = dadr - dxdr this rank builds (rows numbered in the block).

}  // namespace mchrg
