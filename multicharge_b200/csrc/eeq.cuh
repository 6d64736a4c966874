// Device pieces of the EEQ2019 model (model/eeq.f90) and of the erf coordination number
// (mctc_ncoord): host-callable wrappers around the kernels in eeq.cu / eeq_pbc.cu.
#pragma once
#include "model.cuh"

namespace mchrg {

// gather per-species parameter `src` (device, [nid]) to atoms: dst[i] = src[id[i]-1]
void gather_species(mchrg_b200_context* ctx, int nat, const int* id, const double* src, double* dst);

// ---- mctc_ncoord%get_coordination_number -----------------------------------------------------
// cn[nat]; dcndr[3*nat*nat], dcndL[9*nat] optional (both or none).  en_a == nullptr: cn_count%erf,
// else erf_en (directed, weighted by en_j - en_i).  trans: device [3*ntr].
// cn_raw_out (optional) receives the device array of the CN before the log cut (lives in the call's arena).
void ncoord_device(mchrg_b200_context* ctx, const DevMol& mol, const double* rcov_a, const double* en_a,
                   const NcoordPar& par, int ntr, const double* trans, double* cn, double* dcndr, double* dcndL,
                   const double** cn_raw_out = nullptr);

// What the solve needs to contract the CN derivatives on the fly instead of reading dcndr / dcndL
// (single points without dq/dr: the (3, N, N) array is never formed).
struct CnLazy {
   NcoordPar par;
   const double* rcov_a = nullptr;
   const double* cn_raw = nullptr;
   const double* trans = nullptr;
   int ntr = 0;
};
// gradx(3, je - jb) = sum_k thalf_k q_k dcndr(:, j, k) for the rows [jb, je);  sig_cn(3, 3) = sum_k thalf_k q_k dcndL(:, :, k)
// [row0, row1) (row1 < 0: all): the atoms this call evaluates -- one large system on several GPUs: every rank takes a
// share, writes its rows of the (then full-length, zeroed) gradx and its partial sig_cn, and one all-reduce joins them.
void cn_gradient_device(mchrg_b200_context* ctx, const DevMol& mol, const CnLazy& lz, const double* thalf, const double* q,
                        int jb, int je, double* gradx, double* sig_cn, int row0 = 0, int row1 = -1);

// ---- EEQ2019 -----------------------------------------------------------------------------------
// get_xvec (eeq.f90:127-150): xvec[nat+1]; thalf[nat] (optional) = 0.5 * kcnchi / sqrt(cn + reg),
// the factor of get_xvec_derivs (eeq.f90:175-177).
void eeq_xvec(mchrg_b200_context* ctx, const DevMol& mol, const double* chi_a, const double* kcnchi_a,
              const double* cn, double* xvec, double* thalf);

// get_amat_0d (eeq.f90:199-242).  border=true: full (nat+1)^2 matrix with the constraint row/column,
// ld = nat+1.  border=false: the nat x nat Coulomb block J into a buffer of order npad >= nat
// (ld = npad) whose padding is the identity (solver work matrix).  own_panel > 0 (multi-GPU factorisation,
// dpotrf_dist_panel): only the block columns of that width this rank owns block-cyclically are assembled
// (SURVEY.md 8e "assembly": pair tiles sharded by ownership of the factorisation, no collective).
void eeq_amat_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a, double* out,
                 int64_t ld, bool border, int64_t npad, int64_t own_panel = 0);

// One pass over all pairs of row i (0-D): jq[i] = (J q)_i, atrace[3*nat] (eeq.f90:410-411),
// dadL[9*nat] (eeq.f90:414-415).  Any output may be nullptr.
// [row0, row1) (row1 < 0: all): rows evaluated by this call (row shards of one large system, see above).
void eeq_pair_rows_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a,
                      const double* q, double* jq, double* atrace, double* dadL, int row0 = 0, int row1 = -1);

// B(3j+c, k) = dadr(c,j,k) [+ atrace on the diagonal] - thalf_k * dcndr(c,j,k)   (type.F90:270-287)
// written with leading dimension ldb; optionally accumulates
//   gradx[3j+c] += sum_k thalf_k q_k dcndr(c,j,k)      (the dxdr*q term of type.F90:278)
//   bz[3j+c]    += sum_k B(3j+c, k) zvec_k             (rank-one Schur correction of dq/dr)
// dcndr may be nullptr (pure dadr, get_coulomb_derivs); Bout may be nullptr (gradient only).
// pair_in_b: the off-diagonal pair terms are already stored in Bout (periodic path) and are
// updated in place instead of being evaluated from the non-periodic pair formula.
// [jb, je): atom (row) block to build; rows of Bout / gradx / bz are numbered within the block
// (multi-GPU: every rank owns a block of dq/dr rows, the right-side solves act on rows independently).
void eeq_build_b_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* q,
                    const double* atrace, const double* thalf, const double* dcndr, double* Bout, int64_t ldb,
                    double* gradx, const double* zvec, double* bz, bool pair_in_b = false, int jb = 0, int je = -1);

}  // namespace mchrg
