// Periodic EEQ2019 (Ewald) device pieces -- see eeq_pbc.cu.
#pragma once
#include <vector>

#include "model.cuh"

namespace mchrg {

// host copies of the per-call periodic setup (kept alive until the stream has consumed them)
struct PbcHost {
   std::vector<double> wtrans, dtrans, rtrans, rcoef, packed;
   double alpha = 0.0;
};

struct PbcWork {
   int64_t npad = 0;
   void* state = nullptr;     // device-side descriptors (PbcState)
   bool phase_ready = false;  // Phi filled by pbc_amat
};

// update() of eeq.f90:119-123 for periodic structures: WSC candidate translations, Ewald alpha,
// direct / reciprocal translations, G-vector coefficients -> device
void pbc_begin(mchrg_b200_context* ctx, const DevMol& mol, PbcHost& host, PbcWork& work);
void pbc_end(PbcWork& work);

// get_amat_3d (eeq.f90:244-307): Coulomb block into W (order work.npad, ld = npad, identity padding)
// shard (several GPUs, dist_pbc_apply): every rank evaluates its share of the pair tiles and of the rows of the
// reciprocal-space product, rank 0 the diagonal; ONE all-reduce of W joins them (every entry has at most two
// non-zero contributions, so the sum equals the one-GPU result)
void pbc_amat(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a, PbcWork& work,
              double* W, bool shard = false);

// y = A x for a full symmetric column-major matrix
void pbc_symv(mchrg_b200_context* ctx, int nat, const double* A, int64_t ld, const double* x, double* y);

// get_damat_3d (eeq.f90:429-508): atrace[3 nat], dadL[9 nat] (overwritten) and, when B != nullptr, the
// off-diagonal blocks dadr(:, j, k) into B (ld = ldb, mp padded rows; B must be zero-initialised)
// shard: this rank's share of the pair tiles feeds the row sums; rank 0 adds the reciprocal-space rows, the self images
// are row-shared; atrace / dadL must be zero on entry and are PARTIAL sums on return (the caller all-reduces them).
// [jb, je): atom block whose rows of B are built (rows numbered within the block, mp >= 3 (je - jb) + 9).
void pbc_derivs(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* q, PbcWork& work,
                double* B, int64_t ldb, int64_t mp, double* atrace, double* dadL, bool shard = false, int jb = 0,
                int je = -1);

// Periodic EEQ-BC, charge / energy path (get_cmat_3d eeqbc.f90:1065-1128, get_amat_3d :532-598): fills the
// capacitance matrix Cm and the Coulomb block W (both order work.npad, ld = npad; padding: Cm zero, W identity)
// and cself(a) = capacitance of atom a with its own images; `atoms` as produced by bc_atom_kernel (eeqbc.cu)
void bc_pbc_matrices(mchrg_b200_context* ctx, const DevMol& mol, const double* atoms, const double* rvdw, int nid,
                     double kbc, PbcWork& work, double* Cm, double* W, double* cself);
// xvec(a) -= cself(a) xtmp(a)   (the periodic part of get_xvec, eeqbc.f90:283-311)
void bc_pbc_xvec_fix(mchrg_b200_context* ctx, int nat, const double* atoms, const double* cself, double* xvec);

// Periodic EEQ-BC derivatives (get_damat_3d, get_dcmat_3d, periodic get_xvec_derivs): adds everything except the
// dense part dtmpdr C (expected in B already, scaled by -1) to B = dadr + diag(atrace) - dxdr (may be null),
// accumulates ga = dadr q, gx = dxdr q (zero-initialised by the caller), writes dadL, dxdL per atom (zero-initialised
// by the caller) and bz = B z for the rows < 3 nat.  cq = C q, cself from bc_pbc_matrices.
void bc_pbc_derivs(mchrg_b200_context* ctx, const DevMol& mol, const double* atoms, const double* rvdw, int nid, double kbc,
                   const double* kcnchi, const double* kqchi, const double* kqeta, const double* q, const double* cq,
                   const double* Cm, const double* cself, const double* dcndr, const double* dcndL, const double* dqlocdr,
                   const double* dqlocdL, const double* zvec, PbcWork& work, double* B, int64_t ldb, double* ga, double* gx,
                   double* dadL, double* dxdL, double* bz, double sa = 1.0, double sx = 1.0, bool add_atrace = true,
                   double* atrace = nullptr);
// (sa, sx, add_atrace) select B = sa (dadr + add_atrace diag(atrace)) - sx dxdr for the piecewise accessors;
// atrace (zero-initialised by the caller) receives get_damat_3d's atrace when given; zvec / bz may be null.

}  // namespace mchrg
