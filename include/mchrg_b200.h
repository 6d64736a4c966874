/* mchrg_b200 -- C ABI of the B200-native EEQ charge hot path.
 *
 * Drop-in boundary for the ONE path of grimme-lab/multicharge v0.5.0 this library replaces:
 * Coulomb-matrix assembly -> constrained (N+1)x(N+1) solve -> charge derivatives, behind the
 * reference's model API.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  All computation runs in hand-written CUDA kernels for
 * sm_100a; there is NO CPU fallback: every call fails with MCHRG_B200_ERR_CUDA when no device is
 * usable.
 *
 * Conventions (bit-for-bit the reference's Fortran arrays, so an iso_c_binding shim is a pure
 * pass-through -- see INTEGRATION.md):
 *   - real(wp) = double, column-major:  xyz(3,nat)  lattice(3,3) [lattice(:,k) = k-th cell vector]
 *     cn(nat) qloc(nat) dcndr(3,nat,nat) [dcndr(c,j,i) = d cn_i/d r_jc] dcndL(3,3,nat)
 *     qvec(nat) energy(nat) gradient(3,nat) sigma(3,3) dqdr(3,nat,nat) dqdL(3,3,nat)
 *     amat(nat+1,nat+1) xvec(nat+1) dadr(3,nat,nat+1) dadL(3,3,nat+1) atrace(3,nat)
 *     dxdr(3,nat,nat+1) dxdL(3,3,nat+1)
 *   - id(nat) is the 1-based species index mol%id; per-species model arrays have length nid.
 *   - Optional arguments of the Fortran interface are nullable pointers; presence selects the
 *     work exactly as in model/type.F90:199-201.
 *   - energy and sigma are ACCUMULATED INTO, gradient/qvec/dqdr/dqdL are overwritten
 *     (model/type.F90:252,261,276-280,289-290).
 *   - Every array pointer may be a host pointer or a device pointer of the context's GPU
 *     (decided per pointer with cudaPointerGetAttributes); host arrays are staged.
 *   - Return value: 0 = success; > 0 = factorisation failure (the LAPACK-info analogue: 1-based
 *     index of the first non-positive pivot), reported by the reference as
 *     fatal_error("Bunch-Kaufman factorization failed.") (model/type.F90:222-225);
 *     < 0 = MCHRG_B200_ERR_*.  mchrg_b200_last_error() returns the message for the calling thread.
 */
#ifndef MCHRG_B200_H
#define MCHRG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCHRG_B200_ERR_ARG (-1)    /* invalid argument                                   */
#define MCHRG_B200_ERR_CUDA (-2)   /* CUDA runtime / kernel failure, or no usable device */
#define MCHRG_B200_ERR_ALLOC (-3)  /* device or pinned-host allocation failed            */
#define MCHRG_B200_ERR_MODEL (-4)  /* model cannot serve this request (e.g. qloc missing
                                      for EEQ-BC: `error stop "qloc required"`, eeqbc.f90:202) */

typedef struct mchrg_b200_context mchrg_b200_context; /* one GPU, one stream, grow-only workspace */
typedef struct mchrg_b200_model mchrg_b200_model;     /* class(mchrg_model_type), immutable        */

/* type(structure_type) of mctc-lib, reduced to the members the path reads
 * (nat, nid, id, xyz, charge, periodic, lattice; model/type.F90:156, eeq.f90:217-223). */
typedef struct mchrg_b200_structure {
   int32_t nat;
   int32_t nid;
   const int32_t* id;   /* [nat], 1-based, host or device */
   const double* xyz;   /* [3*nat] Bohr, host or device   */
   double charge;
   int32_t periodic[3];
   double lattice[9];
} mchrg_b200_structure;

/* ---- context ---------------------------------------------------------------------------- */
int mchrg_b200_context_create(int device, mchrg_b200_context** ctx);
void mchrg_b200_context_destroy(mchrg_b200_context* ctx);
const char* mchrg_b200_last_error(void);
int mchrg_b200_version(void); /* 10000*major + 100*minor + patch */
/* cudaStreamSynchronize on the context's stream (all entry points are synchronous w.r.t. host
 * output arrays; device output arrays are complete after this call). */
int mchrg_b200_synchronize(mchrg_b200_context* ctx);
/* stream the context launches on (cudaStream_t as void*), for callers that time with events */
void* mchrg_b200_stream(mchrg_b200_context* ctx);
/* second stream of the context; panel exchanges (below) are ordered on it */
void* mchrg_b200_side_stream(mchrg_b200_context* ctx);

/* ---- multi-GPU factorisation of ONE large system ("one exchange per panel") -------------
 * The reference factors the bordered matrix with one LAPACK call (dsytrf, model/type.F90:221,
 * lapack.F90:165-201).  With an exchange installed, every rank of a replicated solve owns the block columns
 * b % nranks == rank of the Cholesky factorisation (block-cyclic, 256 wide): the owner factors panel b,
 * `start` broadcasts it (device pointer, bytes, root rank, slot id) and `wait` orders the context's main
 * stream after the transfer of that slot; ranks update only the columns they own.  The callbacks are the
 * caller's collective (torch.distributed / NCCL broadcast in multicharge_b200/api.py); the library itself
 * links no communication library.  `start` must enqueue on mchrg_b200_side_stream, `wait` on
 * mchrg_b200_stream.  All ranks must make the same sequence of library calls with the same inputs.
 * nranks <= 1 or start == NULL removes the exchange. */
typedef void (*mchrg_b200_exchange_start_fn)(void* user, void* dev_ptr, int64_t bytes, int32_t root, int32_t slot);
typedef void (*mchrg_b200_exchange_wait_fn)(void* user, int32_t slot);
int mchrg_b200_set_exchange(mchrg_b200_context* ctx, int32_t rank, int32_t nranks, mchrg_b200_exchange_start_fn start,
                            mchrg_b200_exchange_wait_fn wait, void* user);

/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
uint64_t mchrg_b200_launch_count(const mchrg_b200_context* ctx);

/* ---- constructors ------------------------------------------------------------------------ */
/* new_eeq_model (model/eeq.f90:62-95): per-species chi, rad, eta, kcnchi, rcov; cutoff, cn_exp,
 * cn_max configure the erf coordination number (mctc_ncoord, cn_count%erf). */
int mchrg_b200_new_eeq_model(mchrg_b200_context* ctx, int32_t nid, const double* chi, const double* rad,
                             const double* eta, const double* kcnchi, const double* rcov, double cutoff,
                             double cn_exp, double cn_max, mchrg_b200_model** model);
/* new_eeq2019_model (param.f90:46-73): element tables of param/eeq2019.f90 for atomic numbers
 * num[nid], D3 covalent radii, cutoff 25, cn_exp 7.5, cn_max 8. */
int mchrg_b200_new_eeq2019_model(mchrg_b200_context* ctx, int32_t nid, const int32_t* num,
                                 mchrg_b200_model** model);
/* new_eeqbc_model (model/eeqbc.f90:100-176).  rvdw is per SPECIES pair [nid*nid] (the reference
 * stores rvdw(iat,jat) = table(num(id(iat)), num(id(jat))), param.f90:114-115). */
int mchrg_b200_new_eeqbc_model(mchrg_b200_context* ctx, int32_t nid, const double* chi, const double* rad,
                               const double* eta, const double* kcnchi, const double* kqchi,
                               const double* kqeta, double kcnrad, const double* cap, const double* avg_cn,
                               const double* rvdw, double kbc, double cutoff, double cn_exp,
                               const double* rcov, const double* en, double norm_exp,
                               mchrg_b200_model** model);
/* new_eeqbc2025_model (param.f90:75-125) with the tables of param/eeqbc2025.f90.  The pairwise D3
 * vdW radii and Pauling EN live in mctc-lib (not vendored): the caller supplies them
 * (rvdw[nid*nid] in ANGSTROM-valued Bohr numbers exactly as get_vdw_rad()*autoaa, en[nid] Pauling). */
int mchrg_b200_new_eeqbc2025_model(mchrg_b200_context* ctx, int32_t nid, const int32_t* num,
                                   const double* rvdw, const double* en_pauling, mchrg_b200_model** model);
void mchrg_b200_model_free(mchrg_b200_model* model);
/* 1 = eeq2019-type (mchrg_model%eeq2019), 2 = eeqbc-type (param.f90:34-42) */
int mchrg_b200_model_kind(const mchrg_b200_model* model);

/* ---- mctc_ncoord: self%ncoord%get_coordination_number(mol, trans, cn[, dcndr, dcndL]) -----
 * (call sites charge.f90:71, app/main.f90:115).  The lattice translations are generated inside
 * from the model's cutoff (get_lattice_points, charge.f90:70).  dcndr and dcndL both NULL or both
 * given. */
int mchrg_b200_get_coordination_number(const mchrg_b200_model* model, const mchrg_b200_structure* mol,
                                       double* cn, double* dcndr, double* dcndL);
/* mchrg_model_type%local_charge (model/type.F90:294-322) */
int mchrg_b200_local_charge(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* qloc,
                            double* dqlocdr, double* dqlocdL);

/* ---- deferred procedures of mchrg_model_type (model/type.F90:78-125), for parity tests ----
 * The per-solve cache (update(), eeq.f90:97-125) is rebuilt inside each call from cn/qloc. */
int mchrg_b200_get_xvec(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                        const double* qloc, double* xvec);
int mchrg_b200_get_xvec_derivs(const mchrg_b200_model* model, const mchrg_b200_structure* mol,
                               const double* cn, const double* qloc, const double* dcndr,
                               const double* dcndL, const double* dqlocdr, const double* dqlocdL,
                               double* dxdr, double* dxdL);
int mchrg_b200_get_coulomb_matrix(const mchrg_b200_model* model, const mchrg_b200_structure* mol,
                                  const double* cn, const double* qloc, double* amat);
int mchrg_b200_get_coulomb_derivs(const mchrg_b200_model* model, const mchrg_b200_structure* mol,
                                  const double* cn, const double* qloc, const double* dcndr,
                                  const double* dcndL, const double* dqlocdr, const double* dqlocdL,
                                  const double* qvec, double* dadr, double* dadL, double* atrace);

/* ---- mchrg_model_type%solve (model/type.F90:151-292) -------------------------------------- */
int mchrg_b200_solve(const mchrg_b200_model* model, const mchrg_b200_structure* mol, const double* cn,
                     const double* qloc, const double* dcndr, const double* dcndL, const double* dqlocdr,
                     const double* dqlocdL, double* energy, double* gradient, double* sigma, double* qvec,
                     double* dqdr, double* dqdL);

/* ---- get_charges (charge.f90:37-77): lattice points -> CN -> local charge -> solve -------- */
/* Without derivatives, for a non-periodic EEQ2019-type structure of at most 208 atoms, this is ONE kernel launch
 * (the fused one-CTA-per-molecule kernel of the batch path). */
int mchrg_b200_get_charges(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* qvec,
                           double* dqdr, double* dqdL);

/* ---- the CLI's call sequence (app/main.f90:114-118): CN(+derivs) -> local_charge -> solve with
 * every requested output; cn / qloc are optional outputs.  CN derivatives never leave the device. */
int mchrg_b200_singlepoint(const mchrg_b200_model* model, const mchrg_b200_structure* mol, double* cn,
                           double* qloc, double* energy, double* gradient, double* sigma, double* qvec,
                           double* dqdr, double* dqdL);

/* ---- one large system over several GPUs: row-sharded derivatives ----------------------
 * mchrg_b200_singlepoint restricted to the atom block [row_begin, row_end) of the derivative outputs:
 *   gradient_rows(3, nrows)      = gradient(:, row_begin+1 : row_end)
 *   dqdr_rows(3, nrows, nat)     = dqdr(:, row_begin+1 : row_end, :)
 * while cn, energy, sigma, qvec and dqdL are complete on every caller.  The right-side solves that
 * produce dq/dr act on its rows independently, so ranks that own disjoint blocks need NO data-path
 * collective; the O(N^3/3) factorisation is replicated (see DESIGN.md "Multi-GPU").  Non-periodic
 * EEQ2019-type models only. */
int mchrg_b200_singlepoint_rows(const mchrg_b200_model* model, const mchrg_b200_structure* mol, int32_t row_begin,
                                int32_t row_end, double* cn, double* energy, double* gradient_rows, double* sigma,
                                double* qvec, double* dqdr_rows, double* dqdL);

/* ---- batches of independent non-periodic molecules (BASELINE.json configs[0..1]) -----------
 * get_charges / singlepoint over nmol molecules in one call: molecule m owns atoms
 * [offsets[m], offsets[m+1]).  num[] are ATOMIC NUMBERS per atom (the model must have been built
 * by *_new_eeq2019_model with num = 1..103, i.e. species index == atomic number).  Outputs are
 * concatenated per atom; energy is per atom and overwritten; status[m] follows the return-value
 * convention above per molecule.  gradient may be NULL (charges + energy only).  Molecules of up to
 * 208 atoms run in the one-CTA-per-molecule kernels; larger ones are solved one after the other by
 * the single-system path (mchrg_b200_singlepoint) inside the same call. */
int mchrg_b200_singlepoint_batch(const mchrg_b200_model* model, int32_t nmol, const int64_t* offsets,
                                 const int32_t* num, const double* xyz, const double* charge,
                                 double* qvec, double* energy, double* gradient, int32_t* status);

/* ---- dense FP64 building blocks (lapack.F90 / blas.F90 replacements), exposed for tests ----
 * Column-major.  dpotrf: lower Cholesky in place (upper triangle untouched), returns LAPACK info.
 * dgemm: C = alpha*A*op(B) + beta*C with A not transposed, transb in {'n','t'}. */
int mchrg_b200_dpotrf(mchrg_b200_context* ctx, int64_t n, double* a, int64_t lda);
int mchrg_b200_dgemm(mchrg_b200_context* ctx, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* a, int64_t lda, const double* b, int64_t ldb, double beta, double* c,
                     int64_t ldc);
/* X <- X * inv(L*L^T) for an m x n row block X given the Cholesky factor L (both TRSM sweeps) */
int mchrg_b200_dpotrs_right(mchrg_b200_context* ctx, int64_t m, int64_t n, const double* l, int64_t ldl,
                            double* x, int64_t ldx);

#ifdef __cplusplus
}
#endif
#endif /* MCHRG_B200_H */
