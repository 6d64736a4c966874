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
/* second stream of the context (batch size classes, pair phase A) */
void* mchrg_b200_side_stream(mchrg_b200_context* ctx);

/* ---- one large system over several GPUs: the communicator ------------------------------
 * The reference factors the bordered matrix with one LAPACK call (dsytrf, model/type.F90:221,
 * lapack.F90:165-201) and has no distributed path.  With a communicator installed on its context, every
 * single-system entry point (solve, singlepoint, singlepoint_rows, get_charges) of a structure with at least
 * MCHRG_B200_DIST_MIN (default 8192) atoms shares the work between the ranks:
 *   - Cholesky: column groups are owned block-cyclically (owner(g) = g % nranks), factored sub-panels are
 *     broadcast, every rank updates only the groups it owns (csrc/dense.cu);
 *   - the O(N^2) row passes (coordination number, (J q) / atrace / dadL, CN-gradient contraction) are
 *     row-sharded and joined by one all-reduce each;
 *   - dq/dr is row-sharded by the caller through mchrg_b200_singlepoint_rows (no collective);
 *   - periodic EEQ2019 structures (eeq.f90 get_amat_3d / get_damat_3d) of at least MCHRG_B200_DIST_MIN_PBC
 *     (default 1024) atoms: the Ewald real-space pair tiles, the rows of the reciprocal-space product and the
 *     self-image rows are shared, one all-reduce of the matrix and one of the derivative rows join them
 *     (q, E, gradient, sigma); dq/dr rows of a periodic structure are sharded through mchrg_b200_singlepoint_rows
 *     like the molecular ones.
 * All ranks must make the same sequence of library calls with the same inputs.
 * Two backends:
 *   NCCL  -- one process per GPU: rank 0 calls comm_unique_id, distributes the 128 bytes by its own means
 *            (MPI_Bcast, torch.distributed, a file), every rank calls comm_init_nccl.  libnccl.so.2 is loaded
 *            with dlopen at that moment (MCHRG_B200_NCCL_LIB overrides the path); the library links no
 *            communication library.
 *   local -- one process, one host thread per context (several GPUs, or several contexts on one GPU):
 *            comm_init_local(ctxs, n) ties n contexts together; collectives are stream-ordered peer copies. */
#define MCHRG_B200_COMM_ID_BYTES 128
int mchrg_b200_comm_unique_id(void* id /* [MCHRG_B200_COMM_ID_BYTES] out */);
int mchrg_b200_comm_init_nccl(mchrg_b200_context* ctx, int32_t rank, int32_t nranks, const void* id);
int mchrg_b200_comm_init_local(mchrg_b200_context** ctxs, int32_t n);
int mchrg_b200_comm_destroy(mchrg_b200_context* ctx);
int32_t mchrg_b200_comm_rank(const mchrg_b200_context* ctx);
int32_t mchrg_b200_comm_size(const mchrg_b200_context* ctx);
/* the collectives themselves on device buffers (tests; gathering row-sharded results): ordered on the
 * context's stream, synchronous for the host */
int mchrg_b200_comm_bcast(mchrg_b200_context* ctx, void* dev_buf, int64_t bytes, int32_t root);
int mchrg_b200_comm_allreduce_sum(mchrg_b200_context* ctx, double* dev_buf, int64_t count);
int mchrg_b200_comm_allgather(mchrg_b200_context* ctx, const void* dev_send, void* dev_recv, int64_t bytes_per_rank);

/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
uint64_t mchrg_b200_launch_count(const mchrg_b200_context* ctx);
/* With MCHRG_B200_TRACE=1 in the environment a large batch call (singlepoint_batch, split path) runs its three
 * kernels one after the other over the whole batch (same results), each timed alone with CUDA events on the
 * launching stream instead of overlapping them chunk by chunk.  This returns those times: ms[3] =
 * pair phase A, factorisation, pair phase C; launches[3] = kernel launches of each (bench.py's per-kernel roofline). */
int mchrg_b200_batch_trace(const mchrg_b200_context* ctx, double* ms, int32_t* launches);

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
 * collective; with a communicator on the context the factorisation and the O(N^2) row passes are shared as
 * described above (see DESIGN.md "Multi-GPU").  EEQ2019: molecular and periodic structures; EEQ-BC: molecular. */
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
