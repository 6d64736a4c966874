/* TEST INFRASTRUCTURE -- CPU oracle for the EEQ charge hot path of grimme-lab/multicharge v0.5.0.
 *
 * This is a plain-C restatement of the reference's Fortran algorithm, used ONLY as the checker
 * in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
 * product (multicharge_b200/, libmchrg_b200.so) never includes, links or calls anything here.
 *
 * Parity status:
 *   - EEQ2019, non-periodic: PINNED by the reference's own golden data (test/validation/01..03
 *     multicharge_ref.json: CN, charges, energy, gradient; test/unit/test_model.f90:934-981
 *     actinides charges) -- see tests/test_oracle_golden.py.
 *   - EEQ2019 periodic, EEQ-BC: "parity unpinned" against reference NUMBERS (the reference's
 *     periodic / EEQ-BC known answers need mstore geometries and mctc-lib tables that are not
 *     available offline); checked by the reference's finite-difference methodology only.
 *
 * All arrays use the Fortran (column-major) layout of the reference:
 *   xyz(3,nat), lattice(3,3) [lattice(:,k) = k-th cell vector], amat(ndim,ndim), ndim = nat+1,
 *   dcndr(3,nat,nat) [dcndr(c,j,i) = d cn_i / d r_jc], dcndL(3,3,nat), dadr(3,nat,ndim),
 *   dadL(3,3,ndim), atrace(3,nat), dqdr(3,nat,nat), dqdL(3,3,nat).  id[] is 1-based (mol%id).
 */
#ifndef MCHRG_ORACLE_H
#define MCHRG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
   int nat;
   int nid;
   const int* id;        /* [nat] 1-based species index          */
   const double* xyz;    /* [3*nat] Bohr                          */
   double charge;        /* total charge                          */
   int periodic[3];
   double lattice[9];    /* column-major 3x3                      */
} orc_mol;

/* erf coordination number of mctc-lib (mctc_ncoord, cn_count%erf / erf_en) */
typedef struct {
   const double* rcov;   /* [nid]                                 */
   const double* en;     /* [nid] or NULL (erf) / given (erf_en)  */
   double kcn;           /* steepness                             */
   double norm_exp;      /* exponent of rc in the argument        */
   double cutoff;        /* real-space cutoff                     */
   double cut;           /* cn_max log-cut, <= 0: none            */
} orc_ncoord;

/* eeq_model, model/eeq.f90:42-54 */
typedef struct {
   int nid;
   const double* chi;
   const double* rad;
   const double* eta;
   const double* kcnchi;
   orc_ncoord ncoord;
} orc_eeq_model;

/* eeqbc_model, model/eeqbc.f90:56-86 */
typedef struct {
   int nid;
   const double* chi;
   const double* rad;
   const double* eta;
   const double* kcnchi;
   const double* kqchi;
   const double* kqeta;
   const double* cap;
   const double* avg_cn;
   const double* rvdw;    /* [nid*nid] pairwise vdW radii per species pair (reference: per atom pair) */
   double kcnrad;
   double kbc;
   double norm_exp;
   orc_ncoord ncoord;     /* erf, no cut                                    */
   orc_ncoord ncoord_en;  /* erf_en                                          */
} orc_eeqbc_model;

/* ---- mctc-lib pieces the path consumes ---- */
int orc_lattice_points_rep(const double* lat, const int rep[3], int origin, double* trans /* 3*ntr */);
int orc_lattice_points_count(const int rep[3], int origin);
/* returns ntr; *trans_out is malloc'ed (caller frees with orc_free) */
int orc_lattice_points_cutoff(const int periodic[3], const double* lat, double cutoff, double** trans_out);
void orc_free(void* p);
void orc_get_dir_trans(const double* lat, double* trans /* 3*125 */);
void orc_get_rec_trans(const double* lat, double* trans /* 3*124 */);

void orc_ncoord_get_cn(const orc_ncoord* nc, const orc_mol* mol, int ntr, const double* trans,
                       double* cn, double* dcndr, double* dcndL);

/* ---- wignerseitz.f90 / ewald.f90 ---- */
typedef struct {
   int ntr;          /* number of WSC translations (27 in 3-D)          */
   double* trans;    /* [3*ntr]                                          */
   int* nimg;        /* [nat*nat]  nimg(jat,iat)                         */
   int* tridx;       /* [ntr*nat*nat] tridx(img,jat,iat), 1-based        */
   int nimg_max;
} orc_wsc;
void orc_wsc_new(const orc_mol* mol, orc_wsc* wsc);
void orc_wsc_free(orc_wsc* wsc);
double orc_ewald_alpha(const double* lattice);

/* ---- EEQ2019 deferred procedures (model/eeq.f90) ---- */
void orc_eeq_get_xvec(const orc_eeq_model* m, const orc_mol* mol, const double* cn, double* xvec);
void orc_eeq_get_xvec_derivs(const orc_eeq_model* m, const orc_mol* mol, const double* cn,
                             const double* dcndr, const double* dcndL, double* dxdr, double* dxdL);
void orc_eeq_get_coulomb_matrix(const orc_eeq_model* m, const orc_mol* mol, double* amat);
void orc_eeq_get_coulomb_derivs(const orc_eeq_model* m, const orc_mol* mol, const double* qvec,
                                double* dadr, double* dadL, double* atrace);

/* ---- solve (model/type.F90:151-292); optional outputs may be NULL.
 *      returns 0, or 1/2/3 for the three fatal_error sites (factorisation / inversion / solution). */
int orc_eeq_solve(const orc_eeq_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                  const double* dcndr, const double* dcndL,
                  double* energy, double* gradient, double* sigma,
                  double* qvec, double* dqdr, double* dqdL);

/* get_charges (charge.f90:37-77): lattice points -> CN -> solve */
int orc_eeq_get_charges(const orc_eeq_model* m, const orc_mol* mol, double* qvec, double* dqdr, double* dqdL);

/* CLI sequence app/main.f90:114-118: CN -> solve(energy, gradient, sigma, qvec); cn_out optional */
int orc_eeq_singlepoint(const orc_eeq_model* m, const orc_mol* mol, double* cn_out, double* qvec,
                        double* energy, double* gradient, double* sigma, double* dqdr, double* dqdL);

/* ---- EEQ-BC (model/eeqbc.f90) ---- */
void orc_eeqbc_local_charge(const orc_eeqbc_model* m, const orc_mol* mol, int ntr, const double* trans,
                            double* qloc, double* dqlocdr, double* dqlocdL);
int orc_eeqbc_solve(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                    const double* dcndr, const double* dcndL, const double* dqlocdr, const double* dqlocdL,
                    double* energy, double* gradient, double* sigma,
                    double* qvec, double* dqdr, double* dqdL);
int orc_eeqbc_singlepoint(const orc_eeqbc_model* m, const orc_mol* mol, double* cn_out, double* qloc_out,
                          double* qvec, double* energy, double* gradient, double* sigma,
                          double* dqdr, double* dqdL);
void orc_eeqbc_get_coulomb_matrix(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn,
                                  const double* qloc, double* amat);
void orc_eeqbc_get_xvec(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn,
                        const double* qloc, double* xvec);

/* threading for the timed CPU baseline */
void orc_set_num_threads(int n);
int orc_get_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
