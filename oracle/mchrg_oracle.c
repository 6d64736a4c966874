/* TEST INFRASTRUCTURE -- CPU oracle, EEQ2019 part (see mchrg_oracle.h for the status header).
 *
 * Restates, loop by loop, what the reference computes on the EEQ hot path:
 *   model/type.F90:131-292  (get_dir_trans, get_rec_trans, solve)
 *   model/eeq.f90:97-570    (update, get_xvec, get_xvec_derivs, get_amat_0d/3d, get_damat_0d/3d)
 *   wignerseitz.f90:40-117  (new_wignerseitz_cell, get_pairs)
 *   ewald.f90:52-203        (get_alpha, search_alpha)
 *   charge.f90:37-77        (get_charges)
 * and the un-vendored mctc-lib v0.5.2 pieces it consumes (erf coordination number,
 * get_lattice_points; published algorithm restated from SURVEY.md 3.4 / Appendix A.3).
 * The linear algebra is the SAME LAPACK/BLAS the reference calls (dsytrf/dsytri/dsytrs/dsymv/
 * dgemv/dgemm, lapack.F90:165-363, blas.F90:250-630), taken from scipy's bundled OpenBLAS.
 *
 * Deviation from the reference (stated in BASELINE.md 3): OpenMP pair loops write disjoint
 * matrix entries directly instead of using one full-size private copy per thread
 * (eeq.f90:215,395-397); only the small reduction targets (atrace, dadL) are thread-private.
 */
#include "mchrg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- LAPACK / BLAS (Fortran ABI, scipy-prefixed symbols) ---- */
#ifndef ORC_BLAS
#define ORC_BLAS(name) scipy_##name##_
#endif
extern void ORC_BLAS(dsytrf)(const char* uplo, const int* n, double* a, const int* lda, int* ipiv,
                             double* work, const int* lwork, int* info, size_t);
extern void ORC_BLAS(dsytri)(const char* uplo, const int* n, double* a, const int* lda, const int* ipiv,
                             double* work, int* info, size_t);
extern void ORC_BLAS(dsytrs)(const char* uplo, const int* n, const int* nrhs, const double* a,
                             const int* lda, const int* ipiv, double* b, const int* ldb, int* info, size_t);
extern void ORC_BLAS(dsymv)(const char* uplo, const int* n, const double* alpha, const double* a,
                            const int* lda, const double* x, const int* incx, const double* beta,
                            double* y, const int* incy, size_t);
extern void ORC_BLAS(dgemv)(const char* trans, const int* m, const int* n, const double* alpha,
                            const double* a, const int* lda, const double* x, const int* incx,
                            const double* beta, double* y, const int* incy, size_t);
extern void ORC_BLAS(dgemm)(const char* ta, const char* tb, const int* m, const int* n, const int* k,
                            const double* alpha, const double* a, const int* lda, const double* b,
                            const int* ldb, const double* beta, double* c, const int* ldc, size_t, size_t);
extern void scipy_openblas_set_num_threads(int);

static const double PI = 3.14159265358979323846264338327950288;
#define SQRTPI 1.77245385090551602729816748334114518   /* sqrt(pi)   eeq.f90:56 */
#define SQRT2PI 0.797884560802865355879892119868763737 /* sqrt(2/pi) eeq.f90:57 */
static double eps_sqrt(void) { return sqrt(DBL_EPSILON); } /* eeq.f90:58, wignerseitz.f90:33 */

void orc_free(void* p) { free(p); }

void orc_set_num_threads(int n) {
#ifdef _OPENMP
   omp_set_num_threads(n);
#endif
   scipy_openblas_set_num_threads(n);
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

/* ================= mctc-lib: 3x3 helpers, lattice points ================= */

static double det3(const double* a) { /* column-major a(i,j) = a[i+3j] */
   return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[3] * (a[1] * a[8] - a[2] * a[7]) +
          a[6] * (a[1] * a[5] - a[2] * a[4]);
}

/* inverse of a 3x3 (mctc_io_math matinv_3x3: adjugate / determinant) */
static void inv3(const double* a, double* b) {
   double d = det3(a), s = 1.0 / d;
#define A(i, j) a[(i) + 3 * (j)]
#define B(i, j) b[(i) + 3 * (j)]
   B(0, 0) = s * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1));
   B(1, 0) = s * (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2));
   B(2, 0) = s * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
   B(0, 1) = s * (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2));
   B(1, 1) = s * (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0));
   B(2, 1) = s * (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1));
   B(0, 2) = s * (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1));
   B(1, 2) = s * (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2));
   B(2, 2) = s * (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
#undef A
#undef B
}

int orc_lattice_points_count(const int rep[3], int origin) {
   return (2 * rep[0] + 1) * (2 * rep[1] + 1) * (2 * rep[2] + 1) - (origin ? 0 : 1);
}

/* mctc_cutoff get_lattice_points(lat, rep, origin, trans): ix,iy,iz = 0..rep, for each the signs
 * (+, then - when the index is non-zero); origin first. */
int orc_lattice_points_rep(const double* lat, const int rep[3], int origin, double* trans) {
   int itr = 0;
   for (int ix = 0; ix <= rep[0]; ++ix)
      for (int iy = 0; iy <= rep[1]; ++iy)
         for (int iz = 0; iz <= rep[2]; ++iz) {
            if (!origin && ix == 0 && iy == 0 && iz == 0) continue;
            for (int jx = 1; jx >= (ix > 0 ? -1 : 1); jx -= 2)
               for (int jy = 1; jy >= (iy > 0 ? -1 : 1); jy -= 2)
                  for (int jz = 1; jz >= (iz > 0 ? -1 : 1); jz -= 2) {
                     for (int c = 0; c < 3; ++c)
                        trans[3 * itr + c] = lat[c] * (ix * jx) + lat[3 + c] * (iy * jy) + lat[6 + c] * (iz * jz);
                     ++itr;
                  }
         }
   return itr;
}

static void cross3(const double* a, const double* b, double* c) {
   c[0] = a[1] * b[2] - a[2] * b[1];
   c[1] = a[2] * b[0] - a[0] * b[2];
   c[2] = a[0] * b[1] - a[1] * b[0];
}

/* mctc_cutoff get_lattice_points(periodic, lat, cutoff, trans): 0-D -> one zero vector; otherwise
 * rep(i) = ceiling(cutoff / spacing of the lattice planes spanned by the other two vectors). */
int orc_lattice_points_cutoff(const int periodic[3], const double* lat, double cutoff, double** trans_out) {
   if (!(periodic[0] || periodic[1] || periodic[2])) {
      *trans_out = (double*)calloc(3, sizeof(double));
      return 1;
   }
   int rep[3];
   double nrm[3];
   for (int k = 0; k < 3; ++k) {
      const double* a = lat + 3 * ((k + 1) % 3);
      const double* b = lat + 3 * ((k + 2) % 3);
      cross3(a, b, nrm);
      double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      double cosk = (nrm[0] * lat[3 * k] + nrm[1] * lat[3 * k + 1] + nrm[2] * lat[3 * k + 2]) / len;
      rep[k] = (int)ceil(fabs(cutoff / cosk));
      if (!periodic[k]) rep[k] = 0;
   }
   int ntr = orc_lattice_points_count(rep, 1);
   *trans_out = (double*)malloc(sizeof(double) * 3 * (size_t)ntr);
   orc_lattice_points_rep(lat, rep, 1, *trans_out);
   return ntr;
}

/* model/type.F90:131-138 */
void orc_get_dir_trans(const double* lat, double* trans) {
   const int rep[3] = {2, 2, 2};
   orc_lattice_points_rep(lat, rep, 1, trans);
}

/* model/type.F90:140-149 : rec_lat = 2 pi transpose(inv(lattice)), origin excluded */
static void rec_lattice(const double* lat, double* rec) {
   double inv[9];
   inv3(lat, inv);
   for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rec[i + 3 * j] = 2.0 * PI * inv[j + 3 * i];
}
void orc_get_rec_trans(const double* lat, double* trans) {
   const int rep[3] = {2, 2, 2};
   double rec[9];
   rec_lattice(lat, rec);
   orc_lattice_points_rep(rec, rep, 0, trans);
}

/* ================= mctc-lib: erf coordination number ================= */

/* cn_count%erf / erf_en with derivatives; log-cut when nc->cut > 0.
 * dcndr / dcndL may both be NULL (charges-only path). */
void orc_ncoord_get_cn(const orc_ncoord* nc, const orc_mol* mol, int ntr, const double* trans,
                       double* cn, double* dcndr, double* dcndL) {
   const int nat = mol->nat;
   const int grad = (dcndr != NULL && dcndL != NULL);
   const double cutoff2 = nc->cutoff * nc->cutoff;
   const double dirf = nc->en ? -1.0 : 1.0; /* directed_factor */
   memset(cn, 0, sizeof(double) * nat);
   if (grad) {
      memset(dcndr, 0, sizeof(double) * 3 * (size_t)nat * nat);
      memset(dcndL, 0, sizeof(double) * 9 * (size_t)nat);
   }
   /* pair loop in the reference's order (iat outer, jat <= iat, translations inner); the
    * off-diagonal dcndr entries of a pair are touched by one thread only, the per-atom targets
    * (cn, dcndr(:,i,i), dcndL) are reduced from thread-private copies */
#pragma omp parallel
   {
      double* cn_p = (double*)calloc((size_t)nat, sizeof(double));
      double* dg_p = grad ? (double*)calloc(3 * (size_t)nat, sizeof(double)) : NULL;
      double* dL_p = grad ? (double*)calloc(9 * (size_t)nat, sizeof(double)) : NULL;
#pragma omp for schedule(dynamic, 8)
      for (int iat = 0; iat < nat; ++iat) {
         const int izp = mol->id[iat] - 1;
         for (int jat = 0; jat <= iat; ++jat) {
            const int jzp = mol->id[jat] - 1;
            const double den = nc->en ? (nc->en[jzp] - nc->en[izp]) : 1.0;
            const double rc = nc->rcov[izp] + nc->rcov[jzp];
            const double rcn = (nc->norm_exp == 1.0) ? rc : pow(rc, nc->norm_exp);
            for (int itr = 0; itr < ntr; ++itr) {
               double rij[3];
               for (int c = 0; c < 3; ++c)
                  rij[c] = mol->xyz[3 * iat + c] - (mol->xyz[3 * jat + c] + trans[3 * itr + c]);
               const double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
               if (r2 > cutoff2 || r2 < 1.0e-12) continue;
               const double r1 = sqrt(r2);
               const double arg = nc->kcn * (r1 - rc) / rcn;
               const double countf = den * 0.5 * (1.0 + erf(-arg));
               cn_p[iat] += countf;
               if (iat != jat) cn_p[jat] += countf * dirf;
               if (grad) {
                  const double dcount = den * (-nc->kcn / SQRTPI / rcn * exp(-arg * arg));
                  double countd[3];
                  for (int c = 0; c < 3; ++c) countd[c] = dcount * rij[c] / r1;
                  double* dij = dcndr + 3 * ((size_t)iat + (size_t)nat * jat); /* dcndr(:,iat,jat) */
                  double* dji = dcndr + 3 * ((size_t)jat + (size_t)nat * iat); /* dcndr(:,jat,iat) */
                  for (int c = 0; c < 3; ++c) {
                     dg_p[3 * iat + c] += countd[c];        /* dcndr(:,iat,iat) */
                     dg_p[3 * jat + c] -= countd[c] * dirf; /* dcndr(:,jat,jat) */
                     if (iat != jat) {
                        dij[c] += countd[c] * dirf;
                        dji[c] -= countd[c];
                     } else { /* same entry as the diagonal: +dirf*c - c */
                        dg_p[3 * iat + c] += countd[c] * dirf - countd[c];
                     }
                  }
                  for (int b = 0; b < 3; ++b)
                     for (int a = 0; a < 3; ++a) {
                        const double st = countd[b] * rij[a];
                        dL_p[a + 3 * b + 9 * (size_t)iat] += st;
                        if (iat != jat) dL_p[a + 3 * b + 9 * (size_t)jat] += st * dirf;
                     }
               }
            }
         }
      }
#pragma omp critical(orc_ncoord)
      {
         for (int i = 0; i < nat; ++i) cn[i] += cn_p[i];
         if (grad) {
            for (int i = 0; i < nat; ++i)
               for (int c = 0; c < 3; ++c) dcndr[c + 3 * ((size_t)i + (size_t)nat * i)] += dg_p[3 * i + c];
            for (size_t k = 0; k < 9 * (size_t)nat; ++k) dcndL[k] += dL_p[k];
         }
      }
      free(cn_p); free(dg_p); free(dL_p);
   }
   if (nc->cut > 0.0) {
      /* cn <- log(1+e^cut) - log(1+e^(cut-cn)); derivatives of atom i scaled with the un-cut cn */
      const double ecut = exp(nc->cut);
      for (int iat = 0; iat < nat; ++iat) {
         const double c0 = cn[iat];
         if (grad) {
            const double f = ecut / (ecut + exp(c0));
            double* d = dcndr + 3 * (size_t)nat * iat;
            for (size_t k = 0; k < 3 * (size_t)nat; ++k) d[k] *= f;
            for (int k = 0; k < 9; ++k) dcndL[k + 9 * (size_t)iat] *= f;
         }
         cn[iat] = log(1.0 + ecut) - log(1.0 + exp(nc->cut - c0));
      }
   }
}

/* ================= wignerseitz.f90 ================= */

/* get_pairs, wignerseitz.f90:75-117.  NOTE the returned indices count only the images that
 * survive the r2 < thr skip (the reference stores positions in the compacted dist() array and
 * later uses them to index wsc%trans directly, eeq.f90:276,287) -- reproduced literally. */
static int wsc_get_pairs(int ntr, const double* trans, const double* rij, int* list, double* dist, char* mask) {
   const double thr = eps_sqrt(), tol = 0.01;
   int img = 0, iws = 0;
   for (int i = 0; i < ntr; ++i) { list[i] = 0; mask[i] = 1; }
   for (int itr = 0; itr < ntr; ++itr) {
      double v0 = rij[0] - trans[3 * itr], v1 = rij[1] - trans[3 * itr + 1], v2 = rij[2] - trans[3 * itr + 2];
      double r2 = v0 * v0 + v1 * v1 + v2 * v2;
      if (r2 < thr) continue;
      dist[img++] = r2;
   }
   if (img == 0) return 0;
   int pos = 0;
   for (int i = 1; i < img; ++i)
      if (dist[i] < dist[pos]) pos = i; /* minloc: first minimum */
   const double r2 = dist[pos];
   mask[pos] = 0;
   list[iws++] = pos + 1;
   if (img <= iws) return iws;
   for (;;) {
      pos = -1;
      for (int i = 0; i < img; ++i)
         if (mask[i] && (pos < 0 || dist[i] < dist[pos])) pos = i;
      /* minloc over an all-false mask yields 0 in Fortran (out-of-bounds read in the reference);
       * it can only happen when every image is equidistant, which a lattice excludes. */
      if (pos < 0) break;
      if (fabs(dist[pos] - r2) > tol) break;
      mask[pos] = 0;
      list[iws++] = pos + 1;
   }
   return iws;
}

void orc_wsc_new(const orc_mol* mol, orc_wsc* wsc) {
   const int nat = mol->nat;
   wsc->ntr = orc_lattice_points_cutoff(mol->periodic, mol->lattice, eps_sqrt(), &wsc->trans);
   const int ntr = wsc->ntr;
   wsc->nimg = (int*)malloc(sizeof(int) * (size_t)nat * nat);
   wsc->tridx = (int*)malloc(sizeof(int) * (size_t)ntr * nat * nat);
   int nimg_max = 0;
#pragma omp parallel reduction(max : nimg_max)
   {
      double* dist = (double*)malloc(sizeof(double) * ntr);
      char* mask = (char*)malloc(ntr);
#pragma omp for schedule(static)
      for (int iat = 0; iat < nat; ++iat)
         for (int jat = 0; jat < nat; ++jat) {
            double vec[3];
            for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * iat + c] - mol->xyz[3 * jat + c];
            int* list = wsc->tridx + (size_t)ntr * ((size_t)jat + (size_t)nat * iat);
            int n = wsc_get_pairs(ntr, wsc->trans, vec, list, dist, mask);
            wsc->nimg[(size_t)jat + (size_t)nat * iat] = n;
            if (n > nimg_max) nimg_max = n;
         }
      free(dist);
      free(mask);
   }
   wsc->nimg_max = nimg_max;
}

void orc_wsc_free(orc_wsc* wsc) {
   free(wsc->trans);
   free(wsc->nimg);
   free(wsc->tridx);
   memset(wsc, 0, sizeof(*wsc));
}

/* ================= ewald.f90 ================= */

static double ew_rec_term(double gg, double alpha, double vol) { /* ewald.f90:187-203 */
   return 4.0 * PI * (exp(-0.25 * gg * gg / (alpha * alpha)) / (vol * gg * gg));
}
static double ew_dir_term(double rr, double alpha) { return erfc(alpha * rr) / rr; } /* :169-182 */
static double ew_diff(double alpha, double rlen, double dlen, double vol) {          /* :141-165 */
   return (ew_rec_term(4 * rlen, alpha, vol) - ew_rec_term(5 * rlen, alpha, vol)) -
          (ew_dir_term(2 * dlen, alpha) - ew_dir_term(3 * dlen, alpha));
}

/* get_alpha + search_alpha, ewald.f90:52-135 */
double orc_ewald_alpha(const double* lattice) {
   const double tol = eps_sqrt(), alpha0 = 1.0e-8;
   const int niter = 30;
   double rec[9];
   const double vol = fabs(det3(lattice));
   rec_lattice(lattice, rec);
   double rmin = HUGE_VAL, dmin = HUGE_VAL;
   for (int k = 0; k < 3; ++k) {
      double r = rec[3 * k] * rec[3 * k] + rec[3 * k + 1] * rec[3 * k + 1] + rec[3 * k + 2] * rec[3 * k + 2];
      double d = lattice[3 * k] * lattice[3 * k] + lattice[3 * k + 1] * lattice[3 * k + 1] +
                 lattice[3 * k + 2] * lattice[3 * k + 2];
      if (r < rmin) rmin = r;
      if (d < dmin) dmin = d;
   }
   const double rlen = sqrt(rmin), dlen = sqrt(dmin);
   int stat = 0;
   double alpha = alpha0, alpl = 0, alpr = 0;
   double diff = ew_diff(alpha, rlen, dlen, vol);
   while (diff < -tol && alpha <= DBL_MAX) {
      alpha *= 2.0;
      diff = ew_diff(alpha, rlen, dlen, vol);
   }
   if (alpha > DBL_MAX) stat = 1;
   else if (alpha == alpha0) stat = 2;
   if (stat == 0) {
      alpl = 0.5 * alpha;
      while (diff < tol && alpha <= DBL_MAX) {
         alpha *= 2.0;
         diff = ew_diff(alpha, rlen, dlen, vol);
      }
      if (alpha > DBL_MAX) stat = 3;
   }
   if (stat == 0) {
      alpr = alpha;
      alpha = (alpl + alpr) * 0.5;
      int ibs = 0;
      diff = ew_diff(alpha, rlen, dlen, vol);
      while (fabs(diff) > tol && ibs <= niter) {
         if (diff < 0) alpl = alpha;
         else alpr = alpha;
         alpha = (alpl + alpr) * 0.5;
         diff = ew_diff(alpha, rlen, dlen, vol);
         ++ibs;
      }
      if (ibs > niter) stat = 4;
   }
   if (stat != 0) alpha = 0.25;
   return alpha;
}

/* ================= model/eeq.f90 ================= */

static int mol_periodic(const orc_mol* mol) { return mol->periodic[0] || mol->periodic[1] || mol->periodic[2]; }

/* get_xvec, eeq.f90:127-150 */
void orc_eeq_get_xvec(const orc_eeq_model* m, const orc_mol* mol, const double* cn, double* xvec) {
   const double reg = 1.0e-14;
   for (int iat = 0; iat < mol->nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      const double tmp = m->kcnchi[izp] / sqrt(cn[iat] + reg);
      xvec[iat] = -m->chi[izp] + tmp * cn[iat];
   }
   xvec[mol->nat] = mol->charge;
}

/* get_xvec_derivs, eeq.f90:152-180 */
void orc_eeq_get_xvec_derivs(const orc_eeq_model* m, const orc_mol* mol, const double* cn,
                             const double* dcndr, const double* dcndL, double* dxdr, double* dxdL) {
   const double reg = 1.0e-14;
   const int nat = mol->nat, ndim = nat + 1;
   memset(dxdr, 0, sizeof(double) * 3 * (size_t)nat * ndim);
   memset(dxdL, 0, sizeof(double) * 9 * (size_t)ndim);
#pragma omp parallel for schedule(static)
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      const double tmp = m->kcnchi[izp] / sqrt(cn[iat] + reg);
      const double* s = dcndr + 3 * (size_t)nat * iat;
      double* d = dxdr + 3 * (size_t)nat * iat;
      for (size_t k = 0; k < 3 * (size_t)nat; ++k) d[k] = 0.5 * tmp * s[k] + d[k];
      for (int k = 0; k < 9; ++k) dxdL[k + 9 * (size_t)iat] = 0.5 * tmp * dcndL[k + 9 * (size_t)iat] + dxdL[k + 9 * (size_t)iat];
   }
}

/* get_amat_dir_3d, eeq.f90:309-329 */
static double amat_dir_3d(const double* rij, double gam, double alp, int ntr, const double* trans) {
   const double eps = eps_sqrt();
   double amat = 0.0;
   for (int itr = 0; itr < ntr; ++itr) {
      const double v0 = rij[0] + trans[3 * itr], v1 = rij[1] + trans[3 * itr + 1], v2 = rij[2] + trans[3 * itr + 2];
      const double r1 = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      if (r1 < eps) continue;
      amat += erf(gam * r1) / r1 - erf(alp * r1) / r1;
   }
   return amat;
}

/* get_amat_rec_3d, eeq.f90:331-352 */
static double amat_rec_3d(const double* rij, double vol, double alp, int ntr, const double* trans) {
   const double eps = eps_sqrt();
   const double fac = 4 * PI / vol;
   double amat = 0.0;
   for (int itr = 0; itr < ntr; ++itr) {
      const double* g = trans + 3 * itr;
      const double g2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
      if (g2 < eps) continue;
      amat += cos(rij[0] * g[0] + rij[1] * g[1] + rij[2] * g[2]) * fac * exp(-0.25 * g2 / (alp * alp)) / g2;
   }
   return amat;
}

static void border(double* amat, int nat) { /* eeq.f90:238-240 */
   const int ndim = nat + 1;
   for (int k = 0; k < ndim; ++k) {
      amat[nat + (size_t)ndim * k] = 1.0;
      amat[k + (size_t)ndim * nat] = 1.0;
   }
   amat[nat + (size_t)ndim * nat] = 0.0;
}

/* get_amat_0d, eeq.f90:199-242 */
static void eeq_amat_0d(const orc_eeq_model* m, const orc_mol* mol, double* amat) {
   const int nat = mol->nat, ndim = nat + 1;
   memset(amat, 0, sizeof(double) * (size_t)ndim * ndim);
#pragma omp parallel for schedule(dynamic, 16)
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         const double v0 = mol->xyz[3 * jat] - mol->xyz[3 * iat];
         const double v1 = mol->xyz[3 * jat + 1] - mol->xyz[3 * iat + 1];
         const double v2 = mol->xyz[3 * jat + 2] - mol->xyz[3 * iat + 2];
         const double r2 = v0 * v0 + v1 * v1 + v2 * v2;
         const double gam = 1.0 / (m->rad[izp] * m->rad[izp] + m->rad[jzp] * m->rad[jzp]);
         const double tmp = erf(sqrt(r2 * gam)) / sqrt(r2);
         amat[jat + (size_t)ndim * iat] += tmp;
         amat[iat + (size_t)ndim * jat] += tmp;
      }
      amat[iat + (size_t)ndim * iat] += m->eta[izp] + SQRT2PI / m->rad[izp];
   }
   border(amat, nat);
}

/* get_amat_3d, eeq.f90:244-307 */
static void eeq_amat_3d(const orc_eeq_model* m, const orc_mol* mol, const orc_wsc* wsc, double alpha, double* amat) {
   const int nat = mol->nat, ndim = nat + 1, ntw = wsc->ntr;
   double dtrans[3 * 125], rtrans[3 * 124];
   const double vol = fabs(det3(mol->lattice));
   orc_get_dir_trans(mol->lattice, dtrans);
   orc_get_rec_trans(mol->lattice, rtrans);
   memset(amat, 0, sizeof(double) * (size_t)ndim * ndim);
#pragma omp parallel for schedule(dynamic, 4)
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         const double gam = 1.0 / sqrt(m->rad[izp] * m->rad[izp] + m->rad[jzp] * m->rad[jzp]);
         const int nimg = wsc->nimg[jat + (size_t)nat * iat];
         const double wsw = 1.0 / (double)nimg;
         const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
         for (int img = 0; img < nimg; ++img) {
            const double* t = wsc->trans + 3 * (tri[img] - 1);
            double vec[3];
            for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c] + t[c];
            const double dtmp = amat_dir_3d(vec, gam, alpha, 125, dtrans);
            const double rtmp = amat_rec_3d(vec, vol, alpha, 124, rtrans);
            amat[jat + (size_t)ndim * iat] += (dtmp + rtmp) * wsw;
            amat[iat + (size_t)ndim * jat] += (dtmp + rtmp) * wsw;
         }
      }
      const double gam = 1.0 / sqrt(2.0 * m->rad[izp] * m->rad[izp]);
      const int nimg = wsc->nimg[iat + (size_t)nat * iat];
      const double wsw = 1.0 / (double)nimg;
      const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
      for (int img = 0; img < nimg; ++img) {
         const double* vec = wsc->trans + 3 * (tri[img] - 1);
         const double dtmp = amat_dir_3d(vec, gam, alpha, 125, dtrans);
         const double rtmp = amat_rec_3d(vec, vol, alpha, 124, rtrans);
         amat[iat + (size_t)ndim * iat] += (dtmp + rtmp) * wsw;
      }
      amat[iat + (size_t)ndim * iat] += m->eta[izp] + SQRT2PI / m->rad[izp] - 2 * alpha / SQRTPI;
   }
   border(amat, nat);
}

void orc_eeq_get_coulomb_matrix(const orc_eeq_model* m, const orc_mol* mol, double* amat) {
   if (mol_periodic(mol)) { /* update(): eeq.f90:119-123 */
      orc_wsc wsc;
      orc_wsc_new(mol, &wsc);
      const double alpha = orc_ewald_alpha(mol->lattice);
      eeq_amat_3d(m, mol, &wsc, alpha, amat);
      orc_wsc_free(&wsc);
   } else {
      eeq_amat_0d(m, mol, amat);
   }
}

/* get_damat_dir_3d, eeq.f90:510-538 */
static void damat_dir_3d(const double* rij, double gam, double alp, int ntr, const double* trans, double* dg, double* ds) {
   const double eps = eps_sqrt();
   const double gam2 = gam * gam, alp2 = alp * alp;
   memset(dg, 0, 3 * sizeof(double));
   memset(ds, 0, 9 * sizeof(double));
   for (int itr = 0; itr < ntr; ++itr) {
      double vec[3] = {rij[0] + trans[3 * itr], rij[1] + trans[3 * itr + 1], rij[2] + trans[3 * itr + 2]};
      const double r1 = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
      if (r1 < eps) continue;
      const double r2 = r1 * r1;
      const double gtmp = +2 * gam * exp(-r2 * gam2) / (SQRTPI * r2) - erf(r1 * gam) / (r2 * r1);
      const double atmp = -2 * alp * exp(-r2 * alp2) / (SQRTPI * r2) + erf(r1 * alp) / (r2 * r1);
      for (int c = 0; c < 3; ++c) dg[c] += (gtmp + atmp) * vec[c];
      for (int b = 0; b < 3; ++b)
         for (int a = 0; a < 3; ++a) ds[a + 3 * b] += (gtmp + atmp) * vec[b] * vec[a];
   }
}

/* get_damat_rec_3d, eeq.f90:540-570 */
static void damat_rec_3d(const double* rij, double vol, double alp, int ntr, const double* trans, double* dg, double* ds) {
   const double eps = eps_sqrt();
   const double fac = 4 * PI / vol, alp2 = alp * alp;
   memset(dg, 0, 3 * sizeof(double));
   memset(ds, 0, 9 * sizeof(double));
   for (int itr = 0; itr < ntr; ++itr) {
      const double* vec = trans + 3 * itr;
      const double g2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
      if (g2 < eps) continue;
      const double gv = rij[0] * vec[0] + rij[1] * vec[1] + rij[2] * vec[2];
      const double etmp = fac * exp(-0.25 * g2 / alp2) / g2;
      const double dtmp = -sin(gv) * etmp;
      for (int c = 0; c < 3; ++c) dg[c] += dtmp * vec[c];
      const double cg = etmp * cos(gv), pre = 2.0 / g2 + 0.5 / alp2;
      for (int b = 0; b < 3; ++b)
         for (int a = 0; a < 3; ++a) ds[a + 3 * b] += cg * (pre * vec[b] * vec[a] - (a == b ? 1.0 : 0.0));
   }
}

/* per-thread reduction scratch for atrace(3,nat) and dadL(3,3,ndim) */
typedef struct { double* atrace; double* dadL; } red_t;

/* get_damat_0d (eeq.f90:372-427) and get_damat_3d (eeq.f90:429-508) */
void orc_eeq_get_coulomb_derivs(const orc_eeq_model* m, const orc_mol* mol, const double* qvec,
                                double* dadr, double* dadL, double* atrace) {
   const int nat = mol->nat, ndim = nat + 1;
   const int pbc = mol_periodic(mol);
   memset(atrace, 0, sizeof(double) * 3 * (size_t)nat);
   memset(dadr, 0, sizeof(double) * 3 * (size_t)nat * ndim);
   memset(dadL, 0, sizeof(double) * 9 * (size_t)ndim);
   orc_wsc wsc;
   double alpha = 0.0, vol = 0.0;
   double dtrans[3 * 125], rtrans[3 * 124];
   if (pbc) {
      orc_wsc_new(mol, &wsc);
      alpha = orc_ewald_alpha(mol->lattice);
      vol = fabs(det3(mol->lattice));
      orc_get_dir_trans(mol->lattice, dtrans);
      orc_get_rec_trans(mol->lattice, rtrans);
   }
#pragma omp parallel
   {
      double* atr = (double*)calloc(3 * (size_t)nat, sizeof(double));
      double* dL = (double*)calloc(9 * (size_t)ndim, sizeof(double));
#pragma omp for schedule(dynamic, 8)
      for (int iat = 0; iat < nat; ++iat) {
         const int izp = mol->id[iat] - 1;
         for (int jat = 0; jat < iat; ++jat) {
            const int jzp = mol->id[jat] - 1;
            double dG[3], dS[9];
            const double gam = 1.0 / sqrt(m->rad[izp] * m->rad[izp] + m->rad[jzp] * m->rad[jzp]);
            if (!pbc) {
               double vec[3];
               for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c];
               const double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
               const double arg = gam * gam * r2;
               const double dtmp = 2.0 * gam * exp(-arg) / (SQRTPI * r2) - erf(sqrt(arg)) / (r2 * sqrt(r2));
               for (int c = 0; c < 3; ++c) dG[c] = dtmp * vec[c];
               for (int b = 0; b < 3; ++b)
                  for (int a = 0; a < 3; ++a) dS[a + 3 * b] = dG[b] * vec[a];
            } else {
               memset(dG, 0, sizeof dG);
               memset(dS, 0, sizeof dS);
               const int nimg = wsc.nimg[jat + (size_t)nat * iat];
               const double wsw = 1.0 / (double)nimg;
               const int* tri = wsc.tridx + (size_t)wsc.ntr * (jat + (size_t)nat * iat);
               for (int img = 0; img < nimg; ++img) {
                  const double* t = wsc.trans + 3 * (tri[img] - 1);
                  double vec[3], dGd[3], dSd[9], dGr[3], dSr[9];
                  for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c] + t[c];
                  damat_dir_3d(vec, gam, alpha, 125, dtrans, dGd, dSd);
                  damat_rec_3d(vec, vol, alpha, 124, rtrans, dGr, dSr);
                  for (int c = 0; c < 3; ++c) dG[c] += (dGd[c] + dGr[c]) * wsw;
                  for (int k = 0; k < 9; ++k) dS[k] += (dSd[k] + dSr[k]) * wsw;
               }
            }
            double* dij = dadr + 3 * ((size_t)iat + (size_t)nat * jat); /* dadr(:,iat,jat) */
            double* dji = dadr + 3 * ((size_t)jat + (size_t)nat * iat); /* dadr(:,jat,iat) */
            for (int c = 0; c < 3; ++c) {
               atr[3 * iat + c] += -dG[c] * qvec[jat];
               atr[3 * jat + c] += +dG[c] * qvec[iat];
               dij[c] += -dG[c] * qvec[iat];
               dji[c] += +dG[c] * qvec[jat];
            }
            for (int k = 0; k < 9; ++k) {
               dL[k + 9 * (size_t)jat] += dS[k] * qvec[iat];
               dL[k + 9 * (size_t)iat] += dS[k] * qvec[jat];
            }
         }
         if (pbc) { /* self-image stress term, eeq.f90:488-497 */
            double dS[9];
            memset(dS, 0, sizeof dS);
            const double gam = 1.0 / sqrt(2.0 * m->rad[izp] * m->rad[izp]);
            const int nimg = wsc.nimg[iat + (size_t)nat * iat];
            const double wsw = 1.0 / (double)nimg;
            const int* tri = wsc.tridx + (size_t)wsc.ntr * (iat + (size_t)nat * iat);
            for (int img = 0; img < nimg; ++img) {
               const double* vec = wsc.trans + 3 * (tri[img] - 1);
               double dGd[3], dSd[9], dGr[3], dSr[9];
               damat_dir_3d(vec, gam, alpha, 125, dtrans, dGd, dSd);
               damat_rec_3d(vec, vol, alpha, 124, rtrans, dGr, dSr);
               for (int k = 0; k < 9; ++k) dS[k] += (dSd[k] + dSr[k]) * wsw;
            }
            for (int k = 0; k < 9; ++k) dL[k + 9 * (size_t)iat] += dS[k] * qvec[iat];
         }
      }
#pragma omp critical(orc_damat)
      {
         for (size_t k = 0; k < 3 * (size_t)nat; ++k) atrace[k] += atr[k];
         for (size_t k = 0; k < 9 * (size_t)ndim; ++k) dadL[k] += dL[k];
      }
      free(atr);
      free(dL);
   }
   if (pbc) orc_wsc_free(&wsc);
}

/* ================= model/type.F90: solve ================= */

/* Shared tail of solve() for both models: everything after amat / xvec exist.
 * dxdr/dxdL and dadr/dadL/atrace are produced by the callbacks so EEQ-BC can reuse this. */
typedef struct {
   void (*xvec_derivs)(const void* model, const orc_mol* mol, const void* cache, double* dxdr, double* dxdL);
   void (*coulomb_derivs)(const void* model, const orc_mol* mol, const void* cache, const double* vrhs,
                          double* dadr, double* dadL, double* atrace);
} orc_solve_vtbl;

int orc_solve_core(const void* model, const orc_mol* mol, const void* cache, const orc_solve_vtbl* vt,
                   const double* amat, const double* xvec, int dcn,
                   double* energy, double* gradient, double* sigma, double* qvec, double* dqdr, double* dqdL) {
   const int nat = mol->nat, ndim = nat + 1, one = 1;
   const int grad = gradient && sigma && dcn;
   const int cpq = dqdr && dqdL && dcn;
   const size_t n2 = (size_t)ndim * ndim;
   int info = 0;
   double* vrhs = (double*)malloc(sizeof(double) * ndim);
   double* ainv = (double*)malloc(sizeof(double) * n2);
   int* ipiv = (int*)malloc(sizeof(int) * ndim);
   memcpy(vrhs, xvec, sizeof(double) * ndim);
   memcpy(ainv, amat, sizeof(double) * n2);

   /* sytrf(ainv, ipiv, uplo='l'), type.F90:221, lapack.F90:165-201 (workspace query, then factor) */
   {
      double test = 0.0;
      int lwork = -1;
      ORC_BLAS(dsytrf)("l", &ndim, ainv, &ndim, ipiv, &test, &lwork, &info, 1);
      if (info == 0) {
         lwork = (int)lround(test);
         double* work = (double*)malloc(sizeof(double) * (size_t)lwork);
         ORC_BLAS(dsytrf)("l", &ndim, ainv, &ndim, ipiv, work, &lwork, &info, 1);
         free(work);
      }
   }
   if (info != 0) { free(vrhs); free(ainv); free(ipiv); return 1; }

   if (cpq) {
      double* work = (double*)malloc(sizeof(double) * ndim);
      ORC_BLAS(dsytri)("l", &ndim, ainv, &ndim, ipiv, work, &info, 1); /* type.F90:229 */
      free(work);
      if (info != 0) { free(vrhs); free(ainv); free(ipiv); return 2; }
      const double a1 = 1.0, b0 = 0.0;
      ORC_BLAS(dsymv)("l", &ndim, &a1, ainv, &ndim, xvec, &one, &b0, vrhs, &one, 1); /* :235 */
      for (int ic = 0; ic < ndim; ++ic)                                                /* :236-240 */
         for (int jc = ic + 1; jc < ndim; ++jc) ainv[ic + (size_t)ndim * jc] = ainv[jc + (size_t)ndim * ic];
   } else {
      ORC_BLAS(dsytrs)("l", &ndim, &one, ainv, &ndim, ipiv, vrhs, &ndim, &info, 1); /* :243 */
      if (info != 0) { free(vrhs); free(ainv); free(ipiv); return 3; }
   }

   if (qvec) memcpy(qvec, vrhs, sizeof(double) * nat);

   if (energy) { /* type.F90:255-262: xtmp = 0.5*J*q - x ; energy += q * xtmp */
      double* jmat = (double*)malloc(sizeof(double) * (size_t)nat * nat);
      double* xtmp = (double*)malloc(sizeof(double) * nat);
      for (int j = 0; j < nat; ++j) memcpy(jmat + (size_t)nat * j, amat + (size_t)ndim * j, sizeof(double) * nat);
      memcpy(xtmp, xvec, sizeof(double) * nat);
      const double ah = 0.5, bm = -1.0;
      ORC_BLAS(dsymv)("l", &nat, &ah, jmat, &nat, vrhs, &one, &bm, xtmp, &one, 1);
      for (int i = 0; i < nat; ++i) energy[i] += vrhs[i] * xtmp[i];
      free(jmat);
      free(xtmp);
   }

   if (grad || cpq) {
      const size_t n3 = 3 * (size_t)nat * ndim;
      double* dadr = (double*)malloc(sizeof(double) * n3);
      double* dadL = (double*)malloc(sizeof(double) * 9 * (size_t)ndim);
      double* atrace = (double*)malloc(sizeof(double) * 3 * (size_t)nat);
      double* dxdr = (double*)malloc(sizeof(double) * n3);
      double* dxdL = (double*)malloc(sizeof(double) * 9 * (size_t)ndim);
      vt->xvec_derivs(model, mol, cache, dxdr, dxdL);
      vt->coulomb_derivs(model, mol, cache, vrhs, dadr, dadL, atrace);
      for (int iat = 0; iat < nat; ++iat) /* type.F90:270-272 */
         for (int c = 0; c < 3; ++c) dadr[c + 3 * ((size_t)iat + (size_t)nat * iat)] += atrace[c + 3 * iat];

      if (grad) { /* type.F90:275-281 */
         const int m3 = 3 * nat, m9 = 9;
         const double ah = 0.5, am = -1.0, b1 = 1.0;
         memset(gradient, 0, sizeof(double) * 3 * (size_t)nat);
         ORC_BLAS(dgemv)("n", &m3, &nat, &ah, dadr, &m3, vrhs, &one, &b1, gradient, &one, 1);
         ORC_BLAS(dgemv)("n", &m3, &nat, &am, dxdr, &m3, vrhs, &one, &b1, gradient, &one, 1);
         ORC_BLAS(dgemv)("n", &m9, &ndim, &ah, dadL, &m9, vrhs, &one, &b1, sigma, &one, 1);
         ORC_BLAS(dgemv)("n", &m9, &ndim, &am, dxdL, &m9, vrhs, &one, &b1, sigma, &one, 1);
      }
      if (cpq) { /* type.F90:283-291 */
         const int m3 = 3 * nat, m9 = 9;
         const double am = -1.0, b0 = 0.0;
         for (size_t k = 0; k < 3 * (size_t)nat * nat; ++k) dadr[k] = -dxdr[k] + dadr[k];
         for (size_t k = 0; k < 9 * (size_t)nat; ++k) dadL[k] = -dxdL[k] + dadL[k];
         ORC_BLAS(dgemm)("n", "n", &m3, &nat, &ndim, &am, dadr, &m3, ainv, &ndim, &b0, dqdr, &m3, 1, 1);
         ORC_BLAS(dgemm)("n", "n", &m9, &nat, &ndim, &am, dadL, &m9, ainv, &ndim, &b0, dqdL, &m9, 1, 1);
      }
      free(dadr); free(dadL); free(atrace); free(dxdr); free(dxdL);
   }
   free(vrhs); free(ainv); free(ipiv);
   return 0;
}

typedef struct { const double* cn; const double* dcndr; const double* dcndL; } eeq_cache;

static void eeq_vt_xvec_derivs(const void* model, const orc_mol* mol, const void* cache, double* dxdr, double* dxdL) {
   const eeq_cache* c = (const eeq_cache*)cache;
   orc_eeq_get_xvec_derivs((const orc_eeq_model*)model, mol, c->cn, c->dcndr, c->dcndL, dxdr, dxdL);
}
static void eeq_vt_coulomb_derivs(const void* model, const orc_mol* mol, const void* cache, const double* vrhs,
                                  double* dadr, double* dadL, double* atrace) {
   (void)cache;
   orc_eeq_get_coulomb_derivs((const orc_eeq_model*)model, mol, vrhs, dadr, dadL, atrace);
}

int orc_eeq_solve(const orc_eeq_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                  const double* dcndr, const double* dcndL,
                  double* energy, double* gradient, double* sigma,
                  double* qvec, double* dqdr, double* dqdL) {
   (void)qloc; /* EEQ2019 ignores the local charge, eeq.f90:97-125 */
   const int ndim = mol->nat + 1;
   const orc_solve_vtbl vt = {eeq_vt_xvec_derivs, eeq_vt_coulomb_derivs};
   eeq_cache cache = {cn, dcndr, dcndL};
   double* amat = (double*)malloc(sizeof(double) * (size_t)ndim * ndim);
   double* xvec = (double*)malloc(sizeof(double) * ndim);
   orc_eeq_get_coulomb_matrix(m, mol, amat);
   orc_eeq_get_xvec(m, mol, cn, xvec);
   int stat = orc_solve_core(m, mol, &cache, &vt, amat, xvec, dcndr && dcndL,
                             energy, gradient, sigma, qvec, dqdr, dqdL);
   free(amat);
   free(xvec);
   return stat;
}

/* get_charges, charge.f90:37-77 */
int orc_eeq_get_charges(const orc_eeq_model* m, const orc_mol* mol, double* qvec, double* dqdr, double* dqdL) {
   const int nat = mol->nat;
   const int grad = dqdr && dqdL;
   double* cn = (double*)malloc(sizeof(double) * nat);
   double *dcndr = NULL, *dcndL = NULL, *trans = NULL;
   if (grad) {
      dcndr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nat);
      dcndL = (double*)malloc(sizeof(double) * 9 * (size_t)nat);
   }
   int ntr = orc_lattice_points_cutoff(mol->periodic, mol->lattice, m->ncoord.cutoff, &trans);
   orc_ncoord_get_cn(&m->ncoord, mol, ntr, trans, cn, dcndr, dcndL);
   int stat = orc_eeq_solve(m, mol, cn, NULL, dcndr, dcndL, NULL, NULL, NULL, qvec, dqdr, dqdL);
   free(cn); free(dcndr); free(dcndL); free(trans);
   return stat;
}

/* app/main.f90:114-118 call sequence (CN with derivatives -> solve with everything requested) */
int orc_eeq_singlepoint(const orc_eeq_model* m, const orc_mol* mol, double* cn_out, double* qvec,
                        double* energy, double* gradient, double* sigma, double* dqdr, double* dqdL) {
   const int nat = mol->nat;
   const int grad = (gradient && sigma) || (dqdr && dqdL);
   double* cn = (double*)malloc(sizeof(double) * nat);
   double *dcndr = NULL, *dcndL = NULL, *trans = NULL;
   if (grad) {
      dcndr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nat);
      dcndL = (double*)malloc(sizeof(double) * 9 * (size_t)nat);
   }
   int ntr = orc_lattice_points_cutoff(mol->periodic, mol->lattice, m->ncoord.cutoff, &trans);
   orc_ncoord_get_cn(&m->ncoord, mol, ntr, trans, cn, dcndr, dcndL);
   if (cn_out) memcpy(cn_out, cn, sizeof(double) * nat);
   int stat = orc_eeq_solve(m, mol, cn, NULL, dcndr, dcndL, energy, gradient, sigma, qvec, dqdr, dqdL);
   free(cn); free(dcndr); free(dcndL); free(trans);
   return stat;
}

/* ================= element tables (param/eeq2019.f90, param/eeqbc2025.f90, mctc_data) ================= */
#include "param_tables_oracle.inc" /* the oracle's own tables (oracle/gen_tables.py), not the product's */

#define ORC_AATOAU (1.0 / 0.529177210903)

/* out = {chi, eta, kcnchi, rad, rcov}; get_eeq_* (param/eeq2019.f90) + get_covalent_rad (mctc_data);
 * returns -1 for an unknown element like the reference getters. */
int orc_param_eeq2019(int z, double* out) {
   if (z < 1 || z > 103) { for (int k = 0; k < 5; ++k) out[k] = -1.0; return -1; }
   out[0] = orc_eeq_chi[z - 1];
   out[1] = orc_eeq_eta[z - 1];
   out[2] = orc_eeq_kcnchi[z - 1];
   out[3] = orc_eeq_rad[z - 1];
   out[4] = orc_d3_rcov_angstrom[z - 1] * ORC_AATOAU * 4.0 / 3.0;
   return 0;
}

/* out = {chi, eta, rad, kcnchi, kqchi, kqeta, cap, cov_radii, avg_cn} (param/eeqbc2025.f90) */
int orc_param_eeqbc2025(int z, double* out) {
   if (z < 1 || z > 103) { for (int k = 0; k < 9; ++k) out[k] = -1.0; return -1; }
   out[0] = orc_eeqbc_chi[z - 1];
   out[1] = orc_eeqbc_eta[z - 1];
   out[2] = orc_eeqbc_rad[z - 1];
   out[3] = orc_eeqbc_kcnchi[z - 1];
   out[4] = orc_eeqbc_kqchi[z - 1];
   out[5] = orc_eeqbc_kqeta[z - 1];
   out[6] = orc_eeqbc_cap[z - 1];
   out[7] = orc_eeqbc_cov_radii[z - 1];
   out[8] = orc_eeqbc_avg_cn[z - 1];
   return 0;
}

/* ================= batch driver for the timed CPU baseline =================
 * One reference single point (app/main.f90:114-118: CN -> solve with energy/gradient) per molecule,
 * molecules distributed over the host threads (the inner OpenMP regions then run serially). */
#include <stdint.h>
int orc_eeq2019_singlepoint_batch(int nmol, const int64_t* offsets, const int* num, const double* xyz,
                                  const double* charge, double* qvec, double* energy, double* gradient,
                                  int* status, int want_gradient) {
   static double chi[103], eta[103], kcnchi[103], rad[103], rcov[103];
   for (int z = 1; z <= 103; ++z) {
      double p[5];
      orc_param_eeq2019(z, p);
      chi[z - 1] = p[0]; eta[z - 1] = p[1]; kcnchi[z - 1] = p[2]; rad[z - 1] = p[3]; rcov[z - 1] = p[4];
   }
   orc_eeq_model model;
   model.nid = 103;
   model.chi = chi; model.rad = rad; model.eta = eta; model.kcnchi = kcnchi;
   model.ncoord.rcov = rcov; model.ncoord.en = NULL; model.ncoord.kcn = 7.5; model.ncoord.norm_exp = 1.0;
   model.ncoord.cutoff = 25.0; model.ncoord.cut = 8.0;
   int nfail = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nfail)
   for (int m = 0; m < nmol; ++m) {
      const int64_t off = offsets[m];
      orc_mol mol;
      memset(&mol, 0, sizeof mol);
      mol.nat = (int)(offsets[m + 1] - off);
      mol.nid = 103;
      mol.id = num + off; /* species index == atomic number */
      mol.xyz = xyz + 3 * off;
      mol.charge = charge[m];
      double sigma[9] = {0};
      for (int i = 0; i < mol.nat; ++i) energy[off + i] = 0.0;
      status[m] = orc_eeq_singlepoint(&model, &mol, NULL, qvec + off, energy + off,
                                      want_gradient ? gradient + 3 * off : NULL, want_gradient ? sigma : NULL, NULL, NULL);
      if (status[m]) ++nfail;
   }
   return nfail;
}
