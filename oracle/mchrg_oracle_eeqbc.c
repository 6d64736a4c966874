/* TEST INFRASTRUCTURE -- CPU oracle, EEQ-BC part (non-periodic), see mchrg_oracle.h.
 *
 * Restates model/eeqbc.f90 for molecular (0-D) structures loop by loop:
 *   update            eeqbc.f90:178-252   (cmat, dcdr, dcdL)
 *   get_xvec          eeqbc.f90:254-312   (xtmp, xvec = cmat * xtmp)
 *   get_xvec_derivs   eeqbc.f90:314-457   (two dgemm + pair terms)
 *   get_amat_0d       eeqbc.f90:475-530
 *   get_damat_0d      eeqbc.f90:656-787   (incl. the per-pair dgamdr array of which only the
 *                                          columns iat / jat are ever read)
 *   get_cmat_0d       eeqbc.f90:1015-1063, get_cpair :1130-1143, get_dcpair :1167-1187,
 *   get_dcmat_0d      eeqbc.f90:1189-1241
 *   local_charge      model/type.F90:294-322 (erf_en coordination number of mctc-lib)
 * The solve itself is the shared orc_solve_core (model/type.F90:151-292).
 *
 * PARITY UNPINNED against reference numbers: the EEQ-BC known answers of the reference need the
 * pairwise D3 vdW radii and Pauling EN tables of mctc-lib, which are not available offline.  The
 * restatement is checked by the reference's finite-difference methodology (tests/test_oracle_eeqbc.py).
 * Periodic EEQ-BC is restated as well: get_cmat_3d eeqbc.f90:1065-1128, get_cpair_dir :1145-1165, the periodic parts
 * of get_xvec :283-311 and get_xvec_derivs :364-418, get_amat_3d / get_amat_dir_3d :532-627, get_damat_3d :789-924 with
 * get_damat_dir :926-957 and get_damat_dc_dir :959-985, get_dcmat_3d :1243-1313, get_dcpair_dir :1315-1333.
 */
#include "mchrg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737

extern void scipy_dgemv_(const char* trans, const int* m, const int* n, const double* alpha, const double* a,
                         const int* lda, const double* x, const int* incx, const double* beta, double* y,
                         const int* incy, size_t);
extern void scipy_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k,
                         const double* alpha, const double* a, const int* lda, const double* b, const int* ldb,
                         const double* beta, double* c, const int* ldc, size_t, size_t);

typedef struct {
   void (*xvec_derivs)(const void* model, const orc_mol* mol, const void* cache, double* dxdr, double* dxdL);
   void (*coulomb_derivs)(const void* model, const orc_mol* mol, const void* cache, const double* vrhs,
                          double* dadr, double* dadL, double* atrace);
} orc_solve_vtbl;
int orc_solve_core(const void* model, const orc_mol* mol, const void* cache, const orc_solve_vtbl* vt,
                   const double* amat, const double* xvec, int dcn, double* energy, double* gradient,
                   double* sigma, double* qvec, double* dqdr, double* dqdL);

/* eeqbc_cache, eeqbc.f90:39-54 */
typedef struct {
   const double *cn, *qloc, *dcndr, *dcndL, *dqlocdr, *dqlocdL;
   double *cmat, *dcdr, *dcdL, *xtmp;
   orc_wsc wsc; /* periodic only */
   int periodic;
} bc_cache;

static double rvdw_of(const orc_eeqbc_model* m, const orc_mol* mol, int iat, int jat) {
   return m->rvdw[(mol->id[iat] - 1) + (size_t)m->nid * (mol->id[jat] - 1)];
}

/* get_cpair, eeqbc.f90:1130-1143 */
static double cpair(double kbc, double r1, double rvdw, double capi, double capj) {
   const double arg = -kbc * (r1 - rvdw) / rvdw;
   return sqrt(capi * capj) * 0.5 * (1.0 + erf(arg));
}

/* get_dcpair, eeqbc.f90:1167-1187: dgpair(3), dspair(a,b) = dgpair(b) * vec(a) */
static void dcpair(double kbc, const double* vec, double rvdw, double capi, double capj, double* dg, double* ds) {
   const double r1 = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
   const double t = kbc * (r1 - rvdw) / rvdw;
   const double arg = -(t * t);
   const double dtmp = -sqrt(capi * capj) * kbc * exp(arg) / (SQRTPI * rvdw);
   for (int c = 0; c < 3; ++c) dg[c] = dtmp * vec[c] / r1;
   for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a) ds[a + 3 * b] = dg[b] * vec[a];
}

#define BC_EPS 1.4901161193847656e-08 /* sqrt(epsilon(0.0_wp)), eeqbc.f90:90 */

/* get_cpair_dir, eeqbc.f90:1145-1165 */
static double cpair_dir(double kbc, const double* rij, const double* dtrans, double rvdw, double capi, double capj) {
   double sum = 0.0;
   for (int itr = 0; itr < 125; ++itr) {
      const double v0 = rij[0] + dtrans[3 * itr], v1 = rij[1] + dtrans[3 * itr + 1], v2 = rij[2] + dtrans[3 * itr + 2];
      const double r1 = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      if (r1 < BC_EPS) continue;
      sum += cpair(kbc, r1, rvdw, capi, capj);
   }
   return sum;
}

/* get_cmat_3d, eeqbc.f90:1065-1128 */
static void cmat_3d(const orc_eeqbc_model* m, const orc_mol* mol, const orc_wsc* wsc, double* cmat) {
   const int nat = mol->nat, nd = nat + 1, ntw = wsc->ntr;
   double dtrans[3 * 125];
   orc_get_dir_trans(mol->lattice, dtrans);
   memset(cmat, 0, sizeof(double) * (size_t)nd * nd);
   for (int iat = 0; iat < nat; ++iat) {
      const double capi = m->cap[mol->id[iat] - 1];
      for (int jat = 0; jat < iat; ++jat) {
         const double capj = m->cap[mol->id[jat] - 1];
         const double rvdw = rvdw_of(m, mol, iat, jat);
         const int nimg = wsc->nimg[jat + (size_t)nat * iat];
         const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double* t = wsc->trans + 3 * (tri[img] - 1);
            double vec[3];
            for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * iat + c] - mol->xyz[3 * jat + c] - t[c];
            const double tmp = cpair_dir(m->kbc, vec, dtrans, rvdw, capi, capj);
            cmat[jat + (size_t)nd * iat] -= tmp * wsw;
            cmat[iat + (size_t)nd * jat] -= tmp * wsw;
            cmat[iat + (size_t)nd * iat] += tmp * wsw;
            cmat[jat + (size_t)nd * jat] += tmp * wsw;
         }
      }
      /* diagonal capacitance (interaction with images) */
      const double rvdw = rvdw_of(m, mol, iat, iat);
      const int nimg = wsc->nimg[iat + (size_t)nat * iat];
      const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
      const double wsw = 1.0 / (double)nimg;
      for (int img = 0; img < nimg; ++img) {
         const double tmp = cpair_dir(m->kbc, wsc->trans + 3 * (tri[img] - 1), dtrans, rvdw, capi, capi);
         cmat[iat + (size_t)nd * iat] += tmp * wsw;
      }
   }
   cmat[nat + (size_t)nd * nat] = 1.0;
}

/* get_dcpair_dir, eeqbc.f90:1315-1333 */
static void dcpair_dir(double kbc, const double* rij, const double* dtrans, double rvdw, double capi, double capj,
                       double* dg, double* ds) {
   memset(dg, 0, 3 * sizeof(double));
   memset(ds, 0, 9 * sizeof(double));
   for (int itr = 0; itr < 125; ++itr) {
      double vec[3], gt[3], st[9];
      for (int c = 0; c < 3; ++c) vec[c] = rij[c] + dtrans[3 * itr + c];
      const double r1 = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
      if (r1 < BC_EPS) continue;
      dcpair(kbc, vec, rvdw, capi, capj, gt, st);
      for (int c = 0; c < 3; ++c) dg[c] += gt[c];
      for (int k = 0; k < 9; ++k) ds[k] += st[k];
   }
}

/* get_dcmat_3d, eeqbc.f90:1243-1313 */
static void dcmat_3d(const orc_eeqbc_model* m, const orc_mol* mol, const orc_wsc* wsc, double* dcdr, double* dcdL) {
   const int nat = mol->nat, nd = nat + 1, ntw = wsc->ntr;
   double dtrans[3 * 125];
   orc_get_dir_trans(mol->lattice, dtrans);
   memset(dcdr, 0, sizeof(double) * 3 * (size_t)nat * nd);
   memset(dcdL, 0, sizeof(double) * 9 * (size_t)nd);
   for (int iat = 0; iat < nat; ++iat) {
      const double capi = m->cap[mol->id[iat] - 1];
      for (int jat = 0; jat < iat; ++jat) {
         const double capj = m->cap[mol->id[jat] - 1];
         const double rvdw = rvdw_of(m, mol, iat, jat);
         const int nimg = wsc->nimg[jat + (size_t)nat * iat];
         const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double* t = wsc->trans + 3 * (tri[img] - 1);
            double vec[3], dG[3], dS[9];
            for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c] + t[c];
            dcpair_dir(m->kbc, vec, dtrans, rvdw, capi, capj, dG, dS);
            for (int c = 0; c < 3; ++c) {
               dcdr[c + 3 * ((size_t)iat + (size_t)nat * jat)] += +dG[c] * wsw;
               dcdr[c + 3 * ((size_t)jat + (size_t)nat * iat)] += -dG[c] * wsw;
               dcdr[c + 3 * ((size_t)iat + (size_t)nat * iat)] += -dG[c] * wsw;
               dcdr[c + 3 * ((size_t)jat + (size_t)nat * jat)] += +dG[c] * wsw;
            }
            for (int k = 0; k < 9; ++k) {
               dcdL[k + 9 * (size_t)jat] += dS[k] * wsw;
               dcdL[k + 9 * (size_t)iat] += dS[k] * wsw;
            }
         }
      }
      const double rvdw = rvdw_of(m, mol, iat, iat);
      const int nimg = wsc->nimg[iat + (size_t)nat * iat];
      const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
      const double wsw = 1.0 / (double)nimg;
      for (int img = 0; img < nimg; ++img) {
         double dG[3], dS[9];
         dcpair_dir(m->kbc, wsc->trans + 3 * (tri[img] - 1), dtrans, rvdw, capi, capi, dG, dS);
         for (int k = 0; k < 9; ++k) dcdL[k + 9 * (size_t)iat] += dS[k] * wsw; /* positive diagonal elements */
      }
   }
}

/* get_cmat_0d, eeqbc.f90:1015-1063 */
static void cmat_0d(const orc_eeqbc_model* m, const orc_mol* mol, double* cmat) {
   const int nat = mol->nat, nd = nat + 1;
   memset(cmat, 0, sizeof(double) * (size_t)nd * nd);
   for (int iat = 0; iat < nat; ++iat) {
      const double capi = m->cap[mol->id[iat] - 1];
      for (int jat = 0; jat < iat; ++jat) {
         const double capj = m->cap[mol->id[jat] - 1];
         const double v0 = mol->xyz[3 * jat] - mol->xyz[3 * iat], v1 = mol->xyz[3 * jat + 1] - mol->xyz[3 * iat + 1],
                      v2 = mol->xyz[3 * jat + 2] - mol->xyz[3 * iat + 2];
         const double r1 = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
         const double tmp = cpair(m->kbc, r1, rvdw_of(m, mol, iat, jat), capi, capj);
         cmat[jat + (size_t)nd * iat] = -tmp;
         cmat[iat + (size_t)nd * jat] = -tmp;
         cmat[iat + (size_t)nd * iat] += tmp;
         cmat[jat + (size_t)nd * jat] += tmp;
      }
   }
   cmat[nat + (size_t)nd * nat] = 1.0;
}

/* get_dcmat_0d, eeqbc.f90:1189-1241 */
static void dcmat_0d(const orc_eeqbc_model* m, const orc_mol* mol, double* dcdr, double* dcdL) {
   const int nat = mol->nat, nd = nat + 1;
   memset(dcdr, 0, sizeof(double) * 3 * (size_t)nat * nd);
   memset(dcdL, 0, sizeof(double) * 9 * (size_t)nd);
   for (int iat = 0; iat < nat; ++iat) {
      const double capi = m->cap[mol->id[iat] - 1];
      for (int jat = 0; jat < iat; ++jat) {
         const double capj = m->cap[mol->id[jat] - 1];
         double vec[3], dG[3], dS[9];
         for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c];
         dcpair(m->kbc, vec, rvdw_of(m, mol, iat, jat), capi, capj, dG, dS);
         for (int c = 0; c < 3; ++c) {
            dcdr[c + 3 * ((size_t)iat + (size_t)nat * jat)] = +dG[c];
            dcdr[c + 3 * ((size_t)jat + (size_t)nat * iat)] = -dG[c];
            dcdr[c + 3 * ((size_t)iat + (size_t)nat * iat)] += -dG[c];
            dcdr[c + 3 * ((size_t)jat + (size_t)nat * jat)] += +dG[c];
         }
         for (int k = 0; k < 9; ++k) {
            dcdL[k + 9 * (size_t)iat] += dS[k];
            dcdL[k + 9 * (size_t)jat] += dS[k];
         }
      }
   }
}

/* update, eeqbc.f90:178-252 (non-periodic branch) */
static void bc_update(const orc_eeqbc_model* m, const orc_mol* mol, bc_cache* c, const double* cn, const double* qloc,
                      const double* dcndr, const double* dcndL, const double* dqlocdr, const double* dqlocdL) {
   const int nat = mol->nat, nd = nat + 1;
   const int grad = dcndr && dcndL && dqlocdr && dqlocdL;
   memset(c, 0, sizeof *c);
   c->cn = cn;
   c->qloc = qloc;
   if (grad) { c->dcndr = dcndr; c->dcndL = dcndL; c->dqlocdr = dqlocdr; c->dqlocdL = dqlocdL; }
   c->xtmp = (double*)calloc(nd, sizeof(double));
   c->cmat = (double*)malloc(sizeof(double) * (size_t)nd * nd);
   c->periodic = mol->periodic[0] || mol->periodic[1] || mol->periodic[2];
   if (c->periodic) { /* eeqbc.f90:222-236 */
      orc_wsc_new(mol, &c->wsc);
      cmat_3d(m, mol, &c->wsc, c->cmat);
      if (grad) {
         c->dcdr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nd);
         c->dcdL = (double*)malloc(sizeof(double) * 9 * (size_t)nd);
         dcmat_3d(m, mol, &c->wsc, c->dcdr, c->dcdL);
      }
      return;
   }
   cmat_0d(m, mol, c->cmat);
   if (grad) {
      c->dcdr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nd);
      c->dcdL = (double*)malloc(sizeof(double) * 9 * (size_t)nd);
      dcmat_0d(m, mol, c->dcdr, c->dcdL);
   }
}
static void bc_free(bc_cache* c) {
   free(c->xtmp); free(c->cmat); free(c->dcdr); free(c->dcdL);
   if (c->periodic) orc_wsc_free(&c->wsc);
}

/* get_xvec, eeqbc.f90:254-312 */
static void bc_get_xvec(const orc_eeqbc_model* m, const orc_mol* mol, bc_cache* c, double* xvec) {
   const int nat = mol->nat, nd = nat + 1, one = 1;
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      c->xtmp[iat] = -m->chi[izp] + m->kcnchi[izp] * c->cn[iat] + m->kqchi[izp] * c->qloc[iat];
   }
   c->xtmp[nat] = mol->charge;
   const double a1 = 1.0, b0 = 0.0;
   memset(xvec, 0, sizeof(double) * nd);
   scipy_dgemv_("n", &nd, &nd, &a1, c->cmat, &nd, c->xtmp, &one, &b0, xvec, &one, 1);
   if (c->periodic) { /* eliminate self-interaction (quasi off-diagonal), eeqbc.f90:283-311 */
      double dtrans[3 * 125];
      orc_get_dir_trans(mol->lattice, dtrans);
      const int ntw = c->wsc.ntr;
      for (int iat = 0; iat < nat; ++iat) {
         const double capi = m->cap[mol->id[iat] - 1];
         const double rvdw = rvdw_of(m, mol, iat, iat);
         const int nimg = c->wsc.nimg[iat + (size_t)nat * iat];
         const int* tri = c->wsc.tridx + (size_t)ntw * (iat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double ctmp = cpair_dir(m->kbc, c->wsc.trans + 3 * (tri[img] - 1), dtrans, rvdw, capi, capi);
            xvec[iat] -= wsw * ctmp * c->xtmp[iat];
         }
      }
   }
}

/* get_xvec_derivs, eeqbc.f90:314-457 (non-periodic branch) */
static void bc_get_xvec_derivs(const orc_eeqbc_model* m, const orc_mol* mol, const bc_cache* c, double* dxdr, double* dxdL) {
   const int nat = mol->nat, nd = nat + 1;
   const size_t n3 = 3 * (size_t)nat;
   double* dtmpdr = (double*)calloc(n3 * nd, sizeof(double));
   double* dtmpdL = (double*)calloc(9 * (size_t)nd, sizeof(double));
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      for (size_t k = 0; k < n3; ++k)
         dtmpdr[k + n3 * iat] = m->kcnchi[izp] * c->dcndr[k + n3 * iat] + m->kqchi[izp] * c->dqlocdr[k + n3 * iat];
      for (int k = 0; k < 9; ++k)
         dtmpdL[k + 9 * (size_t)iat] = m->kcnchi[izp] * c->dcndL[k + 9 * (size_t)iat] + m->kqchi[izp] * c->dqlocdL[k + 9 * (size_t)iat];
   }
   const int m3 = (int)n3, m9 = 9;
   const double a1 = 1.0, b0 = 0.0;
   scipy_dgemm_("n", "n", &m3, &nd, &nd, &a1, dtmpdr, &m3, c->cmat, &nd, &b0, dxdr, &m3, 1, 1); /* :361 */
   scipy_dgemm_("n", "n", &m9, &nd, &nd, &a1, dtmpdL, &m9, c->cmat, &nd, &b0, dxdL, &m9, 1, 1); /* :362 */
   free(dtmpdr);
   free(dtmpdL);
#define DCDR(cc, a, b) c->dcdr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
#define DXDR(cc, a, b) dxdr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
   if (c->periodic) { /* eeqbc.f90:364-418, including its index conventions (nimg(iat, jat) with tridx(img, jat, iat)) */
      double dtrans[3 * 125];
      orc_get_dir_trans(mol->lattice, dtrans);
      const orc_wsc* wsc = &c->wsc;
      const int ntw = wsc->ntr;
      for (int iat = 0; iat < nat; ++iat) {
         const int izp = mol->id[iat] - 1;
         const double capi = m->cap[izp];
         for (int jat = 0; jat < nat; ++jat) {
            const double rvdw = rvdw_of(m, mol, iat, jat);
            const double capj = m->cap[mol->id[jat] - 1];
            for (int cc = 0; cc < 3; ++cc) {
               DXDR(cc, iat, iat) += c->xtmp[jat] * DCDR(cc, iat, jat);
               DXDR(cc, iat, jat) += (c->xtmp[iat] - c->xtmp[jat]) * DCDR(cc, iat, jat);
            }
            const int nimg = wsc->nimg[iat + (size_t)nat * jat];
            const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
            const double wsw = 1.0 / (double)nimg;
            for (int img = 0; img < nimg; ++img) {
               const double* t = wsc->trans + 3 * (tri[img] - 1);
               double vec[3], dG[3], dS[9];
               for (int cc = 0; cc < 3; ++cc) vec[cc] = mol->xyz[3 * jat + cc] - mol->xyz[3 * iat + cc] + t[cc];
               dcpair_dir(m->kbc, vec, dtrans, rvdw, capi, capj, dG, dS);
               for (int k = 0; k < 9; ++k) dxdL[k + 9 * (size_t)iat] -= wsw * dS[k] * c->xtmp[jat];
            }
         }
         for (int k = 0; k < 9; ++k) dxdL[k + 9 * (size_t)iat] += c->xtmp[iat] * c->dcdL[k + 9 * (size_t)iat];
         /* capacitance terms for i = j, T != 0 */
         const double rvdw = rvdw_of(m, mol, iat, iat);
         const int nimg = wsc->nimg[iat + (size_t)nat * iat];
         const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double ctmp = cpair_dir(m->kbc, wsc->trans + 3 * (tri[img] - 1), dtrans, rvdw, capi, capi) * wsw;
            for (size_t k = 0; k < n3; ++k)
               dxdr[k + n3 * iat] -= ctmp * (m->kcnchi[izp] * c->dcndr[k + n3 * iat] + m->kqchi[izp] * c->dqlocdr[k + n3 * iat]);
            for (int k = 0; k < 9; ++k)
               dxdL[k + 9 * (size_t)iat] -= ctmp * (m->kcnchi[izp] * c->dcndL[k + 9 * (size_t)iat] + m->kqchi[izp] * c->dqlocdL[k + 9 * (size_t)iat]);
         }
      }
      return;
   }
   for (int iat = 0; iat < nat; ++iat) {
      for (int jat = 0; jat < iat; ++jat) {
         double vec[3];
         for (int cc = 0; cc < 3; ++cc) {
            DXDR(cc, iat, iat) += c->xtmp[jat] * DCDR(cc, iat, jat);
            DXDR(cc, jat, jat) += c->xtmp[iat] * DCDR(cc, jat, iat);
            DXDR(cc, iat, jat) += (c->xtmp[iat] - c->xtmp[jat]) * DCDR(cc, iat, jat);
            DXDR(cc, jat, iat) += (c->xtmp[jat] - c->xtmp[iat]) * DCDR(cc, jat, iat);
            vec[cc] = mol->xyz[3 * iat + cc] - mol->xyz[3 * jat + cc];
         }
         for (int b = 0; b < 3; ++b)
            for (int a = 0; a < 3; ++a) {
               dxdL[a + 3 * b + 9 * (size_t)iat] += c->xtmp[jat] * DCDR(b, iat, jat) * vec[a];
               dxdL[a + 3 * b + 9 * (size_t)jat] += c->xtmp[iat] * DCDR(b, jat, iat) * (-vec[a]);
            }
      }
      for (int cc = 0; cc < 3; ++cc) DXDR(cc, iat, iat) += c->xtmp[iat] * DCDR(cc, iat, iat);
      for (int k = 0; k < 9; ++k) dxdL[k + 9 * (size_t)iat] += c->xtmp[iat] * c->dcdL[k + 9 * (size_t)iat];
   }
#undef DXDR
}

/* charge width of atom i: the two (numerically different) operation orders of eeqbc.f90:504-506 / 512-513 */
static double rad_i(const orc_eeqbc_model* m, int izp, double cn) {
   const double norm_cn = 1.0 / pow(m->avg_cn[izp], m->norm_exp);
   return m->rad[izp] * (1.0 - m->kcnrad * cn * norm_cn);
}
static double rad_j(const orc_eeqbc_model* m, int jzp, double cn) {
   const double norm_cn = cn / pow(m->avg_cn[jzp], m->norm_exp);
   return m->rad[jzp] * (1.0 - m->kcnrad * norm_cn);
}

/* get_amat_dir_3d, eeqbc.f90:600-627 */
static double amat_dir_3d(const double* rij, double gam, const double* dtrans, double kbc, double rvdw, double capi,
                          double capj) {
   double amat = 0.0;
   for (int itr = 0; itr < 125; ++itr) {
      const double v0 = rij[0] + dtrans[3 * itr], v1 = rij[1] + dtrans[3 * itr + 1], v2 = rij[2] + dtrans[3 * itr + 2];
      const double r1 = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      if (r1 < BC_EPS) continue;
      amat += -cpair(kbc, r1, rvdw, capi, capj) * erf(gam * r1) / r1;
   }
   return amat;
}

/* get_amat_3d, eeqbc.f90:532-598 */
static void bc_amat_3d(const orc_eeqbc_model* m, const orc_mol* mol, const orc_wsc* wsc, const double* cn,
                       const double* qloc, const double* cmat, double* amat) {
   const int nat = mol->nat, nd = nat + 1, ntw = wsc->ntr;
   double dtrans[3 * 125];
   orc_get_dir_trans(mol->lattice, dtrans);
   memset(amat, 0, sizeof(double) * (size_t)nd * nd);
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      const double radi = rad_i(m, izp, cn[iat]);
      const double capi = m->cap[izp];
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         const double rvdw = rvdw_of(m, mol, iat, jat);
         const double radj = rad_j(m, jzp, cn[jat]);
         const double capj = m->cap[jzp];
         const double gam = 1.0 / sqrt(radi * radi + radj * radj);
         const int nimg = wsc->nimg[jat + (size_t)nat * iat];
         const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double* t = wsc->trans + 3 * (tri[img] - 1);
            double vec[3];
            for (int c = 0; c < 3; ++c) vec[c] = mol->xyz[3 * jat + c] - mol->xyz[3 * iat + c] + t[c];
            const double dtmp = amat_dir_3d(vec, gam, dtrans, m->kbc, rvdw, capi, capj);
            amat[jat + (size_t)nd * iat] += dtmp * wsw;
            amat[iat + (size_t)nd * jat] += dtmp * wsw;
         }
      }
      /* diagonal Coulomb interaction terms */
      const double gam = 1.0 / sqrt(2.0 * radi * radi);
      const double rvdw = rvdw_of(m, mol, iat, iat);
      const int nimg = wsc->nimg[iat + (size_t)nat * iat];
      const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
      const double wsw = 1.0 / (double)nimg;
      for (int img = 0; img < nimg; ++img)
         amat[iat + (size_t)nd * iat] += amat_dir_3d(wsc->trans + 3 * (tri[img] - 1), gam, dtrans, m->kbc, rvdw, capi, capi) * wsw;
      /* effective hardness */
      const double dtmp = m->eta[izp] + m->kqeta[izp] * qloc[iat] + SQRT2PI / radi;
      amat[iat + (size_t)nd * iat] += cmat[iat + (size_t)nd * iat] * dtmp + 1.0;
   }
   for (int k = 0; k <= nat; ++k) {
      amat[nat + (size_t)nd * k] = 1.0;
      amat[k + (size_t)nd * nat] = 1.0;
   }
   amat[nat + (size_t)nd * nat] = 0.0;
}

/* get_amat_0d, eeqbc.f90:475-530 */
static void bc_amat_0d(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                       const double* cmat, double* amat) {
   const int nat = mol->nat, nd = nat + 1;
   memset(amat, 0, sizeof(double) * (size_t)nd * nd);
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      const double radi = rad_i(m, izp, cn[iat]);
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         const double v0 = mol->xyz[3 * jat] - mol->xyz[3 * iat], v1 = mol->xyz[3 * jat + 1] - mol->xyz[3 * iat + 1],
                      v2 = mol->xyz[3 * jat + 2] - mol->xyz[3 * iat + 2];
         const double r2 = v0 * v0 + v1 * v1 + v2 * v2;
         const double radj = rad_j(m, jzp, cn[jat]);
         const double gam2 = 1.0 / (radi * radi + radj * radj);
         const double tmp = erf(sqrt(r2 * gam2)) / sqrt(r2) * cmat[jat + (size_t)nd * iat];
         amat[jat + (size_t)nd * iat] = tmp;
         amat[iat + (size_t)nd * jat] = tmp;
      }
      const double tmp = m->eta[izp] + m->kqeta[izp] * qloc[iat] + SQRT2PI / radi;
      amat[iat + (size_t)nd * iat] += tmp * cmat[iat + (size_t)nd * iat] + 1.0;
   }
   for (int k = 0; k < nd; ++k) {
      amat[nat + (size_t)nd * k] = 1.0;
      amat[k + (size_t)nd * nat] = 1.0;
   }
   amat[nat + (size_t)nd * nat] = 0.0;
}

/* get_damat_0d, eeqbc.f90:656-787 */
static void bc_damat_0d(const orc_eeqbc_model* m, const orc_mol* mol, const bc_cache* c, const double* qvec,
                        double* dadr, double* dadL, double* atrace) {
   const int nat = mol->nat, nd = nat + 1;
   const size_t n3 = 3 * (size_t)nat;
   const double* cmat = c->cmat;
   memset(atrace, 0, sizeof(double) * n3);
   memset(dadr, 0, sizeof(double) * n3 * nd);
   memset(dadL, 0, sizeof(double) * 9 * (size_t)nd);
#define CM(a, b) cmat[(a) + (size_t)nd * (b)]
#define DADR(cc, a, b) dadr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
#define DCN(cc, a, b) c->dcndr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      double norm_cn = 1.0 / pow(m->avg_cn[izp], m->norm_exp);
      const double radi = m->rad[izp] * (1.0 - m->kcnrad * c->cn[iat] * norm_cn);
      const double dradi = -m->rad[izp] * m->kcnrad * norm_cn;
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         double vec[3], dG[3], dS[9], dgam_i[3], dgam_j[3], dgamdL[9];
         for (int cc = 0; cc < 3; ++cc) vec[cc] = mol->xyz[3 * jat + cc] - mol->xyz[3 * iat + cc];
         const double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
         norm_cn = 1.0 / pow(m->avg_cn[jzp], m->norm_exp);
         const double radj = m->rad[jzp] * (1.0 - m->kcnrad * c->cn[jat] * norm_cn);
         const double dradj = -m->rad[jzp] * m->kcnrad * norm_cn;
         const double gam = 1.0 / sqrt(radi * radi + radj * radj);
         const double gam3 = pow(gam, 3.0);
         /* dgamdr(:, k) for the only two columns that are read, k = iat and k = jat (eeqbc.f90:716-741) */
         for (int cc = 0; cc < 3; ++cc) {
            dgam_i[cc] = -(radi * dradi * DCN(cc, iat, iat) + radj * dradj * DCN(cc, iat, jat)) * gam3;
            dgam_j[cc] = -(radi * dradi * DCN(cc, jat, iat) + radj * dradj * DCN(cc, jat, jat)) * gam3;
         }
         for (int k = 0; k < 9; ++k)
            dgamdL[k] = -(radi * dradi * c->dcndL[k + 9 * (size_t)iat] + radj * dradj * c->dcndL[k + 9 * (size_t)jat]) * gam3;

         /* explicit derivative */
         const double arg = gam * gam * r2;
         double dtmp = 2.0 * gam * exp(-arg) / (SQRTPI * r2) - erf(sqrt(arg)) / (r2 * sqrt(r2));
         for (int cc = 0; cc < 3; ++cc) dG[cc] = dtmp * vec[cc];
         for (int b = 0; b < 3; ++b)
            for (int a = 0; a < 3; ++a) dS[a + 3 * b] = dG[b] * vec[a];
         for (int cc = 0; cc < 3; ++cc) {
            atrace[cc + 3 * iat] += -dG[cc] * qvec[jat] * CM(jat, iat);
            atrace[cc + 3 * jat] += +dG[cc] * qvec[iat] * CM(iat, jat);
            DADR(cc, iat, jat) += -dG[cc] * qvec[iat] * CM(iat, jat);
            DADR(cc, jat, iat) += +dG[cc] * qvec[jat] * CM(jat, iat);
         }
         for (int k = 0; k < 9; ++k) {
            dadL[k + 9 * (size_t)iat] += dS[k] * qvec[jat] * CM(jat, iat);
            dadL[k + 9 * (size_t)jat] += dS[k] * qvec[iat] * CM(iat, jat);
         }
         /* effective charge width derivative */
         dtmp = 2.0 * exp(-arg) / SQRTPI;
         for (int cc = 0; cc < 3; ++cc) {
            atrace[cc + 3 * iat] += -dtmp * qvec[jat] * dgam_j[cc] * CM(jat, iat);
            atrace[cc + 3 * jat] += -dtmp * qvec[iat] * dgam_i[cc] * CM(iat, jat);
            DADR(cc, iat, jat) += +dtmp * qvec[iat] * dgam_i[cc] * CM(iat, jat);
            DADR(cc, jat, iat) += +dtmp * qvec[jat] * dgam_j[cc] * CM(jat, iat);
         }
         for (int k = 0; k < 9; ++k) {
            dadL[k + 9 * (size_t)iat] += dtmp * qvec[jat] * dgamdL[k] * CM(jat, iat);
            dadL[k + 9 * (size_t)jat] += dtmp * qvec[iat] * dgamdL[k] * CM(iat, jat);
         }
         /* capacitance derivative, off-diagonal */
         dtmp = erf(sqrt(r2) * gam) / sqrt(r2);
         for (int cc = 0; cc < 3; ++cc) {
            atrace[cc + 3 * iat] += -dtmp * qvec[jat] * DCDR(cc, jat, iat);
            atrace[cc + 3 * jat] += -dtmp * qvec[iat] * DCDR(cc, iat, jat);
            DADR(cc, iat, jat) += +dtmp * qvec[iat] * DCDR(cc, iat, jat);
            DADR(cc, jat, iat) += +dtmp * qvec[jat] * DCDR(cc, jat, iat);
         }
         for (int b = 0; b < 3; ++b)
            for (int a = 0; a < 3; ++a) { /* spread(dcdr(:,iat,jat),2,3)*spread(vec,1,3): (a,b) -> dcdr(a) vec(b) */
               const double t = DCDR(a, iat, jat) * vec[b];
               dadL[a + 3 * b + 9 * (size_t)iat] += -dtmp * qvec[jat] * t;
               dadL[a + 3 * b + 9 * (size_t)jat] += -dtmp * qvec[iat] * t;
            }
         /* capacitance derivative, diagonal */
         dtmp = (m->eta[izp] + m->kqeta[izp] * c->qloc[iat] + SQRT2PI / radi) * qvec[iat];
         for (int cc = 0; cc < 3; ++cc) DADR(cc, jat, iat) += -dtmp * DCDR(cc, jat, iat);
         dtmp = (m->eta[jzp] + m->kqeta[jzp] * c->qloc[jat] + SQRT2PI / radj) * qvec[jat];
         for (int cc = 0; cc < 3; ++cc) DADR(cc, iat, jat) += -dtmp * DCDR(cc, iat, jat);
      }
      /* hardness derivative */
      double dtmp = m->kqeta[izp] * qvec[iat] * CM(iat, iat);
      for (size_t k = 0; k < n3; ++k) dadr[k + n3 * iat] += dtmp * c->dqlocdr[k + n3 * iat];
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dqlocdL[k + 9 * (size_t)iat];
      /* effective charge width derivative */
      dtmp = -SQRT2PI * dradi / (radi * radi) * qvec[iat] * CM(iat, iat);
      for (size_t k = 0; k < n3; ++k) dadr[k + n3 * iat] += dtmp * c->dcndr[k + n3 * iat];
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dcndL[k + 9 * (size_t)iat];
      /* capacitance derivative */
      dtmp = (m->eta[izp] + m->kqeta[izp] * c->qloc[iat] + SQRT2PI / radi) * qvec[iat];
      for (int cc = 0; cc < 3; ++cc) DADR(cc, iat, iat) += dtmp * DCDR(cc, iat, iat);
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dcdL[k + 9 * (size_t)iat];
   }
#undef CM
#undef DADR
#undef DCN
#undef DCDR
}

/* get_damat_dir, eeqbc.f90:926-957 */
static void damat_dir(const double* rij, const double* dtrans, double capi, double capj, double rvdw, double kbc,
                      double gam, double* dG, double* dS, double* dgam) {
   memset(dG, 0, 3 * sizeof(double));
   memset(dS, 0, 9 * sizeof(double));
   *dgam = 0.0;
   const double gam2 = gam * gam;
   for (int itr = 0; itr < 125; ++itr) {
      double vec[3];
      for (int c = 0; c < 3; ++c) vec[c] = rij[c] + dtrans[3 * itr + c];
      const double r1 = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
      if (r1 < BC_EPS) continue;
      const double r2 = r1 * r1;
      const double cm = cpair(kbc, r1, rvdw, capi, capj);
      const double gtmp = 2.0 * gam * exp(-r2 * gam2) / (SQRTPI * r2) - erf(r1 * gam) / (r2 * r1);
      for (int c = 0; c < 3; ++c) dG[c] -= cm * gtmp * vec[c];
      for (int b = 0; b < 3; ++b)
         for (int a = 0; a < 3; ++a) dS[a + 3 * b] -= cm * gtmp * vec[b] * vec[a]; /* spread(vec,1,3)*spread(vec,2,3) */
      *dgam += cm * 2.0 * exp(-gam2 * r2) / SQRTPI;
   }
}

/* get_damat_dc_dir, eeqbc.f90:959-985 */
static void damat_dc_dir(const double* rij, const double* dtrans, double capi, double capj, double rvdw, double kbc,
                         double gam, double* dG, double* dS) {
   memset(dG, 0, 3 * sizeof(double));
   memset(dS, 0, 9 * sizeof(double));
   for (int itr = 0; itr < 125; ++itr) {
      double vec[3], gt[3], st[9];
      for (int c = 0; c < 3; ++c) vec[c] = rij[c] + dtrans[3 * itr + c];
      const double r1 = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
      if (r1 < BC_EPS) continue;
      dcpair(kbc, vec, rvdw, capi, capj, gt, st);
      const double tmp = erf(gam * r1) / r1;
      for (int c = 0; c < 3; ++c) dG[c] += tmp * gt[c];
      for (int k = 0; k < 9; ++k) dS[k] += tmp * st[k];
   }
}

/* get_damat_3d, eeqbc.f90:789-924 */
static void bc_damat_3d(const orc_eeqbc_model* m, const orc_mol* mol, const bc_cache* c, const double* qvec,
                        double* dadr, double* dadL, double* atrace) {
   const int nat = mol->nat, nd = nat + 1;
   const size_t n3 = 3 * (size_t)nat;
   const double* cmat = c->cmat;
   const orc_wsc* wsc = &c->wsc;
   const int ntw = wsc->ntr;
   double dtrans[3 * 125];
   orc_get_dir_trans(mol->lattice, dtrans);
   memset(atrace, 0, sizeof(double) * n3);
   memset(dadr, 0, sizeof(double) * n3 * nd);
   memset(dadL, 0, sizeof(double) * 9 * (size_t)nd);
#define CM(a, b) cmat[(a) + (size_t)nd * (b)]
#define DADR(cc, a, b) dadr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
#define DCN(cc, a, b) c->dcndr[(cc) + 3 * ((size_t)(a) + (size_t)nat * (b))]
   for (int iat = 0; iat < nat; ++iat) {
      const int izp = mol->id[iat] - 1;
      double norm_cn = 1.0 / pow(m->avg_cn[izp], m->norm_exp);
      const double radi = m->rad[izp] * (1.0 - m->kcnrad * c->cn[iat] * norm_cn);
      const double dradi = -m->rad[izp] * m->kcnrad * norm_cn;
      const double capi = m->cap[izp];
      for (int jat = 0; jat < iat; ++jat) {
         const int jzp = mol->id[jat] - 1;
         const double capj = m->cap[jzp];
         const double rvdw = rvdw_of(m, mol, iat, jat);
         norm_cn = 1.0 / pow(m->avg_cn[jzp], m->norm_exp);
         const double radj = m->rad[jzp] * (1.0 - m->kcnrad * c->cn[jat] * norm_cn);
         const double dradj = -m->rad[jzp] * m->kcnrad * norm_cn;
         const double gam = 1.0 / sqrt(radi * radi + radj * radj);
         const double gam3 = pow(gam, 3.0);
         double dgam_i[3], dgam_j[3], dgamdL[9];
         for (int cc = 0; cc < 3; ++cc) { /* the two columns of dgamdr that are read */
            dgam_i[cc] = -(radi * dradi * DCN(cc, iat, iat) + radj * dradj * DCN(cc, iat, jat)) * gam3;
            dgam_j[cc] = -(radi * dradi * DCN(cc, jat, iat) + radj * dradj * DCN(cc, jat, jat)) * gam3;
         }
         for (int k = 0; k < 9; ++k)
            dgamdL[k] = -(radi * dradi * c->dcndL[k + 9 * (size_t)iat] + radj * dradj * c->dcndL[k + 9 * (size_t)jat]) * gam3;
         const int nimg = wsc->nimg[jat + (size_t)nat * iat];
         const int* tri = wsc->tridx + (size_t)ntw * (jat + (size_t)nat * iat);
         const double wsw = 1.0 / (double)nimg;
         for (int img = 0; img < nimg; ++img) {
            const double* t = wsc->trans + 3 * (tri[img] - 1);
            double vec[3], dG[3], dS[9], dgam;
            for (int cc = 0; cc < 3; ++cc) vec[cc] = mol->xyz[3 * jat + cc] - mol->xyz[3 * iat + cc] + t[cc];
            damat_dir(vec, dtrans, capi, capj, rvdw, m->kbc, gam, dG, dS, &dgam);
            for (int cc = 0; cc < 3; ++cc) dG[cc] *= wsw;
            for (int k = 0; k < 9; ++k) dS[k] *= wsw;
            dgam *= wsw;
            /* explicit derivative */
            for (int cc = 0; cc < 3; ++cc) {
               atrace[cc + 3 * iat] += -dG[cc] * qvec[jat];
               atrace[cc + 3 * jat] += +dG[cc] * qvec[iat];
               DADR(cc, iat, jat) += -dG[cc] * qvec[iat];
               DADR(cc, jat, iat) += +dG[cc] * qvec[jat];
            }
            for (int k = 0; k < 9; ++k) {
               dadL[k + 9 * (size_t)jat] += dS[k] * qvec[iat];
               dadL[k + 9 * (size_t)iat] += dS[k] * qvec[jat];
            }
            /* effective charge width derivative */
            for (int cc = 0; cc < 3; ++cc) {
               atrace[cc + 3 * iat] += +dgam * qvec[jat] * dgam_j[cc];
               atrace[cc + 3 * jat] += +dgam * qvec[iat] * dgam_i[cc];
               DADR(cc, iat, jat) += -dgam * qvec[iat] * dgam_i[cc];
               DADR(cc, jat, iat) += -dgam * qvec[jat] * dgam_j[cc];
            }
            for (int k = 0; k < 9; ++k) {
               dadL[k + 9 * (size_t)iat] += -dgam * qvec[jat] * dgamdL[k];
               dadL[k + 9 * (size_t)jat] += -dgam * qvec[iat] * dgamdL[k];
            }
            damat_dc_dir(vec, dtrans, capi, capj, rvdw, m->kbc, gam, dG, dS);
            for (int cc = 0; cc < 3; ++cc) dG[cc] *= wsw;
            for (int k = 0; k < 9; ++k) dS[k] *= wsw;
            /* capacitance derivative, off-diagonal */
            for (int cc = 0; cc < 3; ++cc) {
               atrace[cc + 3 * iat] += +qvec[jat] * dG[cc];
               atrace[cc + 3 * jat] += -qvec[iat] * dG[cc];
               DADR(cc, jat, iat) += -qvec[jat] * dG[cc];
               DADR(cc, iat, jat) += +qvec[iat] * dG[cc];
            }
            for (int k = 0; k < 9; ++k) {
               dadL[k + 9 * (size_t)jat] += -qvec[iat] * dS[k];
               dadL[k + 9 * (size_t)iat] += -qvec[jat] * dS[k];
            }
            dcpair_dir(m->kbc, vec, dtrans, rvdw, capi, capj, dG, dS);
            for (int cc = 0; cc < 3; ++cc) dG[cc] *= wsw;
            /* capacitance derivative, diagonal */
            double dtmp = (m->eta[izp] + m->kqeta[izp] * c->qloc[iat] + SQRT2PI / radi) * qvec[iat];
            for (int cc = 0; cc < 3; ++cc) DADR(cc, jat, iat) += +dtmp * dG[cc];
            dtmp = (m->eta[jzp] + m->kqeta[jzp] * c->qloc[jat] + SQRT2PI / radj) * qvec[jat];
            for (int cc = 0; cc < 3; ++cc) DADR(cc, iat, jat) += -dtmp * dG[cc];
         }
      }
      /* diagonal explicit, charge width, and capacitance derivative terms */
      const double gam = 1.0 / sqrt(2.0 * radi * radi);
      double dtmp = -SQRT2PI * dradi / (radi * radi) * qvec[iat];
      const double rvdw = rvdw_of(m, mol, iat, iat);
      const int nimg = wsc->nimg[iat + (size_t)nat * iat];
      const int* tri = wsc->tridx + (size_t)ntw * (iat + (size_t)nat * iat);
      const double wsw = 1.0 / (double)nimg;
      for (int img = 0; img < nimg; ++img) {
         const double* vec = wsc->trans + 3 * (tri[img] - 1);
         double dG[3], dS[9], dgam;
         damat_dir(vec, dtrans, capi, capi, rvdw, m->kbc, gam, dG, dS, &dgam);
         dgam *= wsw;
         for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dS[k] * wsw * qvec[iat];
         for (int cc = 0; cc < 3; ++cc) {
            atrace[cc + 3 * iat] += dtmp * DCN(cc, iat, iat) * dgam;
            DADR(cc, iat, iat) += -dtmp * DCN(cc, iat, iat) * dgam;
         }
         for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += -dtmp * c->dcndL[k + 9 * (size_t)iat] * dgam;
         damat_dc_dir(vec, dtrans, capi, capi, rvdw, m->kbc, gam, dG, dS);
         for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += -qvec[iat] * dS[k] * wsw;
      }
      /* hardness derivative */
      dtmp = m->kqeta[izp] * qvec[iat] * CM(iat, iat);
      for (size_t k = 0; k < n3; ++k) dadr[k + n3 * iat] += dtmp * c->dqlocdr[k + n3 * iat];
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dqlocdL[k + 9 * (size_t)iat];
      /* effective charge width derivative */
      dtmp = -SQRT2PI * dradi / (radi * radi) * qvec[iat] * CM(iat, iat);
      for (size_t k = 0; k < n3; ++k) dadr[k + n3 * iat] += dtmp * c->dcndr[k + n3 * iat];
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dcndL[k + 9 * (size_t)iat];
      dtmp = (m->eta[izp] + m->kqeta[izp] * c->qloc[iat] + SQRT2PI / radi) * qvec[iat];
      for (int cc = 0; cc < 3; ++cc) DADR(cc, iat, iat) += dtmp * c->dcdr[cc + 3 * ((size_t)iat + (size_t)nat * iat)];
      for (int k = 0; k < 9; ++k) dadL[k + 9 * (size_t)iat] += dtmp * c->dcdL[k + 9 * (size_t)iat];
   }
#undef CM
#undef DADR
#undef DCN
}

/* ---- public wrappers -------------------------------------------------------------------- */

void orc_eeqbc_local_charge(const orc_eeqbc_model* m, const orc_mol* mol, int ntr, const double* trans,
                            double* qloc, double* dqlocdr, double* dqlocdL) { /* type.F90:294-322 */
   orc_ncoord_get_cn(&m->ncoord_en, mol, ntr, trans, qloc, dqlocdr, dqlocdL);
   for (int i = 0; i < mol->nat; ++i) qloc[i] += mol->charge / (double)mol->nat;
}

void orc_eeqbc_get_xvec(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc, double* xvec) {
   bc_cache c;
   bc_update(m, mol, &c, cn, qloc, NULL, NULL, NULL, NULL);
   bc_get_xvec(m, mol, &c, xvec);
   bc_free(&c);
}

void orc_eeqbc_get_coulomb_matrix(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn,
                                  const double* qloc, double* amat) {
   bc_cache c;
   bc_update(m, mol, &c, cn, qloc, NULL, NULL, NULL, NULL);
   if (c.periodic) bc_amat_3d(m, mol, &c.wsc, cn, qloc, c.cmat, amat);
   else bc_amat_0d(m, mol, cn, qloc, c.cmat, amat);
   bc_free(&c);
}

void orc_eeqbc_get_cmat(const orc_eeqbc_model* m, const orc_mol* mol, double* cmat, double* dcdr, double* dcdL) {
   if (mol->periodic[0] || mol->periodic[1] || mol->periodic[2]) {
      orc_wsc wsc;
      orc_wsc_new(mol, &wsc);
      cmat_3d(m, mol, &wsc, cmat);
      if (dcdr && dcdL) dcmat_3d(m, mol, &wsc, dcdr, dcdL);
      orc_wsc_free(&wsc);
      return;
   }
   cmat_0d(m, mol, cmat);
   if (dcdr && dcdL) dcmat_0d(m, mol, dcdr, dcdL);
}

void orc_eeqbc_get_xvec_derivs(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                               const double* dcndr, const double* dcndL, const double* dqlocdr,
                               const double* dqlocdL, double* dxdr, double* dxdL) {
   bc_cache c;
   double* xvec = (double*)malloc(sizeof(double) * (mol->nat + 1));
   bc_update(m, mol, &c, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL);
   bc_get_xvec(m, mol, &c, xvec); /* fills xtmp, as in solve() where get_xvec precedes get_xvec_derivs */
   bc_get_xvec_derivs(m, mol, &c, dxdr, dxdL);
   free(xvec);
   bc_free(&c);
}

void orc_eeqbc_get_coulomb_derivs(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                                  const double* dcndr, const double* dcndL, const double* dqlocdr,
                                  const double* dqlocdL, const double* qvec, double* dadr, double* dadL,
                                  double* atrace) {
   bc_cache c;
   bc_update(m, mol, &c, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL);
   if (c.periodic) bc_damat_3d(m, mol, &c, qvec, dadr, dadL, atrace);
   else bc_damat_0d(m, mol, &c, qvec, dadr, dadL, atrace);
   bc_free(&c);
}

static void bc_vt_xvec_derivs(const void* model, const orc_mol* mol, const void* cache, double* dxdr, double* dxdL) {
   bc_get_xvec_derivs((const orc_eeqbc_model*)model, mol, (const bc_cache*)cache, dxdr, dxdL);
}
static void bc_vt_coulomb_derivs(const void* model, const orc_mol* mol, const void* cache, const double* vrhs,
                                 double* dadr, double* dadL, double* atrace) {
   const bc_cache* c = (const bc_cache*)cache;
   if (c->periodic) bc_damat_3d((const orc_eeqbc_model*)model, mol, c, vrhs, dadr, dadL, atrace);
   else bc_damat_0d((const orc_eeqbc_model*)model, mol, c, vrhs, dadr, dadL, atrace);
}

int orc_eeqbc_solve(const orc_eeqbc_model* m, const orc_mol* mol, const double* cn, const double* qloc,
                    const double* dcndr, const double* dcndL, const double* dqlocdr, const double* dqlocdL,
                    double* energy, double* gradient, double* sigma, double* qvec, double* dqdr, double* dqdL) {
   if (!qloc) return -2; /* error stop, eeqbc.f90:202 */
   const int nd = mol->nat + 1;
   const orc_solve_vtbl vt = {bc_vt_xvec_derivs, bc_vt_coulomb_derivs};
   bc_cache c;
   bc_update(m, mol, &c, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL);
   double* amat = (double*)malloc(sizeof(double) * (size_t)nd * nd);
   double* xvec = (double*)malloc(sizeof(double) * nd);
   if (c.periodic) bc_amat_3d(m, mol, &c.wsc, cn, qloc, c.cmat, amat);
   else bc_amat_0d(m, mol, cn, qloc, c.cmat, amat);
   bc_get_xvec(m, mol, &c, xvec);
   /* solve()'s `grad`/`cpq` only test dcndr/dcndL (type.F90:199-201); the EEQ-BC derivative routines
    * additionally need the local-charge derivatives */
   const int dcn = dcndr && dcndL && dqlocdr && dqlocdL;
   int stat = orc_solve_core(m, mol, &c, &vt, amat, xvec, dcn, energy, gradient, sigma, qvec, dqdr, dqdL);
   free(amat);
   free(xvec);
   bc_free(&c);
   return stat;
}

/* CN -> local charge -> solve, app/main.f90:114-118 */
int orc_eeqbc_singlepoint(const orc_eeqbc_model* m, const orc_mol* mol, double* cn_out, double* qloc_out,
                          double* qvec, double* energy, double* gradient, double* sigma, double* dqdr, double* dqdL) {
   const int nat = mol->nat;
   const int grad = (gradient && sigma) || (dqdr && dqdL);
   double* cn = (double*)malloc(sizeof(double) * nat);
   double* qloc = (double*)malloc(sizeof(double) * nat);
   double *dcndr = NULL, *dcndL = NULL, *dqlocdr = NULL, *dqlocdL = NULL, *trans = NULL;
   if (grad) {
      dcndr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nat);
      dcndL = (double*)malloc(sizeof(double) * 9 * (size_t)nat);
      dqlocdr = (double*)malloc(sizeof(double) * 3 * (size_t)nat * nat);
      dqlocdL = (double*)malloc(sizeof(double) * 9 * (size_t)nat);
   }
   int ntr = orc_lattice_points_cutoff(mol->periodic, mol->lattice, m->ncoord.cutoff, &trans);
   orc_ncoord_get_cn(&m->ncoord, mol, ntr, trans, cn, dcndr, dcndL);
   orc_eeqbc_local_charge(m, mol, ntr, trans, qloc, dqlocdr, dqlocdL);
   if (cn_out) memcpy(cn_out, cn, sizeof(double) * nat);
   if (qloc_out) memcpy(qloc_out, qloc, sizeof(double) * nat);
   int stat = orc_eeqbc_solve(m, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, energy, gradient, sigma, qvec, dqdr, dqdL);
   free(cn); free(qloc); free(dcndr); free(dcndL); free(dqlocdr); free(dqlocdL); free(trans);
   return stat;
}
