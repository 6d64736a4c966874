// EEQ-BC 2025 bond-capacity model, molecular (0-D) structures, on the device (model/eeqbc.f90); the periodic
// branch (get_*_3d) lives in eeq_pbc.cu next to the Wigner-Seitz / lattice-sum helpers it shares with EEQ2019.
//
// Everything is expressed per ORDERED pair (a, b) so that one warp owns one row a and one thread
// owns one row of B -- the same decomposition as the EEQ2019 kernels in eeq.cu.  With v = r_b - r_a,
// r = |v|, gamma_ab = 1/sqrt(rad_a^2 + rad_b^2) (CN dependent widths), C the Maxwell capacitance
// matrix (C_ab = -c_ab, C_aa = sum_b c_ab) and D_ab = dcdr(:, a, b) = dc_ab/dr * v / r:
//
//   A_ab      = erf(gamma r)/r * C_ab ,  A_aa = h_a C_aa + 1 ,  h_a = eta_a + kqeta_a qloc_a + sqrt(2/pi)/rad_a
//   xvec      = C xtmp ,  xtmp_a = -chi_a + kcnchi_a cn_a + kqchi_a qloc_a              (eeqbc.f90:254-282)
//   dadr(:,a,b) = t1 (r_a - r_b) q_a C_ab + t2 q_a dgam_ab(:,a) C_ab + t3 q_a D_ab - h_b q_b D_ab
//                 + (kqeta_b q_b C_bb) dqlocdr(:,a,b) + (-sqrt(2/pi) drad_b/rad_b^2 q_b C_bb) dcndr(:,a,b)
//        t1 = 2 gamma e/(sqrt(pi) r^2) - erf/(r^3), t2 = 2 e/sqrt(pi), t3 = erf/r, e = exp(-gamma^2 r^2),
//        dgam_ab(:,k) = -(rad_a drad_a dcndr(:,k,a) + rad_b drad_b dcndr(:,k,b)) gamma^3 for k in {a, b} only
//        (the reference reads exactly these two columns of its per-pair dgamdr array, eeqbc.f90:716-741)
//   dxdr(:,a,b) = [dtmpdr C](:,a,b) + (xtmp_a - xtmp_b) D_ab                           (eeqbc.f90:361, 438-439)
// and the matching row sums for atrace, dadL, dxdL and the diagonal blocks (eeqbc.f90:423-455, 744-776).
// The only dense contraction, dxdr = dtmpdr * C (3N x N x N, eeqbc.f90:361), runs on the DMMA GEMM.
#include "dense.cuh"
#include "eeq.cuh"
#include "eeqbc.cuh"
#include "erfgauss.cuh"

namespace mchrg {

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// per-atom quantities: [0] rad_eff, [1] drad/dcn, [2] h (effective hardness), [3] xtmp, [4] cap
__global__ void bc_atom_kernel(int nat, BcAtoms at, double kcnrad, double norm_exp, const double* __restrict__ cn,
                               const double* __restrict__ qloc, double charge, double* __restrict__ out) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i > nat) return;
   double* o = out + (size_t)5 * i;
   if (i == nat) {  // xtmp(nat+1) = charge
      o[0] = o[1] = o[2] = o[4] = 0.0;
      o[3] = charge;
      return;
   }
   const double norm_cn = 1.0 / pow(at.avg_cn[i], norm_exp);
   const double radi = at.rad[i] * (1.0 - kcnrad * cn[i] * norm_cn);
   o[0] = radi;
   o[1] = -at.rad[i] * kcnrad * norm_cn;
   o[2] = at.eta[i] + at.kqeta[i] * qloc[i] + SQRT2PI / radi;
   o[3] = -at.chi[i] + at.kcnchi[i] * cn[i] + at.kqchi[i] * qloc[i];
   o[4] = at.cap[i];
}

__device__ __forceinline__ double rvdw_pair(const double* __restrict__ rvdw, int nid, const int* __restrict__ id, int a,
                                            int b) {
   return rvdw[(id[a] - 1) + (size_t)nid * (id[b] - 1)];
}

// get_cmat_0d (eeqbc.f90:1015-1063) and the row sums of get_dcmat_0d (1189-1241):
// warp per row a.  C(a, b) = -c_ab, C(a, a) = sum_b c_ab; optional dcdr(:, a, a), dcdL(:, :, a)
__global__ void __launch_bounds__(256) bc_cmat_kernel(int nat, int64_t npad, const double* __restrict__ xyz,
                                                      const int* __restrict__ id, const double* __restrict__ atoms,
                                                      const double* __restrict__ rvdw, int nid, double kbc,
                                                      double* __restrict__ Cm, double* __restrict__ dcdr_diag,
                                                      double* __restrict__ dcdL, const double2* __restrict__ etab) {
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __syncthreads();
   const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (a >= npad) return;
   double* col = Cm + (int64_t)a * npad;  // symmetric: column a == row a, contiguous in b
   if (a >= nat) {
      for (int b = lane; b < npad; b += 32) col[b] = 0.0;
      return;
   }
   const double xa = xyz[3 * a], ya = xyz[3 * a + 1], za = xyz[3 * a + 2], capa = atoms[5 * a + 4];
   double csum = 0.0, dg[3] = {0.0, 0.0, 0.0}, dl[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   for (int b = lane; b < npad; b += 32) {
      if (b >= nat) {
         col[b] = 0.0;
         continue;
      }
      if (b == a) continue;
      const double vx = xyz[3 * b] - xa, vy = xyz[3 * b + 1] - ya, vz = xyz[3 * b + 2] - za;
      const double r2 = vx * vx + vy * vy + vz * vz;
      const double rinv = eg::fast_rsqrt(r2), r1 = r2 * rinv;
      const double rv = rvdw_pair(rvdw, nid, id, a, b);
      const double sc = sqrt(capa * atoms[5 * b + 4]);
      const double t = kbc * (r1 - rv) / rv;
      double et;
      const double c = sc * (0.5 - copysign(0.5, t) * eg::erf_gauss<6, true>(setab, fabs(t), et));  // get_cpair, eeqbc.f90:1130-1143
      col[b] = -c;
      csum += c;
      if (dcdr_diag) {
         const double d = -sc * kbc * et / (SQRTPI * rv) * rinv;  // get_dcpair: D = d * v
         const double gx = d * vx, gy = d * vy, gz = d * vz;
         dg[0] -= gx; dg[1] -= gy; dg[2] -= gz;  // dcdr(:, a, a) = -sum_b D_ab
         dl[0] += gx * vx; dl[1] += gx * vy; dl[2] += gx * vz;  // dS(al, be) = D(be) v(al)
         dl[3] += gy * vx; dl[4] += gy * vy; dl[5] += gy * vz;
         dl[6] += gz * vx; dl[7] += gz * vy; dl[8] += gz * vz;
      }
   }
   csum = wsum(csum);
   if (dcdr_diag) {
#pragma unroll
      for (int c = 0; c < 3; ++c) dg[c] = wsum(dg[c]);
#pragma unroll
      for (int c = 0; c < 9; ++c) dl[c] = wsum(dl[c]);
   }
   if (lane == 0) {
      col[a] = csum;
      if (dcdr_diag) {
#pragma unroll
         for (int c = 0; c < 3; ++c) dcdr_diag[3 * a + c] = dg[c];
#pragma unroll
         for (int c = 0; c < 9; ++c) dcdL[(size_t)9 * a + c] = dl[c];
      }
   }
}

// get_amat_0d (eeqbc.f90:475-530) into the padded work matrix (identity padding)
__global__ void __launch_bounds__(128) bc_amat_kernel(int nat, int64_t npad, const double* __restrict__ xyz,
                                                      const double* __restrict__ atoms, const double* __restrict__ Cm,
                                                      double* __restrict__ W, const double2* __restrict__ etab,
                                                      int own_panel, int nranks, int rank) {
   // multi-GPU factorisation: only the column groups this rank owns (block-cyclic, own_panel columns wide)
   if (own_panel > 0 && (int)(((int64_t)blockIdx.y * 8) / own_panel) % nranks != rank) return;
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __syncthreads();
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= npad) return;
   const int64_t c0 = (int64_t)blockIdx.y * 8;
#pragma unroll
   for (int cc = 0; cc < 8; ++cc) {
      const int64_t c = c0 + cc;
      if (c >= npad) break;
      double v;
      if (r < nat && c < nat) {
         const double cm = Cm[r + c * npad];
         if (r == c) {
            v = atoms[5 * r + 2] * cm + 1.0;
         } else {
            const double dx = xyz[3 * c] - xyz[3 * r], dy = xyz[3 * c + 1] - xyz[3 * r + 1], dz = xyz[3 * c + 2] - xyz[3 * r + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double ra = atoms[5 * r], rb = atoms[5 * c];
            const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(ra * ra + rb * rb);
            double g;
            v = eg::erf_gauss<5, false>(setab, r2 * rinv * gam, g) * rinv * cm;
         }
      } else {
         v = (r == c) ? 1.0 : 0.0;
      }
      W[r + c * npad] = v;
   }
}

// pair quantities shared by the row and the B kernels
struct BcPair {
   double vx, vy, vz, r1, t1, t2, t3, gam3;
   double Dx, Dy, Dz;  // dcdr(:, a, b)
};
// (erf, exp(-x^2) from the node table `tab` in shared memory, 1/sqrt from the hardware seed: erfgauss.cuh)
__device__ __forceinline__ BcPair bc_pair(const double2* __restrict__ tab, double xa, double ya, double za, double rada,
                                          double capa, double xb, double yb, double zb, double radb, double capb, double rv,
                                          double kbc) {
   BcPair p;
   p.vx = xb - xa; p.vy = yb - ya; p.vz = zb - za;
   const double r2 = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
   const double rinv = eg::fast_rsqrt(r2);
   p.r1 = r2 * rinv;
   const double gam = eg::fast_rsqrt(rada * rada + radb * radb);
   double e;
   const double ef = eg::erf_gauss<6, true>(tab, gam * p.r1, e);
   p.t1 = (2.0 / SQRTPI * gam * e - ef * rinv) * rinv * rinv;
   p.t2 = 2.0 * e / SQRTPI;
   p.t3 = ef * rinv;
   p.gam3 = gam * gam * gam;
   const double t = kbc * (p.r1 - rv) / rv;
   double et;
   eg::erf_gauss<6, true>(tab, fabs(t), et);  // et = exp(-t^2)
   const double d = -sqrt(capa * capb) * kbc * et / (SQRTPI * rv) * rinv;
   p.Dx = d * p.vx; p.Dy = d * p.vy; p.Dz = d * p.vz;
   return p;
}

// Row pass (warp per atom a): atrace_a, the pair part of dadL_a, and the pair sums of dxdr(:,a,a), dxdL(:,:,a)
//   rows[a] = { atrace(3), dadL(9), sxd(3) = sum_b xtmp_b D_ab, sxl(9) = sum_b xtmp_b D(be) v(al) }
__global__ void __launch_bounds__(256) bc_rows_kernel(int nat, int64_t npad, const double* __restrict__ xyz,
                                                      const int* __restrict__ id, const double* __restrict__ atoms,
                                                      const double* __restrict__ rvdw, int nid, double kbc,
                                                      const double* __restrict__ Cm, const double* __restrict__ q,
                                                      const double* __restrict__ dcndr,
                                                      const double* __restrict__ ddiag,
                                                      const double* __restrict__ dcndL, double* __restrict__ rows,
                                                      const double2* __restrict__ etab, int row0, int row1) {
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __syncthreads();
   const int a = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (a >= row1) return;
   const double xa = xyz[3 * a], ya = xyz[3 * a + 1], za = xyz[3 * a + 2];
   const double rada = atoms[5 * a], drada = atoms[5 * a + 1], capa = atoms[5 * a + 4];
   double La[9];
#pragma unroll
   for (int c = 0; c < 9; ++c) La[c] = dcndL[(size_t)9 * a + c];
   double acc[24];
#pragma unroll
   for (int c = 0; c < 24; ++c) acc[c] = 0.0;
   const double* dcn_a = dcndr + (size_t)3 * nat * a;  // dcndr(:, b, a): derivative of cn_a w.r.t. atom b
   for (int b = lane; b < nat; b += 32) {
      if (b == a) continue;
      const double radb = atoms[5 * b], dradb = atoms[5 * b + 1];
      const BcPair p = bc_pair(setab, xa, ya, za, rada, capa, xyz[3 * b], xyz[3 * b + 1], xyz[3 * b + 2], radb, atoms[5 * b + 4],
                               rvdw_pair(rvdw, nid, id, a, b), kbc);
      const double cab = Cm[b + (int64_t)a * npad], qb = q[b], xb = atoms[5 * b + 3];
      // dgam_ab(:, b) = -(rad_a drad_a dcndr(:, b, a) + rad_b drad_b dcndr(:, b, b)) gamma^3
      const double* dbb = ddiag + 3 * b;  // dcndr(:, b, b), gathered once (in place it is one 32-byte sector per lane, N^2 times)
      const double wa = rada * drada, wb = radb * dradb;
      const double g0 = -(wa * dcn_a[3 * b] + wb * dbb[0]) * p.gam3;
      const double g1 = -(wa * dcn_a[3 * b + 1] + wb * dbb[1]) * p.gam3;
      const double g2 = -(wa * dcn_a[3 * b + 2] + wb * dbb[2]) * p.gam3;
      // atrace_a (eeqbc.f90:727-728, 737-738, 746-747):  t1 (r_a - r_b) q_b C - t2 q_b dgam(:, b) C - t3 q_b D_ba
      const double e1 = -p.t1 * qb * cab, e2 = p.t2 * qb * cab, e3 = p.t3 * qb;
      acc[0] += e1 * p.vx - e2 * g0 + e3 * p.Dx;
      acc[1] += e1 * p.vy - e2 * g1 + e3 * p.Dy;
      acc[2] += e1 * p.vz - e2 * g2 + e3 * p.Dz;
      // dadL_a pair part (eeqbc.f90:731-732, 741-742, 750-753)
      const double s1 = p.t1 * qb * cab;  // dS(al, be) = t1 v(be) v(al)
      const double v[3] = {p.vx, p.vy, p.vz}, D[3] = {p.Dx, p.Dy, p.Dz};
#pragma unroll
      for (int be = 0; be < 3; ++be)
#pragma unroll
         for (int al = 0; al < 3; ++al) {
            const int k = al + 3 * be;
            const double dgl = -(wa * La[k] + wb * dcndL[(size_t)9 * b + k]) * p.gam3;
            acc[3 + k] += s1 * v[be] * v[al] + e2 * dgl - e3 * D[al] * v[be];
            acc[15 + k] += xb * D[be] * v[al];  // sum_b xtmp_b D(be) v(al)   (v = r_b - r_a)
         }
      acc[12] += xb * p.Dx; acc[13] += xb * p.Dy; acc[14] += xb * p.Dz;
   }
#pragma unroll
   for (int c = 0; c < 24; ++c) acc[c] = wsum(acc[c]);
   if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 24; ++c) rows[(size_t)24 * a + c] = acc[c];
   }
}

// per-atom totals:  dadL(:,:,a) and dxdL(:,:,a) complete (eeqbc.f90:361-362, 446, 455, 760-776)
//   gemmL_a = sum_i (kcnchi_i dcndL_i + kqchi_i dqlocdL_i) C_ia ; warp per atom a
__global__ void __launch_bounds__(256) bc_strain_kernel(int nat, int64_t npad, BcAtoms at,
                                                        const double* __restrict__ atoms,
                                                        const double* __restrict__ Cm, const double* __restrict__ q,
                                                        const double* __restrict__ rows,
                                                        const double* __restrict__ dcdL,
                                                        const double* __restrict__ dcndL,
                                                        const double* __restrict__ dqlocdL, double* __restrict__ dadL,
                                                        double* __restrict__ dxdL, int row0, int row1) {
   const int a = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (a >= row1) return;
   double g[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   for (int i = lane; i < nat; i += 32) {
      const double cia = Cm[i + (int64_t)a * npad], k1 = at.kcnchi[i] * cia, k2 = at.kqchi[i] * cia;
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] += k1 * dcndL[(size_t)9 * i + k] + k2 * dqlocdL[(size_t)9 * i + k];
   }
#pragma unroll
   for (int k = 0; k < 9; ++k) g[k] = wsum(g[k]);
   if (lane == 0) {
      const double caa = Cm[a + (int64_t)a * npad], qa = q[a];
      const double rada = atoms[5 * a], drada = atoms[5 * a + 1], ha = atoms[5 * a + 2], xa = atoms[5 * a + 3];
      const double w_h = at.kqeta[a] * qa * caa, w_r = -SQRT2PI * drada / (rada * rada) * qa * caa, w_c = ha * qa;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
         dadL[(size_t)9 * a + k] = rows[(size_t)24 * a + 3 + k] + w_h * dqlocdL[(size_t)9 * a + k] +
                                   w_r * dcndL[(size_t)9 * a + k] + w_c * dcdL[(size_t)9 * a + k];
         // dxdL: + xtmp_b D(be) (r_a - r_b)(al) = - sxl ; + xtmp_a dcdL_a
         dxdL[(size_t)9 * a + k] = g[k] - rows[(size_t)24 * a + 15 + k] + xa * dcdL[(size_t)9 * a + k];
      }
   }
}

// dtmpdr (eeqbc.f90:344-348) into the padded GEMM operand T(3a+c, i) = kcnchi_i dcndr(c,a,i) + kqchi_i dqlocdr(c,a,i)
// rows of the atom block [jb, je): r = 3 (a - jb) + c
__global__ void bc_dtmpdr_kernel(int nat, int64_t mp, int64_t npad, BcAtoms at, const double* __restrict__ dcndr,
                                 const double* __restrict__ dqlocdr, double* __restrict__ T, int jb, int je) {
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int64_t i = blockIdx.y;
   if (r >= mp || i >= npad) return;
   double v = 0.0;
   if (r < 3 * (int64_t)(je - jb) && i < nat) {
      const size_t src = (size_t)3 * jb + r + (size_t)3 * nat * i;
      v = at.kcnchi[i] * dcndr[src] + at.kqchi[i] * dqlocdr[src];
   }
   T[r + i * mp] = v;
}

// B(3a+c, b) = sa * dadr(c, a, b) - sx * dxdr(c, a, b); thread = atom a, CTA = 128 atoms x 64 columns.
// The GEMM part of dxdr is expected in Bout already, scaled by -sx (Bout = -sx dtmpdr C), when sx != 0.
//   ga[3a+c] += sum_b dadr(c,a,b) q_b   (incl. atrace on the diagonal)       type.F90:277
//   gx[3a+c] += sum_b [dxdr - gemm part](c,a,b) q_b + sum_i dtmpdr(c,a,i) (C q)_i   type.F90:278
//   bz[3a+c] += sum_b B(3a+c, b) z_b
constexpr int BCK = 64;
__global__ void __launch_bounds__(128) bc_build_b_kernel(int nat, int64_t npad, const double* __restrict__ xyz,
                                                         const int* __restrict__ id, const double* __restrict__ atoms,
                                                         BcAtoms at, const double* __restrict__ rvdw, int nid,
                                                         double kbc, const double* __restrict__ Cm,
                                                         const double* __restrict__ q, const double* __restrict__ cq,
                                                         const double* __restrict__ rows,
                                                         const double* __restrict__ dcdr_diag,
                                                         const double* __restrict__ dcndr,
                                                         const double* __restrict__ dqlocdr, double sa, double sx,
                                                         int add_atrace, double* __restrict__ Bout, int64_t ldb,
                                                         double* __restrict__ ga, double* __restrict__ gx,
                                                         const double* __restrict__ zvec, double* __restrict__ bz,
                                                         const double2* __restrict__ etab, int jb, int je) {
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __shared__ double sk[BCK][12];  // x y z rad drad h xtmp cap | q C_bb-weights(2) z
   // atom a of the row block [jb, je); rows of Bout / ga / gx / bz are numbered within the block
   const int al = blockIdx.x * blockDim.x + threadIdx.x;
   const int a = jb + al;
   const int b0 = blockIdx.y * BCK;
   const int bn = min(BCK, nat - b0);
   for (int idx = threadIdx.x; idx < bn; idx += blockDim.x) {
      const int b = b0 + idx;
      sk[idx][0] = xyz[3 * b]; sk[idx][1] = xyz[3 * b + 1]; sk[idx][2] = xyz[3 * b + 2];
      const double radb = atoms[5 * b], dradb = atoms[5 * b + 1];
      sk[idx][3] = radb; sk[idx][4] = dradb; sk[idx][5] = atoms[5 * b + 2]; sk[idx][6] = atoms[5 * b + 3];
      sk[idx][7] = atoms[5 * b + 4];
      const double qb = q[b], cbb = Cm[b + (int64_t)b * npad];
      sk[idx][8] = qb;
      sk[idx][9] = at.kqeta[b] * qb * cbb;                         // hardness derivative weight (eeqbc.f90:759)
      sk[idx][10] = -SQRT2PI * dradb / (radb * radb) * qb * cbb;    // charge width weight (eeqbc.f90:764)
      sk[idx][11] = zvec ? zvec[b] : 0.0;
   }
   __syncthreads();
   if (a >= je) return;
   const double xa = xyz[3 * a], ya = xyz[3 * a + 1], za = xyz[3 * a + 2];
   const double rada = atoms[5 * a], drada = atoms[5 * a + 1], xtmpa = atoms[5 * a + 3], capa = atoms[5 * a + 4];
   const double qa = q[a];
   const double* daa = dcndr + ((size_t)nat * a + a) * 3;  // dcndr(:, a, a)
   const double daa0 = daa[0], daa1 = daa[1], daa2 = daa[2];
   double sga[3] = {0.0, 0.0, 0.0}, sgx[3] = {0.0, 0.0, 0.0}, sbz[3] = {0.0, 0.0, 0.0};
   for (int kk = 0; kk < bn; ++kk) {
      const int b = b0 + kk;
      const double* dcn = dcndr + ((size_t)nat * b + a) * 3;    // dcndr(:, a, b)
      const double* dql = dqlocdr + ((size_t)nat * b + a) * 3;  // dqlocdr(:, a, b)
      const double c0 = dcn[0], c1 = dcn[1], c2 = dcn[2], l0 = dql[0], l1 = dql[1], l2 = dql[2];
      const double qb = sk[kk][8], wh = sk[kk][9], wr = sk[kk][10];
      double da0 = wh * l0 + wr * c0, da1 = wh * l1 + wr * c1, da2 = wh * l2 + wr * c2;  // row copies, all a
      double dx0, dx1, dx2;
      if (b == a) {
         const double* rw = rows + (size_t)24 * a;
         const double hq = sk[kk][5] * qb;
         da0 += hq * dcdr_diag[3 * a]; da1 += hq * dcdr_diag[3 * a + 1]; da2 += hq * dcdr_diag[3 * a + 2];
         if (add_atrace) { da0 += rw[0]; da1 += rw[1]; da2 += rw[2]; }
         dx0 = rw[12] + xtmpa * dcdr_diag[3 * a];
         dx1 = rw[13] + xtmpa * dcdr_diag[3 * a + 1];
         dx2 = rw[14] + xtmpa * dcdr_diag[3 * a + 2];
      } else {
         const double radb = sk[kk][3], dradb = sk[kk][4];
         const BcPair p = bc_pair(setab, xa, ya, za, rada, capa, sk[kk][0], sk[kk][1], sk[kk][2], radb, sk[kk][7],
                                  rvdw_pair(rvdw, nid, id, a, b), kbc);
         const double cab = Cm[a + (int64_t)b * npad];
         const double wa = rada * drada, wb = radb * dradb;
         // dgam_ab(:, a) = -(rad_a drad_a dcndr(:, a, a) + rad_b drad_b dcndr(:, a, b)) gamma^3
         const double g0 = -(wa * daa0 + wb * c0) * p.gam3, g1 = -(wa * daa1 + wb * c1) * p.gam3,
                      g2 = -(wa * daa2 + wb * c2) * p.gam3;
         const double e1 = -p.t1 * qa * cab;  // t1 (r_a - r_b) q_a C_ab
         const double e2 = p.t2 * qa * cab;
         const double e3 = p.t3 * qa - sk[kk][5] * qb;  // t3 q_a D_ab - h_b q_b D_ab
         da0 += e1 * p.vx + e2 * g0 + e3 * p.Dx;
         da1 += e1 * p.vy + e2 * g1 + e3 * p.Dy;
         da2 += e1 * p.vz + e2 * g2 + e3 * p.Dz;
         const double dxt = xtmpa - sk[kk][6];
         dx0 = dxt * p.Dx; dx1 = dxt * p.Dy; dx2 = dxt * p.Dz;
      }
      sga[0] = fma(da0, qb, sga[0]); sga[1] = fma(da1, qb, sga[1]); sga[2] = fma(da2, qb, sga[2]);
      // dxdr . q : pair part + gemm part through w = C q
      const double wq = cq ? cq[b] : 0.0;
      const double tk = at.kcnchi[b] * wq, tl = at.kqchi[b] * wq;
      sgx[0] += dx0 * qb + tk * c0 + tl * l0;
      sgx[1] += dx1 * qb + tk * c1 + tl * l1;
      sgx[2] += dx2 * qb + tk * c2 + tl * l2;
      if (Bout) {
         double* o = Bout + (size_t)b * ldb + 3 * al;
         double b0v = sa * da0 - sx * dx0, b1v = sa * da1 - sx * dx1, b2v = sa * da2 - sx * dx2;
         if (sx != 0.0) { b0v += o[0]; b1v += o[1]; b2v += o[2]; }
         o[0] = b0v; o[1] = b1v; o[2] = b2v;
         const double zk = sk[kk][11];
         sbz[0] = fma(b0v, zk, sbz[0]); sbz[1] = fma(b1v, zk, sbz[1]); sbz[2] = fma(b2v, zk, sbz[2]);
      }
   }
   if (ga) { atomicAdd(ga + 3 * al, sga[0]); atomicAdd(ga + 3 * al + 1, sga[1]); atomicAdd(ga + 3 * al + 2, sga[2]); }
   if (gx) { atomicAdd(gx + 3 * al, sgx[0]); atomicAdd(gx + 3 * al + 1, sgx[1]); atomicAdd(gx + 3 * al + 2, sgx[2]); }
   if (bz && Bout) { atomicAdd(bz + 3 * al, sbz[0]); atomicAdd(bz + 3 * al + 1, sbz[1]); atomicAdd(bz + 3 * al + 2, sbz[2]); }
}

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
void bc_prepare(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcAtoms& at,
                const double* cn, const double* qloc, bool grad, BcWork& w) {
   const int nat = mol.nat;
   w.npad = round_up(nat, TILE);
   w.atoms = ctx->alloc<double>((size_t)5 * (nat + 1));
   MCHRG_LAUNCH(ctx, bc_atom_kernel, (nat + 1 + 127) / 128, 128, 0, nat, at, model->kcnrad, model->norm_exp, cn, qloc,
                mol.charge, w.atoms);
   w.Cm = ctx->alloc<double>((size_t)w.npad * w.npad);
   if (mol.periodic) return;  // the periodic capacitance matrix comes from bc_pbc_matrices (eeq_pbc.cu)
   w.dcdr_diag = grad ? ctx->alloc<double>((size_t)3 * nat) : nullptr;
   w.dcdL = grad ? ctx->alloc<double>((size_t)9 * nat) : nullptr;
   MCHRG_LAUNCH(ctx, bc_cmat_kernel, (unsigned)((w.npad + 7) / 8), 256, 0, nat, w.npad, mol.xyz, mol.id, w.atoms,
                model->dev_rvdw(), model->nid, model->kbc, w.Cm, w.dcdr_diag, w.dcdL, ctx->erf_tab);
}

void bc_amat(mchrg_b200_context* ctx, const DevMol& mol, const BcWork& w, double* W, int64_t own_panel) {
   dim3 grid((unsigned)((w.npad + 127) / 128), (unsigned)((w.npad + 7) / 8));
   MCHRG_LAUNCH(ctx, bc_amat_kernel, grid, 128, 0, mol.nat, w.npad, mol.xyz, w.atoms, w.Cm, W, ctx->erf_tab, (int)own_panel,
                ctx->nranks(), ctx->rank());
}

// (A q)_a = sum_b A_ab q_b re-evaluated from the pair formula (no copy of the matrix is kept for the energy,
// type.F90:257-259); warp per row a in [row0, row1)
__global__ void __launch_bounds__(256) bc_jq_rows_kernel(int nat, int64_t npad, const double* __restrict__ xyz,
                                                         const double* __restrict__ atoms, const double* __restrict__ Cm,
                                                         const double* __restrict__ q, double* __restrict__ jq,
                                                         const double2* __restrict__ etab, int row0, int row1) {
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __syncthreads();
   const int a = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (a >= row1) return;
   const double xa = xyz[3 * a], ya = xyz[3 * a + 1], za = xyz[3 * a + 2], ra = atoms[5 * a];
   const double* col = Cm + (int64_t)a * npad;  // symmetric
   double acc = 0.0;
   for (int b = lane; b < nat; b += 32) {
      if (b == a) continue;
      const double dx = xyz[3 * b] - xa, dy = xyz[3 * b + 1] - ya, dz = xyz[3 * b + 2] - za;
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double rb = atoms[5 * b];
      const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(ra * ra + rb * rb);
      double g;
      acc = fma(eg::erf_gauss<5, false>(setab, r2 * rinv * gam, g) * rinv * col[b], q[b], acc);
   }
   acc = wsum(acc);
   if (lane == 0) jq[a] = acc + (atoms[5 * a + 2] * col[a] + 1.0) * q[a];
}
void bc_jq_rows(mchrg_b200_context* ctx, const DevMol& mol, const BcWork& w, const double* q, double* jq, int row0, int row1) {
   if (row1 < 0) { row0 = 0; row1 = mol.nat; }
   if (row1 <= row0) return;
   MCHRG_LAUNCH(ctx, bc_jq_rows_kernel, (unsigned)((row1 - row0 + 7) / 8), 256, 0, mol.nat, w.npad, mol.xyz, w.atoms, w.Cm, q, jq,
                ctx->erf_tab, row0, row1);
}

__global__ void bc_diag3_kernel(int nat, const double* __restrict__ d, double* __restrict__ out) {
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= 3 * nat) return;
   const int b = t / 3, c = t - 3 * b;
   out[t] = d[((size_t)nat * b + b) * 3 + c];
}

void bc_rows(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcWork& w,
             const double* q, const double* dcndr, const double* dcndL, double* rows, int row0, int row1) {
   if (row1 < 0) { row0 = 0; row1 = mol.nat; }
   if (row1 <= row0) return;
   double* ddiag = ctx->alloc<double>((size_t)3 * mol.nat);
   MCHRG_LAUNCH(ctx, bc_diag3_kernel, (unsigned)((3 * mol.nat + 255) / 256), 256, 0, mol.nat, dcndr, ddiag);
   MCHRG_LAUNCH(ctx, bc_rows_kernel, (unsigned)((row1 - row0 + 7) / 8), 256, 0, mol.nat, w.npad, mol.xyz, mol.id, w.atoms,
                model->dev_rvdw(), model->nid, model->kbc, w.Cm, q, dcndr, ddiag, dcndL, rows, ctx->erf_tab, row0, row1);
}

void bc_strain(mchrg_b200_context* ctx, const DevMol& mol, const BcAtoms& at, const BcWork& w, const double* q,
               const double* rows, const double* dcndL, const double* dqlocdL, double* dadL, double* dxdL, int row0,
               int row1) {
   if (row1 < 0) { row0 = 0; row1 = mol.nat; }
   if (row1 <= row0) return;
   MCHRG_LAUNCH(ctx, bc_strain_kernel, (unsigned)((row1 - row0 + 7) / 8), 256, 0, mol.nat, w.npad, at, w.atoms, w.Cm, q, rows,
                w.dcdL, dcndL, dqlocdL, dadL, dxdL, row0, row1);
}

void bc_dxdr_gemm(mchrg_b200_context* ctx, const DevMol& mol, const BcAtoms& at, const BcWork& w, const double* dcndr,
                  const double* dqlocdr, double alpha, double* Bm, int64_t mp, int jb, int je) {
   if (je < 0) { jb = 0; je = mol.nat; }
   double* T = ctx->alloc<double>((size_t)mp * w.npad);
   dim3 grid((unsigned)((mp + 255) / 256), (unsigned)w.npad);
   MCHRG_LAUNCH(ctx, bc_dtmpdr_kernel, grid, 256, 0, mol.nat, mp, w.npad, at, dcndr, dqlocdr, T, jb, je);
   // C is symmetric: op(B) = C^T = C
   dgemm_padded(ctx, true, mp, w.npad, w.npad, alpha, T, mp, w.Cm, w.npad, 0.0, Bm, mp, false);
}

void bc_build_b(mchrg_b200_context* ctx, const DevMol& mol, const mchrg_b200_model* model, const BcAtoms& at,
                const BcWork& w, const double* q, const double* cq, const double* rows, const double* dcndr,
                const double* dqlocdr, double sa, double sx, bool add_atrace, double* Bout, int64_t ldb, double* ga,
                double* gx, const double* zvec, double* bz, int jb, int je) {
   if (je < 0) { jb = 0; je = mol.nat; }
   dim3 grid((unsigned)((je - jb + 127) / 128), (unsigned)((mol.nat + BCK - 1) / BCK));
   MCHRG_LAUNCH(ctx, bc_build_b_kernel, grid, 128, 0, mol.nat, w.npad, mol.xyz, mol.id, w.atoms, at, model->dev_rvdw(),
                model->nid, model->kbc, w.Cm, q, cq, rows, w.dcdr_diag, dcndr, dqlocdr, sa, sx, add_atrace ? 1 : 0, Bout,
                ldb, ga, gx, zvec, bz, ctx->erf_tab, jb, je);
}

}  // namespace mchrg
