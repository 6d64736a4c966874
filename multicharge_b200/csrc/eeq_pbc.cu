// Periodic EEQ2019 on the device: Wigner-Seitz images on the fly, Ewald real-space lattice sums
// per unique pair, reciprocal space as a G-vector factorisation (model/eeq.f90:244-352, 429-570;
// wignerseitz.f90:40-117; model/type.F90:131-149).
//
// Real space: for every unique pair (i > j) the WSC images are re-derived from the 27 candidate
// translations (no N^2 x 27 index array in HBM) and the 125-term direct sums are evaluated once;
// both matrix entries / both dadr blocks / both rows' reductions are fed from that one evaluation.
// Reciprocal space: the reference evaluates sum_G w_G cos(G.(r_j - r_i + T)) per pair and image.
// T is a lattice vector, so G.T is a multiple of 2 pi and cos(G.(r_j - r_i)) = c_i c_j + s_i s_j with
// c_i = cos(G.r_i), s_i = sin(G.r_i): the N x N reciprocal matrix is Phi W Phi^T (one DMMA GEMM with
// K = 248 -> 256), the dadr block is Psi Phi^T, and all row reductions ((Jq)_i, atrace_i, dadL_i)
// collapse to O(N * 124) sums over the structure factors  C_G = sum_b q_b c_b, S_G = sum_b q_b s_b.
#include "dense.cuh"
#include "eeq.cuh"
#include "eeq_pbc.cuh"
#include "erfgauss.cuh"
#include <cfloat>

#include "lattice.hpp"

namespace mchrg {

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737
#define EPS_SQRT 1.4901161193847656e-08 /* sqrt(epsilon(1.0_wp)): eeq.f90:58, wignerseitz.f90:33 */
#define WSC_TOL 0.01                    /* wignerseitz.f90:36 */
// beyond this argument erf() is 1 to double precision and exp(-x^2) < 5e-19: terms that cannot
// change any sum at the 1e-16 level are skipped (exact for the matrix, < 1e-18 relative for derivatives)
#define PRUNE_ARG 6.5

constexpr int NDT = 125, NRT = 124, KREC = 256;

__device__ __forceinline__ double warp_sum_p(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// get_pairs (wignerseitz.f90:75-117) for vecw = r_i - r_j: bit p of the result is set when the p-th
// image that survives the r2 < thr skip lies within tol of the minimum.  The reference later uses
// that COMPACTED position p as an index into the full translation list (eeq.f90:276,287); callers
// reproduce this by reading wtrans[p].
__device__ __forceinline__ unsigned wsc_select(const double* __restrict__ wt, int nw, double vx, double vy,
                                               double vz) {
   double dmin = 1.0e300;
   for (int itr = 0; itr < nw; ++itr) {
      const double x = vx - wt[3 * itr], y = vy - wt[3 * itr + 1], z = vz - wt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      if (r2 < EPS_SQRT) continue;
      dmin = fmin(dmin, r2);
   }
   unsigned mask = 0;
   int p = 0;
   for (int itr = 0; itr < nw; ++itr) {
      const double x = vx - wt[3 * itr], y = vy - wt[3 * itr + 1], z = vz - wt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      if (r2 < EPS_SQRT) continue;
      if (fabs(r2 - dmin) <= WSC_TOL) mask |= 1u << p;
      ++p;
   }
   return mask;
}

// get_amat_dir_3d (eeq.f90:309-329)
__device__ __forceinline__ double dir_amat(const double* __restrict__ dt, double vx, double vy, double vz,
                                           double gam, double alp) {
   double acc = 0.0;
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r1 = sqrt(x * x + y * y + z * z);
      if (r1 < EPS_SQRT) continue;
      if (gam * r1 > PRUNE_ARG && alp * r1 > PRUNE_ARG) continue;
      acc += erf(gam * r1) / r1 - erf(alp * r1) / r1;
   }
   return acc;
}

// get_damat_dir_3d (eeq.f90:510-538): dg(3), ds as xx, xy, xz, yy, yz, zz
__device__ __forceinline__ void dir_damat(const double* __restrict__ dt, double vx, double vy, double vz,
                                          double gam, double alp, double* g, double* s) {
   const double gam2 = gam * gam, alp2 = alp * alp;
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      const double r1 = sqrt(r2);
      if (r1 < EPS_SQRT) continue;
      if (gam * r1 > PRUNE_ARG && alp * r1 > PRUNE_ARG) continue;
      const double gtmp = +2 * gam * exp(-r2 * gam2) / (SQRTPI * r2) - erf(r1 * gam) / (r2 * r1);
      const double atmp = -2 * alp * exp(-r2 * alp2) / (SQRTPI * r2) + erf(r1 * alp) / (r2 * r1);
      const double f = gtmp + atmp;
      g[0] += f * x; g[1] += f * y; g[2] += f * z;
      s[0] += f * x * x; s[1] += f * x * y; s[2] += f * x * z;
      s[3] += f * y * y; s[4] += f * y * z; s[5] += f * z * z;
   }
}

// the same sums with the table-driven erf / exp(-x^2) and the hardware-seeded rsqrt of erfgauss.cuh (pair tiles)
__device__ __forceinline__ double dir_amat_fast(const double2* __restrict__ tab, const double* __restrict__ dt, double vx,
                                                double vy, double vz, double gam, double alp) {
   double acc = 0.0;
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      if (r2 < EPS_SQRT * EPS_SQRT) continue;
      const double rinv = eg::fast_rsqrt(r2), r1 = r2 * rinv;
      if (gam * r1 > PRUNE_ARG && alp * r1 > PRUNE_ARG) continue;
      double g;
      acc += (eg::erf_gauss<5, false>(tab, gam * r1, g) - eg::erf_gauss<5, false>(tab, alp * r1, g)) * rinv;
   }
   return acc;
}

__device__ __forceinline__ void dir_damat_fast(const double2* __restrict__ tab, const double* __restrict__ dt, double vx,
                                               double vy, double vz, double gam, double alp, double* g, double* s) {
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      if (r2 < EPS_SQRT * EPS_SQRT) continue;
      const double rinv = eg::fast_rsqrt(r2), r1 = r2 * rinv;
      if (gam * r1 > PRUNE_ARG && alp * r1 > PRUNE_ARG) continue;
      double eg_, ea_;
      const double fg = eg::erf_gauss<6, true>(tab, gam * r1, eg_), fa = eg::erf_gauss<6, true>(tab, alp * r1, ea_);
      // gtmp + atmp of get_damat_dir_3d: (2 gam e_g - 2 alp e_a) / (sqrt(pi) r^2) - (erf_g - erf_a) / r^3
      const double f = ((2.0 / SQRTPI) * (gam * eg_ - alp * ea_) - (fg - fa) * rinv) * rinv * rinv;
      g[0] += f * x; g[1] += f * y; g[2] += f * z;
      s[0] += f * x * x; s[1] += f * x * y; s[2] += f * x * z;
      s[3] += f * y * y; s[4] += f * y * z; s[5] += f * z * z;
   }
}

struct PbcDev {
   const double* wtrans;  // WSC candidate translations [3 * nw]
   int nw;
   const double* dtrans;  // [3 * 125]
   const double* rtrans;  // [3 * 124]
   const double* rcoef;   // [124]  4 pi / vol * exp(-g2 / (4 alpha^2)) / g2   (0 where g2 < eps)
   double alpha;
   const double2* etab;   // erf / exp(-x^2) node table (erfgauss.cuh)
};

// ---------------------------------------------------------------------------------------------
// phases: Phi(a, G) = cos(G.r_a), Phi(a, 124 + G) = sin(G.r_a); PhiW = Phi * coef; padding zero
// ---------------------------------------------------------------------------------------------
__global__ void pbc_phase_kernel(int nat, int64_t npad, const double* __restrict__ xyz, PbcDev pb,
                                 double* __restrict__ Phi, double* __restrict__ PhiW) {
   const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int col = blockIdx.y;  // 0..255
   if (a >= npad) return;
   double v = 0.0, w = 0.0;
   if (a < nat && col < 2 * NRT) {
      const int gidx = col % NRT;
      const double* gv = pb.rtrans + 3 * gidx;
      const double ph = gv[0] * xyz[3 * a] + gv[1] * xyz[3 * a + 1] + gv[2] * xyz[3 * a + 2];
      v = (col < NRT) ? cos(ph) : sin(ph);
      w = v * pb.rcoef[gidx];
   }
   Phi[a + (int64_t)col * npad] = v;
   if (PhiW) PhiW[a + (int64_t)col * npad] = w;
}

// structure factors SF[col] = sum_b q_b Phi(b, col); one warp per column
__global__ void __launch_bounds__(256) pbc_sfac_kernel(int nat, int64_t npad, const double* __restrict__ Phi,
                                                       const double* __restrict__ q, double* __restrict__ sf) {
   const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (col >= KREC) return;
   double acc = 0.0;
   for (int b = lane; b < nat; b += 32) acc = fma(q[b], Phi[b + (int64_t)col * npad], acc);
   acc = warp_sum_p(acc);
   if (lane == 0) sf[col] = acc;
}

// reciprocal-space row sums from the structure factors: atrace_a, dadL_a (written, not accumulated)
__global__ void pbc_rec_rows_kernel(int nat, int64_t npad, const double* __restrict__ Phi,
                                    const double* __restrict__ sf, PbcDev pb, double* __restrict__ atrace,
                                    double* __restrict__ dadL) {
   const int a = blockIdx.x * blockDim.x + threadIdx.x;
   if (a >= nat) return;
   const double alp2 = pb.alpha * pb.alpha;
   double at[3] = {0.0, 0.0, 0.0}, dl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   for (int gi = 0; gi < NRT; ++gi) {
      const double coef = pb.rcoef[gi];
      if (coef == 0.0) continue;
      const double gx = pb.rtrans[3 * gi], gy = pb.rtrans[3 * gi + 1], gz = pb.rtrans[3 * gi + 2];
      const double g2 = gx * gx + gy * gy + gz * gz;
      const double ca = Phi[a + (int64_t)gi * npad], sa = Phi[a + (int64_t)(NRT + gi) * npad];
      const double CG = sf[gi], SG = sf[NRT + gi];
      // sum_b q_b sin(G.(r_a - r_b)) and sum_b q_b cos(G.(r_a - r_b))
      const double ssum = sa * CG - ca * SG, csum = ca * CG + sa * SG;
      const double t = -coef * ssum;  // dG_rec = -sin(G.v) etmp G   (eeq.f90:564-565)
      at[0] += t * gx; at[1] += t * gy; at[2] += t * gz;
      const double pre = 2.0 / g2 + 0.5 / alp2, u = coef * csum;  // eeq.f90:566-567
      dl[0] += u * (pre * gx * gx - 1.0); dl[1] += u * (pre * gx * gy); dl[2] += u * (pre * gx * gz);
      dl[3] += u * (pre * gy * gy - 1.0); dl[4] += u * (pre * gy * gz); dl[5] += u * (pre * gz * gz - 1.0);
   }
   atrace[3 * a] = at[0]; atrace[3 * a + 1] = at[1]; atrace[3 * a + 2] = at[2];
   double* d = dadL + (size_t)9 * a;
   d[0] = dl[0]; d[1] = dl[1]; d[2] = dl[2];
   d[3] = dl[1]; d[4] = dl[3]; d[5] = dl[4];
   d[6] = dl[2]; d[7] = dl[4]; d[8] = dl[5];
}

// Psi((3j+c), G) = -q_j coef_G G_c s_Gj ; Psi((3j+c), 124+G) = +q_j coef_G G_c c_Gj ; rest zero
// (rows of the atom block [jb, jb + nj), numbered within the block)
__global__ void pbc_psi_kernel(int jb, int nj, int64_t npad, int64_t mp, const double* __restrict__ Phi,
                               const double* __restrict__ q, PbcDev pb, double* __restrict__ Psi) {
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int col = blockIdx.y;
   if (r >= mp) return;
   double v = 0.0;
   if (r < 3 * (int64_t)nj && col < 2 * NRT) {
      const int j = jb + (int)(r / 3), c = (int)(r % 3), gi = col % NRT;
      const double f = q[j] * pb.rcoef[gi] * pb.rtrans[3 * gi + c];
      v = (col < NRT) ? -f * Phi[j + (int64_t)(NRT + gi) * npad] : f * Phi[j + (int64_t)gi * npad];
   }
   Psi[r + (int64_t)col * mp] = v;
}

// ---------------------------------------------------------------------------------------------
// real-space pair tiles: 32 x 32 atoms per CTA over the lower block triangle; warp = row i, lane = j
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_coords(int b, int& ti, int& tj) {
   int r = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
   while ((int64_t)(r + 1) * (r + 2) / 2 <= b) ++r;
   while ((int64_t)r * (r + 1) / 2 > b) --r;
   ti = r;
   tj = b - r * (r + 1) / 2;
}

// W(i, j) += direct part (both triangles); W is pre-filled with the reciprocal matrix
__global__ void __launch_bounds__(256) pbc_amat_tiles_kernel(int nat, const double* __restrict__ xyz,
                                                             const double* __restrict__ rad, PbcDev pb,
                                                             double* __restrict__ W, int64_t ld, int tile0, int tstride) {
   __shared__ double sdt[3 * NDT];
   __shared__ double swt[3 * 27];
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, pb.etab, (int)threadIdx.x, (int)blockDim.x);
   for (int k = threadIdx.x; k < 3 * NDT; k += blockDim.x) sdt[k] = pb.dtrans[k];
   for (int k = threadIdx.x; k < 3 * pb.nw; k += blockDim.x) swt[k] = pb.wtrans[k];
   __syncthreads();
   int ti, tj;
   tile_coords((int)blockIdx.x * tstride + tile0, ti, tj);  // several GPUs: tiles tile0, tile0 + tstride, ...
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int j = tj * 32 + lane;
   for (int rr = 0; rr < 4; ++rr) {
      const int i = ti * 32 + warp + 8 * rr;
      if (i >= nat || j >= i) continue;  // strictly lower triangle
      const double vx = xyz[3 * j] - xyz[3 * i], vy = xyz[3 * j + 1] - xyz[3 * i + 1], vz = xyz[3 * j + 2] - xyz[3 * i + 2];
      const double gam = 1.0 / sqrt(rad[i] * rad[i] + rad[j] * rad[j]);
      const unsigned mask = wsc_select(swt, pb.nw, -vx, -vy, -vz);  // wsc vec = r_i - r_j
      const double wsw = 1.0 / (double)__popc(mask);
      double acc = 0.0;
      for (unsigned m = mask; m; m &= m - 1) {
         const int p = __ffs(m) - 1;
         acc += dir_amat_fast(setab, sdt, vx + swt[3 * p], vy + swt[3 * p + 1], vz + swt[3 * p + 2], gam, pb.alpha) * wsw;
      }
      W[i + (int64_t)j * ld] += acc;
      W[j + (int64_t)i * ld] += acc;
   }
}

// diagonal: self images (eeq.f90:284-294); padding rows become the identity
__global__ void pbc_amat_diag_kernel(int nat, int64_t npad, const double* __restrict__ rad,
                                     const double* __restrict__ eta, PbcDev pb, double* __restrict__ W, int64_t ld,
                                     int row0, int row1) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npad) return;
   if (i >= nat) {
      if (row0 == 0) W[i + i * ld] = 1.0;  // (several GPUs: the rank that owns the first rows)
      return;
   }
   if (i < row0 || i >= row1) return;
   const double gam = 1.0 / sqrt(2.0 * rad[i] * rad[i]);
   const unsigned mask = wsc_select(pb.wtrans, pb.nw, 0.0, 0.0, 0.0);
   const double wsw = 1.0 / (double)__popc(mask);
   double acc = 0.0;
   for (unsigned m = mask; m; m &= m - 1) {
      const int p = __ffs(m) - 1;
      acc += dir_amat(pb.dtrans, pb.wtrans[3 * p], pb.wtrans[3 * p + 1], pb.wtrans[3 * p + 2], gam, pb.alpha) * wsw;
   }
   W[i + i * ld] += acc + eta[i] + SQRT2PI / rad[i] - 2.0 * pb.alpha / SQRTPI;
}

// derivative tiles: one evaluation of dG, dS per unique pair feeds
//   B(3j+c, i) = +dG q_j, B(3i+c, j) = -dG q_i        (eeq.f90:482-483)
//   atrace_i -= dG q_j, atrace_j += dG q_i            (eeq.f90:480-481)
//   dadL_j += dS q_i, dadL_i += dS q_j                (eeq.f90:484-485)
__global__ void __launch_bounds__(256) pbc_damat_tiles_kernel(int nat, const double* __restrict__ xyz,
                                                              const double* __restrict__ rad,
                                                              const double* __restrict__ q, PbcDev pb,
                                                              double* __restrict__ B, int64_t ldb,
                                                              double* __restrict__ atrace, double* __restrict__ dadL,
                                                              int tile0, int tstride, int jb, int je) {
   __shared__ double sdt[3 * NDT];
   __shared__ double swt[3 * 27];
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, pb.etab, (int)threadIdx.x, (int)blockDim.x);
   for (int k = threadIdx.x; k < 3 * NDT; k += blockDim.x) sdt[k] = pb.dtrans[k];
   for (int k = threadIdx.x; k < 3 * pb.nw; k += blockDim.x) swt[k] = pb.wtrans[k];
   __syncthreads();
   int ti, tj;
   tile_coords((int)blockIdx.x * tstride + tile0, ti, tj);
   // B is written for the rows of the atom block [jb, je) only (rows numbered within the block); a pass without the
   // row sums (atrace == nullptr) skips the tiles that hold no atom of the block
   const bool sums = atrace != nullptr;
   if (!sums && (ti * 32 >= je || ti * 32 + 32 <= jb) && (tj * 32 >= je || tj * 32 + 32 <= jb)) return;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int j = tj * 32 + lane;
   double aj[3] = {0.0, 0.0, 0.0}, lj[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   const double qj = (j < nat) ? q[j] : 0.0;
   for (int rr = 0; rr < 4; ++rr) {
      const int i = ti * 32 + warp + 8 * rr;  // warp-uniform
      if (i >= nat) continue;
      double g[3] = {0.0, 0.0, 0.0}, s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      const bool active = (j < i);
      const double qi = q[i];
      if (active) {
         const double vx = xyz[3 * j] - xyz[3 * i], vy = xyz[3 * j + 1] - xyz[3 * i + 1], vz = xyz[3 * j + 2] - xyz[3 * i + 2];
         const double gam = 1.0 / sqrt(rad[i] * rad[i] + rad[j] * rad[j]);
         const unsigned mask = wsc_select(swt, pb.nw, -vx, -vy, -vz);
         const double wsw = 1.0 / (double)__popc(mask);
         for (unsigned m = mask; m; m &= m - 1) {
            const int p = __ffs(m) - 1;
            dir_damat_fast(setab, sdt, vx + swt[3 * p], vy + swt[3 * p + 1], vz + swt[3 * p + 2], gam, pb.alpha, g, s);
         }
#pragma unroll
         for (int c = 0; c < 3; ++c) g[c] *= wsw;
#pragma unroll
         for (int c = 0; c < 6; ++c) s[c] *= wsw;
         if (B) {
            double* bji = B + (int64_t)i * ldb + 3 * (j - jb);  // dadr(:, j, i)
            double* bij = B + (int64_t)j * ldb + 3 * (i - jb);  // dadr(:, i, j)
            const bool jin = j >= jb && j < je, iin = i >= jb && i < je;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
               if (jin) bji[c] = g[c] * qj;
               if (iin) bij[c] = -g[c] * qi;
            }
         }
#pragma unroll
         for (int c = 0; c < 3; ++c) aj[c] += g[c] * qi;
#pragma unroll
         for (int c = 0; c < 6; ++c) lj[c] += s[c] * qi;
      }
      if (!sums) continue;  // (warp-uniform)
      // row-i reductions over the lanes
      double ri[9];
#pragma unroll
      for (int c = 0; c < 3; ++c) ri[c] = warp_sum_p(active ? -g[c] * qj : 0.0);
#pragma unroll
      for (int c = 0; c < 6; ++c) ri[3 + c] = warp_sum_p(active ? s[c] * qj : 0.0);
      if (lane == 0) {
#pragma unroll
         for (int c = 0; c < 3; ++c) atomicAdd(atrace + 3 * i + c, ri[c]);
         double* d = dadL + (size_t)9 * i;
         atomicAdd(d + 0, ri[3]); atomicAdd(d + 1, ri[4]); atomicAdd(d + 2, ri[5]);
         atomicAdd(d + 3, ri[4]); atomicAdd(d + 4, ri[6]); atomicAdd(d + 5, ri[7]);
         atomicAdd(d + 6, ri[5]); atomicAdd(d + 7, ri[7]); atomicAdd(d + 8, ri[8]);
      }
   }
   if (sums && j < nat) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
         if (aj[c] != 0.0) atomicAdd(atrace + 3 * j + c, aj[c]);
      if (lj[0] != 0.0 || lj[3] != 0.0 || lj[5] != 0.0) {
         double* d = dadL + (size_t)9 * j;
         atomicAdd(d + 0, lj[0]); atomicAdd(d + 1, lj[1]); atomicAdd(d + 2, lj[2]);
         atomicAdd(d + 3, lj[1]); atomicAdd(d + 4, lj[3]); atomicAdd(d + 5, lj[4]);
         atomicAdd(d + 6, lj[2]); atomicAdd(d + 7, lj[4]); atomicAdd(d + 8, lj[5]);
      }
   }
}

// self-image stress term (eeq.f90:488-497), real-space part: dadL_i += dS_self q_i
__global__ void pbc_damat_diag_kernel(int nat, const double* __restrict__ rad, const double* __restrict__ q,
                                      PbcDev pb, double* __restrict__ dadL, int row0, int row1) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nat || i < row0 || i >= row1) return;
   const double gam = 1.0 / sqrt(2.0 * rad[i] * rad[i]);
   const unsigned mask = wsc_select(pb.wtrans, pb.nw, 0.0, 0.0, 0.0);
   const double wsw = 1.0 / (double)__popc(mask);
   double g[3] = {0.0, 0.0, 0.0}, s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   for (unsigned m = mask; m; m &= m - 1) {
      const int p = __ffs(m) - 1;
      dir_damat(pb.dtrans, pb.wtrans[3 * p], pb.wtrans[3 * p + 1], pb.wtrans[3 * p + 2], gam, pb.alpha, g, s);
   }
   const double f = wsw * q[i];
   double* d = dadL + (size_t)9 * i;
   atomicAdd(d + 0, s[0] * f); atomicAdd(d + 1, s[1] * f); atomicAdd(d + 2, s[2] * f);
   atomicAdd(d + 3, s[1] * f); atomicAdd(d + 4, s[3] * f); atomicAdd(d + 5, s[4] * f);
   atomicAdd(d + 6, s[2] * f); atomicAdd(d + 7, s[4] * f); atomicAdd(d + 8, s[5] * f);
}

// y_i = sum_j A(i, j) x_j for a full symmetric matrix (column-major): thread per row, coalesced
__global__ void __launch_bounds__(256) symv_full_kernel(int nat, const double* __restrict__ A, int64_t ld,
                                                        const double* __restrict__ x, double* __restrict__ y) {
   __shared__ double part[8][32];
   // CTA = 32 rows x 8 column slices
   const int r = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
   double acc = 0.0;
   if (r < nat)
      for (int j = slice; j < nat; j += 8) acc = fma(A[r + (int64_t)j * ld], x[j], acc);
   part[slice][threadIdx.x & 31] = acc;
   __syncthreads();
   if (slice == 0 && r < nat) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x & 31];
      y[r] = s;
   }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PbcDev upload_pbc(mchrg_b200_context* ctx, const DevMol& mol, PbcHost& host) {
   int periodic[3] = {1, 1, 1};
   host.wtrans = lat::points_cutoff(periodic, mol.lattice, std::sqrt(DBL_EPSILON));  // wignerseitz.f90:53
   if (host.wtrans.size() / 3 > 27) fail(MCHRG_B200_ERR_ARG, "unexpected number of Wigner-Seitz translations");
   host.dtrans = lat::dir_trans(mol.lattice);
   host.rtrans = lat::rec_trans(mol.lattice);
   host.alpha = lat::ewald_alpha(mol.lattice);
   const double vol = std::fabs(lat::det3(mol.lattice));
   const double pi = 3.14159265358979323846264338327950288;
   const double fac = 4.0 * pi / vol;
   host.rcoef.assign(NRT, 0.0);
   for (int g = 0; g < NRT; ++g) {
      const double* v = &host.rtrans[3 * g];
      const double g2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (g2 < std::sqrt(DBL_EPSILON)) continue;  // eeq.f90:347
      host.rcoef[g] = fac * std::exp(-0.25 * g2 / (host.alpha * host.alpha)) / g2;
   }
   host.packed.clear();
   host.packed.insert(host.packed.end(), host.wtrans.begin(), host.wtrans.end());
   host.packed.resize(81, 0.0);
   host.packed.insert(host.packed.end(), host.dtrans.begin(), host.dtrans.end());
   host.packed.insert(host.packed.end(), host.rtrans.begin(), host.rtrans.end());
   host.packed.insert(host.packed.end(), host.rcoef.begin(), host.rcoef.end());
   double* d = ctx->alloc<double>(host.packed.size());
   MCHRG_CUDA(cudaMemcpyAsync(d, host.packed.data(), host.packed.size() * sizeof(double), cudaMemcpyHostToDevice,
                              ctx->stream));
   PbcDev pb;
   pb.wtrans = d;
   pb.nw = (int)(host.wtrans.size() / 3);
   pb.dtrans = d + 81;
   pb.rtrans = pb.dtrans + 3 * NDT;
   pb.rcoef = pb.rtrans + 3 * NRT;
   pb.alpha = host.alpha;
   pb.etab = ctx->erf_tab;
   return pb;
}

struct PbcState {
   PbcDev pb;
   double* Phi;
};

void pbc_begin(mchrg_b200_context* ctx, const DevMol& mol, PbcHost& host, PbcWork& work) {
   const int64_t npad = round_up(mol.nat, TILE);
   auto* st = new PbcState();
   st->pb = upload_pbc(ctx, mol, host);
   st->Phi = ctx->alloc<double>((size_t)npad * KREC);
   work.npad = npad;
   work.state = st;
}

void pbc_end(PbcWork& work) {
   delete static_cast<PbcState*>(work.state);
   work.state = nullptr;
}

void pbc_amat(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a, PbcWork& work,
              double* W, bool shard) {
   auto* st = static_cast<PbcState*>(work.state);
   const int64_t npad = work.npad;
   const int nat = mol.nat;
   const int R = shard ? ctx->nranks() : 1, r = shard ? ctx->rank() : 0;
   double* PhiW = ctx->alloc<double>((size_t)npad * KREC);
   dim3 gp((unsigned)((npad + 255) / 256), KREC);
   MCHRG_LAUNCH(ctx, pbc_phase_kernel, gp, 256, 0, nat, npad, mol.xyz, st->pb, st->Phi, PhiW);
   // reciprocal space Phi W Phi^T: row blocks of 128 shared between the ranks
   const int64_t nrt = npad / TILE, t0 = nrt * r / R, t1 = nrt * (r + 1) / R;
   if (shard) MCHRG_CUDA(cudaMemsetAsync(W, 0, (size_t)npad * npad * sizeof(double), ctx->active));
   if (t1 > t0)
      dgemm_padded(ctx, true, (t1 - t0) * TILE, npad, KREC, 1.0, PhiW + t0 * TILE, npad, st->Phi, npad, 0.0, W + t0 * TILE, npad,
                   false);
   const int nt = (nat + 31) / 32, ntiles = nt * (nt + 1) / 2;
   if (ntiles > r)
      MCHRG_LAUNCH(ctx, pbc_amat_tiles_kernel, (unsigned)((ntiles - r + R - 1) / R), 256, 0, nat, mol.xyz, rad_a, st->pb, W, npad,
                   r, R);
   int row0 = 0, row1 = nat;  // self images of the diagonal: rows shared like the other row passes
   if (shard) row_share(ctx, nat, &row0, &row1);
   MCHRG_LAUNCH(ctx, pbc_amat_diag_kernel, (unsigned)((npad + 127) / 128), 128, 0, nat, npad, rad_a, eta_a, st->pb, W, npad,
                row0, row1);
   if (shard) comm_allreduce_sum(ctx, W, (size_t)npad * npad);
   work.phase_ready = true;
}

// ---------------------------------------------------------------------------------------------
// Periodic EEQ-BC, charge / energy path (eeqbc.f90:532-627, 1065-1165, 283-311).  The bond capacity
// c(r) = sqrt(cap_i cap_j) / 2 (1 + erf(-kbc (r - rvdw) / rvdw)) is short ranged, so the reference uses plain
// direct sums (no Ewald split): per WSC image the 125 direct translations.  One evaluation per unique pair
// feeds C(i, j) = C(j, i) = -sum c and A(i, j) = A(j, i) = -sum c erf(gamma r) / r.
// atoms[5 a + 0] = rad_eff(a), [2] = effective hardness h_a, [4] = cap_a (bc_atom_kernel in eeqbc.cu).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bc_dir_sums(const double* __restrict__ dt, double vx, double vy, double vz, double gam,
                                            double kbc, double rv, double sc, double& csum, double& asum) {
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r1 = sqrt(x * x + y * y + z * z);
      if (r1 < EPS_SQRT) continue;
      const double targ = kbc * (r1 - rv) / rv;
      if (targ > PRUNE_ARG) continue;  // 1 + erf(-t) == 0 to double precision beyond
      const double c = sc * 0.5 * (1.0 + erf(-targ));  // get_cpair, eeqbc.f90:1130-1143
      csum += c;
      asum -= c * erf(gam * r1) / r1;  // get_amat_dir_3d, eeqbc.f90:600-627
   }
}

__global__ void __launch_bounds__(256) bc_pbc_pair_tiles_kernel(int nat, const double* __restrict__ xyz,
                                                                const int* __restrict__ id,
                                                                const double* __restrict__ atoms,
                                                                const double* __restrict__ rvdw, int nid, double kbc,
                                                                PbcDev pb, double* __restrict__ Cm,
                                                                double* __restrict__ W, int64_t ld) {
   __shared__ double sdt[3 * NDT];
   __shared__ double swt[3 * 27];
   for (int k = threadIdx.x; k < 3 * NDT; k += blockDim.x) sdt[k] = pb.dtrans[k];
   for (int k = threadIdx.x; k < 3 * pb.nw; k += blockDim.x) swt[k] = pb.wtrans[k];
   __syncthreads();
   int ti, tj;
   tile_coords(blockIdx.x, ti, tj);
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int j = tj * 32 + lane;
   for (int rr = 0; rr < 4; ++rr) {
      const int i = ti * 32 + warp + 8 * rr;
      if (i >= nat || j >= i) continue;  // strictly lower triangle
      const double vx = xyz[3 * j] - xyz[3 * i], vy = xyz[3 * j + 1] - xyz[3 * i + 1], vz = xyz[3 * j + 2] - xyz[3 * i + 2];
      const double ri = atoms[5 * i], rj = atoms[5 * j];
      const double gam = 1.0 / sqrt(ri * ri + rj * rj);
      const double sc = sqrt(atoms[5 * i + 4] * atoms[5 * j + 4]);
      const double rv = rvdw[(id[i] - 1) + (size_t)nid * (id[j] - 1)];
      const unsigned mask = wsc_select(swt, pb.nw, -vx, -vy, -vz);  // wsc vec = r_i - r_j
      const double wsw = 1.0 / (double)__popc(mask);
      double csum = 0.0, asum = 0.0;
      for (unsigned m = mask; m; m &= m - 1) {
         const int p = __ffs(m) - 1;
         bc_dir_sums(sdt, vx + swt[3 * p], vy + swt[3 * p + 1], vz + swt[3 * p + 2], gam, kbc, rv, sc, csum, asum);
      }
      Cm[i + (int64_t)j * ld] = Cm[j + (int64_t)i * ld] = -csum * wsw;
      W[i + (int64_t)j * ld] = W[j + (int64_t)i * ld] = asum * wsw;
   }
}

// diagonal: C(a, a) = sum_b c_ab + self images, A(a, a) = self images + h_a C(a, a) + 1; cself(a) for get_xvec;
// padding: C = 0, A = identity.  One warp per row (the off-diagonal entries of the row are summed).
__global__ void __launch_bounds__(256) bc_pbc_diag_kernel(int nat, int64_t npad, const int* __restrict__ id,
                                                          const double* __restrict__ atoms,
                                                          const double* __restrict__ rvdw, int nid, double kbc,
                                                          PbcDev pb, double* __restrict__ Cm, double* __restrict__ W,
                                                          int64_t ld, double* __restrict__ cself) {
   const int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (a >= npad) return;
   if (a >= nat) {
      for (int64_t b = lane; b < npad; b += 32) {
         Cm[b + a * ld] = 0.0;
         Cm[a + b * ld] = 0.0;
         W[b + a * ld] = (b == a) ? 1.0 : 0.0;
         W[a + b * ld] = (b == a) ? 1.0 : 0.0;
      }
      return;
   }
   double rowsum = 0.0;
   for (int b = lane; b < nat; b += 32)
      if (b != a) rowsum -= Cm[b + a * ld];  // C(a, b) = -c_ab
   rowsum = warp_sum_p(rowsum);
   // self images: the 125 direct translations are split over the lanes
   const double ra = atoms[5 * a], capa = atoms[5 * a + 4];
   const double gam = 1.0 / sqrt(2.0 * ra * ra);
   const double rv = rvdw[(id[a] - 1) + (size_t)nid * (id[a] - 1)];
   const unsigned mask = wsc_select(pb.wtrans, pb.nw, 0.0, 0.0, 0.0);
   const double wsw = 1.0 / (double)__popc(mask);
   double cs = 0.0, as = 0.0;
   for (unsigned m = mask; m; m &= m - 1) {
      const int p = __ffs(m) - 1;
      for (int itr = lane; itr < NDT; itr += 32) {
         const double x = pb.wtrans[3 * p] + pb.dtrans[3 * itr], y = pb.wtrans[3 * p + 1] + pb.dtrans[3 * itr + 1],
                      z = pb.wtrans[3 * p + 2] + pb.dtrans[3 * itr + 2];
         const double r1 = sqrt(x * x + y * y + z * z);
         if (r1 < EPS_SQRT) continue;
         const double c = capa * 0.5 * (1.0 + erf(-kbc * (r1 - rv) / rv));
         cs += c;
         as -= c * erf(gam * r1) / r1;
      }
   }
   cs = warp_sum_p(cs) * wsw;
   as = warp_sum_p(as) * wsw;
   if (lane == 0) {
      const double caa = rowsum + cs;
      Cm[a + a * ld] = caa;
      W[a + a * ld] = as + caa * atoms[5 * a + 2] + 1.0;
      cself[a] = cs;
   }
}

// xvec(a) -= cself(a) xtmp(a): eliminate the self interaction (eeqbc.f90:283-311)
__global__ void bc_pbc_xvec_fix_kernel(int nat, const double* __restrict__ atoms, const double* __restrict__ cself,
                                       double* __restrict__ xvec) {
   const int a = blockIdx.x * blockDim.x + threadIdx.x;
   if (a < nat) xvec[a] -= cself[a] * atoms[5 * a + 3];
}

void bc_pbc_matrices(mchrg_b200_context* ctx, const DevMol& mol, const double* atoms, const double* rvdw, int nid,
                     double kbc, PbcWork& work, double* Cm, double* W, double* cself) {
   auto* st = static_cast<PbcState*>(work.state);
   const int64_t npad = work.npad;
   const int nat = mol.nat;
   const int nt = (nat + 31) / 32;
   MCHRG_LAUNCH(ctx, bc_pbc_pair_tiles_kernel, (unsigned)(nt * (nt + 1) / 2), 256, 0, nat, mol.xyz, mol.id, atoms, rvdw, nid,
                kbc, st->pb, Cm, W, npad);
   MCHRG_LAUNCH(ctx, bc_pbc_diag_kernel, (unsigned)((npad + 7) / 8), 256, 0, nat, npad, mol.id, atoms, rvdw, nid, kbc, st->pb,
                Cm, W, npad, cself);
}

void bc_pbc_xvec_fix(mchrg_b200_context* ctx, int nat, const double* atoms, const double* cself, double* xvec) {
   MCHRG_LAUNCH(ctx, bc_pbc_xvec_fix_kernel, (unsigned)((nat + 255) / 256), 256, 0, nat, atoms, cself, xvec);
}

// ---------------------------------------------------------------------------------------------
// Periodic EEQ-BC derivatives (get_damat_3d eeqbc.f90:789-924 with get_damat_dir :926-957 and get_damat_dc_dir
// :959-985, get_dcmat_3d :1243-1313, the periodic branch of get_xvec_derivs :364-418), assembled directly into
//   B(3k+c, i) = [dadr + diag(atrace)](c, k, i) - dxdr(c, k, i)          (type.F90:267-269, 283-286)
// and the contractions  ga = dadr q, gx = dxdr q  (type.F90:277-278), dadL, dxdL per atom.
// Per unique pair (i > j) ONE pass over the images x 125 translations yields the seven lattice sums
//   D1, S1, g1 (get_damat_dir), D2, S2 (get_damat_dc_dir), D3, S3 (get_dcpair_dir);   all S are symmetric.
// ---------------------------------------------------------------------------------------------
struct BcSums {
   double d1[3], s1[6], g1, d2[3], s2[6], d3[3], s3[6];
};

__device__ __forceinline__ void bc_dir_deriv_sums(const double* __restrict__ dt, double vx, double vy, double vz,
                                                  double gam, double kbc, double rv, double sc, double w, BcSums& o) {
   const double gam2 = gam * gam;
   for (int itr = 0; itr < NDT; ++itr) {
      const double x = vx + dt[3 * itr], y = vy + dt[3 * itr + 1], z = vz + dt[3 * itr + 2];
      const double r2 = x * x + y * y + z * z;
      const double r1 = sqrt(r2);
      if (r1 < EPS_SQRT) continue;
      const double t = kbc * (r1 - rv) / rv;
      if (t > PRUNE_ARG) continue;  // capacity and its derivative vanish to double precision
      const double c = sc * 0.5 * (1.0 + erf(-t));                             // get_cpair
      const double dc = -sc * kbc * exp(-(t * t)) / (SQRTPI * rv) / r1;        // get_dcpair: gradient = dc * vec
      const double e = exp(-r2 * gam2), ef = erf(r1 * gam);
      const double gtmp = 2.0 * gam * e / (SQRTPI * r2) - ef / (r2 * r1);
      const double f1 = -c * gtmp * w, f2 = ef / r1 * dc * w, f3 = dc * w;
      const double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
      o.d1[0] += f1 * x; o.d1[1] += f1 * y; o.d1[2] += f1 * z;
      o.s1[0] += f1 * xx; o.s1[1] += f1 * xy; o.s1[2] += f1 * xz; o.s1[3] += f1 * yy; o.s1[4] += f1 * yz; o.s1[5] += f1 * zz;
      o.g1 += c * 2.0 * e / SQRTPI * w;
      o.d2[0] += f2 * x; o.d2[1] += f2 * y; o.d2[2] += f2 * z;
      o.s2[0] += f2 * xx; o.s2[1] += f2 * xy; o.s2[2] += f2 * xz; o.s2[3] += f2 * yy; o.s2[4] += f2 * yz; o.s2[5] += f2 * zz;
      o.d3[0] += f3 * x; o.d3[1] += f3 * y; o.d3[2] += f3 * z;
      o.s3[0] += f3 * xx; o.s3[1] += f3 * xy; o.s3[2] += f3 * xz; o.s3[3] += f3 * yy; o.s3[4] += f3 * yz; o.s3[5] += f3 * zz;
   }
}

// symmetric 6-vector (xx, xy, xz, yy, yz, zz) -> entry k = a + 3 b of the 3 x 3 array
__device__ __forceinline__ double sym9(const double* s6, int k) {
   constexpr int map[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
   return s6[map[k]];
}

// per-atom accumulators of the derivative path
struct BcDerivAcc {
   double* bdiag;   // [3 nat]  diagonal blocks B(3i+c, i)
   double* ga;      // [3 nat]  dadr q (with atrace on the diagonal)
   double* gx;      // [3 nat]  dxdr q
   double* dadL;    // [9 nat]
   double* dxdL;    // [9 nat]
   double* dcdrd;   // [3 nat]  dcdr(:, i, i)
   double* dcdLs;   // [9 nat]  dcdL(:, :, i)
   double* atr;     // [3 nat]  atrace (may be null)
   // B = sa (dadr + add_atrace diag(atrace)) - sx dxdr: the solve uses (1, 1, true); the piecewise accessors
   // get_coulomb_derivs (1, 0, false) and get_xvec_derivs (0, -1, false; q = 0)
   double sa, sx;
   int add_atrace;
};

__global__ void __launch_bounds__(256) bc_pbc_deriv_tiles_kernel(int nat, const double* __restrict__ xyz,
                                                                 const int* __restrict__ id,
                                                                 const double* __restrict__ atoms,
                                                                 const double* __restrict__ rvdw, int nid, double kbc,
                                                                 const double* __restrict__ q,
                                                                 const double* __restrict__ dcndr,
                                                                 const double* __restrict__ dcndL, PbcDev pb,
                                                                 double* __restrict__ B, int64_t ldb, BcDerivAcc acc) {
   __shared__ double sdt[3 * NDT];
   __shared__ double swt[3 * 27];
   for (int k = threadIdx.x; k < 3 * NDT; k += blockDim.x) sdt[k] = pb.dtrans[k];
   for (int k = threadIdx.x; k < 3 * pb.nw; k += blockDim.x) swt[k] = pb.wtrans[k];
   __syncthreads();
   int ti, tj;
   tile_coords(blockIdx.x, ti, tj);
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int j = tj * 32 + lane;
   const size_t n3 = 3 * (size_t)nat;
   for (int rr = 0; rr < 4; ++rr) {
      const int i = ti * 32 + warp + 8 * rr;
      if (i >= nat || j >= i) continue;  // strictly lower triangle
      const double vx = xyz[3 * j] - xyz[3 * i], vy = xyz[3 * j + 1] - xyz[3 * i + 1], vz = xyz[3 * j + 2] - xyz[3 * i + 2];
      const double ri = atoms[5 * i], dri = atoms[5 * i + 1], hi = atoms[5 * i + 2], xti = atoms[5 * i + 3];
      const double rj = atoms[5 * j], drj = atoms[5 * j + 1], hj = atoms[5 * j + 2], xtj = atoms[5 * j + 3];
      const double gam = 1.0 / sqrt(ri * ri + rj * rj);
      const double sc = sqrt(atoms[5 * i + 4] * atoms[5 * j + 4]);
      const double rv = rvdw[(id[i] - 1) + (size_t)nid * (id[j] - 1)];
      const unsigned mask = wsc_select(swt, pb.nw, -vx, -vy, -vz);  // wsc vec = r_i - r_j
      const double wsw = 1.0 / (double)__popc(mask);
      BcSums sm = {};
      for (unsigned m = mask; m; m &= m - 1) {
         const int p = __ffs(m) - 1;
         bc_dir_deriv_sums(sdt, vx + swt[3 * p], vy + swt[3 * p + 1], vz + swt[3 * p + 2], gam, kbc, rv, sc, wsw, sm);
      }
      const double qi = q[i], qj = q[j], gam3 = gam * gam * gam;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
         // the two columns of the pair's dgamdr that the reference reads (eeqbc.f90:847-850, 873-878)
         const double dgi = -(ri * dri * dcndr[c + 3 * (i + (size_t)nat * i)] + rj * drj * dcndr[c + 3 * (i + (size_t)nat * j)]) * gam3;
         const double dgj = -(ri * dri * dcndr[c + 3 * (j + (size_t)nat * i)] + rj * drj * dcndr[c + 3 * (j + (size_t)nat * j)]) * gam3;
         const double a_ij = -sm.d1[c] * qi - sm.g1 * qi * dgi + qi * sm.d2[c] - hj * qj * sm.d3[c];  // dadr(c, i, j)
         const double a_ji = +sm.d1[c] * qj - sm.g1 * qj * dgj - qj * sm.d2[c] + hi * qi * sm.d3[c];  // dadr(c, j, i)
         const double at_i = -sm.d1[c] * qj + sm.g1 * qj * dgj + qj * sm.d2[c];                      // atrace(c, i)
         const double at_j = +sm.d1[c] * qi + sm.g1 * qi * dgi - qi * sm.d2[c];                      // atrace(c, j)
         const double x_off = (xti - xtj) * sm.d3[c];                                               // dxdr(c, i, j) = dxdr(c, j, i)
         const double x_ii = xtj * sm.d3[c], x_jj = -xti * sm.d3[c];                                // dxdr(c, i, i), dxdr(c, j, j)
         if (B) {
            B[(3 * (size_t)i + c) + (size_t)j * ldb] += acc.sa * a_ij - acc.sx * x_off;
            B[(3 * (size_t)j + c) + (size_t)i * ldb] += acc.sa * a_ji - acc.sx * x_off;
         }
         atomicAdd(acc.bdiag + 3 * i + c, (acc.add_atrace ? acc.sa * at_i : 0.0) - acc.sx * x_ii);
         atomicAdd(acc.bdiag + 3 * j + c, (acc.add_atrace ? acc.sa * at_j : 0.0) - acc.sx * x_jj);
         if (acc.atr) {
            atomicAdd(acc.atr + 3 * i + c, at_i);
            atomicAdd(acc.atr + 3 * j + c, at_j);
         }
         atomicAdd(acc.ga + 3 * i + c, a_ij * qj + at_i * qi);
         atomicAdd(acc.ga + 3 * j + c, a_ji * qi + at_j * qj);
         atomicAdd(acc.gx + 3 * i + c, x_off * qj + x_ii * qi);
         atomicAdd(acc.gx + 3 * j + c, x_off * qi + x_jj * qj);
         atomicAdd(acc.dcdrd + 3 * i + c, -sm.d3[c]);
         atomicAdd(acc.dcdrd + 3 * j + c, +sm.d3[c]);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
         const double dgl = -(ri * dri * dcndL[k + 9 * (size_t)i] + rj * drj * dcndL[k + 9 * (size_t)j]) * gam3;
         const double s1 = sym9(sm.s1, k), s2 = sym9(sm.s2, k), s3 = sym9(sm.s3, k);
         atomicAdd(acc.dadL + 9 * (size_t)j + k, s1 * qi - sm.g1 * qi * dgl - qi * s2);
         atomicAdd(acc.dadL + 9 * (size_t)i + k, s1 * qj - sm.g1 * qj * dgl - qj * s2);
         atomicAdd(acc.dxdL + 9 * (size_t)i + k, -s3 * xtj);
         atomicAdd(acc.dxdL + 9 * (size_t)j + k, -s3 * xti);
         atomicAdd(acc.dcdLs + 9 * (size_t)i + k, s3);
         atomicAdd(acc.dcdLs + 9 * (size_t)j + k, s3);
      }
      (void)n3;
   }
}

// per atom (one warp): self-image sums, the diagonal / per-atom terms of get_damat_3d and of the periodic
// get_xvec_derivs, the dense part of dxdL (dtmpdL C).  Runs after the pair kernel (needs the complete
// dcdr(:, i, i), dcdL(:, :, i)).
__global__ void __launch_bounds__(256) bc_pbc_deriv_self_kernel(int nat, int64_t npad, const int* __restrict__ id,
                                                                const double* __restrict__ atoms,
                                                                const double* __restrict__ rvdw, int nid, double kbc,
                                                                const double* __restrict__ q, const double* __restrict__ Cm,
                                                                const double* __restrict__ cself,
                                                                const double* __restrict__ kcnchi,
                                                                const double* __restrict__ kqchi,
                                                                const double* __restrict__ kqeta,
                                                                const double* __restrict__ dcndr,
                                                                const double* __restrict__ dcndL,
                                                                const double* __restrict__ dqlocdL, PbcDev pb,
                                                                BcDerivAcc acc) {
   const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (i >= nat) return;
   const double ri = atoms[5 * i], dri = atoms[5 * i + 1], hi = atoms[5 * i + 2], xti = atoms[5 * i + 3], capi = atoms[5 * i + 4];
   const double qi = q[i];
   const double gam = 1.0 / sqrt(2.0 * ri * ri);
   const double rv = rvdw[(id[i] - 1) + (size_t)nid * (id[i] - 1)];
   const unsigned mask = wsc_select(pb.wtrans, pb.nw, 0.0, 0.0, 0.0);
   const double wsw = 1.0 / (double)__popc(mask);
   // self images, the 125 translations split over the lanes (same sums as bc_dir_deriv_sums)
   double s1[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0}, s3[6] = {0, 0, 0, 0, 0, 0}, g1 = 0.0;
   const double gam2 = gam * gam;
   for (unsigned m = mask; m; m &= m - 1) {
      const int p = __ffs(m) - 1;
      for (int itr = lane; itr < NDT; itr += 32) {
         const double x = pb.wtrans[3 * p] + pb.dtrans[3 * itr], y = pb.wtrans[3 * p + 1] + pb.dtrans[3 * itr + 1],
                      z = pb.wtrans[3 * p + 2] + pb.dtrans[3 * itr + 2];
         const double r2 = x * x + y * y + z * z, r1 = sqrt(r2);
         if (r1 < EPS_SQRT) continue;
         const double t = kbc * (r1 - rv) / rv;
         const double c = capi * 0.5 * (1.0 + erf(-t));
         const double dc = -capi * kbc * exp(-(t * t)) / (SQRTPI * rv) / r1;
         const double e = exp(-r2 * gam2), ef = erf(r1 * gam);
         const double gtmp = 2.0 * gam * e / (SQRTPI * r2) - ef / (r2 * r1);
         const double f1 = -c * gtmp * wsw, f2 = ef / r1 * dc * wsw, f3 = dc * wsw;
         const double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
         s1[0] += f1 * xx; s1[1] += f1 * xy; s1[2] += f1 * xz; s1[3] += f1 * yy; s1[4] += f1 * yz; s1[5] += f1 * zz;
         s2[0] += f2 * xx; s2[1] += f2 * xy; s2[2] += f2 * xz; s2[3] += f2 * yy; s2[4] += f2 * yz; s2[5] += f2 * zz;
         s3[0] += f3 * xx; s3[1] += f3 * xy; s3[2] += f3 * xz; s3[3] += f3 * yy; s3[4] += f3 * yz; s3[5] += f3 * zz;
         g1 += c * 2.0 * e / SQRTPI * wsw;
      }
   }
#pragma unroll
   for (int k = 0; k < 6; ++k) { s1[k] = warp_sum_p(s1[k]); s2[k] = warp_sum_p(s2[k]); s3[k] = warp_sum_p(s3[k]); }
   g1 = warp_sum_p(g1);
   // dense part of dxdL: dxdL(:, :, i) = sum_a dtmpdL(:, :, a) C(a, i)   (eeqbc.f90:362)
   double dl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
   for (int a = lane; a < nat; a += 32) {
      const double ca = Cm[a + (int64_t)i * npad];
#pragma unroll
      for (int k = 0; k < 9; ++k) dl[k] += (kcnchi[a] * dcndL[k + 9 * (size_t)a] + kqchi[a] * dqlocdL[k + 9 * (size_t)a]) * ca;
   }
#pragma unroll
   for (int k = 0; k < 9; ++k) dl[k] = warp_sum_p(dl[k]);
   if (lane != 0) return;
   const double cii = Cm[i + (int64_t)i * npad];
   const double dtw = -SQRT2PI * dri / (ri * ri) * qi;
   const double s1c = kqeta[i] * qi * cii, s2c = dtw * cii, cs = cself[i];
#pragma unroll
   for (int k = 0; k < 9; ++k) {
      const double dcdl = acc.dcdLs[9 * (size_t)i + k] + sym9(s3, k);  // + positive diagonal elements (self images)
      const double cnl = dcndL[k + 9 * (size_t)i], qll = dqlocdL[k + 9 * (size_t)i];
      acc.dadL[9 * (size_t)i + k] += sym9(s1, k) * qi - dtw * cnl * g1 - qi * sym9(s2, k)  // self images
                                     + s1c * qll + s2c * cnl                               // hardness, charge width
                                     + hi * qi * dcdl;                                     // capacitance
      acc.dxdL[9 * (size_t)i + k] += dl[k] - sym9(s3, k) * xti + xti * dcdl - cs * (kcnchi[i] * cnl + kqchi[i] * qll);
   }
#pragma unroll
   for (int c = 0; c < 3; ++c) {
      const double dd = acc.dcdrd[3 * i + c];
      // dadr(c, i, i) += h_i q_i dcdr(c, i, i); dxdr(c, i, i) += xtmp_i dcdr(c, i, i);
      // self images: atrace(c, i) += t, dadr(c, i, i) -= t with t = dtw dcndr(c, i, i) g1 (they cancel in the solve)
      const double tself = dtw * dcndr[c + 3 * (i + (size_t)nat * i)] * g1;
      acc.bdiag[3 * i + c] += acc.sa * (hi * qi * dd - tself + (acc.add_atrace ? tself : 0.0)) - acc.sx * xti * dd;
      if (acc.atr) acc.atr[3 * i + c] += tself;
      acc.ga[3 * i + c] += hi * qi * dd * qi;
      acc.gx[3 * i + c] += xti * dd * qi;
   }
}

// thread per row r = 3k+c: the column-scaled terms of dadr / dxdr, the diagonal fold, ga / gx / bz.
//   dadr(r, i) += s1_i dqlocdr(r, i) + s2_i dcndr(r, i),  s1_i = kqeta_i q_i C_ii, s2_i = -sqrt(2/pi) drad_i / rad_i^2 q_i C_ii
//   dxdr(r, i) += -cself_i (kcnchi_i dcndr(r, i) + kqchi_i dqlocdr(r, i))           (self images, eeqbc.f90:407-416)
//   gx(r)      += sum_a dtmpdr(r, a) (C q)_a                                          (the dense part, contracted)
__global__ void __launch_bounds__(128) bc_pbc_deriv_rows_kernel(int nat, int64_t npad, const double* __restrict__ atoms,
                                                                const double* __restrict__ q, const double* __restrict__ cq,
                                                                const double* __restrict__ Cm,
                                                                const double* __restrict__ cself,
                                                                const double* __restrict__ kcnchi,
                                                                const double* __restrict__ kqchi,
                                                                const double* __restrict__ kqeta,
                                                                const double* __restrict__ dcndr,
                                                                const double* __restrict__ dqlocdr,
                                                                const double* __restrict__ zvec, double* __restrict__ B,
                                                                int64_t ldb, BcDerivAcc acc, double* __restrict__ bz) {
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const size_t n3 = 3 * (size_t)nat;
   if (r >= (int64_t)n3) return;
   const int k = (int)(r / 3);
   double ga = 0.0, gx = 0.0, z = 0.0;
   for (int i = 0; i < nat; ++i) {
      const double ri = atoms[5 * i], dri = atoms[5 * i + 1], qi = q[i];
      const double cii = Cm[i + (int64_t)i * npad];
      const double cn = dcndr[r + n3 * i], ql = dqlocdr[r + n3 * i];
      const double v = kqeta[i] * qi * cii * ql + (-SQRT2PI * dri / (ri * ri) * qi * cii) * cn;
      const double u = -cself[i] * (kcnchi[i] * cn + kqchi[i] * ql);
      ga = fma(v, qi, ga);
      gx = fma(u, qi, gx);
      gx = fma(kcnchi[i] * cn + kqchi[i] * ql, cq[i], gx);
      if (B) {
         double b = B[r + (size_t)i * ldb] + acc.sa * v - acc.sx * u;
         if (i == k) b += acc.bdiag[r];
         B[r + (size_t)i * ldb] = b;
         if (zvec) z = fma(b, zvec[i], z);
      }
   }
   acc.ga[r] += ga;
   acc.gx[r] += gx;
   if (B && bz) bz[r] = z;
}

void bc_pbc_derivs(mchrg_b200_context* ctx, const DevMol& mol, const double* atoms, const double* rvdw, int nid, double kbc,
                   const double* kcnchi, const double* kqchi, const double* kqeta, const double* q, const double* cq,
                   const double* Cm, const double* cself, const double* dcndr, const double* dcndL, const double* dqlocdr,
                   const double* dqlocdL, const double* zvec, PbcWork& work, double* B, int64_t ldb, double* ga, double* gx,
                   double* dadL, double* dxdL, double* bz, double sa, double sx, bool add_atrace, double* atrace) {
   auto* st = static_cast<PbcState*>(work.state);
   const int nat = mol.nat;
   BcDerivAcc acc;
   acc.sa = sa;
   acc.sx = sx;
   acc.add_atrace = add_atrace ? 1 : 0;
   acc.atr = atrace;
   acc.bdiag = ctx->alloc_zero<double>((size_t)3 * nat);
   acc.dcdrd = ctx->alloc_zero<double>((size_t)3 * nat);
   acc.dcdLs = ctx->alloc_zero<double>((size_t)9 * nat);
   acc.ga = ga;
   acc.gx = gx;
   acc.dadL = dadL;
   acc.dxdL = dxdL;
   const int nt = (nat + 31) / 32;
   MCHRG_LAUNCH(ctx, bc_pbc_deriv_tiles_kernel, (unsigned)(nt * (nt + 1) / 2), 256, 0, nat, mol.xyz, mol.id, atoms, rvdw, nid,
                kbc, q, dcndr, dcndL, st->pb, B, ldb, acc);
   MCHRG_LAUNCH(ctx, bc_pbc_deriv_self_kernel, (unsigned)((nat + 7) / 8), 256, 0, nat, work.npad, mol.id, atoms, rvdw, nid,
                kbc, q, Cm, cself, kcnchi, kqchi, kqeta, dcndr, dcndL, dqlocdL, st->pb, acc);
   MCHRG_LAUNCH(ctx, bc_pbc_deriv_rows_kernel, (unsigned)((3 * (int64_t)nat + 127) / 128), 128, 0, nat, work.npad, atoms, q, cq,
                Cm, cself, kcnchi, kqchi, kqeta, dcndr, dqlocdr, zvec, B, ldb, acc, bz);
}

void pbc_symv(mchrg_b200_context* ctx, int nat, const double* A, int64_t ld, const double* x, double* y) {
   MCHRG_LAUNCH(ctx, symv_full_kernel, (unsigned)((nat + 31) / 32), 256, 0, nat, A, ld, x, y);
}

void pbc_derivs(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* q, PbcWork& work,
                double* B, int64_t ldb, int64_t mp, double* atrace, double* dadL, bool shard, int jb, int je) {
   auto* st = static_cast<PbcState*>(work.state);
   const int64_t npad = work.npad;
   const int nat = mol.nat;
   if (je < 0) je = nat;
   const int nj = je - jb;
   const int R = shard ? ctx->nranks() : 1, r = shard ? ctx->rank() : 0;
   if (!work.phase_ready) {
      dim3 gp((unsigned)((npad + 255) / 256), KREC);
      MCHRG_LAUNCH(ctx, pbc_phase_kernel, gp, 256, 0, nat, npad, mol.xyz, st->pb, st->Phi, (double*)nullptr);
   }
   if (r == 0) {  // reciprocal-space rows (these OVERWRITE atrace / dadL)
      double* sf = ctx->alloc<double>(KREC);
      MCHRG_LAUNCH(ctx, pbc_sfac_kernel, KREC / 8, 256, 0, nat, npad, st->Phi, q, sf);
      MCHRG_LAUNCH(ctx, pbc_rec_rows_kernel, (unsigned)((nat + 127) / 128), 128, 0, nat, npad, st->Phi, sf, st->pb, atrace,
                   dadL);
   }
   int row0 = 0, row1 = nat;  // self images: rows shared between the ranks
   if (shard) row_share(ctx, nat, &row0, &row1);
   MCHRG_LAUNCH(ctx, pbc_damat_diag_kernel, (unsigned)((nat + 127) / 128), 128, 0, nat, rad_a, q, st->pb, dadL, row0, row1);
   const int nt = (nat + 31) / 32, ntiles = nt * (nt + 1) / 2;
   // One pass over the pair tiles feeds the row sums AND B when this context evaluates every pair and wants every row
   // of B anyway.  Otherwise two passes: the row sums of ALL atoms (tiles shared round robin between the ranks when
   // `shard`; the caller all-reduces them), then B for the atom block [jb, je) from the tiles that hold one of its atoms.
   const bool fused = B && R == 1 && nj == nat;
   if (ntiles > r)
      MCHRG_LAUNCH(ctx, pbc_damat_tiles_kernel, (unsigned)((ntiles - r + R - 1) / R), 256, 0, nat, mol.xyz, rad_a, q, st->pb,
                   fused ? B : (double*)nullptr, ldb, atrace, dadL, r, R, jb, je);
   if (B && !fused)
      MCHRG_LAUNCH(ctx, pbc_damat_tiles_kernel, (unsigned)ntiles, 256, 0, nat, mol.xyz, rad_a, q, st->pb, B, ldb,
                   (double*)nullptr, (double*)nullptr, 0, 1, jb, je);
   if (B) {
      double* Psi = ctx->alloc<double>((size_t)mp * KREC);
      dim3 gs((unsigned)((mp + 255) / 256), KREC);
      MCHRG_LAUNCH(ctx, pbc_psi_kernel, gs, 256, 0, jb, nj, npad, mp, st->Phi, q, st->pb, Psi);
      dgemm_padded(ctx, true, mp, npad, KREC, 1.0, Psi, mp, st->Phi, npad, 1.0, B, ldb, false);
   }
}

}  // namespace mchrg
