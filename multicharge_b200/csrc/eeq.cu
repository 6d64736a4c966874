// EEQ2019 pair kernels (non-periodic) and the erf coordination number, FP64, sm_100a.
//
// Work decomposition: one warp owns one matrix row (atom i) and strides its lanes over the
// partner atoms j, so every per-row reduction (cn_i, atrace_i, dadL_i, (Jq)_i) is a warp shuffle
// and every N^2-sized array (dcndr, amat, B) is written with lanes on the contiguous index.
// These kernels are FP64-ALU bound (erf/exp are polynomial sequences on the FMA pipe) until the
// 24 N^2-byte stores dominate.
#include "eeq.cuh"
#include "dense.cuh"
#include "erfgauss.cuh"

namespace mchrg {

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// ---------------------------------------------------------------------------------------------
__global__ void gather_kernel(int nat, const int* __restrict__ id, const double* __restrict__ src,
                              double* __restrict__ dst) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nat) dst[i] = src[id[i] - 1];
}

void gather_species(mchrg_b200_context* ctx, int nat, const int* id, const double* src, double* dst) {
   MCHRG_LAUNCH(ctx, gather_kernel, (nat + 255) / 256, 256, 0, nat, id, src, dst);
}

// ---------------------------------------------------------------------------------------------
// erf coordination number (mctc_ncoord; restated in SURVEY.md 3.4)
//   count(r)  = 1/2 (1 + erf(-kcn (r - rc) / rc^norm_exp)),  rc = rcov_i + rcov_j
//   dcount/dr = -kcn / (sqrt(pi) rc^norm_exp) exp(-(kcn (r - rc) / rc^norm_exp)^2)
// pass 1: raw cn (row sums).  pass 2: derivatives, scaled by the log-cut factor of the raw cn.
// ---------------------------------------------------------------------------------------------
struct NcDev {
   double kcn, norm_exp, cutoff2, cut;
   int use_en;
};

// Range of the erf counting function: for kcn (r - rc) / rcn > XMAX the count 1/2 (1 - erf) is zero to the last
// bit and its derivative below 5e-19 (the table clamps there), so pairs beyond rc + XMAX rcn / kcn (~ 2 rc for
// kcn = 7.5, far inside the 25 Bohr cutoff of the reference) contribute nothing.  Returns the squared range.
__device__ __forceinline__ double count_range2(const NcDev& p, double rc, double rcn) {
   const double rmax = fma(eg::XMAX / p.kcn, rcn, rc);
   return fmin(p.cutoff2, rmax * rmax);
}

// erf / exp(-x^2) through the node table of erfgauss.cuh (shared-memory copy per CTA) and 1/sqrt through the
// hardware seed + one third-order step: ~4x fewer instructions per pair than the library functions
#define MCHRG_LOAD_ERF_TABLE(etab_global)                                            \
   __shared__ double2 setab[eg::NTAB];                                               \
   eg::load_table(setab, etab_global, (int)threadIdx.x, (int)blockDim.x);            \
   __syncthreads()

constexpr int CN_TJ = 256;  // atoms j per shared-memory tile of the row kernels
__global__ void __launch_bounds__(256) cn_rows_kernel(int nat, const double* __restrict__ xyz,
                                                      const double* __restrict__ rcov, const double* __restrict__ en,
                                                      NcDev p, int ntr, const double* __restrict__ trans,
                                                      double* __restrict__ cn_raw, const double2* __restrict__ etab,
                                                      int row0, int row1) {
   MCHRG_LOAD_ERF_TABLE(etab);
   __shared__ double sj[5][CN_TJ];  // x, y, z, rcov, en of the tile's atoms: the CTA's eight rows share the loads
   const int i = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   const bool on = i < row1;
   const int ic = on ? i : row1 - 1;
   const double xi = xyz[3 * ic], yi = xyz[3 * ic + 1], zi = xyz[3 * ic + 2];
   const double rci = rcov[ic];
   const double eni = p.use_en ? en[ic] : 0.0;
   double acc = 0.0;
   for (int j0 = 0; j0 < nat; j0 += CN_TJ) {
      __syncthreads();
      {
         const int t = threadIdx.x, j = min(j0 + t, nat - 1);
         sj[0][t] = xyz[3 * j]; sj[1][t] = xyz[3 * j + 1]; sj[2][t] = xyz[3 * j + 2];
         sj[3][t] = rcov[j];
         sj[4][t] = p.use_en ? en[j] : 0.0;
      }
      __syncthreads();
      if (!on) continue;
      const int tn = min(CN_TJ, nat - j0);
      for (int t = lane; t < tn; t += 32) {
         const double rc = rci + sj[3][t];
         const double rcn = (p.norm_exp == 1.0) ? rc : pow(rc, p.norm_exp);
         const double den = p.use_en ? (sj[4][t] - eni) : 1.0;
         const double dx = xi - sj[0][t], dy = yi - sj[1][t], dz = zi - sj[2][t];
         const double cut2 = count_range2(p, rc, rcn);
         for (int itr = 0; itr < ntr; ++itr) {
            const double rx = dx - trans[3 * itr], ry = dy - trans[3 * itr + 1], rz = dz - trans[3 * itr + 2];
            const double r2 = rx * rx + ry * ry + rz * rz;
            if (r2 > cut2 || r2 < 1.0e-12) continue;
            const double r1 = r2 * eg::fast_rsqrt(r2);
            const double a = p.kcn * (r1 - rc) / rcn;
            double g;
            acc += den * (0.5 - copysign(0.5, a) * eg::erf_gauss<6, false>(setab, fabs(a), g));  // 1/2 (1 + erf(-a))
         }
      }
   }
   if (!on) return;
   acc = warp_sum(acc);
   if (lane == 0) cn_raw[i] = acc;
}

// dcndr(:, j, i) for all j and dcndL(:, :, i): one warp per atom i, the CTA's eight rows walk the atoms j together in
// tiles of 128 staged in shared memory (coalesced loads instead of eight strided passes over xyz), and every warp hands
// its 32 x 3 results through a shared-memory slab so that the (3, N, N) array -- 6.4 GB at N = 16384, the reason this
// kernel exists -- is written with fully coalesced 256-byte stores (three strided 8-byte stores per lane before:
// 0.9 TB/s).
constexpr int CND_TJ = 128;
__global__ void __launch_bounds__(256) cn_derivs_kernel(int nat, const double* __restrict__ xyz,
                                                        const double* __restrict__ rcov, const double* __restrict__ en,
                                                        NcDev p, int ntr, const double* __restrict__ trans,
                                                        const double* __restrict__ cn_raw, double* __restrict__ dcndr,
                                                        double* __restrict__ dcndL, const double2* __restrict__ etab) {
   MCHRG_LOAD_ERF_TABLE(etab);
   __shared__ double sj[5][CND_TJ];  // x, y, z, rcov, en of the tile's atoms
   __shared__ double wbuf[8][96];   // per warp: 32 atoms x 3 components on their way out
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int i = blockIdx.x * 8 + warp;
   const bool on = i < nat;
   const int ic = on ? i : nat - 1;
   const double xi = xyz[3 * ic], yi = xyz[3 * ic + 1], zi = xyz[3 * ic + 2];
   const double rci = rcov[ic];
   const double eni = p.use_en ? en[ic] : 0.0;
   double fcut = 1.0;
   if (p.cut > 0.0) {
      const double ecut = exp(p.cut);
      fcut = ecut / (ecut + exp(cn_raw[ic]));
   }
   double dg[3] = {0.0, 0.0, 0.0};
   double dl[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   double* row = dcndr + (size_t)3 * nat * ic;  // dcndr(:, :, i)
   for (int j0 = 0; j0 < nat; j0 += CND_TJ) {
      __syncthreads();
      for (int t = threadIdx.x; t < CND_TJ; t += blockDim.x) {
         const int j = min(j0 + t, nat - 1);
         sj[0][t] = xyz[3 * j]; sj[1][t] = xyz[3 * j + 1]; sj[2][t] = xyz[3 * j + 2];
         sj[3][t] = rcov[j];
         sj[4][t] = p.use_en ? en[j] : 0.0;
      }
      __syncthreads();
      if (!on) continue;
      for (int t0 = 0; t0 < CND_TJ && j0 + t0 < nat; t0 += 32) {
         const int t = t0 + lane, j = j0 + t;
         double g0 = 0.0, g1 = 0.0, g2 = 0.0;
         if (j < nat) {
            const double rc = rci + sj[3][t];
            const double rcn = (p.norm_exp == 1.0) ? rc : pow(rc, p.norm_exp);
            const double den = p.use_en ? (sj[4][t] - eni) : 1.0;
            const double dx = xi - sj[0][t], dy = yi - sj[1][t], dz = zi - sj[2][t];
            const double cut2 = count_range2(p, rc, rcn);
            for (int itr = 0; itr < ntr; ++itr) {
               const double rx = dx - trans[3 * itr], ry = dy - trans[3 * itr + 1], rz = dz - trans[3 * itr + 2];
               const double r2 = rx * rx + ry * ry + rz * rz;
               if (r2 > cut2 || r2 < 1.0e-12) continue;
               const double rinv = eg::fast_rsqrt(r2);
               const double arg = p.kcn * (r2 * rinv - rc) / rcn;
               double g;
               eg::erf_gauss<6, true>(setab, fabs(arg), g);  // g = exp(-arg^2)
               const double dc = den * (-p.kcn / SQRTPI / rcn * g) * rinv;
               const double c0 = dc * rx, c1 = dc * ry, c2 = dc * rz;  // countd
               g0 += c0; g1 += c1; g2 += c2;
               dl[0] += c0 * rx; dl[1] += c0 * ry; dl[2] += c0 * rz;
               dl[3] += c1 * rx; dl[4] += c1 * ry; dl[5] += c1 * rz;
               dl[6] += c2 * rx; dl[7] += c2 * ry; dl[8] += c2 * rz;
            }
            if (j == i) g0 = g1 = g2 = 0.0;  // self images cancel in dcndr; the diagonal entry is written at the end
            dg[0] += g0; dg[1] += g1; dg[2] += g2;
         }
         // d cn_i / d r_j = -countd
         __syncwarp();
         wbuf[warp][3 * lane] = -g0 * fcut;
         wbuf[warp][3 * lane + 1] = -g1 * fcut;
         wbuf[warp][3 * lane + 2] = -g2 * fcut;
         __syncwarp();
         const size_t base = (size_t)3 * (j0 + t0);
         const int nval = 3 * min(32, nat - (j0 + t0));
#pragma unroll
         for (int c = 0; c < 3; ++c)
            if (lane + 32 * c < nval) row[base + lane + 32 * c] = wbuf[warp][lane + 32 * c];
      }
   }
   if (!on) return;
#pragma unroll
   for (int c = 0; c < 3; ++c) dg[c] = warp_sum(dg[c]);
#pragma unroll
   for (int c = 0; c < 9; ++c) dl[c] = warp_sum(dl[c]);
   __syncwarp();
   if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) row[3 * i + c] = dg[c] * fcut;
#pragma unroll
      for (int c = 0; c < 9; ++c) dcndL[(size_t)9 * i + c] = dl[c] * fcut;
   }
}

__global__ void cn_cut_kernel(int nat, double cut, const double* __restrict__ cn_raw, double* __restrict__ cn) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nat) return;
   const double c0 = cn_raw[i];
   cn[i] = (cut > 0.0) ? log(1.0 + exp(cut)) - log(1.0 + exp(cut - c0)) : c0;
}

void ncoord_device(mchrg_b200_context* ctx, const DevMol& mol, const double* rcov_a, const double* en_a,
                   const NcoordPar& par, int ntr, const double* trans, double* cn, double* dcndr, double* dcndL,
                   const double** cn_raw_out) {
   const int nat = mol.nat;
   NcDev p{par.kcn, par.norm_exp, par.cutoff * par.cutoff, par.cut, en_a != nullptr};
   double* cn_raw = ctx->alloc<double>(nat);
   if (cn_raw_out) *cn_raw_out = cn_raw;
   const int rows_per_cta = 8;
   const unsigned grid = (nat + rows_per_cta - 1) / rows_per_cta;
   // one large system on several GPUs: every rank counts its share of the rows, one all-reduce joins them
   // (every entry has exactly one non-zero contribution, so the sum is exact and identical on all ranks)
   int row0 = 0, row1 = nat;
   const bool shard = dist_rows_apply(ctx, nat) && ctx->active == ctx->stream;
   if (shard) {
      row_share(ctx, nat, &row0, &row1);
      MCHRG_CUDA(cudaMemsetAsync(cn_raw, 0, (size_t)nat * sizeof(double), ctx->active));
   }
   if (row1 > row0)
      MCHRG_LAUNCH(ctx, cn_rows_kernel, (unsigned)((row1 - row0 + rows_per_cta - 1) / rows_per_cta), 256, 0, nat, mol.xyz, rcov_a,
                   en_a, p, ntr, trans, cn_raw, ctx->erf_tab, row0, row1);
   if (shard) comm_allreduce_sum(ctx, cn_raw, (size_t)nat);
   if (dcndr && dcndL)
      MCHRG_LAUNCH(ctx, cn_derivs_kernel, grid, 256, 0, nat, mol.xyz, rcov_a, en_a, p, ntr, trans, cn_raw, dcndr, dcndL,
                   ctx->erf_tab);
   MCHRG_LAUNCH(ctx, cn_cut_kernel, (nat + 255) / 256, 256, 0, nat, par.cut, cn_raw, cn);
}

// ---------------------------------------------------------------------------------------------
// The dcndr / dcndL contractions of the gradient and the stress WITHOUT the (3, N, N) array (6.4 GB at
// N = 16384): with w_k = thalf_k q_k (type.F90:277-280 with dxdr of eeq.f90:172-177) and the log-cut factor f_k,
//    gradx_j  = sum_k w_k dcndr(:, j, k) = sum_{k != j} (w_j f_j + w_k f_k) G_jk ,
//    sig_cn   = sum_k w_k dcndL(:, :, k) = sum_j w_j f_j sum_k sum_t c_jk(t) (x) r_jk(t) ,
// G_jk = sum_t c_jk(t), c_jk(t) = countd(|r_jk(t)|) r_jk(t) / |r_jk(t)|, r_jk(t) = r_j - r_k - t
// (dcn_k/dr_j = -f_k G_kj = +f_k G_jk: the translation set of get_lattice_points is symmetric, t <-> -t).
// One warp per atom j; same pair arithmetic as cn_derivs_kernel.  cn_count%erf only (no EN weights).
// ---------------------------------------------------------------------------------------------
__global__ void cn_weight_kernel(int nat, double cut, const double* __restrict__ cn_raw, const double* __restrict__ thalf,
                                 const double* __restrict__ q, double* __restrict__ wf) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nat) return;
   double fcut = 1.0;
   if (cut > 0.0) {
      const double ecut = exp(cut);
      fcut = ecut / (ecut + exp(cn_raw[i]));
   }
   wf[i] = thalf[i] * q[i] * fcut;
}

__global__ void __launch_bounds__(256) cn_grad_rows_kernel(int nat, const double* __restrict__ xyz,
                                                           const double* __restrict__ rcov, NcDev p, int ntr,
                                                           const double* __restrict__ trans,
                                                           const double* __restrict__ wf, int jb, int je,
                                                           double* __restrict__ gradx, double* __restrict__ sig_cn,
                                                           const double2* __restrict__ etab, int row0, int row1) {
   MCHRG_LOAD_ERF_TABLE(etab);
   __shared__ double sred[8][9];
   __shared__ double sk[5][CN_TJ];  // x, y, z, rcov, wf of the tile's atoms k (shared by the CTA's eight rows)
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int j = row0 + blockIdx.x * 8 + warp;
   const bool on = j < row1;
   const int jc = on ? j : row1 - 1;
   double gs[3] = {0.0, 0.0, 0.0};
   double dl[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   const double xj = xyz[3 * jc], yj = xyz[3 * jc + 1], zj = xyz[3 * jc + 2];
   const double rcj = rcov[jc], wj = wf[jc];
   for (int k0 = 0; k0 < nat; k0 += CN_TJ) {
      __syncthreads();
      {
         const int t = threadIdx.x, k = min(k0 + t, nat - 1);
         sk[0][t] = xyz[3 * k]; sk[1][t] = xyz[3 * k + 1]; sk[2][t] = xyz[3 * k + 2];
         sk[3][t] = rcov[k];
         sk[4][t] = wf[k];
      }
      __syncthreads();
      if (!on) continue;
      const int tn = min(CN_TJ, nat - k0);
      for (int t = lane; t < tn; t += 32) {
         const int k = k0 + t;
         const double rc = rcj + sk[3][t];
         const double rcn = (p.norm_exp == 1.0) ? rc : pow(rc, p.norm_exp);
         const double dx = xj - sk[0][t], dy = yj - sk[1][t], dz = zj - sk[2][t];
         const double wjk = (k != j) ? wj + sk[4][t] : 0.0;  // self images cancel in dcndr
         const double cut2 = count_range2(p, rc, rcn);
         for (int itr = 0; itr < ntr; ++itr) {
            const double rx = dx - trans[3 * itr], ry = dy - trans[3 * itr + 1], rz = dz - trans[3 * itr + 2];
            const double r2 = rx * rx + ry * ry + rz * rz;
            if (r2 > cut2 || r2 < 1.0e-12) continue;
            const double rinv = eg::fast_rsqrt(r2);
            const double arg = p.kcn * (r2 * rinv - rc) / rcn;
            double g;
            eg::erf_gauss<6, true>(setab, fabs(arg), g);  // g = exp(-arg^2)
            const double dc = -p.kcn / SQRTPI / rcn * g * rinv;  // (the division only for pairs in range)
            const double c0 = dc * rx, c1 = dc * ry, c2 = dc * rz;  // countd
            gs[0] = fma(wjk, c0, gs[0]); gs[1] = fma(wjk, c1, gs[1]); gs[2] = fma(wjk, c2, gs[2]);
            dl[0] += c0 * rx; dl[1] += c0 * ry; dl[2] += c0 * rz;
            dl[3] += c1 * rx; dl[4] += c1 * ry; dl[5] += c1 * rz;
            dl[6] += c2 * rx; dl[7] += c2 * ry; dl[8] += c2 * rz;
         }
      }
   }
   if (on) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gs[c] = warp_sum(gs[c]);
#pragma unroll
      for (int c = 0; c < 9; ++c) dl[c] = warp_sum(dl[c]) * wj;
      if (lane == 0 && j >= jb && j < je) {
#pragma unroll
         for (int c = 0; c < 3; ++c) gradx[3 * (size_t)(j - jb) + c] = gs[c];
      }
   }
   if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 9; ++c) sred[warp][c] = dl[c];
   }
   __syncthreads();
   if (threadIdx.x < 9) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += sred[w][threadIdx.x];
      atomicAdd(sig_cn + threadIdx.x, acc);
   }
}

void cn_gradient_device(mchrg_b200_context* ctx, const DevMol& mol, const CnLazy& lz, const double* thalf, const double* q,
                        int jb, int je, double* gradx, double* sig_cn, int row0, int row1) {
   const int nat = mol.nat;
   if (je < 0) je = nat;
   if (row1 < 0) { row0 = 0; row1 = nat; }
   NcDev p{lz.par.kcn, lz.par.norm_exp, lz.par.cutoff * lz.par.cutoff, lz.par.cut, 0};
   double* wf = ctx->alloc<double>(nat);
   MCHRG_LAUNCH(ctx, cn_weight_kernel, (nat + 255) / 256, 256, 0, nat, lz.par.cut, lz.cn_raw, thalf, q, wf);
   MCHRG_CUDA(cudaMemsetAsync(sig_cn, 0, 9 * sizeof(double), ctx->active));
   if (row1 > row0)
      MCHRG_LAUNCH(ctx, cn_grad_rows_kernel, (unsigned)((row1 - row0 + 7) / 8), 256, 0, nat, mol.xyz, lz.rcov_a, p, lz.ntr,
                   lz.trans, wf, jb, je, gradx, sig_cn, ctx->erf_tab, row0, row1);
}

// ---------------------------------------------------------------------------------------------
// get_xvec (eeq.f90:127-150) and the dx factor of get_xvec_derivs (eeq.f90:175)
// ---------------------------------------------------------------------------------------------
__global__ void eeq_xvec_kernel(int nat, const double* __restrict__ chi, const double* __restrict__ kcnchi,
                                const double* __restrict__ cn, double charge, double* __restrict__ xvec,
                                double* __restrict__ thalf) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nat) {
      const double tmp = kcnchi[i] / sqrt(cn[i] + 1.0e-14);
      xvec[i] = -chi[i] + tmp * cn[i];
      if (thalf) thalf[i] = 0.5 * tmp;
   } else if (i == nat) {
      xvec[nat] = charge;
   }
}

void eeq_xvec(mchrg_b200_context* ctx, const DevMol& mol, const double* chi_a, const double* kcnchi_a,
              const double* cn, double* xvec, double* thalf) {
   MCHRG_LAUNCH(ctx, eeq_xvec_kernel, (mol.nat + 1 + 255) / 256, 256, 0, mol.nat, chi_a, kcnchi_a, cn, mol.charge,
                xvec, thalf);
}

// ---------------------------------------------------------------------------------------------
// get_amat_0d (eeq.f90:199-242): thread = row index (contiguous), 8 columns per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) eeq_amat0d_kernel(int nat, const double* __restrict__ xyz,
                                                         const double* __restrict__ rad, const double* __restrict__ eta,
                                                         double* __restrict__ out, int64_t ld, int border, int64_t npad,
                                                         const double2* __restrict__ etab, int own_panel, int nranks,
                                                         int rank) {
   // multi-GPU factorisation: only the block columns this rank owns (block-cyclic, own_panel columns wide)
   if (own_panel > 0 && (int)(((int64_t)blockIdx.y * 8) / own_panel) % nranks != rank) return;
   MCHRG_LOAD_ERF_TABLE(etab);
   const int64_t nrow = border ? nat + 1 : npad;
   const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= nrow) return;
   double xr = 0, yr = 0, zr = 0, rr = 0;
   if (r < nat) {
      xr = xyz[3 * r]; yr = xyz[3 * r + 1]; zr = xyz[3 * r + 2];
      rr = rad[r] * rad[r];
   }
   const int64_t c0 = (int64_t)blockIdx.y * 8;
#pragma unroll
   for (int cc = 0; cc < 8; ++cc) {
      const int64_t c = c0 + cc;
      if (c >= nrow) break;
      double v;
      if (r < nat && c < nat) {
         if (r == c) {
            v = eta[r] + SQRT2PI / rad[r];
         } else {
            const double dx = xyz[3 * c] - xr, dy = xyz[3 * c + 1] - yr, dz = xyz[3 * c + 2] - zr;
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(rr + rad[c] * rad[c]);
            double g;
            const double x = r2 * rinv * gam;  // far field: erf == 1 to the last bit beyond the table range
            v = (x < eg::XMAX) ? eg::erf_gauss<5, false>(setab, x, g) * rinv : rinv;  // erf(gamma r) / r
         }
      } else if (border) {
         v = (r == nat && c == nat) ? 0.0 : 1.0;
      } else {
         v = (r == c) ? 1.0 : 0.0;
      }
      out[r + c * ld] = v;
   }
}

void eeq_amat_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a, double* out,
                 int64_t ld, bool border, int64_t npad, int64_t own_panel) {
   const int64_t nrow = border ? mol.nat + 1 : npad;
   dim3 grid((unsigned)((nrow + 127) / 128), (unsigned)((nrow + 7) / 8));
   if (own_panel % 8) own_panel = 0;
   MCHRG_LAUNCH(ctx, eeq_amat0d_kernel, grid, 128, 0, mol.nat, mol.xyz, rad_a, eta_a, out, ld, border ? 1 : 0, npad,
                ctx->erf_tab, (int)own_panel, ctx->nranks(), ctx->rank());
}

// ---------------------------------------------------------------------------------------------
// Row pass over all pairs: (J q)_i, atrace_i, dadL_i   (eeq.f90:199-242, 372-427; type.F90:259)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eeq_pair_rows0d_kernel(int nat, const double* __restrict__ xyz,
                                                              const double* __restrict__ rad,
                                                              const double* __restrict__ eta,
                                                              const double* __restrict__ q, double* __restrict__ jq,
                                                              double* __restrict__ atrace, double* __restrict__ dadL,
                                                              const double2* __restrict__ etab, int row0, int row1) {
   MCHRG_LOAD_ERF_TABLE(etab);
   const int i = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (i >= row1) return;
   const double xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2];
   const double ri2 = rad[i] * rad[i];
   const bool want_d = (atrace != nullptr) || (dadL != nullptr);
   double sj = 0.0, at[3] = {0.0, 0.0, 0.0}, dl[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
   for (int j = lane; j < nat; j += 32) {
      if (j == i) continue;
      // vec = r_j - r_i as in eeq.f90:221,403 (every term is even or odd in vec, the sign is kept)
      const double vx = xyz[3 * j] - xi, vy = xyz[3 * j + 1] - yi, vz = xyz[3 * j + 2] - zi;
      const double r2 = vx * vx + vy * vy + vz * vz;
      const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(ri2 + rad[j] * rad[j]);
      // erf(x), exp(-x^2), x = gamma r; far field (x beyond the table range): erf == 1 to the last bit, exp(-x^2) < 5e-19
      const double x = r2 * rinv * gam;
      double gs = 0.0, ef = 1.0;
      if (x < eg::XMAX) ef = eg::erf_gauss<6, true>(setab, x, gs);
      const double qj = q[j];
      sj = fma(ef * rinv, qj, sj);
      if (want_d) {
         const double dtmp = (2.0 / SQRTPI * gam * gs - ef * rinv) * rinv * rinv;
         // atrace(:, i) = sum_j dtmp (r_i - r_j) q_j = -dG q_j with dG = dtmp * vec
         const double gx = dtmp * vx, gy = dtmp * vy, gz = dtmp * vz;
         at[0] -= gx * qj; at[1] -= gy * qj; at[2] -= gz * qj;
         // dadL(a, b, i) += dS(a, b) q_j,  dS(a, b) = vec(a) dG(b)
         dl[0] += vx * gx * qj; dl[1] += vy * gx * qj; dl[2] += vz * gx * qj;
         dl[3] += vx * gy * qj; dl[4] += vy * gy * qj; dl[5] += vz * gy * qj;
         dl[6] += vx * gz * qj; dl[7] += vy * gz * qj; dl[8] += vz * gz * qj;
      }
   }
   sj = warp_sum(sj);
   if (want_d) {
#pragma unroll
      for (int c = 0; c < 3; ++c) at[c] = warp_sum(at[c]);
#pragma unroll
      for (int c = 0; c < 9; ++c) dl[c] = warp_sum(dl[c]);
   }
   if (lane == 0) {
      if (jq) jq[i] = sj + (eta[i] + SQRT2PI / rad[i]) * q[i];
      if (atrace) {
#pragma unroll
         for (int c = 0; c < 3; ++c) atrace[3 * i + c] = at[c];
      }
      if (dadL) {
#pragma unroll
         for (int c = 0; c < 9; ++c) dadL[(size_t)9 * i + c] = dl[c];
      }
   }
}

void eeq_pair_rows_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* eta_a,
                      const double* q, double* jq, double* atrace, double* dadL, int row0, int row1) {
   if (row1 < 0) { row0 = 0; row1 = mol.nat; }
   if (row1 <= row0) return;
   const unsigned grid = (row1 - row0 + 7) / 8;
   MCHRG_LAUNCH(ctx, eeq_pair_rows0d_kernel, grid, 256, 0, mol.nat, mol.xyz, rad_a, eta_a, q, jq, atrace, dadL, ctx->erf_tab,
                row0, row1);
}

// ---------------------------------------------------------------------------------------------
// B = dadr (+ atrace on the diagonal) - dxdr, fused with the dcndr-contractions of the gradient
// and the rank-one Schur vector (see eeq.cuh).  Thread = atom j (3 contiguous doubles of a column),
// CTA = 128 atoms j x KCHUNK columns k.
// ---------------------------------------------------------------------------------------------
constexpr int BB_KCHUNK = 64;

__global__ void __launch_bounds__(128) eeq_build_b0d_kernel(int nat, const double* __restrict__ xyz,
                                                            const double* __restrict__ rad,
                                                            const double* __restrict__ q,
                                                            const double* __restrict__ atrace,
                                                            const double* __restrict__ thalf,
                                                            const double* __restrict__ dcndr, double* __restrict__ Bout,
                                                            int64_t ldb, double* __restrict__ gradx,
                                                            const double* __restrict__ zvec, double* __restrict__ bz,
                                                            int pair_in_b, int jb, int je, const double2* __restrict__ etab) {
   __shared__ double2 setab[eg::NTAB];
   eg::load_table(setab, etab, (int)threadIdx.x, (int)blockDim.x);
   __shared__ double sk[BB_KCHUNK][8];  // x, y, z, rad^2, thalf, thalf*q, z, unused
   // atom j of the row block [jb, je); rows of Bout / gradx / bz are numbered within the block
   const int jl = blockIdx.x * blockDim.x + threadIdx.x;
   const int j = jb + jl;
   const int k0 = blockIdx.y * BB_KCHUNK;
   const int kn = min(BB_KCHUNK, nat - k0);
   for (int idx = threadIdx.x; idx < kn; idx += blockDim.x) {
      const int k = k0 + idx;
      sk[idx][0] = xyz[3 * k]; sk[idx][1] = xyz[3 * k + 1]; sk[idx][2] = xyz[3 * k + 2];
      sk[idx][3] = rad[k] * rad[k];
      const double th = thalf ? thalf[k] : 0.0;
      sk[idx][4] = th;
      sk[idx][5] = th * q[k];
      sk[idx][6] = zvec ? zvec[k] : 0.0;
   }
   __syncthreads();
   if (j >= je) return;
   const double xj = xyz[3 * j], yj = xyz[3 * j + 1], zj = xyz[3 * j + 2];
   const double rj2 = rad[j] * rad[j], qj = q[j];
   double gx[3] = {0.0, 0.0, 0.0}, uz[3] = {0.0, 0.0, 0.0};
   const bool need_pair = (Bout != nullptr) || (bz != nullptr);
   for (int kk = 0; kk < kn; ++kk) {
      const int k = k0 + kk;
      double b0, b1, b2;
      if (k == j) {
         b0 = b1 = b2 = 0.0;
         if (atrace) { b0 = atrace[3 * j]; b1 = atrace[3 * j + 1]; b2 = atrace[3 * j + 2]; }
      } else if (pair_in_b) {  // periodic: the pair term was written by the Ewald kernels
         const double* o = Bout + (size_t)k * ldb + 3 * jl;
         b0 = o[0]; b1 = o[1]; b2 = o[2];
      } else if (!need_pair) {
         b0 = b1 = b2 = 0.0;
      } else {
         // dadr(:, j, k) = dtmp (r_j - r_k) q_j      (eeq.f90:412-413 for both orderings of the pair)
         const double vx = xj - sk[kk][0], vy = yj - sk[kk][1], vz = zj - sk[kk][2];
         const double r2 = vx * vx + vy * vy + vz * vz;
         const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(rj2 + sk[kk][3]);
         const double x = r2 * rinv * gam;
         double gs = 0.0, ef = 1.0;
         if (x < eg::XMAX) ef = eg::erf_gauss<6, true>(setab, x, gs);  // far field: erf == 1, exp(-x^2) < 5e-19
         const double dtmp = (2.0 / SQRTPI * gam * gs - ef * rinv) * rinv * rinv * qj;
         b0 = dtmp * vx; b1 = dtmp * vy; b2 = dtmp * vz;
      }
      if (dcndr) {
         const double* d = dcndr + ((size_t)nat * k + j) * 3;
         const double d0 = d[0], d1 = d[1], d2 = d[2];
         const double th = sk[kk][4], tq = sk[kk][5];
         b0 = fma(-th, d0, b0); b1 = fma(-th, d1, b1); b2 = fma(-th, d2, b2);
         gx[0] = fma(tq, d0, gx[0]); gx[1] = fma(tq, d1, gx[1]); gx[2] = fma(tq, d2, gx[2]);
      }
      if (Bout) {
         double* o = Bout + (size_t)k * ldb + 3 * jl;
         o[0] = b0; o[1] = b1; o[2] = b2;
      }
      if (bz) {
         const double zk = sk[kk][6];
         uz[0] = fma(b0, zk, uz[0]); uz[1] = fma(b1, zk, uz[1]); uz[2] = fma(b2, zk, uz[2]);
      }
   }
   if (gradx && dcndr) {
      atomicAdd(gradx + 3 * jl, gx[0]); atomicAdd(gradx + 3 * jl + 1, gx[1]); atomicAdd(gradx + 3 * jl + 2, gx[2]);
   }
   if (bz) {
      atomicAdd(bz + 3 * jl, uz[0]); atomicAdd(bz + 3 * jl + 1, uz[1]); atomicAdd(bz + 3 * jl + 2, uz[2]);
   }
}

void eeq_build_b_0d(mchrg_b200_context* ctx, const DevMol& mol, const double* rad_a, const double* q,
                    const double* atrace, const double* thalf, const double* dcndr, double* Bout, int64_t ldb,
                    double* gradx, const double* zvec, double* bz, bool pair_in_b, int jb, int je) {
   if (je < 0) je = mol.nat;
   dim3 grid((unsigned)((je - jb + 127) / 128), (unsigned)((mol.nat + BB_KCHUNK - 1) / BB_KCHUNK));
   MCHRG_LAUNCH(ctx, eeq_build_b0d_kernel, grid, 128, 0, mol.nat, mol.xyz, rad_a, q, atrace, thalf, dcndr, Bout, ldb,
                gradx, zvec, bz, pair_in_b ? 1 : 0, jb, je, ctx->erf_tab);
}

}  // namespace mchrg
