// Batches of independent non-periodic molecules (BASELINE.json configs[0..1]): ONE CTA PER MOLECULE,
// the whole EEQ2019 single point (coordination number -> Coulomb block -> Cholesky + Schur
// complement -> energy / gradient) inside one launch with the matrix resident in shared memory.
//
//   shared memory: packed lower triangle of J (row-major, order padded to a multiple of 16),
//                  a 16 x (NP+4) k-major panel buffer, the 16x16 diagonal block, 13 per-atom vectors
//   factorisation: right-looking blocked Cholesky, block 16: diagonal block in one warp (register
//                  rows + shuffles), panel rows one thread each, trailing update on FP64 tensor
//                  cores (mma.sync m8n8k4 DMMA, 16x16 tiles per warp, operands from the panel buffer)
//   reference path replaced: app/main.f90:114-118 / charge.f90:37-77 per molecule, i.e.
//                  mctc_ncoord (erf CN, log-cut), eeq.f90 get_xvec / get_amat_0d / get_damat_0d,
//                  type.F90 solve (sytrf/sytrs -> Cholesky + Schur), energy and gradient.
// Molecules are grouped by padded order NP (one launch per class) so shared memory, and with it the
// number of resident CTAs per SM, matches the molecule size.
#include <algorithm>
#include <vector>

#include "model.cuh"

namespace mchrg {
namespace batch {

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737

constexpr int KB = 16;        // Cholesky block
constexpr int NP_MAX = 208;   // largest padded order that fits in 227 KB of shared memory
constexpr int NVEC = 13;

struct Args {
   const int64_t* offsets;
   const int32_t* num;
   const double* xyz;
   const double* charge;
   double* qvec;
   double* energy;
   double* gradient;
   int32_t* status;
   const int32_t* perm;  // molecule ids of this class
   const double* par;    // model parameters [slot * nid + (Z-1)]
   int nid;
   double kcn, cutoff2, cut;
   int np;               // padded order of the class (multiple of 16)
};

__host__ __device__ inline size_t smem_doubles(int np) {
   return (size_t)np * (np + 1) / 2 + (size_t)KB * (np + 4) + KB * (KB + 1) + KB + (size_t)NVEC * np + 8;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// register-only broadcast of a double (the library overload goes through memory and forces arrays that are
// shuffled element-wise into local memory)
__device__ __forceinline__ double shfl_d(double v, int src) {
   const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
   const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
   return __hiloint2double(hi, lo);
}

// decode a packed-lower index e -> (i, j), j <= i
__device__ __forceinline__ void tri_decode(int e, int& i, int& j) {
   i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);  // e < 2^15: the float estimate is within 1
   while (tri(i + 1) <= e) ++i;
   while (tri(i) > e) --i;
   j = e - tri(i);
}

// Cholesky of the 16 x 16 diagonal block at k0 by one warp, register resident: lane = (row r = lane & 15,
// column parity h = lane >> 4) holds the 8 entries m[k] = T(r, 2k + h).  Per column c the scaled column is
// broadcast with shuffles from the lanes of parity c & 1 (register c >> 1) and every lane applies the
// rank-one update to its 8 entries independently (no serial chain inside a column step).
// Publishes L into A and Ld (row stride 17) and the reciprocal pivots rd[k0 + c].
// Returns 0 or the 1-based index of the first non-positive pivot.
__device__ __forceinline__ int diag_block(double* __restrict__ A, double* __restrict__ Ld, double* __restrict__ rd, int k0,
                                          int lane) {
   const int r = lane & 15, h = lane >> 4;
   double m[8];
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      const int c = 2 * k + h;
      m[k] = (c <= r) ? A[tri(k0 + r) + k0 + c] : 0.0;
   }
   int bad = 0;
#pragma unroll
   for (int c = 0; c < KB; ++c) {
      const int src_h = (c & 1) << 4;  // lanes holding column c
      const double piv = shfl_d(m[c >> 1], c + src_h);
      bad = (!(piv > 0.0) && bad == 0) ? (k0 + c + 1) : bad;
      const double rinv = rsqrt(piv);
      const double lrc = shfl_d(m[c >> 1], r + src_h) * rinv;  // L(r, c) for r > c
#pragma unroll
      for (int k = 0; k < 8; ++k) {
         const int c2 = 2 * k + h;
         const double lc2 = shfl_d(m[c >> 1], (c2 & 15) + src_h) * rinv;  // L(c2, c)
         if (c2 > c && c2 <= r) m[k] = fma(-lrc, lc2, m[k]);
      }
      if (h == (c & 1)) {  // owner lanes: store the scaled column entry / the pivot
         if (r == c) m[c >> 1] = piv * rinv;
         else if (r > c) m[c >> 1] = lrc;
      }
      if (lane == 0) rd[k0 + c] = rinv;
   }
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      const int c = 2 * k + h;
      if (c <= r) A[tri(k0 + r) + k0 + c] = m[k];
      Ld[r * (KB + 1) + c] = (c <= r) ? m[k] : 0.0;
   }
   return bad;
}

// one 16 x 16 tile of the trailing update A(i0.., j0..) -= P(:, i0..)^T P(:, j0..) on DMMA
__device__ __forceinline__ void update_tile(double* __restrict__ A, const double* __restrict__ P, int ldp, int i0, int j0,
                                            int lane) {
   const int g = lane >> 2, t = lane & 3;
   double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
   for (int kk = 0; kk < KB / 4; ++kk) {
      const double* pk = P + (4 * kk + t) * ldp;
      const double a0 = pk[i0 + g], a1 = pk[i0 + 8 + g];
      const double c0 = pk[j0 + g], c1 = pk[j0 + 8 + g];
      dmma884(acc[0][0][0], acc[0][0][1], a0, c0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, c1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, c0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, c1);
   }
#pragma unroll
   for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ntl = 0; ntl < 2; ++ntl) {
         const int row = i0 + 8 * mt + g;
         const int col = j0 + 8 * ntl + 2 * t;
         double* dst = A + tri(row) + col;
         if (col <= row) dst[0] -= acc[mt][ntl][0];
         if (col + 1 <= row) dst[1] -= acc[mt][ntl][1];
      }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) eeq_batch_kernel(Args p) {
   extern __shared__ __align__(16) double sm[];
   const int np = p.np;
   const int ldp = np + 4;
   double* A = sm;                           // packed lower, row-major: A[tri(i) + j], j <= i
   double* P = A + (size_t)np * (np + 1) / 2;  // panel, k-major: P[c * ldp + i]
   double* Ld = P + (size_t)KB * ldp;        // diagonal block Ld[r * 17 + c]
   double* spare = Ld + KB * (KB + 1);       // KB doubles, unused
   double* vec = spare + KB;
   double* X = vec;            double* Y = X + np;        double* Z = Y + np;
   double* rad2 = Z + np;      double* dself = rad2 + np; double* rcov = dself + np;
   double* fcut = rcov + np;   double* xv = fcut + np;    double* th = xv + np;
   double* b0 = th + np;       double* b1 = b0 + np;      double* qv = b1 + np;
   double* rd = qv + np;       // reciprocal pivots 1 / L(i, i); later reused for thalf * q * fcut
   double* scal = rd + np;     // [0] lambda, [1] s
   __shared__ int sfail;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = NT / 32;
   const int mol = p.perm[blockIdx.x];
   const int64_t off = p.offsets[mol];
   const int n = (int)(p.offsets[mol + 1] - off);
   const double charge = p.charge[mol];

   // ---- load atoms and per-element parameters (chi, rad, eta, kcnchi, rcov = slots 0..4) ----
   if (tid == 0) sfail = 0;
   for (int i = tid; i < np; i += NT) {
      if (i < n) {
         const int z = p.num[off + i] - 1;
         X[i] = p.xyz[3 * (off + i)];
         Y[i] = p.xyz[3 * (off + i) + 1];
         Z[i] = p.xyz[3 * (off + i) + 2];
         const double rad = p.par[1 * p.nid + z];
         rad2[i] = rad * rad;
         dself[i] = p.par[2 * p.nid + z] + SQRT2PI / rad;  // eta + sqrt(2/pi)/rad  (eeq.f90:228)
         rcov[i] = p.par[4 * p.nid + z];
         xv[i] = p.par[0 * p.nid + z];   // chi, turned into xvec below
         th[i] = p.par[3 * p.nid + z];   // kcnchi, turned into thalf below
      } else {
         X[i] = Y[i] = Z[i] = 0.0;
         rad2[i] = 1.0; dself[i] = 1.0; rcov[i] = 0.0; xv[i] = 0.0; th[i] = 0.0;
      }
   }
   __syncthreads();

   // ---- erf coordination number with log-cut (mctc_ncoord): every unique pair once, the counts are
   //      parked in the (still unused) matrix storage and summed per row afterwards ----
   const int ntri_n = tri(n);
   for (int e = tid; e < ntri_n; e += NT) {
      int i, j;
      tri_decode(e, i, j);
      double c = 0.0;
      if (i != j) {
         const double dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
         const double r2 = dx * dx + dy * dy + dz * dz;
         if (!(r2 > p.cutoff2 || r2 < 1.0e-12)) {
            // 1/2 (1 + erf(-kcn (r - rc)/rc)) = 1/2 erfc(kcn (r/rc - 1))
            const double r1 = r2 * rsqrt(r2);
            c = 0.5 * erfc(p.kcn * (r1 / (rcov[i] + rcov[j]) - 1.0));
         }
      }
      A[e] = c;
   }
   __syncthreads();
   const double ecut = exp(p.cut);
   for (int i = warp; i < n; i += NW) {
      double acc = 0.0;
      const double* row = A + tri(i);
      for (int j = lane; j < i; j += 32) acc += row[j];
      for (int j = i + 1 + lane; j < n; j += 32) acc += A[tri(j) + i];
      acc = warp_sum(acc);
      if (lane == 0) {
         double cn = acc, f = 1.0;
         if (p.cut > 0.0) {
            f = ecut / (ecut + exp(acc));
            cn = log(1.0 + ecut) - log(1.0 + exp(p.cut - acc));
         }
         const double tmp = th[i] / sqrt(cn + 1.0e-14);
         xv[i] = -xv[i] + tmp * cn;  // get_xvec, eeq.f90:127-150
         th[i] = 0.5 * tmp;
         fcut[i] = f;
      }
   }
   __syncthreads();

   // ---- Coulomb block J, packed lower (eeq.f90:199-242); padding = identity ----
   const int ntri = tri(np);
   for (int e = tid; e < ntri; e += NT) {
      int i, j;
      tri_decode(e, i, j);
      double v;
      if (i == j) {
         v = dself[i];
      } else if (i < n) {
         const double dx = X[j] - X[i], dy = Y[j] - Y[i], dz = Z[j] - Z[i];
         const double r2 = dx * dx + dy * dy + dz * dz;
         const double rinv = rsqrt(r2), gam = rsqrt(rad2[i] + rad2[j]);
         v = erf(r2 * rinv * gam) * rinv;  // erf(gamma r) / r
      } else {
         v = 0.0;
      }
      A[e] = v;
   }
   __syncthreads();

   // ---- blocked Cholesky, block 16, with look-ahead: the next diagonal block is factored by warp 0
   //      while the other warps finish the trailing update ----
   if (warp == 0) {
      const int bad = diag_block(A, Ld, rd, 0, lane);
      if (bad && lane == 0) sfail = bad;
   }
   __syncthreads();
   for (int k0 = 0; k0 + KB < np && !sfail; k0 += KB) {
      const int t0 = k0 + KB;
      // panel rows: one thread per row, a <- a * inv(Ld)^T (right-looking inside the row)
      for (int i = t0 + tid; i < np; i += NT) {
         double a[KB];
         double* row = A + tri(i) + k0;
#pragma unroll
         for (int c = 0; c < KB; ++c) a[c] = row[c];
#pragma unroll
         for (int c = 0; c < KB; ++c) {
            a[c] *= rd[k0 + c];
#pragma unroll
            for (int c2 = c + 1; c2 < KB; ++c2) a[c2] = fma(-a[c], Ld[c2 * (KB + 1) + c], a[c2]);
         }
#pragma unroll
         for (int c = 0; c < KB; ++c) {
            row[c] = a[c];
            P[c * ldp + i] = a[c];
         }
      }
      __syncthreads();
      const int nt = (np - t0) / KB;
      // first block column of the trailing matrix (needed by the next diagonal block and panel)
      for (int ti = warp; ti < nt; ti += NW) update_tile(A, P, ldp, t0 + KB * ti, t0, lane);
      __syncthreads();
      // warp 0: next diagonal block; other warps: the rest of the trailing update
      if (warp == 0) {
         const int bad = diag_block(A, Ld, rd, t0, lane);
         if (bad && lane == 0) sfail = bad;
         if (NW == 1) {
            const int nrest = tri(nt - 1);
            for (int tile = 0; tile < nrest; ++tile) {
               int a, b;
               tri_decode(tile, a, b);
               update_tile(A, P, ldp, t0 + KB * (a + 1), t0 + KB * (b + 1), lane);
            }
         }
      } else {
         const int nrest = tri(nt - 1);  // tiles (ti, tj) with 1 <= tj <= ti <= nt - 1
         for (int tile = warp - 1; tile < nrest; tile += NW - 1) {
            int a, b;
            tri_decode(tile, a, b);
            update_tile(A, P, ldp, t0 + KB * (a + 1), t0 + KB * (b + 1), lane);
         }
      }
      __syncthreads();
   }
   if (sfail) {  // not positive definite: report the pivot, poison the outputs
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = tid; i < n; i += NT) {
         p.qvec[off + i] = nan;
         if (p.energy) p.energy[off + i] = nan;
         if (p.gradient) p.gradient[3 * (off + i)] = p.gradient[3 * (off + i) + 1] = p.gradient[3 * (off + i) + 2] = nan;
      }
      if (tid == 0) p.status[mol] = sfail;
      return;
   }

   // ---- (L L^T) [y z] = [x 1] ----
   for (int i = tid; i < np; i += NT) {
      b0[i] = (i < n) ? xv[i] : 0.0;
      b1[i] = (i < n) ? 1.0 : 0.0;
   }
   __syncthreads();
   for (int k0 = 0; k0 < np; k0 += KB) {  // forward
      if (warp == 0) {
         const int r = lane & 15;
         double v0 = b0[k0 + r], v1 = b1[k0 + r];
#pragma unroll
         for (int c = 0; c < KB; ++c) {
            const double dinv = rd[k0 + c];
            const double w0 = shfl_d(v0, c) * dinv;
            const double w1 = shfl_d(v1, c) * dinv;
            if (r == c) { v0 = w0; v1 = w1; }
            else if (r > c) {
               const double l = A[tri(k0 + r) + k0 + c];
               v0 = fma(-l, w0, v0);
               v1 = fma(-l, w1, v1);
            }
         }
         if (lane < KB) { b0[k0 + r] = v0; b1[k0 + r] = v1; }
      }
      __syncthreads();
      for (int i = k0 + KB + tid; i < np; i += NT) {
         const double* row = A + tri(i) + k0;
         double s0 = b0[i], s1 = b1[i];
#pragma unroll
         for (int c = 0; c < KB; ++c) {
            s0 = fma(-row[c], b0[k0 + c], s0);
            s1 = fma(-row[c], b1[k0 + c], s1);
         }
         b0[i] = s0; b1[i] = s1;
      }
      __syncthreads();
   }
   for (int k0 = np - KB; k0 >= 0; k0 -= KB) {  // backward
      if (warp == 0) {
         const int r = lane & 15;
         double v0 = b0[k0 + r], v1 = b1[k0 + r];
#pragma unroll
         for (int c = KB - 1; c >= 0; --c) {
            const double dinv = rd[k0 + c];
            const double w0 = shfl_d(v0, c) * dinv;
            const double w1 = shfl_d(v1, c) * dinv;
            if (r == c) { v0 = w0; v1 = w1; }
            else if (r < c) {
               const double l = A[tri(k0 + c) + k0 + r];
               v0 = fma(-l, w0, v0);
               v1 = fma(-l, w1, v1);
            }
         }
         if (lane < KB) { b0[k0 + r] = v0; b1[k0 + r] = v1; }
      }
      __syncthreads();
      for (int j = tid; j < k0; j += NT) {
         double s0 = b0[j], s1 = b1[j];
#pragma unroll
         for (int c = 0; c < KB; ++c) {
            const double l = A[tri(k0 + c) + j];
            s0 = fma(-l, b0[k0 + c], s0);
            s1 = fma(-l, b1[k0 + c], s1);
         }
         b0[j] = s0; b1[j] = s1;
      }
      __syncthreads();
   }
   // Schur complement of the charge constraint: lambda = (1^T y - Q) / (1^T z), q = y - lambda z
   if (warp == 0) {
      double sy = 0.0, sz = 0.0;
      for (int i = lane; i < n; i += 32) { sy += b0[i]; sz += b1[i]; }
      sy = warp_sum(sy);
      sz = warp_sum(sz);
      if (lane == 0) { scal[0] = (sy - charge) / sz; scal[1] = sz; }
   }
   __syncthreads();
   const double lam = scal[0];
   double* wf = rd;  // pivots are no longer needed
   for (int i = tid; i < n; i += NT) {
      const double q = b0[i] - lam * b1[i];
      qv[i] = q;
      wf[i] = th[i] * q * fcut[i];
      p.qvec[off + i] = q;
      // energy (type.F90:255-262): q_i (0.5 (J q)_i - x_i) with (J q)_i = x_i - lambda from the solved system
      // J q + lambda 1 = x (residual of the Cholesky solve ~ n eps |J| |q|)
      if (p.energy) p.energy[off + i] = q * (0.5 * (xv[i] - lam) - xv[i]);
   }
   if (tid == 0) p.status[mol] = 0;
   if (!p.gradient) return;
   __syncthreads();

   // ---- gradient (type.F90:275-278): g_i = sum_j f_ij (r_i - r_j) with the symmetric pair scalar
   //        f_ij = dtmp_ij q_i q_j - dcount_ij / r_ij (w_i + w_j),   w = thalf * q * fcut
   //      (q_i atrace_i minus the dcndr contraction; eeq.f90:403-411, mctc_ncoord); one evaluation per
   //      unique pair, parked in the matrix storage, then summed per row ----
   for (int e = tid; e < ntri_n; e += NT) {
      int i, j;
      tri_decode(e, i, j);
      double f = 0.0;
      if (i != j) {
         const double vx = X[i] - X[j], vy = Y[i] - Y[j], vz = Z[i] - Z[j];
         const double r2 = vx * vx + vy * vy + vz * vz;
         const double rinv = rsqrt(r2), gam = rsqrt(rad2[i] + rad2[j]);
         const double r1 = r2 * rinv, x = r1 * gam;  // x = gamma r
         // dtmp = 2 gamma exp(-x^2) / (sqrt(pi) r^2) - erf(x) / r^3
         f = rinv * rinv * (2.0 / SQRTPI * gam * exp(-x * x) - erf(x) * rinv) * qv[i] * qv[j];
         if (r2 <= p.cutoff2 && r2 >= 1.0e-12) {
            const double rci = 1.0 / (rcov[i] + rcov[j]);
            const double a = p.kcn * (r1 * rci - 1.0);
            f += p.kcn / SQRTPI * rci * exp(-a * a) * rinv * (wf[i] + wf[j]);
         }
      }
      A[e] = f;
   }
   __syncthreads();
   for (int i = warp; i < n; i += NW) {
      const double xi = X[i], yi = Y[i], zi = Z[i];
      double gx = 0.0, gy = 0.0, gz = 0.0;
      const double* row = A + tri(i);
      for (int j = lane; j < n; j += 32) {
         if (j == i) continue;
         const double f = (j < i) ? row[j] : A[tri(j) + i];
         gx = fma(f, xi - X[j], gx); gy = fma(f, yi - Y[j], gy); gz = fma(f, zi - Z[j], gz);
      }
      gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
      if (lane == 0) {
         p.gradient[3 * (off + i)] = gx;
         p.gradient[3 * (off + i) + 1] = gy;
         p.gradient[3 * (off + i) + 2] = gz;
      }
   }
}

template <int NT>
static void launch_class(mchrg_b200_context* ctx, const Args& a, int count) {
   const size_t smem = smem_doubles(a.np) * sizeof(double);
   static size_t configured = 0;
   if (smem > configured) {
      MCHRG_CUDA(cudaFuncSetAttribute(eeq_batch_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(smem_doubles(NP_MAX) * sizeof(double))));
      configured = smem_doubles(NP_MAX) * sizeof(double);
   }
   MCHRG_LAUNCH(ctx, eeq_batch_kernel<NT>, (unsigned)count, NT, smem, a);
}



// bucket the molecules [m0, m1) by padded order; returns the class-sorted permutation (chunk-local ids)
// and the start of every class in it
static void bucket(const std::vector<int64_t>& hoff, int32_t m0, int32_t m1, std::vector<int32_t>& perm,
                   std::vector<size_t>& cls_start, std::vector<int>& cls_count) {
   constexpr int NCLASS = NP_MAX / KB;
   std::vector<std::vector<int32_t>> cls(NCLASS + 1);
   for (int32_t m = m0; m < m1; ++m) {
      const int64_t n = hoff[m + 1] - hoff[m];
      cls[(n + KB - 1) / KB].push_back(m - m0);
   }
   perm.clear();
   cls_start.assign(NCLASS + 1, 0);
   cls_count.assign(NCLASS + 1, 0);
   for (int c = 1; c <= NCLASS; ++c) {
      cls_start[c] = perm.size();
      cls_count[c] = (int)cls[c].size();
      perm.insert(perm.end(), cls[c].begin(), cls[c].end());
   }
}

// Launch every size class of one chunk: the shared-memory-heavy classes (one CTA per SM, latency bound in the
// factorisation) on `big`, the light classes on `small`; their CTAs fit beside each other on an SM, so the
// light ones fill the idle FP64 issue slots.  Largest classes first.
static void launch_chunk(mchrg_b200_context* ctx, Args a, const int32_t* dperm, const std::vector<size_t>& cls_start,
                         const std::vector<int>& cls_count, cudaStream_t big, cudaStream_t small) {
   for (int c = NP_MAX / KB; c >= 1; --c) {
      if (cls_count[c] == 0) continue;
      a.np = c * KB;
      a.perm = dperm + cls_start[c];
      StreamScope on(ctx, a.np > 112 ? big : small);
      if (a.np <= 64) launch_class<128>(ctx, a, cls_count[c]);
      else if (a.np <= 128) launch_class<256>(ctx, a, cls_count[c]);
      else launch_class<512>(ctx, a, cls_count[c]);
   }
}

}  // namespace batch
}  // namespace mchrg

using namespace mchrg;

extern "C" int mchrg_b200_singlepoint_batch(const mchrg_b200_model* model, int32_t nmol, const int64_t* offsets,
                                            const int32_t* num, const double* xyz, const double* charge, double* qvec,
                                            double* energy, double* gradient, int32_t* status) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   return guarded(ctx, [&]() {
      if (nmol <= 0 || !offsets || !num || !xyz || !charge || !qvec || !status)
         fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: missing argument");
      if (model->kind != 1) fail(MCHRG_B200_ERR_MODEL, "singlepoint_batch serves EEQ2019-type models only");
      // molecule sizes are needed on the host to form the size classes
      std::vector<int64_t> hoff((size_t)nmol + 1);
      if (mchrg_b200_context::is_device_ptr(offsets))
         MCHRG_CUDA(cudaMemcpy(hoff.data(), offsets, hoff.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
      else
         std::copy(offsets, offsets + nmol + 1, hoff.begin());
      const size_t ntot = (size_t)hoff[nmol];
      for (int32_t m = 0; m < nmol; ++m) {
         const int64_t n = hoff[m + 1] - hoff[m];
         if (n <= 0) fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: molecule %d is empty", m);
         if (n > batch::NP_MAX)
            fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: molecule %d has %lld atoms (> %d); use mchrg_b200_singlepoint", m,
                 (long long)n, batch::NP_MAX);
      }
      batch::Args a;
      a.par = model->d_par;
      a.nid = model->nid;
      a.kcn = model->ncoord.kcn;
      a.cutoff2 = model->ncoord.cutoff * model->ncoord.cutoff;
      a.cut = model->ncoord.cut;
      std::vector<int32_t> perm;
      std::vector<size_t> cls_start;
      std::vector<int> cls_count;

      const bool host_io = !mchrg_b200_context::is_device_ptr(xyz) && !mchrg_b200_context::is_device_ptr(num) &&
                           !mchrg_b200_context::is_device_ptr(charge) && !mchrg_b200_context::is_device_ptr(offsets) &&
                           !mchrg_b200_context::is_device_ptr(qvec) && !mchrg_b200_context::is_device_ptr(status) &&
                           (!energy || !mchrg_b200_context::is_device_ptr(energy)) &&
                           (!gradient || !mchrg_b200_context::is_device_ptr(gradient));
      if (!host_io || ntot < (size_t)1 << 18) {
         // device-resident (or small) batch: one chunk, inputs staged on the main stream
         batch::bucket(hoff, 0, nmol, perm, cls_start, cls_count);
         int32_t* dperm = ctx->alloc<int32_t>(nmol);
         MCHRG_CUDA(cudaMemcpyAsync(dperm, perm.data(), perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
         a.offsets = ctx->in(offsets, (size_t)nmol + 1);
         a.num = ctx->in(num, ntot);
         a.xyz = ctx->in(xyz, 3 * ntot);
         a.charge = ctx->in(charge, (size_t)nmol);
         a.qvec = ctx->out(qvec, ntot);
         a.energy = ctx->out(energy, ntot);
         a.gradient = ctx->out(gradient, 3 * ntot);
         a.status = ctx->out(status, (size_t)nmol);
         cudaEvent_t staged = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(staged, ctx->stream));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_aux, staged, 0));
         batch::launch_chunk(ctx, a, dperm, cls_start, cls_count, ctx->stream, ctx->stream_aux);
         cudaEvent_t aux_done = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(aux_done, ctx->stream_aux));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, aux_done, 0));
         MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));  // perm (host vector) must outlive the async copy
         return 0;
      }

      // Host buffers: pipeline chunks of molecules so that the H2D copy of chunk c+1 and the D2H copy of
      // chunk c-1 overlap the kernels of chunk c (asynchronous when the caller's buffers are pinned).
      const int nchunk = (int)std::min<size_t>(8, std::max<size_t>(2, ntot >> 21));
      std::vector<int32_t> bounds(nchunk + 1, 0);
      for (int c = 1; c < nchunk; ++c) {  // split by atoms
         const int64_t target = (int64_t)(ntot * (size_t)c / nchunk);
         bounds[c] = (int32_t)(std::lower_bound(hoff.begin(), hoff.end(), target) - hoff.begin());
         if (bounds[c] < bounds[c - 1]) bounds[c] = bounds[c - 1];
      }
      bounds[nchunk] = nmol;
      std::vector<std::vector<int64_t>> loc_off(nchunk);
      std::vector<std::vector<int32_t>> loc_perm(nchunk);
      cudaEvent_t start = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(start, ctx->stream));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_in, start, 0));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_out, start, 0));
      for (int c = 0; c < nchunk; ++c) {
         const int32_t m0 = bounds[c], m1 = bounds[c + 1], mc = m1 - m0;
         if (mc <= 0) continue;
         const size_t a0 = (size_t)hoff[m0], na = (size_t)(hoff[m1] - hoff[m0]);
         loc_off[c].resize((size_t)mc + 1);
         for (int32_t m = 0; m <= mc; ++m) loc_off[c][m] = hoff[m0 + m] - hoff[m0];
         batch::bucket(hoff, m0, m1, loc_perm[c], cls_start, cls_count);
         int64_t* d_off = ctx->alloc<int64_t>((size_t)mc + 1);
         int32_t* d_perm = ctx->alloc<int32_t>(mc);
         int32_t* d_num = ctx->alloc<int32_t>(na);
         double* d_xyz = ctx->alloc<double>(3 * na);
         double* d_chg = ctx->alloc<double>(mc);
         double* d_q = ctx->alloc<double>(na);
         double* d_e = energy ? ctx->alloc<double>(na) : nullptr;
         double* d_g = gradient ? ctx->alloc<double>(3 * na) : nullptr;
         int32_t* d_st = ctx->alloc<int32_t>(mc);
         cudaStream_t in = ctx->stream_in, out = ctx->stream_out;
         MCHRG_CUDA(cudaMemcpyAsync(d_off, loc_off[c].data(), ((size_t)mc + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_perm, loc_perm[c].data(), (size_t)mc * sizeof(int32_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_num, num + a0, na * sizeof(int32_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_xyz, xyz + 3 * a0, 3 * na * sizeof(double), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_chg, charge + m0, (size_t)mc * sizeof(double), cudaMemcpyHostToDevice, in));
         cudaEvent_t ev_in = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(ev_in, in));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ev_in, 0));
         a.offsets = d_off; a.num = d_num; a.xyz = d_xyz; a.charge = d_chg;
         a.qvec = d_q; a.energy = d_e; a.gradient = d_g; a.status = d_st;
         batch::launch_chunk(ctx, a, d_perm, cls_start, cls_count, ctx->stream, ctx->stream_aux);
         cudaEvent_t aux_done = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(aux_done, ctx->stream_aux));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, aux_done, 0));
         cudaEvent_t ev_done = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(ev_done, ctx->stream));
         MCHRG_CUDA(cudaStreamWaitEvent(out, ev_done, 0));
         MCHRG_CUDA(cudaMemcpyAsync(qvec + a0, d_q, na * sizeof(double), cudaMemcpyDeviceToHost, out));
         if (energy) MCHRG_CUDA(cudaMemcpyAsync(energy + a0, d_e, na * sizeof(double), cudaMemcpyDeviceToHost, out));
         if (gradient) MCHRG_CUDA(cudaMemcpyAsync(gradient + 3 * a0, d_g, 3 * na * sizeof(double), cudaMemcpyDeviceToHost, out));
         MCHRG_CUDA(cudaMemcpyAsync(status + m0, d_st, (size_t)mc * sizeof(int32_t), cudaMemcpyDeviceToHost, out));
      }
      cudaEvent_t out_done = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(out_done, ctx->stream_out));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, out_done, 0));
      MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));  // host staging vectors must outlive the async copies
      return 0;
   });
}
