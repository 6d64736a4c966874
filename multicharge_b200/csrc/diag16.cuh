// Cholesky + explicit inverse of one 16 x 16 diagonal tile by ONE warp, shared by the batch factorisation
// (batch.cu, eeq_factor_ll_kernel) and the 128 x 128 leaf of the single-system Cholesky (dense.cu).
// Replaces the innermost level of dsytrf / dsytri (lapack.F90:165-201, 335-363) on the device.
#pragma once
#include <cuda_runtime.h>

namespace mchrg {
namespace d16 {

constexpr int KB = 16;   // tile order
constexpr int CBW = 20;  // column buffer: 16 entries (parity-split) + [16] pivot + [17] next diagonal entry

// 1 / x for the pivots (normal, positive x; anything else ends in the failed-pivot report anyway).
// D16_RCP = 1 (default): hardware seed (MUFU.RCP64H, > 20 good bits) and two Newton steps, 5 dependent instructions;
// the IEEE division (D16_RCP = 0) is a ~20-instruction sequence with a slow-path call and sits on the serial pivot
// chain of every column step (measured: tools/micro/diag_variants.cu).  Relative error < 2^-52 x a few.
#ifndef D16_RCP
#define D16_RCP 1
#endif
__device__ __forceinline__ double pivot_rcp(double x) {
#if D16_RCP
   double y;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   double e = fma(-x, y, 1.0);
   y = fma(y, e, y);
   e = fma(-x, y, 1.0);
   return fma(y, e, y);
#else
   return 1.0 / x;
#endif
}

// conflict-free 16 x 16 layout for the FP64 mma fragment loads of batch.cu
__device__ __forceinline__ int swz(int row, int col) { return row * KB + (col ^ ((row & 3) << 2)); }

// Cholesky + inverse of the 16 x 16 tile T (row stride TSR; lower triangle valid) by one warp.
// lane = (row r = lane & 15, column parity h = lane >> 4) holds m[k] = A(r, 2k + h) and x[k] = X(r, 2k + h).
// The elimination runs on UNSCALED columns (f_r = a_rc / a_cc, L(r, c) = a_rc / sqrt(a_cc) only at the end),
// X collects inv(L) by the same row operations on [A | I] with its rows scaled at the end as well.  Per step
// the owners publish column c (and row c of X) through shared memory; every lane then computes the NEXT
// pivot a_{c+1,c+1} - a_{c+1,c}^2 / a_cc and its reciprocal redundantly, so the only serial dependence from
// step to step is one FMA + one reciprocal (~90 cycles) and the shared-memory broadcast of the next column
// runs in its shadow (measured: tools/micro/diag_latency.cu).
// Columns >= nfree are kept at identity (the right-hand-side rows of an augmented matrix; nfree = 14 or 16).
// Outputs: inv(L) in Li (LI_SWZ: swz() layout, else row-major with stride LIS); L in T when want_l.
// Scratch: cbuf[2][20], rbuf[2][16], dvec[16].
// Returns 0 or the 1-based local index of the first non-positive pivot.
// (No __restrict__ on the broadcast buffers: they are written by OTHER lanes between the warp barriers.)
template <int TSR, bool LI_SWZ, int LIS>
__device__ __forceinline__ int diag_factor_inv(double* T, double* Li, double* cbuf, double* rbuf, double* dvec, int lane,
                                               int nfree, bool want_l) {
   const int r = lane & 15, h = lane >> 4;
   double m[8], x[8];
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      const int c = 2 * k + h;
      m[k] = (c <= r) ? T[r * TSR + c] : 0.0;
      if (c >= nfree) m[k] = (c == r) ? 1.0 : 0.0;
      x[k] = (c == r) ? 1.0 : 0.0;
   }
   const int myslot = (r & 1) * 8 + (r >> 1);  // parity-split position of row r in the column buffer
   auto publish = [&](const int c) {
      double* cb = cbuf + (c & 1) * CBW;
      const bool forced = (c >= KB - 2) && (c >= nfree);  // right-hand-side rows: their own columns stay identity
      if (h == (c & 1)) {
         cb[myslot] = (r > c && !forced) ? m[c >> 1] : 0.0;
         if (r == c) cb[16] = forced ? 1.0 : m[c >> 1];
      }
      if (c + 1 < KB && r == c + 1 && h == ((c + 1) & 1)) cb[17] = m[(c + 1) >> 1];
      if (r == c) {
         double2* rb = reinterpret_cast<double2*>(rbuf + (c & 1) * KB + h * 8);
#pragma unroll
         for (int k = 0; k < 4; ++k) rb[k] = make_double2(x[2 * k], x[2 * k + 1]);
      }
   };
   publish(0);
   __syncwarp();
   double piv = cbuf[16];
   double d = pivot_rcp(piv);
   int bad = !(piv > 0.0) ? 1 : 0;
#pragma unroll
   for (int c = 0; c < KB; ++c) {
      const double* cb = cbuf + (c & 1) * CBW;
      const double* rb = rbuf + (c & 1) * KB;
      if (lane == 0) dvec[c] = d;
      double dn = 0.0;
      if (c + 1 < KB) {  // next pivot, redundantly in every lane
         const bool forced = (c + 1 >= KB - 2) && (c + 1 >= nfree);
         const double u = cb[((c + 1) & 1) * 8 + ((c + 1) >> 1)];
         piv = forced ? 1.0 : fma(-u * u, d, cb[17]);
         bad = (!(piv > 0.0) && bad == 0) ? (c + 2) : bad;
         dn = pivot_rcp(piv);
      }
      const double f = cb[myslot] * d;  // 0 for r <= c: the updates below leave those rows alone
      // (one warp issues in order: every instruction of a step delays the next pivot, so the triangular structure is
      //  used -- columns <= c of A are final, row c of X is zero right of column c; c, k are compile-time here)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
         if (2 * k + 1 > c) m[k] = fma(-f, cb[h * 8 + k], m[k]);  // entries right of the diagonal collect garbage; never used
         if (2 * k <= c) x[k] = fma(-f, rb[h * 8 + k], x[k]);
      }
      if (c + 1 < KB) {
         publish(c + 1);
         __syncwarp();
         d = dn;
      }
   }
   __syncwarp();
   if (lane < KB) dvec[lane] = sqrt(dvec[lane]);  // 1 / L(c, c)
   __syncwarp();
   const double rs = dvec[r];
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      const int c = 2 * k + h;
      Li[LI_SWZ ? swz(r, c) : r * LIS + c] = (c <= r) ? x[k] * rs : 0.0;
      if (want_l) T[r * TSR + c] = (c <= r) ? m[k] * dvec[c] : 0.0;
   }
   return bad;
}

}  // namespace d16
}  // namespace mchrg
