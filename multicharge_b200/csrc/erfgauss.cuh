// erf(x) and exp(-x^2) for the pair kernels: a 417-entry table of (erf(x_k), exp(-x_k^2)) at x_k = k / 64
// plus the Taylor series about the nearest node, generated with the Hermite recurrence
//    d^n/dx^n exp(-x^2) = (-1)^n H_n(x) exp(-x^2),   H_{n+1} = 2 x H_n - 2 n H_{n-1}:
//    t_n = (-1)^n H_n(x_k) h^n / n!   =>   t_{n+1} = -2/(n+1) (x_k h t_n + h^2 t_{n-1}),  t_0 = 1, t_1 = -2 x_k h
//    exp(-(x_k+h)^2) = g_k sum_n t_n ,      erf(x_k+h) = e_k + 2/sqrt(pi) g_k h sum_n t_n / (n+1)
// |h| <= 1/128, so 6 terms beyond t_0 give |error| <= 1.2e-16 (erf) and 4e-16 (gauss) absolute on [0, 6.5] --
// the rounding level of the CUDA library functions -- with ~30 FP64 instructions and one 16-byte load
// instead of ~145 (erf) + ~45 (exp) instructions.  Arguments above 6.5 are clamped: erf == 1 to the last
// bit there and exp(-x^2) < 5e-19.  (The Coulomb and coordination-number kernels of the reference call
// libm erf/exp: eeq.f90:220-226, 403-411; mctc_ncoord erf counting function.)
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

namespace mchrg {
namespace eg {

constexpr int W = 64;                    // nodes per unit
constexpr double XMAX = 6.5;
constexpr int NTAB = (int)(W * XMAX) + 1;  // 417
constexpr double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer (low word)
constexpr double TWO_OVER_SQRTPI = 1.1283791670955125738961589031215452;

inline std::vector<double2> host_table() {
   std::vector<double2> t(NTAB);
   for (int k = 0; k < NTAB; ++k) {
      const double x = (double)k / W;
      t[k].x = std::erf(x);
      t[k].y = std::exp(-x * x);
   }
   return t;
}

// cooperative copy of the table into shared memory (caller synchronises)
__device__ __forceinline__ void load_table(double2* __restrict__ dst, const double2* __restrict__ src, int tid, int nt) {
   for (int k = tid; k < NTAB; k += nt) dst[k] = src[k];
}

// 1 / sqrt(x) for normal positive x: hardware seed (MUFU.RSQ64H, > 20 good bits) and ONE third-order
// correction y (1 + e/2 + 3 e^2/8), e = 1 - x y^2  (error ~ 5/16 e^3 < 2^-60): 6 instructions instead of the
// ~14 of the library routine (two Newton steps plus special-case handling the pair kernels do not need).
__device__ __forceinline__ double fast_rsqrt(double x) {
   double y;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   const double e = fma(-x * y, y, 1.0);
   const double q = e * fma(0.375, e, 0.5);
   return fma(y, q, y);
}

// x >= 0.  Returns erf(x); `g` receives exp(-x^2) when WANT_G.  NTERM = highest series term.
template <int NTERM, bool WANT_G>
__device__ __forceinline__ double erf_gauss(const double2* __restrict__ tab, double x, double& g) {
   x = fmin(x, XMAX);
   const double s = fma(x, (double)W, MAGIC);
   const int k = __double2loint(s);
   const double xk = (s - MAGIC) * (1.0 / W);
   const double h = x - xk;
   const double2 eg = tab[k];
   const double a = xk * h, b = h * h;
   double tm = 1.0, t = -2.0 * a;
   double sg = 1.0 + t;
   double se = fma(0.5, t, 1.0);
#pragma unroll
   for (int n = 1; n < NTERM; ++n) {
      const double tn = (-2.0 / (n + 1)) * fma(a, t, b * tm);
      tm = t;
      t = tn;
      if (WANT_G) sg += t;
      se = fma(t, 1.0 / (n + 2), se);
   }
   if (WANT_G) g = eg.y * sg;
   return fma(TWO_OVER_SQRTPI * eg.y * h, se, eg.x);
}

}  // namespace eg
}  // namespace mchrg
