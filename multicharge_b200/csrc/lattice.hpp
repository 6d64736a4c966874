// Host-side scalar pieces of the periodic path that stay on the CPU:
// lattice translations (mctc_cutoff get_lattice_points, model/type.F90:131-149) and the Ewald
// splitting parameter (ewald.f90:52-135).  O(1) work per call.
#pragma once
#include <cfloat>
#include <cmath>
#include <vector>

namespace mchrg {
namespace lat {

inline double det3(const double* a) {
   return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[3] * (a[1] * a[8] - a[2] * a[7]) + a[6] * (a[1] * a[5] - a[2] * a[4]);
}

inline void inv3(const double* a, double* b) {
   const double s = 1.0 / det3(a);
   auto A = [&](int i, int j) { return a[i + 3 * j]; };
   b[0] = s * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1));
   b[1] = s * (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2));
   b[2] = s * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
   b[3] = s * (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2));
   b[4] = s * (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0));
   b[5] = s * (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1));
   b[6] = s * (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1));
   b[7] = s * (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2));
   b[8] = s * (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
}

// get_lattice_points(lat, rep, origin, trans): loop order ix, iy, iz = 0..rep with the signs
// (+, -) of each non-zero index innermost; the origin comes first when requested.
inline std::vector<double> points_rep(const double* lattice, const int rep[3], bool origin) {
   std::vector<double> t;
   for (int ix = 0; ix <= rep[0]; ++ix)
      for (int iy = 0; iy <= rep[1]; ++iy)
         for (int iz = 0; iz <= rep[2]; ++iz) {
            if (!origin && ix == 0 && iy == 0 && iz == 0) continue;
            for (int sx = 1; sx >= (ix > 0 ? -1 : 1); sx -= 2)
               for (int sy = 1; sy >= (iy > 0 ? -1 : 1); sy -= 2)
                  for (int sz = 1; sz >= (iz > 0 ? -1 : 1); sz -= 2)
                     for (int c = 0; c < 3; ++c)
                        t.push_back(lattice[c] * (ix * sx) + lattice[3 + c] * (iy * sy) + lattice[6 + c] * (iz * sz));
         }
   return t;
}

// get_lattice_points(periodic, lat, cutoff, trans)
inline std::vector<double> points_cutoff(const int periodic[3], const double* lattice, double cutoff) {
   if (!(periodic[0] || periodic[1] || periodic[2])) return std::vector<double>(3, 0.0);
   int rep[3];
   for (int k = 0; k < 3; ++k) {
      const double* a = lattice + 3 * ((k + 1) % 3);
      const double* b = lattice + 3 * ((k + 2) % 3);
      const double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      const double spacing = (n[0] * lattice[3 * k] + n[1] * lattice[3 * k + 1] + n[2] * lattice[3 * k + 2]) / len;
      rep[k] = periodic[k] ? (int)std::ceil(std::fabs(cutoff / spacing)) : 0;
   }
   return points_rep(lattice, rep, true);
}

inline void reciprocal(const double* lattice, double* rec) {  // 2 pi transpose(inv(lattice))
   double inv[9];
   inv3(lattice, inv);
   const double twopi = 6.283185307179586476925286766559;
   for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rec[i + 3 * j] = twopi * inv[j + 3 * i];
}

inline std::vector<double> dir_trans(const double* lattice) {  // type.F90:131-138
   const int rep[3] = {2, 2, 2};
   return points_rep(lattice, rep, true);
}
inline std::vector<double> rec_trans(const double* lattice) {  // type.F90:140-149
   const int rep[3] = {2, 2, 2};
   double rec[9];
   reciprocal(lattice, rec);
   return points_rep(rec, rep, false);
}

// get_alpha / search_alpha (ewald.f90:52-135): match the decay of the real- and reciprocal-space
// terms by doubling + bisection; 0.25 on any failure.
inline double ewald_alpha(const double* lattice) {
   const double pi = 3.14159265358979323846264338327950288;
   const double tol = std::sqrt(DBL_EPSILON), alpha0 = 1.0e-8;
   const double vol = std::fabs(det3(lattice));
   double rec[9];
   reciprocal(lattice, rec);
   double rmin = HUGE_VAL, dmin = HUGE_VAL;
   for (int k = 0; k < 3; ++k) {
      rmin = std::fmin(rmin, rec[3 * k] * rec[3 * k] + rec[3 * k + 1] * rec[3 * k + 1] + rec[3 * k + 2] * rec[3 * k + 2]);
      dmin = std::fmin(dmin, lattice[3 * k] * lattice[3 * k] + lattice[3 * k + 1] * lattice[3 * k + 1] +
                                lattice[3 * k + 2] * lattice[3 * k + 2]);
   }
   const double rlen = std::sqrt(rmin), dlen = std::sqrt(dmin);
   auto rterm = [&](double g, double a) { return 4.0 * pi * (std::exp(-0.25 * g * g / (a * a)) / (vol * g * g)); };
   auto dterm = [&](double r, double a) { return std::erfc(a * r) / r; };
   auto diff = [&](double a) {
      return (rterm(4 * rlen, a) - rterm(5 * rlen, a)) - (dterm(2 * dlen, a) - dterm(3 * dlen, a));
   };
   double alpha = alpha0, lo = 0.0, hi = 0.0, d = diff(alpha);
   int stat = 0;
   while (d < -tol && alpha <= DBL_MAX) { alpha *= 2.0; d = diff(alpha); }
   if (alpha > DBL_MAX) stat = 1;
   else if (alpha == alpha0) stat = 2;
   if (stat == 0) {
      lo = 0.5 * alpha;
      while (d < tol && alpha <= DBL_MAX) { alpha *= 2.0; d = diff(alpha); }
      if (alpha > DBL_MAX) stat = 3;
   }
   if (stat == 0) {
      hi = alpha;
      alpha = 0.5 * (lo + hi);
      int it = 0;
      d = diff(alpha);
      while (std::fabs(d) > tol && it <= 30) {
         if (d < 0) lo = alpha; else hi = alpha;
         alpha = 0.5 * (lo + hi);
         d = diff(alpha);
         ++it;
      }
      if (it > 30) stat = 4;
   }
   return stat == 0 ? alpha : 0.25;
}

}  // namespace lat
}  // namespace mchrg
