// Batches of independent non-periodic molecules (BASELINE.json configs[0..1]): ONE CTA PER MOLECULE.
//
//   reference path replaced: app/main.f90:114-118 / charge.f90:37-77 per molecule, i.e.
//                  mctc_ncoord (erf CN, log-cut), eeq.f90 get_xvec / get_amat_0d / get_damat_0d,
//                  type.F90 solve (sytrf/sytrs -> Cholesky + Schur complement), energy and gradient.
//
//   fused path  (small batches)   eeq_batch_kernel<NT, 0>: the whole single point in one launch with the
//                  packed matrix in shared memory (right-looking blocked Cholesky, block 16, trailing
//                  update on FP64 tensor cores); one launch per size class of 16.
//   split path  (>= 4096 molecules) three kernels per chunk of molecules on three streams:
//                  eeq_batch_kernel<256, 1>  pair phase A: CN, xvec, Coulomb block -> global scratch
//                  eeq_factor_ll_kernel      left-looking Cholesky of the augmented matrix in the scratch
//                                            (L2 resident), forward/backward substitution, Schur complement
//                  eeq_batch_kernel<256, 3>  pair phase C: energy, gradient
//                  The scratch matrices are stored in FP64-mma fragment order (lay()).
//                  What depends on the molecule offsets alone (size classes, launch order, scratch layout) is kept
//                  on the context (SplitPlan) while a caller steps the same set of molecules.
//                  eeq_factor_tma_kernel is the measured-slower alternative of the middle kernel
//                  (MCHRG_B200_FACTOR=tma), kept under test.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "model.cuh"
#include "erfgauss.cuh"
#include "diag16.cuh"

namespace mchrg {
namespace batch {

#define SQRTPI 1.77245385090551602729816748334114518
#define SQRT2PI 0.797884560802865355879892119868763737

constexpr int KB = 16;        // Cholesky block
constexpr int NP_MAX = 208;   // largest padded order that fits in 227 KB of shared memory
constexpr int NVEC = 13;
// split path: the matrix is augmented by two right-hand-side rows (see eeq_factor_ll_kernel)
constexpr int NPA_MAX = NP_MAX + KB;
static_assert(NP_MAX == BATCH_MAX_ATOMS, "eeq.cuh: BATCH_MAX_ATOMS must match NP_MAX");
__host__ __device__ inline int aug_order(int n) { return ((n + 2 + KB - 1) / KB) * KB; }
// Scratch layout of the split path: 8 x 8 blocks of the lower block triangle, block (R, C) at 64 (R (R + 1) / 2 + C),
// inside a block the entry (g, c) sits at (g * 4 + (c & 3)) * 2 + (c >> 2) -- the order in which the 32 lanes
// of an FP64 mma fragment (lane = g * 4 + t) consume it: one 16-byte load per lane fetches the operands of two
// k-steps and a warp reads 512 contiguous bytes.
__host__ __device__ inline int64_t lay_size(int npa) { return 32 * (int64_t)(npa / 8) * (npa / 8 + 1); }
__device__ __forceinline__ int lay(int i, int j) {
   const int R = i >> 3, C = j >> 3;
   return 64 * (R * (R + 1) / 2 + C) + ((((i & 7) << 2) + (j & 3)) << 1) + ((j >> 2) & 1);
}

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
   const double2* etab;  // erfgauss.cuh node table
   int nid;
   double kcn, cutoff2, cut;
   int np;               // padded order of the class (multiple of 16); split pair phases: capacity of the launch
   // split path (PH != 0): the pair phases and the factorisation run as separate kernels
   double* scratch;      // packed matrices of all molecules of the chunk
   const int64_t* soff;  // [nmol_chunk] start of molecule m in scratch
   double* vecg;         // [3 * ntot_chunk]: xvec | thalf | fcut per atom
   double* lamg;         // [nmol_chunk] Lagrange multiplier
   int64_t ntot;         // atoms in the chunk
};

// pair-phase auxiliaries: the erf/gauss node table and one 64-entry ring of deferred near pairs per warp
constexpr int AUX_RING = 64;                                   // ints per warp
constexpr int AUX_DOUBLES = 2 * eg::NTAB + 2 + 8 * AUX_RING / 2;  // table (417 double2) + 8 warps * 64 ints
// In the heavy (matrix in shared memory) layout the auxiliaries live in the Cholesky panel buffer, which the
// pair phases do not use, when it is large enough; small orders append them.
__host__ __device__ inline bool aux_in_panel(int np) { return KB * (np + 4) >= AUX_DOUBLES; }
__host__ __device__ inline size_t smem_doubles(int np) {
   return (size_t)np * (np + 1) / 2 + (size_t)KB * (np + 4) + KB * (KB + 1) + KB + (size_t)NVEC * np + 8 +
          (aux_in_panel(np) ? 0 : AUX_DOUBLES);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// Sums of six values over the warp with a butterfly that halves the number of values per lane at every step
// (8 shuffles of a double instead of 30): value order (a0, a1, a2, 0, b0, b1, b2, 0); lane 4 v returns the total of
// value v, the other lanes partial garbage.
__device__ __forceinline__ double warp_sum_3x2(double a0, double a1, double a2, double b0, double b1, double b2, int lane) {
   constexpr unsigned FULL = 0xffffffffu;
   const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
   // step 1: (a_k, b_k) pairs; upper half-warp keeps b
   const double w0 = (u16 ? b0 : a0) + __shfl_xor_sync(FULL, u16 ? a0 : b0, 16);
   const double w1 = (u16 ? b1 : a1) + __shfl_xor_sync(FULL, u16 ? a1 : b1, 16);
   const double w2 = (u16 ? b2 : a2) + __shfl_xor_sync(FULL, u16 ? a2 : b2, 16);
   // step 2: (w0, w2), (w1, 0)
   const double v0 = (u8 ? w2 : w0) + __shfl_xor_sync(FULL, u8 ? w0 : w2, 8);
   const double v1 = (u8 ? 0.0 : w1) + __shfl_xor_sync(FULL, u8 ? w1 : 0.0, 8);
   // step 3: (v0, v1)
   double t = (u4 ? v1 : v0) + __shfl_xor_sync(FULL, u4 ? v0 : v1, 4);
   t += __shfl_xor_sync(FULL, t, 2);
   t += __shfl_xor_sync(FULL, t, 1);
   return t;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
   asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

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

// PH selects the phases a launch executes:
//   0  everything in one kernel, matrix in shared memory (small batches)
//   1  pair phase A: coordination number, xvec, Coulomb block -> global scratch           (light smem)
//   3  pair phase C: energy and gradient, pair scalars parked in scratch                  (light smem)
// (between 1 and 3 the split path runs eeq_factor_ll_kernel; the former phase 2 -- the in-kernel factorisation
//  with the matrix reloaded into shared memory -- was replaced by it)
template <int NT, int PH>
__global__ void __launch_bounds__(NT, 1024 / NT) eeq_batch_kernel(Args p) {
   extern __shared__ __align__(16) double sm[];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = NT / 32;
   const int mol = p.perm[blockIdx.x];
   const int64_t off = p.offsets[mol];
   const int n = (int)(p.offsets[mol + 1] - off);
   const double charge = p.charge[mol];
   constexpr bool MAT_IN_SMEM = (PH == 0);
   // pair phases work on the molecule's own padded order, the class launches on the class order
   const int np = MAT_IN_SMEM ? p.np : aug_order(n);  // split path: augmented order (see eeq_factor_ll_kernel)
   const int npv = MAT_IN_SMEM ? p.np : NPA_MAX;  // stride of the per-atom vectors (compile-time in the split path)
   const int ldp = np + 4;
   // fused: packed lower, row-major A[tri(i) + j]; split: phase A writes A[lay(i, j)], phase C parks at A[tri(i) + j]
   double* A = MAT_IN_SMEM ? sm : p.scratch + p.soff[mol];
   double* P = sm + (MAT_IN_SMEM ? (size_t)np * (np + 1) / 2 : 0);  // panel, k-major: P[c * ldp + i]
   double* Ld = P + (MAT_IN_SMEM ? (size_t)KB * ldp : 0);            // diagonal block Ld[r * 17 + c]
   double* spare = Ld + (MAT_IN_SMEM ? KB * (KB + 1) : 0);           // KB doubles, unused
   double* vec = spare + (MAT_IN_SMEM ? KB : 0);
   double* X = vec;            double* Y = X + npv;        double* Z = Y + npv;
   double* rad2 = Z + npv;     double* dself = rad2 + npv; double* rcov = dself + npv;
   double* fcut = rcov + npv;  double* xv = fcut + npv;    double* th = xv + npv;
   double* b0 = th + npv;      double* b1 = b0 + npv;      double* qv = b1 + npv;
   double* rd = qv + npv;      // reciprocal pivots 1 / L(i, i); later reused for thalf * q * fcut
   double* scal = rd + npv;    // [0] lambda, [1] s
   double* aux = (MAT_IN_SMEM && aux_in_panel(np)) ? P : scal + 8;
   double2* etab = reinterpret_cast<double2*>(aux);                 // erf / gauss nodes
   int* ring = reinterpret_cast<int*>(aux + 2 * eg::NTAB + 2) + warp * AUX_RING;  // this warp's deferred pairs
   unsigned* cnfix = reinterpret_cast<unsigned*>(b0);  // fixed-point CN sums, three 20-bit limbs [3][npv] (b0, b1; phase A)
   __shared__ int sfail;
   double* xv_g = p.vecg + off;
   double* th_g = p.vecg + p.ntot + off;
   double* fc_g = p.vecg + 2 * p.ntot + off;

   // ---- load atoms and per-element parameters (chi, rad, eta, kcnchi, rcov = slots 0..4) ----
   if (tid == 0) sfail = 0;
   int badz = 0;
   for (int i = tid; i < np; i += NT) {
      if (i < n) {
         int z = p.num[off + i] - 1;
         if ((unsigned)z >= (unsigned)p.nid) {  // not a species of the model: the molecule is rejected below
            badz = 1;
            z = 0;
         }
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
      if (PH == 0 || PH == 1) cnfix[i] = cnfix[npv + i] = cnfix[2 * npv + i] = 0u;
   }
   eg::load_table(etab, p.etab, tid, NT);
   if (__syncthreads_or(badz) && PH != 3) {  // atomic number outside the model's species: status MCHRG_B200_ERR_ARG, outputs poisoned
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = tid; i < n; i += NT) {
         p.qvec[off + i] = nan;
         if (p.energy) p.energy[off + i] = nan;
         if (p.gradient) p.gradient[3 * (off + i)] = p.gradient[3 * (off + i) + 1] = p.gradient[3 * (off + i) + 2] = nan;
      }
      if (tid == 0) p.status[mol] = MCHRG_B200_ERR_ARG;
      return;
   }

   const int ntri_n = tri(n);
   constexpr unsigned FULL = 0xffffffffu;
   const unsigned lt_mask = (1u << lane) - 1u;
   if (PH == 0 || PH == 1) {
   // ---- one pass over the unique pairs: Coulomb block J, packed lower (eeq.f90:199-242; padding = identity),
   //      and the erf coordination number (mctc_ncoord).  The counting function is exactly zero beyond
   //      r = (1 + 6 / kcn) rc, so only the few bonded pairs need its erf: they are deferred into the warp's
   //      ring and evaluated 32 at a time with all lanes busy.  Counts are summed in 2^-52 fixed point with
   //      integer shared-memory atomics: associative, so the result does not depend on the arrival order.
   const double near = 1.0 + 6.0 / p.kcn;
   const double near2 = near * near;
   constexpr double FIX = 4503599627370496.0;  // 2^52
   int wr = 0, rdp = 0;
   auto count_pairs = [&](int slot) {
      const int ij = ring[slot & (AUX_RING - 1)];
      const int i = ij >> 16, j = ij & 0xffff;
      const double dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double r1 = r2 * rsqrt(r2);
      // 1/2 (1 + erf(-kcn (r - rc)/rc))
      const double a = p.kcn * (r1 / (rcov[i] + rcov[j]) - 1.0);
      double g;
      const double ea = eg::erf_gauss<6, false>(etab, fabs(a), g);
      const double c = 0.5 - copysign(0.5, a) * ea;
      // (64-bit shared-memory atomics are compare-and-swap loops; 32-bit adds are native: each limb sums at
      //  most 224 values < 2^20, so nothing carries)
      const unsigned long long f = (unsigned long long)__double2ll_rn(c * FIX);
      const unsigned f0 = (unsigned)f & 0xfffffu, f1 = (unsigned)(f >> 20) & 0xfffffu, f2 = (unsigned)(f >> 40);
      // (tried: summing the row atom's limbs in the warp first with match.any + redux, because 32 atomics on one
      //  address serialise -- slower, 6.15 vs 5.72 ms per 100k molecules)
      atomicAdd(&cnfix[i], f0); atomicAdd(&cnfix[npv + i], f1); atomicAdd(&cnfix[2 * npv + i], f2);
      atomicAdd(&cnfix[j], f0); atomicAdd(&cnfix[npv + j], f1); atomicAdd(&cnfix[2 * npv + j], f2);
   };
   // diagonal and padding rows
   for (int i = tid; i < n; i += NT) A[MAT_IN_SMEM ? tri(i) + i : lay(i, i)] = dself[i];
   for (int i = n + warp; i < np; i += NW)
      for (int j = lane; j <= i; j += 32) A[MAT_IN_SMEM ? tri(i) + j : lay(i, j)] = (i == j) ? 1.0 : 0.0;
   // off-diagonal entries by rows, one warp per PAIR of rows (n-1-k, k): n-1 entries together, so every warp
   // gets the same amount of work and the lanes stay busy; no index decoding, the row data are warp-uniform
   auto col_off = [](int j) { return 64 * (j >> 3) + ((j & 3) << 1) + ((j >> 2) & 1); };  // lay() = row part + this
   for (int k = warp; 2 * k < n; k += NW) {
      const int ia = n - 1 - k, ib = k;
      const int len = (ia == ib) ? ia : ia + ib;
      // storage index of (i, j): both column walks advance by 32 columns = 256 doubles per step
      const int offA = MAT_IN_SMEM ? tri(ia) + lane : lay(ia, 0) + col_off(lane);
      const int offB = MAT_IN_SMEM ? tri(ib) + lane - ia : lay(ib, 0) + col_off(lane - ia);
      for (int t0 = 0; t0 < len; t0 += 32) {
         const int tt = t0 + lane;
         bool hit = false;
         const int i = (tt < ia) ? ia : ib;
         const int j = (tt < ia) ? tt : tt - ia;
         const int idx = ((tt < ia) ? offA : offB) + (MAT_IN_SMEM ? t0 : 8 * t0);
         if (tt < len) {
            const double dx = X[j] - X[i], dy = Y[j] - Y[i], dz = Z[j] - Z[i];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(rad2[i] + rad2[j]);
            double g;
            const double v = eg::erf_gauss<5, false>(etab, r2 * rinv * gam, g) * rinv;  // erf(gamma r) / r
            const double rc = rcov[i] + rcov[j];
            hit = r2 < near2 * rc * rc && !(r2 > p.cutoff2 || r2 < 1.0e-12);
            A[idx] = v;
         }
         const unsigned m = __ballot_sync(FULL, hit);
         if (hit) ring[(wr + __popc(m & lt_mask)) & (AUX_RING - 1)] = (i << 16) | j;
         wr += __popc(m);
         if (wr - rdp >= 32) {
            __syncwarp();
            count_pairs(rdp + lane);
            rdp += 32;
            __syncwarp();
         }
      }
   }
   __syncwarp();
   if (lane < wr - rdp) count_pairs(rdp + lane);
   __syncthreads();
   const double ecut = exp(p.cut);
   for (int i = tid; i < n; i += NT) {
      const double acc = (double)((unsigned long long)cnfix[i] + ((unsigned long long)cnfix[npv + i] << 20) +
                                  ((unsigned long long)cnfix[2 * npv + i] << 40)) * (1.0 / FIX);
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
   __syncthreads();
   }  // PH 0 / 1
   if (PH == 1) {  // hand xvec, thalf, fcut to the later phases; right-hand sides [x | 1] = last two matrix rows
      for (int i = tid; i < n; i += NT) {
         xv_g[i] = xv[i]; th_g[i] = th[i]; fc_g[i] = fcut[i];
         A[lay(np - 2, i)] = xv[i];
         A[lay(np - 1, i)] = 1.0;
      }
      if (tid == 0) p.status[mol] = 0;  // (a negative status tells the factorisation kernel to leave the molecule alone)
      return;
   }
   if (PH == 3) {
      if (p.status[mol] != 0) return;  // not positive definite: the factorisation kernel left NaNs
      for (int i = tid; i < n; i += NT) { xv[i] = xv_g[i]; th[i] = th_g[i]; fcut[i] = fc_g[i]; qv[i] = p.qvec[off + i]; }
      if (tid == 0) scal[0] = p.lamg[mol];
      __syncthreads();
   }

   if (PH == 0) {
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
   {
      const double lam2 = scal[0];
      for (int i = tid; i < n; i += NT) {
         const double q = b0[i] - lam2 * b1[i];
         qv[i] = q;
         p.qvec[off + i] = q;
      }
      if (tid == 0) {
         p.status[mol] = 0;
      }
   }
   __syncthreads();
   }  // PH 0

   if (PH == 0 && aux_in_panel(np)) eg::load_table(etab, p.etab, tid, NT);  // the panel buffer held it before
   const double lam = scal[0];
   double* wf = rd;  // pivots are no longer needed
   for (int i = tid; i < n; i += NT) {
      const double q = qv[i];
      wf[i] = th[i] * q * fcut[i];
      // energy (type.F90:255-262): q_i (0.5 (J q)_i - x_i) with (J q)_i = x_i - lambda from the solved system
      // J q + lambda 1 = x (residual of the Cholesky solve ~ n eps |J| |q|)
      if (p.energy) p.energy[off + i] = q * (0.5 * (xv[i] - lam) - xv[i]);
   }
   if (!p.gradient) return;
   __syncthreads();

   // ---- gradient (type.F90:275-278): g_i = sum_j f_ij (r_i - r_j) with the symmetric pair scalar
   //        f_ij = dtmp_ij q_i q_j - dcount_ij / r_ij (w_i + w_j),   w = thalf * q * fcut
   //      (q_i atrace_i minus the dcndr contraction; eeq.f90:403-411, mctc_ncoord); one evaluation per
   //      unique pair.  Rows are paired as in phase A.  The share of the ROW atom accumulates in registers;
   //      f is parked in the matrix storage (packed, row-major) and the shares of the COLUMN atoms are
   //      collected afterwards with coalesced reads ----
   double* G = th;  // [3][npv] row shares: th, b0, b1 are free now
   {
      const double near = 1.0 + eg::XMAX / p.kcn;  // exp(-a^2) < 5e-19 beyond
      const double near2 = near * near;
      for (int k = warp; 2 * k < n; k += NW) {
         const int ia = n - 1 - k, ib = k;
         const int len = (ia == ib) ? ia : ia + ib;
         double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;  // shares of the rows ia / ib
         int wr = 0, rdp = 0;
         auto dcount_pairs = [&](int slot) {  // coordination-number part of f for a deferred (bonded) pair
            const int ij = ring[slot & (AUX_RING - 1)];
            const int i = ij >> 16, j = ij & 0xffff;
            const double vx = X[i] - X[j], vy = Y[i] - Y[j], vz = Z[i] - Z[j];
            const double r2 = vx * vx + vy * vy + vz * vz;
            const double rinv = rsqrt(r2);
            const double rci = 1.0 / (rcov[i] + rcov[j]);
            const double a = p.kcn * (r2 * rinv * rci - 1.0);
            double g;
            eg::erf_gauss<6, true>(etab, fabs(a), g);
            const double f = p.kcn / SQRTPI * rci * g * rinv * (wf[i] + wf[j]);
            A[tri(i) + j] += f;
            if (i == ia) { ax = fma(f, vx, ax); ay = fma(f, vy, ay); az = fma(f, vz, az); }
            else { bx = fma(f, vx, bx); by = fma(f, vy, by); bz = fma(f, vz, bz); }
         };
         unsigned hmask = 0;  // near pairs of this lane, one bit per step (len <= 207: at most 7 steps); packed into the ring
                              // once per pair of rows (a ballot per step as in phase A costs more here: 9.04 -> 8.65 ms
                              // per 100k molecules together with the six-value warp sum below)
         for (int t0 = 0, st = 0; t0 < len; t0 += 32, ++st) {
            const int tt = t0 + lane;
            const bool first = tt < ia;
            const int i = first ? ia : ib;
            const int j = first ? tt : tt - ia;
            if (tt < len) {
               const double vx = X[i] - X[j], vy = Y[i] - Y[j], vz = Z[i] - Z[j];
               const double r2 = vx * vx + vy * vy + vz * vz;
               const double rinv = eg::fast_rsqrt(r2), gam = eg::fast_rsqrt(rad2[i] + rad2[j]);
               // dtmp = 2 gamma exp(-x^2) / (sqrt(pi) r^2) - erf(x) / r^3,  x = gamma r
               double g;
               const double ex = eg::erf_gauss<6, true>(etab, r2 * rinv * gam, g);
               const double dt = 2.0 / SQRTPI * gam * g - ex * rinv;
               const double f = rinv * rinv * dt * qv[i] * qv[j];
               A[tri(i) + j] = f;
               if (first) { ax = fma(f, vx, ax); ay = fma(f, vy, ay); az = fma(f, vz, az); }
               else { bx = fma(f, vx, bx); by = fma(f, vy, by); bz = fma(f, vz, bz); }
               const double rc = rcov[i] + rcov[j];
               if (r2 < near2 * rc * rc && r2 <= p.cutoff2 && r2 >= 1.0e-12) hmask |= 1u << st;
            }
         }
         for (;;) {
            const bool h = hmask != 0u;
            const unsigned m = __ballot_sync(FULL, h);
            if (m == 0u) break;
            if (h) {
               const int tt = 32 * (__ffs(hmask) - 1) + lane;
               hmask &= hmask - 1u;
               const int i = (tt < ia) ? ia : ib;
               const int j = (tt < ia) ? tt : tt - ia;
               ring[(wr + __popc(m & lt_mask)) & (AUX_RING - 1)] = (i << 16) | j;
            }
            wr += __popc(m);
            if (wr - rdp >= 32) {
               __syncwarp();
               dcount_pairs(rdp + lane);
               rdp += 32;
               __syncwarp();
            }
         }
         __syncwarp();
         if (lane < wr - rdp) dcount_pairs(rdp + lane);  // the rest of this pair of rows
         __syncwarp();
         // the six row shares at once: lane 4 v ends with the sum of value v (0-2: row ia, 4-6: row ib)
         const double tot = warp_sum_3x2(ax, ay, az, bx, by, bz, lane);
         if ((lane & 3) == 0 && (lane & 12) != 12) {
            const int row = (lane & 16) ? ib : ia;
            if (!(lane & 16) || ia != ib) G[((lane >> 2) & 3) * npv + row] = tot;
         }
      }
   }
   __syncthreads();
   // shares of the column atoms: -sum_{i > j} f_ij (r_i - r_j); lanes over 32 consecutive atoms j (contiguous
   // reads for a fixed row i), the warps share the rows and combine their partial sums through shared memory
   double* part = aux;       // [NW][3][32] partial sums (the node table is no longer needed)
   for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const int jc = min(j, n - 1);
      const double xj = X[jc], yj = Y[jc], zj = Z[jc];
      double gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll 4
      for (int i = j0 + 1 + warp; i < n; i += NW) {
         if (j < i) {
            const double f = A[tri(i) + j];
            gx = fma(f, xj - X[i], gx); gy = fma(f, yj - Y[i], gy); gz = fma(f, zj - Z[i], gz);
         }
      }
      part[(warp * 3 + 0) * 32 + lane] = gx;
      part[(warp * 3 + 1) * 32 + lane] = gy;
      part[(warp * 3 + 2) * 32 + lane] = gz;
      __syncthreads();
      if (tid < 96) {
         const int c = tid >> 5;
         double acc = 0.0;
#pragma unroll
         for (int w = 0; w < NW; ++w) acc += part[(w * 3 + c) * 32 + lane];
         if (j < n) p.gradient[3 * (off + j) + c] = G[c * npv + j] + acc;
      }
      __syncthreads();
   }
}

// ------------------------------------------------------------------------------------------------
// Split path, factorisation phase: LEFT-LOOKING blocked Cholesky (block 16) of the AUGMENTED matrix that
// leaves the factor in the global scratch (L2 resident) and keeps one block column in shared memory.
//
//   augmented order npa = roundup(n + 2, 16): rows n .. npa-3 are identity padding, rows npa-2 / npa-1
//   carry the right-hand sides [x | 1] (written by pair phase A), so the forward substitution
//   w = L^-1 [x 1] falls out of the factorisation as the last two rows of L (their own columns are kept
//   at identity).  Per block column k0:
//     U  Cb = A(k0:, k0:k0+16) - L(k0:, :k0) L(k0:k0+16, :k0)^T          DMMA, operands from global / L2
//     D  the 16 x 16 diagonal tile: Cholesky AND explicit inverse in one warp (lanes = row x column parity,
//        column / row broadcasts through shared memory); the INVERSE replaces the diagonal block in global
//        memory (later block columns never read the diagonal blocks of L)
//     S  rows below: L(i, blk) = Cb(i, :) inv(L_kk)^T                     DMMA, operands from shared memory
//   The next diagonal tile is updated one block column ahead (job 0 -> T2, completed by the warp that solves
//   the first row tile), so D(k0) runs while the other warps work on the row tiles; the D warp rotates over
//   the four SM sub-partitions (measured: tools/micro/diag_latency.cu).
//   backward substitution, right-looking: y_blk = inv(L_kk)^T w_blk, then w(:k0) -= L(blk, :k0)^T y_blk with
//   coalesced row reads.  ~32 KB of shared memory per CTA, so 6-7 molecules are resident per SM and hide
//   each other's latencies.
// ------------------------------------------------------------------------------------------------
constexpr int TS = KB + 1;            // row stride of the diagonal tile
#ifndef LL_MINB
#define LL_MINB 6
#endif

#ifndef LL_ASYNC_BACKWARD
#define LL_ASYNC_BACKWARD 1  // backward substitution fed by cp.async-staged strips (0: direct loads)
#endif
#ifndef LL_UNROLL
#define LL_UNROLL 2  // block pairs of the left-looking update in flight per warp (each: four 16-byte loads from L2)
#endif
constexpr int LLU = LL_UNROLL;
using d16::CBW;              // column buffer of the diagonal-tile routine (diag16.cuh)
using d16::swz;


// doubles: block column below the diagonal tile (swizzled, stride 16) | diagonal tile | its inverse (swizzled)
//          | column / row broadcast buffers (double buffered) | w of the last block | scalars
__host__ __device__ inline size_t smem_doubles_ll(int npa) {
   return (size_t)(npa > KB ? npa - KB : 2) * KB + KB * TS + KB * KB + 2 * CBW + 2 * KB + KB + 2 * KB + 8 +
          (npa <= 2 * KB ? 2 * (size_t)npa : 0);  // tiny orders: room for the two solution vectors
}

// Cholesky + inverse of the 16 x 16 diagonal tile by one warp: d16::diag_factor_inv (diag16.cuh), tile stride TS,
// inverse in the swizzled layout the DMMA row solves read
__device__ __forceinline__ int diag_factor_inv(double* T, double* Li, double* cbuf, double* rbuf, double* dvec, int lane,
                                               int nfree, bool want_l) {
   return d16::diag_factor_inv<TS, true, KB>(T, Li, cbuf, rbuf, dvec, lane, nfree, want_l);
}

template <int NT>
__global__ void __launch_bounds__(NT, LL_MINB) eeq_factor_ll_kernel(Args p) {
   extern __shared__ __align__(16) double sm[];
#ifdef LL_PROF  // per-warp phase clocks of sampled molecules (tools/run_llprof.sh)
   long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // D | U jobs | barrier 1 | S | barrier 2 | backward | - | total
   const long long pf_t0 = clock64();
#define LP_BEGIN(v) const long long v = clock64()
#define LP_ADD(i, v) pf[i] += clock64() - v
#else
#define LP_BEGIN(v)
#define LP_ADD(i, v)
#endif
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = NT / 32;
   const int mol = p.perm[blockIdx.x];
   const int64_t off = p.offsets[mol];
   const int n = (int)(p.offsets[mol + 1] - off);
   const int npa = aug_order(n);
   const double charge = p.charge[mol];
   double* S = p.scratch + p.soff[mol];  // fragment-order blocks, S[lay(i, j)]
   double* Cb = sm;                      // rows below the diagonal tile: Cb[swz(i - k0 - 16, c)]
   double* T = Cb + (size_t)(npa > KB ? npa - KB : 2) * KB;
   double* Li = T + KB * TS;
   double* cbuf = Li + KB * KB;
   double* rbuf = cbuf + 2 * CBW;
   double* dvec = rbuf + 2 * KB;
   double* wl = dvec + KB;               // w of the last block: wl[h * 16 + c]
   double* scal = wl + 2 * KB;
   double* b0 = (npa <= 2 * KB) ? scal + 8 : Cb;  // solution vectors alias the block column buffer
   double* b1 = b0 + npa;
   __shared__ int sfail, next_job;
   __shared__ double T2[KB * TS];  // next diagonal tile, updated one block column ahead
   if (p.status[mol] < 0) return;   // rejected by pair phase A (atomic number outside the model)
   if (tid == 0) { sfail = 0; next_job = 0; }
   const int g = lane >> 2, t = lane & 3;
   if (warp == 0) {  // first diagonal tile: nothing to its left
      const int r = lane & 15, h = lane >> 4;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
         const int c = 2 * k + h;
         T[r * TS + c] = (c <= r) ? S[lay(r, c)] : 0.0;
      }
   }
   __syncthreads();

   for (int k0 = 0; k0 < npa; k0 += KB) {
      const int nrt = (npa - k0) / KB;
      // ---- U: left-looking updates of this block column's row tiles (jobs 1 .. nrt-1 -> Cb) and, one block
      //      column ahead, of the NEXT diagonal tile with the columns [0, k0) (job 0 -> T2), handed out
      //      dynamically; meanwhile warp 0 factors the current diagonal tile (D) ----
      auto u_job = [&](const int job) {
         const int i0 = k0 + KB * (job == 0 ? 1 : job);  // rows of the tile
         const int j0 = (job == 0) ? k0 + KB : k0;       // columns of the tile = rows of the B strip
         double acc[2][2][2];
         // original entries first (their loads overlap the first round of panel loads); entries right of the
         // diagonal do not exist in the storage (and are never used)
#pragma unroll
         for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ntl = 0; ntl < 2; ++ntl) {
               const int row = i0 + 8 * mt + g, col = j0 + 8 * ntl + 2 * t;
               const double* src = S + lay(row, (col <= row) ? col : j0);
               acc[mt][ntl][0] = (col <= row) ? -src[0] : 0.0;
               acc[mt][ntl][1] = (col + 1 <= row) ? -src[2] : 0.0;  // (g, c + 1) is 2 doubles on (c even)
            }
         // fragments of two k-steps per 16-byte load, 512 contiguous bytes per warp and block
         const double2* pa0 = reinterpret_cast<const double2*>(S + 64 * tri(i0 >> 3)) + lane;
         const double2* pa1 = reinterpret_cast<const double2*>(S + 64 * tri((i0 >> 3) + 1)) + lane;
         const double2* pb0 = reinterpret_cast<const double2*>(S + 64 * tri(j0 >> 3)) + lane;
         const double2* pb1 = reinterpret_cast<const double2*>(S + 64 * tri((j0 >> 3) + 1)) + lane;
#pragma unroll LLU
         for (int cb = 0; cb < (k0 >> 3); ++cb) {
            const double2 a0 = pa0[32 * cb], a1 = pa1[32 * cb], c0 = pb0[32 * cb], c1 = pb1[32 * cb];
            dmma884(acc[0][0][0], acc[0][0][1], a0.x, c0.x);
            dmma884(acc[0][1][0], acc[0][1][1], a0.x, c1.x);
            dmma884(acc[1][0][0], acc[1][0][1], a1.x, c0.x);
            dmma884(acc[1][1][0], acc[1][1][1], a1.x, c1.x);
            dmma884(acc[0][0][0], acc[0][0][1], a0.y, c0.y);
            dmma884(acc[0][1][0], acc[0][1][1], a0.y, c1.y);
            dmma884(acc[1][0][0], acc[1][0][1], a1.y, c0.y);
            dmma884(acc[1][1][0], acc[1][1][1], a1.y, c1.y);
         }
         // acc = L L^T - A: store the negative
#pragma unroll
         for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ntl = 0; ntl < 2; ++ntl) {
               const int lr = 8 * mt + g, lc = 8 * ntl + 2 * t;  // within the tile
               if (job == 0) {
                  T2[lr * TS + lc] = (lc <= lr) ? -acc[mt][ntl][0] : 0.0;
                  T2[lr * TS + lc + 1] = (lc + 1 <= lr) ? -acc[mt][ntl][1] : 0.0;
               } else {
                  *reinterpret_cast<double2*>(Cb + swz(KB * (job - 1) + lr, lc)) = make_double2(-acc[mt][ntl][0], -acc[mt][ntl][1]);
               }
            }
      };
      // (the warps of a CTA sit on different SM sub-partitions: rotate the serial FP64-heavy D step over them so
      //  that the D chains of the co-resident molecules do not all queue on one FP64 pipe)
      if (warp == ((blockIdx.x + (k0 >> 4)) % NW)) {
         // ---- D: diagonal tile, Cholesky + inverse; the inverse goes to global in place of the block ----
         const bool last = (k0 + KB == npa);
         LP_BEGIN(td);
         const int bad = diag_factor_inv(T, Li, cbuf, rbuf, dvec, lane, last ? KB - 2 : KB, last);
         LP_ADD(0, td);
         if (bad && lane == 0) sfail = k0 + bad;
         __syncwarp();
         const int r = lane & 15, h = lane >> 4;
#pragma unroll
         for (int k = 0; k < 8; ++k) {
            const int c = 2 * k + h;
            if (c <= r) S[lay(k0 + r, k0 + c)] = Li[swz(r, c)];
         }
         if (last && lane < KB) {  // forward-substituted right-hand sides of the last block
            wl[lane] = (lane < KB - 2) ? T[(KB - 2) * TS + lane] : 0.0;
            wl[KB + lane] = (lane < KB - 2) ? T[(KB - 1) * TS + lane] : 0.0;
         }
      }
      LP_BEGIN(tu);
      if (nrt > 1) {
         for (;;) {
            int job = 0;
            if (lane == 0) job = atomicAdd(&next_job, 1);
            job = __shfl_sync(0xffffffffu, job, 0);
            if (job >= nrt) break;
            u_job(job);
         }
      }
      LP_ADD(1, tu);
      LP_BEGIN(tb1);
      __syncthreads();
      LP_ADD(2, tb1);
      if (sfail) break;
      LP_BEGIN(ts);
      // ---- S: rows below the diagonal tile, L(i, blk) = Cb(i, :) inv(L_kk)^T, straight to the scratch; the warp
      //      of the first tile also completes the next diagonal tile, T = T2 - L(tile, blk) L(tile, blk)^T ----
      for (int ti = 1 + warp; ti < nrt; ti += NW) {
         const int lr0 = KB * (ti - 1);
         double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
         for (int kk = 0; kk < 4; ++kk) {
            const int kc = (4 * kk + t) ^ ((g & 3) << 2);
            const double a0 = Cb[(lr0 + g) * KB + kc], a1 = Cb[(lr0 + 8 + g) * KB + kc];
            const double c1 = Li[(8 + g) * KB + kc];
            if (kk < 2) {  // inv(L)(n, k) = 0 for k > n
               const double c0 = Li[g * KB + kc];
               dmma884(acc[0][0][0], acc[0][0][1], a0, c0);
               dmma884(acc[1][0][0], acc[1][0][1], a1, c0);
            }
            dmma884(acc[0][1][0], acc[0][1][1], a0, c1);
            dmma884(acc[1][1][0], acc[1][1][1], a1, c1);
         }
#pragma unroll
         for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ntl = 0; ntl < 2; ++ntl) {
               double* dst = S + lay(k0 + KB + lr0 + 8 * mt + g, k0 + 8 * ntl + 2 * t);
               dst[0] = acc[mt][ntl][0];
               dst[2] = acc[mt][ntl][1];
            }
         if (ti == 1) {
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl)
                  *reinterpret_cast<double2*>(Cb + swz(8 * mt + g, 8 * ntl + 2 * t)) = make_double2(acc[mt][ntl][0], acc[mt][ntl][1]);
            __syncwarp();
            double d2[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
               const int kc = (4 * kk + t) ^ ((g & 3) << 2);
               const double a0 = Cb[g * KB + kc], a1 = Cb[(8 + g) * KB + kc];  // A(m, k) and B(k, n) = A(n, k)
               dmma884(d2[0][0][0], d2[0][0][1], a0, a0);
               dmma884(d2[0][1][0], d2[0][1][1], a0, a1);
               dmma884(d2[1][0][0], d2[1][0][1], a1, a0);
               dmma884(d2[1][1][0], d2[1][1][1], a1, a1);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const int lr = 8 * mt + g, lc = 8 * ntl + 2 * t;
                  T[lr * TS + lc] = (lc <= lr) ? T2[lr * TS + lc] - d2[mt][ntl][0] : 0.0;
                  T[lr * TS + lc + 1] = (lc + 1 <= lr) ? T2[lr * TS + lc + 1] - d2[mt][ntl][1] : 0.0;
               }
         }
      }
      if (tid == 0) next_job = 0;
      LP_ADD(3, ts);
      LP_BEGIN(tb2);
      __syncthreads();
      LP_ADD(4, tb2);
   }
   if (sfail) {
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = tid; i < n; i += NT) {
         p.qvec[off + i] = nan;
         if (p.energy) p.energy[off + i] = nan;
         if (p.gradient) p.gradient[3 * (off + i)] = p.gradient[3 * (off + i) + 1] = p.gradient[3 * (off + i) + 2] = nan;
      }
      if (tid == 0) {
         p.status[mol] = sfail;
         p.lamg[mol] = nan;
      }
      return;
   }

   // ---- backward substitution L^T [y z] = [w_x w_1], right-looking by blocks ----
   LP_BEGIN(tbw);
   {
      const int kl = npa - KB;
      for (int i = tid; i < npa; i += NT) {
         b0[i] = (i < kl) ? S[lay(npa - 2, i)] : wl[i - kl];
         b1[i] = (i < kl) ? S[lay(npa - 1, i)] : wl[KB + i - kl];
      }
   }
   __syncthreads();
   // Every operand of the sweep is known in advance -- only the FMAs wait for y -- so the strips of L and the inverted
   // diagonal tiles are staged through shared memory with cp.async (no registers: fetching them one step ahead into
   // registers pushed the kernel over its budget, 11.2 vs 10.1 ms per 100k molecules), two 64-column chunks and one tile
   // ahead of their use; a step then costs shared-memory latencies and a few barriers instead of two dependent L2 /
   // HBM round trips.  The chunks live in whatever the molecule leaves free of the launch's shared memory: behind
   // b0 | b1 in the block-column buffer and / or behind its own arrays (two 8 KB stages fit for every order when the
   // launch is sized for the largest class; otherwise the direct loads below).
   constexpr int STG = 2 * 8 * 64;  // doubles per stage: 16 rows x 64 columns as 2 x 8 fragment-order blocks
   unsigned dyn_bytes;
   asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
   const int own = (int)smem_doubles_ll(npa);
   const int n_cb = (npa > 2 * KB) ? ((npa - KB) * KB - 2 * npa) / STG : 0;
   const int n_af = ((int)(dyn_bytes / sizeof(double)) - own) / STG;
   if (LL_ASYNC_BACKWARD && NT == 128 && n_cb + n_af >= 2) {  // (the update below maps 128 threads to 64 columns x 2 right-hand sides)
      double* stage[2];
      stage[0] = (n_cb >= 1) ? Cb + 2 * npa : sm + own;
      stage[1] = (n_cb >= 2) ? Cb + 2 * npa + STG : sm + own + (n_cb >= 1 ? 0 : STG);
      int ik0 = npa - KB, ic0 = 0;  // next chunk to fetch: columns [ic0, ic0 + 64) of the strip of block row ik0
      auto issue_tile = [&](int k0, double* dst) {  // blocks (R,R) | (R+1,R) (R+1,R+1) of the inverted diagonal tile
         if (k0 >= 0 && tid < 96) {
            const int R0 = k0 >> 3;
            cp_async16(dst + 2 * tid, tid < 32 ? S + 64 * (tri(R0) + R0) + 2 * tid : S + 64 * (tri(R0 + 1) + R0) + 2 * (tid - 32));
         }
         cp_async_commit();
      };
      auto issue_chunk = [&](double* dst) {
         if (ik0 >= KB) {
            const int R0 = ik0 >> 3, C0 = ic0 >> 3, nbc = min(8, R0 - C0);
            for (int i = tid; i < 64 * nbc; i += NT) {  // 2 block rows x nbc blocks x 32 chunks of 16 bytes
               const int blk = i >> 5, w = i & 31, rb = (blk >= nbc) ? 1 : 0, cbk = blk - rb * nbc;
               // odd blocks swap the halves of their rows: neighbouring blocks read conflict-free (see the reads)
               cp_async16(dst + (rb * 8 + cbk) * 64 + 2 * (w ^ ((cbk & 1) << 2)), S + 64 * (tri(R0 + rb) + C0 + cbk) + 2 * w);
            }
            ic0 += 64;
            if (ic0 >= ik0) { ik0 -= KB; ic0 = 0; }
         }
         cp_async_commit();  // (an empty group keeps the group count in step)
      };
      // groups are numbered in commit order; a wait names how many NEWER groups may stay in flight
      int ncommit = 0;
      auto wait_for = [&](int group) {
         const int newer = ncommit - 1 - group;
         if (newer <= 0) cp_async_wait<0>();
         else if (newer == 1) cp_async_wait<1>();
         else if (newer == 2) cp_async_wait<2>();
         else cp_async_wait<3>();
      };
      issue_tile(npa - KB, T);
      int g_t0 = ncommit++, g_t1 = -1;  // groups of the tiles in T / Li
      issue_chunk(stage[0]);
      int g_s0 = ncommit++;
      issue_chunk(stage[1]);
      int g_s1 = ncommit++;
      int use = 0;
      for (int k0 = npa - KB, tb = 0; k0 >= 0; k0 -= KB, tb ^= 1) {
         wait_for(tb ? g_t1 : g_t0);
         __syncthreads();  // tile visible; the updates of the previous step are complete, its last stage is free
         if (k0 < npa - KB) {
            issue_chunk(stage[use ^ 1]);
            (use ? g_s0 : g_s1) = ncommit++;
         }
         issue_tile(k0 - KB, tb ? T : Li);
         (tb ? g_t0 : g_t1) = ncommit++;
         const double* Tk = tb ? Li : T;
         if (warp == ((blockIdx.x + (k0 >> 4)) % NW)) {  // y_blk = inv(L_kk)^T w_blk ; lane = (r, rhs h)
            const int r = lane & 15, h = lane >> 4;
            const double* wv = (h ? b1 : b0) + k0;
            double v0 = 0.0, v1 = 0.0;  // two chains: the step waits for this sum
#pragma unroll
            for (int c = 0; c < KB; c += 2) {
               if (c >= r)
                  v0 = fma(Tk[(c < 8 ? 0 : 64 + 64 * (r >> 3)) + ((((c & 7) << 2) + (r & 3)) << 1) + ((r >> 2) & 1)], wv[c], v0);
               if (c + 1 >= r)
                  v1 = fma(Tk[(c + 1 < 8 ? 0 : 64 + 64 * (r >> 3)) + (((((c + 1) & 7) << 2) + (r & 3)) << 1) + ((r >> 2) & 1)], wv[c + 1], v1);
            }
            __syncwarp();
            (h ? b1 : b0)[k0 + r] = v0 + v1;
         }
         __syncthreads();
         for (int c0 = 0; c0 < k0; c0 += 64) {  // w(:k0) -= L(blk, :k0)^T y_blk ; thread = (column, rhs)
            wait_for(use ? g_s1 : g_s0);
            __syncthreads();  // chunk visible; (for c0 > 0) everybody is done with the stage refilled below
            if (c0 > 0) {
               issue_chunk(stage[use ^ 1]);
               (use ? g_s0 : g_s1) = ncommit++;
            }
            const int jl = tid & 63, h = tid >> 6, j = c0 + jl;
            if (j < k0) {
               double* bh = h ? b1 : b0;
               const double* st = stage[use];
               const int cbk = jl >> 3, cc = jl & 7, sw = (cbk & 1) << 3;
               double acc = bh[j], acc1 = 0.0;
#pragma unroll
               for (int r = 0; r < 8; ++r) {
                  const int e = ((((r << 2) + (cc & 3)) << 1) ^ sw) + (cc >> 2);
                  acc = fma(-st[cbk * 64 + e], bh[k0 + r], acc);
                  acc1 = fma(-st[(8 + cbk) * 64 + e], bh[k0 + 8 + r], acc1);
               }
               bh[j] = acc + acc1;
            }
            use ^= 1;
         }
         // (the stage of this step's last chunk is refilled behind the first barrier of the next step)
      }
      cp_async_wait<0>();
   } else {
   for (int k0 = npa - KB; k0 >= 0; k0 -= KB) {
      if (warp == ((blockIdx.x + (k0 >> 4)) % NW)) {  // y_blk = inv(L_kk)^T w_blk ; lane = (r, rhs h)
         const int r = lane & 15, h = lane >> 4;
         const double* wv = (h ? b1 : b0) + k0;
         double v = 0.0;
#pragma unroll
         for (int c = 0; c < KB; ++c)
            if (c >= r) v = fma(S[lay(k0 + c, k0 + r)], wv[c], v);
         __syncwarp();
         (h ? b1 : b0)[k0 + r] = v;
      }
      __syncthreads();
      for (int j = tid; j < k0; j += NT) {  // w(:k0) -= L(blk, :k0)^T y_blk
         double s0 = b0[j], s1 = b1[j];
#pragma unroll
         for (int r = 0; r < KB; ++r) {
            const double l = S[lay(k0 + r, j)];
            s0 = fma(-l, b0[k0 + r], s0);
            s1 = fma(-l, b1[k0 + r], s1);
         }
         b0[j] = s0;
         b1[j] = s1;
      }
      __syncthreads();
   }
   }
   LP_ADD(5, tbw);
#ifdef LL_PROF
   pf[7] = clock64() - pf_t0;
   if (lane == 0 && blockIdx.x % 250 == 7)
      printf("[ll] blk %d n %d npa %d warp %d: D %lld U %lld bar1 %lld S %lld bar2 %lld backward %lld total %lld\n", (int)blockIdx.x, n,
             npa, warp, pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[7]);
#endif
   if (warp == 0) {
      double sy = 0.0, sz = 0.0;
      for (int i = lane; i < n; i += 32) { sy += b0[i]; sz += b1[i]; }
      sy = warp_sum(sy);
      sz = warp_sum(sz);
      if (lane == 0) scal[0] = (sy - charge) / sz;
   }
   __syncthreads();
   const double lam = scal[0];
   for (int i = tid; i < n; i += NT) p.qvec[off + i] = b0[i] - lam * b1[i];
   if (tid == 0) {
      p.status[mol] = 0;
      p.lamg[mol] = lam;
   }
}

// ------------------------------------------------------------------------------------------------
// Split path, factorisation phase, TMA-fed variant of eeq_factor_ll_kernel (same algorithm, same scratch layout).
//
// What bounds the kernel above is latency: every 16 x 16 update tile walks its two row strips of L through
// dependent 16-byte loads from L2 (~0.3 us each round), and the backward substitution does the same per block row.
// Here a PRODUCER warp streams those strips -- contiguous runs of fragment-ordered 8 x 8 blocks -- into shared
// memory with bulk asynchronous copies (cp.async.bulk, completion on mbarriers) one block column AHEAD of the
// consumers, so the DMMA warps read their operands from shared memory and never wait for L2:
//   * left-looking update of block column k0 = [old part: columns < k0 - 16, streamed] + [newest block column
//     k0 - 16, still in shared memory: the row solves write L back into the block column buffer Cb];
//   * chunk-major order: per chunk of CHB * 8 columns one B chunk (rows of the current block column) is shared by
//     all row tiles, whose A chunks follow; every updating warp keeps the accumulators of its (<= 4) row tiles in
//     registers across the chunks;
//   * per step ONE consumer warp (rotating) factors the diagonal tile (D) and takes no update work; the other
//     three are the updating ROLES 0..2, role r owns the row tiles 1 + r + 3 m and a PRIVATE ring of A chunks (so a
//     warp only ever waits for its own operands, and the producer runs ahead by the ring depth for each of them);
//     the B ring is shared (its empty barriers count three arrivals); row solves (S) use all four warps;
//   * backward substitution: the strip L(blk, :k0) of every step arrives through the private rings (chunk c to
//     role c mod 3), the inverted diagonal tile through a small ring of its own; the solving warp rotates as D does.
// Rings: 3 x NAW A slots and NB B slots of CHB KB (two block rows x CHB blocks), a 1.5 KB diagonal-tile ring.
// A failed pivot does not stop the pipeline (the copies in flight must land); the molecule is finished with NaNs
// and reports the pivot.
// ------------------------------------------------------------------------------------------------
namespace tma {
constexpr int NCONS = 4;                        // consumer warps
constexpr int NTHREADS = 32 * (NCONS + 1);      // + producer warp
#ifndef TMA_CHB
#define TMA_CHB 4
#endif
#ifndef TMA_NAW
#define TMA_NAW 3
#endif
constexpr int CHB = TMA_CHB;                    // 8 x 8 blocks per chunk and block row (CHB * 8 columns)
constexpr int SLOT = 2 * CHB * 64;              // doubles per ring slot
constexpr int NROLE = NCONS - 1;                // updating warps per step (one of the four factors the diagonal tile)
constexpr int NAW = TMA_NAW;                    // slots of the private A ring of each updating role
constexpr int NA = NROLE * NAW, NB = 2, ND = 2; // ring depths
constexpr int DSLOT = 3 * 64;                   // diagonal tile: blocks (R,R), (R+1,R), (R+1,R+1)
constexpr int NBAR = 2 * (NA + NB + ND);
__host__ __device__ inline size_t smem_doubles(int npa) {
   return (size_t)(NA + NB) * SLOT + (size_t)ND * DSLOT + (size_t)npa * KB + KB * TS + KB * KB + 2 * CBW + 2 * KB + KB +
          2 * KB + 8 + KB * TS + NBAR + 8;
}
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
   unsigned ok = 0;
   for (unsigned spin = 0; !ok; ++spin) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(sa(bar)), "r"(parity) : "memory");
      if (spin > (1u << 24)) __trap();  // a broken pipeline must not hang the device
   }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(sa(dst)), "l"(src), "r"(bytes), "r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
}  // namespace tma

#ifndef TMA_MINB
#define TMA_MINB 3
#endif
__global__ void __launch_bounds__(tma::NTHREADS, TMA_MINB) eeq_factor_tma_kernel(Args p) {
   using namespace tma;
   extern __shared__ __align__(16) double sm[];  // (ring slots: multiples of 4 KB from the base)
#ifdef TMA_PROF
   long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // wait full | U total | D | S | fence | syncs | backward | total
   const long long pf_t0 = clock64();
#define PF_BEGIN(v) const long long v = clock64()
#define PF_ADD(i, v) pf[i] += clock64() - v
#else
#define PF_BEGIN(v)
#define PF_ADD(i, v)
#endif
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int mol = p.perm[blockIdx.x];
   const int64_t off = p.offsets[mol];
   const int n = (int)(p.offsets[mol + 1] - off);
   const int npa = aug_order(n);
   const int nblk = npa / KB;
   const double charge = p.charge[mol];
   double* S = p.scratch + p.soff[mol];  // fragment-order blocks, S[lay(i, j)]
   double* ringA = sm;                   // [role][NAW] slots
   double* ringB = ringA + NA * SLOT;
   double* ringD = ringB + NB * SLOT;
   double* Cb = ringD + ND * DSLOT;      // block column buffer, ABSOLUTE rows: Cb[swz(row, c)]; later b0 | b1
   double* T = Cb + (size_t)p.np * KB;   // p.np: capacity (padded order) of this launch
   double* Li = T + KB * TS;
   double* cbuf = Li + KB * KB;
   double* rbuf = cbuf + 2 * CBW;
   double* dvec = rbuf + 2 * KB;
   double* wl = dvec + KB;
   double* scal = wl + 2 * KB;
   double* T2 = scal + 8;
   uint64_t* bars = reinterpret_cast<uint64_t*>(T2 + KB * TS);
   uint64_t *fullA = bars, *emptyA = fullA + NA, *fullB = emptyA + NA, *emptyB = fullB + NB, *fullD = emptyB + NB,
            *emptyD = fullD + ND;
   __shared__ int sfail;
   __shared__ int steps_done;  // factor steps whose row solves are complete and visible to the bulk copies
   if (p.status[mol] < 0) return;  // rejected by pair phase A (atomic number outside the model)
   if (tid == 0) {
      sfail = 0;
      steps_done = 0;
      for (int i = 0; i < NA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 1); }
      for (int i = 0; i < NB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, NROLE); }
      for (int i = 0; i < ND; ++i) { mbar_init(fullD + i, 1); mbar_init(emptyD + i, 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
   }
   if (warp == 0) {  // first diagonal tile: nothing to its left
      const int r = lane & 15, h = lane >> 4;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
         const int c = 2 * k + h;
         T[r * TS + c] = (c <= r) ? S[lay(r, c)] : 0.0;
      }
   }
   __syncthreads();
   // row tiles of role r among the tiles 1 .. nrt-1
   auto tiles_of = [](int nrt, int r) { return (nrt - 1 > r) ? (nrt - 1 - r + NROLE - 1) / NROLE : 0; };

   // =================================== producer ===================================
   if (warp == NCONS) {
      if (lane != 0) return;
      unsigned ia[NROLE] = {0, 0, 0}, ib = 0, id = 0;
      auto wait_steps = [&](int want) {
         for (unsigned spin = 0;; ++spin) {
            int v;
            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(sa(&steps_done)) : "memory");
            if (v >= want) break;
            if (spin > (1u << 26)) __trap();
         }
      };
      // one chunk of the two block rows R, R + 1 (blocks cb0 .. cb0 + CHB of each row, below block `lim`) into a slot
      auto load_chunk = [&](double* slot, uint64_t* full, int R, int cb0, int lim) {
         const int nb = min(CHB, lim - cb0);
         mbar_expect_tx(full, (unsigned)(2 * nb * 512));
         bulk_load(slot, S + 64 * (tri(R) + cb0), (unsigned)(nb * 512), full);
         bulk_load(slot + CHB * 64, S + 64 * (tri(R + 1) + cb0), (unsigned)(nb * 512), full);
      };
      auto load_a = [&](int role, int R, int cb0, int lim) {
         const unsigned slot = role * NAW + ia[role] % NAW, use = ia[role] / NAW;
         if (use > 0) mbar_wait(emptyA + slot, (use - 1) & 1);
         load_chunk(ringA + slot * SLOT, fullA + slot, R, cb0, lim);
         ++ia[role];
      };
      for (int k0 = 2 * KB; k0 < npa; k0 += KB) {  // factor steps with an old part: columns [0, k0 - 16)
         const int nrt = (npa - k0) / KB;
         if (nrt <= 1) break;
         const int kob = (k0 - KB) >> 3;  // blocks of the old part per block row
         wait_steps((k0 >> 4) - 1);       // block column k0 - 32 is final in global memory
         asm volatile("fence.proxy.async;" ::: "memory");
         for (int cb0 = 0; cb0 < kob; cb0 += CHB) {
            {
               const unsigned slot = ib % NB, use = ib / NB;
               if (use > 0) mbar_wait(emptyB + slot, (use - 1) & 1);
               load_chunk(ringB + slot * SLOT, fullB + slot, k0 >> 3, cb0, kob);
               ++ib;
            }
            for (int ti = 1; ti < nrt; ++ti) load_a((ti - 1) % NROLE, (k0 + KB * ti) >> 3, cb0, kob);
         }
      }
      // backward substitution: per step the diagonal tile (inverse) and the strip L(blk, :k0), highest columns first
      wait_steps(nblk);
      asm volatile("fence.proxy.async;" ::: "memory");
      for (int k0 = npa - KB; k0 >= 0; k0 -= KB) {
         const int R = k0 >> 3;
         {
            const unsigned slot = id % ND, use = id / ND;
            if (use > 0) mbar_wait(emptyD + slot, (use - 1) & 1);
            double* d = ringD + slot * DSLOT;
            mbar_expect_tx(fullD + slot, 3 * 512);
            bulk_load(d, S + 64 * (tri(R) + R), 512, fullD + slot);
            bulk_load(d + 64, S + 64 * (tri(R + 1) + R), 1024, fullD + slot);
            ++id;
         }
         const int nchb = (R + CHB - 1) / CHB;
         for (int c = 0; c < nchb; ++c) load_a(c % NROLE, R, (nchb - 1 - c) * CHB, R);
      }
      return;
   }

   // =================================== consumers ===================================
   const int g = lane >> 2, t = lane & 3;
   unsigned ia0 = 0, ia1 = 0, ia2 = 0, ib = 0, id = 0;  // ring positions, in lockstep with the producer
   for (int k0 = 0; k0 < npa; k0 += KB) {
      const int nrt = (npa - k0) / KB;
      const int kob = (nrt > 1 && k0 >= 2 * KB) ? (k0 - KB) >> 3 : 0;
      const int nch = (kob + CHB - 1) / CHB;
      const int dwarp = (blockIdx.x + (k0 >> 4)) & 3;
      if (warp == dwarp) {
         // ---- D: diagonal tile, Cholesky + inverse; the inverse goes to global in place of the block ----
         const bool last = (k0 + KB == npa);
         PF_BEGIN(td);
         const int bad = diag_factor_inv(T, Li, cbuf, rbuf, dvec, lane, last ? KB - 2 : KB, last);
         PF_ADD(2, td);
         if (bad && lane == 0 && sfail == 0) sfail = k0 + bad;
         __syncwarp();
         const int r = lane & 15, h = lane >> 4;
#pragma unroll
         for (int k = 0; k < 8; ++k) {
            const int c = 2 * k + h;
            if (c <= r) S[lay(k0 + r, k0 + c)] = Li[swz(r, c)];
         }
         if (last && lane < KB) {  // forward-substituted right-hand sides of the last block
            wl[lane] = (lane < KB - 2) ? T[(KB - 2) * TS + lane] : 0.0;
            wl[KB + lane] = (lane < KB - 2) ? T[(KB - 1) * TS + lane] : 0.0;
         }
      } else if (k0 == 0) {
         // ---- first block column: nothing to its left, the row tiles are the original entries ----
         const int idx = (warp - dwarp - 1) & 3;
         for (int ti = 1 + idx; ti < nrt; ti += NROLE) {
            const int i0 = KB * ti;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const double* src = S + lay(i0 + 8 * mt + g, 8 * ntl + 2 * t);
                  *reinterpret_cast<double2*>(Cb + swz(i0 + 8 * mt + g, 8 * ntl + 2 * t)) = make_double2(src[0], src[2]);
               }
         }
         if (idx == 0 && nrt > 1) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const int lr = 8 * mt + g, lc = 8 * ntl + 2 * t;
                  const double* src = S + lay(KB + lr, KB + ((lc <= lr) ? lc : 0));
                  T2[lr * TS + lc] = (lc <= lr) ? src[0] : 0.0;
                  T2[lr * TS + lc + 1] = (lc + 1 <= lr) ? src[2] : 0.0;
               }
         }
      } else if (nrt > 1) {
         // ---- U: left-looking update of the row tiles ti = 1 + idx + 3 m of this block column (and, with tile 1, of
         //      the next diagonal tile -> T2); idx = role of this warp among the three updating warps ----
         const int idx = (warp - dwarp - 1) & 3;
         unsigned mya = (idx == 0) ? ia0 : ((idx == 1) ? ia1 : ia2);
         PF_BEGIN(tu);
         double acc[4][2][2][2], acc0[2][2][2];
         // original entries first: their loads overlap the wait for the first chunk
#pragma unroll
         for (int m = 0; m < 4; ++m) {
            const int ti = 1 + idx + NROLE * m;
            const int i0 = k0 + KB * ti;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const int row = i0 + 8 * mt + g, col = k0 + 8 * ntl + 2 * t;
                  if (ti < nrt) {
                     const double* src = S + lay(row, col);
                     acc[m][mt][ntl][0] = -src[0];
                     acc[m][mt][ntl][1] = -src[2];  // (g, c + 1) is 2 doubles on (c even); rows are below the block
                  } else {
                     acc[m][mt][ntl][0] = acc[m][mt][ntl][1] = 0.0;
                  }
               }
         }
#pragma unroll
         for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ntl = 0; ntl < 2; ++ntl) {  // next diagonal tile, rows and columns k0 + 16 .. (role 0 only)
               const int row = k0 + KB + 8 * mt + g, col = k0 + KB + 8 * ntl + 2 * t;
               const double* src = S + lay(row, (col <= row) ? col : k0 + KB);
               acc0[mt][ntl][0] = (idx == 0 && col <= row) ? -src[0] : 0.0;
               acc0[mt][ntl][1] = (idx == 0 && col + 1 <= row) ? -src[2] : 0.0;
            }
         // old part: columns [0, k0 - 16) through the rings
         for (int c = 0; c < nch; ++c) {
            const int nb = min(CHB, kob - c * CHB);
            const unsigned sb = ib % NB;
            PF_BEGIN(tw);
            mbar_wait(fullB + sb, (ib / NB) & 1);
            PF_ADD(0, tw);
            const double2* Bs = reinterpret_cast<const double2*>(ringB + sb * SLOT) + lane;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
               const int ti = 1 + idx + NROLE * m;
               if (ti < nrt) {
                  const unsigned sl = idx * NAW + mya % NAW;
                  PF_BEGIN(tw2);
                  mbar_wait(fullA + sl, (mya / NAW) & 1);
                  PF_ADD(0, tw2);
                  const double2* As = reinterpret_cast<const double2*>(ringA + sl * SLOT) + lane;
                  for (int b = 0; b < nb; ++b) {
                     const double2 a0 = As[32 * b], a1 = As[32 * (CHB + b)];
                     const double2 c0 = Bs[32 * b], c1 = Bs[32 * (CHB + b)];
                     dmma884(acc[m][0][0][0], acc[m][0][0][1], a0.x, c0.x);
                     dmma884(acc[m][0][1][0], acc[m][0][1][1], a0.x, c1.x);
                     dmma884(acc[m][1][0][0], acc[m][1][0][1], a1.x, c0.x);
                     dmma884(acc[m][1][1][0], acc[m][1][1][1], a1.x, c1.x);
                     dmma884(acc[m][0][0][0], acc[m][0][0][1], a0.y, c0.y);
                     dmma884(acc[m][0][1][0], acc[m][0][1][1], a0.y, c1.y);
                     dmma884(acc[m][1][0][0], acc[m][1][0][1], a1.y, c0.y);
                     dmma884(acc[m][1][1][0], acc[m][1][1][1], a1.y, c1.y);
                     if (m == 0 && idx == 0) {  // tile 1 with itself: the next diagonal tile
                        dmma884(acc0[0][0][0], acc0[0][0][1], a0.x, a0.x);
                        dmma884(acc0[0][1][0], acc0[0][1][1], a0.x, a1.x);
                        dmma884(acc0[1][0][0], acc0[1][0][1], a1.x, a0.x);
                        dmma884(acc0[1][1][0], acc0[1][1][1], a1.x, a1.x);
                        dmma884(acc0[0][0][0], acc0[0][0][1], a0.y, a0.y);
                        dmma884(acc0[0][1][0], acc0[0][1][1], a0.y, a1.y);
                        dmma884(acc0[1][0][0], acc0[1][0][1], a1.y, a0.y);
                        dmma884(acc0[1][1][0], acc0[1][1][1], a1.y, a1.y);
                     }
                  }
                  __syncwarp();
                  if (lane == 0) mbar_arrive(emptyA + sl);
                  ++mya;
               }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyB + sb);
            ++ib;
         }
         // newest part: block column k0 - 16, whose rows (L after the row solves) are still in Cb
         {
            double bn[4][2];  // rows k0 .. k0 + 15 of it: the B operand of every tile
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
               const int kc = (4 * kk + t) ^ ((g & 3) << 2);
               bn[kk][0] = Cb[(k0 + g) * KB + kc];
               bn[kk][1] = Cb[(k0 + 8 + g) * KB + kc];
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
               const int ti = 1 + idx + NROLE * m;
               if (ti < nrt) {
                  const int i0 = k0 + KB * ti;
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                     const int kc = (4 * kk + t) ^ ((g & 3) << 2);
                     const double a0 = Cb[(i0 + g) * KB + kc], a1 = Cb[(i0 + 8 + g) * KB + kc];
                     dmma884(acc[m][0][0][0], acc[m][0][0][1], a0, bn[kk][0]);
                     dmma884(acc[m][0][1][0], acc[m][0][1][1], a0, bn[kk][1]);
                     dmma884(acc[m][1][0][0], acc[m][1][0][1], a1, bn[kk][0]);
                     dmma884(acc[m][1][1][0], acc[m][1][1][1], a1, bn[kk][1]);
                     if (m == 0 && idx == 0) {
                        dmma884(acc0[0][0][0], acc0[0][0][1], a0, a0);
                        dmma884(acc0[0][1][0], acc0[0][1][1], a0, a1);
                        dmma884(acc0[1][0][0], acc0[1][0][1], a1, a0);
                        dmma884(acc0[1][1][0], acc0[1][1][1], a1, a1);
                     }
                  }
               }
            }
         }
         __syncwarp();
         // acc = L L^T - A: store the negative into Cb (own rows) / T2
#pragma unroll
         for (int m = 0; m < 4; ++m) {
            const int ti = 1 + idx + NROLE * m;
            if (ti < nrt) {
               const int i0 = k0 + KB * ti;
#pragma unroll
               for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int ntl = 0; ntl < 2; ++ntl)
                     *reinterpret_cast<double2*>(Cb + swz(i0 + 8 * mt + g, 8 * ntl + 2 * t)) =
                        make_double2(-acc[m][mt][ntl][0], -acc[m][mt][ntl][1]);
            }
         }
         if (idx == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const int lr = 8 * mt + g, lc = 8 * ntl + 2 * t;
                  T2[lr * TS + lc] = (lc <= lr) ? -acc0[mt][ntl][0] : 0.0;
                  T2[lr * TS + lc + 1] = (lc + 1 <= lr) ? -acc0[mt][ntl][1] : 0.0;
               }
         }
         PF_ADD(1, tu);
      }
      ia0 += (unsigned)(nch * tiles_of(nrt, 0));
      ia1 += (unsigned)(nch * tiles_of(nrt, 1));
      ia2 += (unsigned)(nch * tiles_of(nrt, 2));
      if (warp == dwarp || k0 == 0 || nrt <= 1) ib += (unsigned)nch;  // (updating warps counted their B chunks above)
      PF_BEGIN(ts1);
      cons_sync();
      PF_ADD(5, ts1);
      PF_BEGIN(tsp);
      // ---- S: rows below the diagonal tile, L(i, blk) = Cb(i, :) inv(L_kk)^T, to the scratch AND back into Cb (the
      //      newest part of the next step); the warp of the first tile also completes the next diagonal tile ----
      for (int ti = 1 + warp; ti < nrt; ti += NCONS) {
         const int i0 = k0 + KB * ti;
         double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
         for (int kk = 0; kk < 4; ++kk) {
            const int kc = (4 * kk + t) ^ ((g & 3) << 2);
            const double a0 = Cb[(i0 + g) * KB + kc], a1 = Cb[(i0 + 8 + g) * KB + kc];
            const double c1 = Li[(8 + g) * KB + kc];
            if (kk < 2) {  // inv(L)(n, k) = 0 for k > n
               const double c0 = Li[g * KB + kc];
               dmma884(acc[0][0][0], acc[0][0][1], a0, c0);
               dmma884(acc[1][0][0], acc[1][0][1], a1, c0);
            }
            dmma884(acc[0][1][0], acc[0][1][1], a0, c1);
            dmma884(acc[1][1][0], acc[1][1][1], a1, c1);
         }
         __syncwarp();
#pragma unroll
         for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ntl = 0; ntl < 2; ++ntl) {
               double* dst = S + lay(i0 + 8 * mt + g, k0 + 8 * ntl + 2 * t);
               dst[0] = acc[mt][ntl][0];
               dst[2] = acc[mt][ntl][1];
               *reinterpret_cast<double2*>(Cb + swz(i0 + 8 * mt + g, 8 * ntl + 2 * t)) = make_double2(acc[mt][ntl][0], acc[mt][ntl][1]);
            }
         if (ti == 1) {
            __syncwarp();
            double d2[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
               const int kc = (4 * kk + t) ^ ((g & 3) << 2);
               const double a0 = Cb[(i0 + g) * KB + kc], a1 = Cb[(i0 + 8 + g) * KB + kc];  // A(m, k) and B(k, n) = A(n, k)
               dmma884(d2[0][0][0], d2[0][0][1], a0, a0);
               dmma884(d2[0][1][0], d2[0][1][1], a0, a1);
               dmma884(d2[1][0][0], d2[1][0][1], a1, a0);
               dmma884(d2[1][1][0], d2[1][1][1], a1, a1);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
               for (int ntl = 0; ntl < 2; ++ntl) {
                  const int lr = 8 * mt + g, lc = 8 * ntl + 2 * t;
                  T[lr * TS + lc] = (lc <= lr) ? T2[lr * TS + lc] - d2[mt][ntl][0] : 0.0;
                  T[lr * TS + lc + 1] = (lc + 1 <= lr) ? T2[lr * TS + lc + 1] - d2[mt][ntl][1] : 0.0;
               }
         }
      }
      PF_ADD(3, tsp);
      // what this step wrote to the scratch must be visible to the bulk copies (async proxy) the producer issues next
      PF_BEGIN(tf);
      asm volatile("fence.proxy.async;" ::: "memory");
      PF_ADD(4, tf);
      PF_BEGIN(ts2);
      cons_sync();
      PF_ADD(5, ts2);
      if (tid == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(sa(&steps_done)), "r"((k0 >> 4) + 1) : "memory");
   }

   // ---- backward substitution L^T [y z] = [w_x w_1], right-looking by blocks; Cb is free now: b0 | b1 ----
   PF_BEGIN(tbw);
   double* b0 = Cb;
   double* b1 = Cb + npa;
   {
      const int kl = npa - KB;
      for (int i = tid; i < npa; i += 32 * NCONS) {
         b0[i] = (i < kl) ? S[lay(npa - 2, i)] : wl[i - kl];
         b1[i] = (i < kl) ? S[lay(npa - 1, i)] : wl[KB + i - kl];
      }
   }
   cons_sync();
   for (int k0 = npa - KB; k0 >= 0; k0 -= KB) {
      const int R = k0 >> 3;
      const int solver = (blockIdx.x + (k0 >> 4)) & 3;
      const int nchb = (R + CHB - 1) / CHB;
      if (warp == solver) {  // y_blk = inv(L_kk)^T w_blk ; lane = (r, rhs h); inverse tile from the diagonal ring
         const unsigned sd = id % ND;
         mbar_wait(fullD + sd, (id / ND) & 1);
         const double* dt = ringD + sd * DSLOT;
         const int r = lane & 15, h = lane >> 4;
         const double* wv = (h ? b1 : b0) + k0;
         double v = 0.0;
#pragma unroll
         for (int c = 0; c < KB; ++c)
            if (c >= r) {  // entry (k0 + c, k0 + r): block (R,R) | (R+1,R) | (R+1,R+1)
               const int blk = (c < 8) ? 0 : ((r < 8) ? 1 : 2);
               v = fma(dt[64 * blk + ((((c & 7) << 2) + (r & 3)) << 1) + ((r >> 2) & 1)], wv[c], v);
            }
         __syncwarp();
         (h ? b1 : b0)[k0 + r] = v;
         __syncwarp();
         if (lane == 0) mbar_arrive(emptyD + sd);
      }
      ++id;
      cons_sync();
      if (warp != solver) {  // w(:k0) -= L(blk, :k0)^T y_blk ; chunk c (highest columns first) goes to role c mod 3
         const int idx = (warp - solver - 1) & 3;
         unsigned mya = (idx == 0) ? ia0 : ((idx == 1) ? ia1 : ia2);
         const int cc = lane & 7, rg = lane >> 3;  // column within the block, group of 4 rows
         for (int c = idx; c < nchb; c += NROLE) {
            const int cb0 = (nchb - 1 - c) * CHB;
            const int nb = min(CHB, R - cb0);
            const unsigned sl = idx * NAW + mya % NAW;
            mbar_wait(fullA + sl, (mya / NAW) & 1);
            const double* As = ringA + sl * SLOT;
            for (int b = 0; b < nb; ++b) {
               double s0 = 0.0, s1 = 0.0;
#pragma unroll
               for (int q = 0; q < 4; ++q) {
                  const int r = 4 * rg + q;  // row within the 16
                  const double l = As[(r >> 3) * CHB * 64 + 64 * b + ((((r & 7) << 2) + (cc & 3)) << 1) + ((cc >> 2) & 1)];
                  s0 = fma(l, b0[k0 + r], s0);
                  s1 = fma(l, b1[k0 + r], s1);
               }
               s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
               s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
               s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
               s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
               if (rg == 0) {
                  const int j = 8 * (cb0 + b) + cc;
                  b0[j] -= s0;
                  b1[j] -= s1;
               }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyA + sl);
            ++mya;
         }
      }
      ia0 += (unsigned)((nchb + NROLE - 1) / NROLE);
      ia1 += (unsigned)((nchb + NROLE - 2) / NROLE);
      ia2 += (unsigned)((nchb + NROLE - 3) / NROLE);
      cons_sync();
   }
   PF_ADD(6, tbw);
#ifdef TMA_PROF
   pf[7] = clock64() - pf_t0;
   if (lane == 0 && blockIdx.x % 1000 == 7)
      printf("[tma] blk %d n %d npa %d warp %d: wait_full %lld U %lld D %lld S %lld fence %lld syncs %lld backward %lld total %lld\n",
             (int)blockIdx.x, n, npa, warp, pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7]);
#endif
   const int bad = sfail;
   if (bad) {
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = tid; i < n; i += 32 * NCONS) {
         p.qvec[off + i] = nan;
         if (p.energy) p.energy[off + i] = nan;
         if (p.gradient) p.gradient[3 * (off + i)] = p.gradient[3 * (off + i) + 1] = p.gradient[3 * (off + i) + 2] = nan;
      }
      if (tid == 0) {
         p.status[mol] = bad;
         p.lamg[mol] = nan;
      }
      return;
   }
   if (warp == 0) {
      double sy = 0.0, sz = 0.0;
      for (int i = lane; i < n; i += 32) { sy += b0[i]; sz += b1[i]; }
      sy = warp_sum(sy);
      sz = warp_sum(sz);
      if (lane == 0) scal[0] = (sy - charge) / sz;
   }
   cons_sync();
   const double lam = scal[0];
   for (int i = tid; i < n; i += 32 * NCONS) p.qvec[off + i] = b0[i] - lam * b1[i];
   if (tid == 0) {
      p.status[mol] = 0;
      p.lamg[mol] = lam;
   }
}

static void launch_factor_tma(mchrg_b200_context* ctx, const Args& a, int count, int npa_max) {
   const size_t smem = tma::smem_doubles(npa_max) * sizeof(double);
   ctx->ensure_smem(eeq_factor_tma_kernel, tma::smem_doubles(NPA_MAX) * sizeof(double), /*max_carveout=*/true);
   Args b = a;
   b.np = npa_max;
   MCHRG_LAUNCH(ctx, eeq_factor_tma_kernel, (unsigned)count, tma::NTHREADS, smem, b);
}

template <int NT>
static void launch_factor_ll(mchrg_b200_context* ctx, const Args& a, int count, int npa_max) {
   // MCHRG_B200_FACTOR_SMEM_KB: ask for at least this much shared memory per CTA (lowers the number of resident
   // molecules per SM; residency experiments only)
   const size_t floor_b = (size_t)std::max<long long>(0, env_ll("MCHRG_B200_FACTOR_SMEM_KB", 0)) * 1024;
   const size_t smem = std::max(smem_doubles_ll(npa_max) * sizeof(double), floor_b);
   ctx->ensure_smem(eeq_factor_ll_kernel<NT>, std::max(smem_doubles_ll(NPA_MAX) * sizeof(double), floor_b), /*max_carveout=*/true);
   MCHRG_LAUNCH(ctx, eeq_factor_ll_kernel<NT>, (unsigned)count, NT, smem, a);
}

#ifndef LL_NT_DEF
#define LL_NT_DEF 128
#endif

constexpr int LL_NT = LL_NT_DEF;  // threads per molecule in the factorisation kernel

// Factorisation launches of one chunk: the class-sorted permutation is cut into a few groups of size classes so
// that the block column buffer (and with it the number of resident molecules per SM) matches the molecule size.
static void launch_factor_groups(mchrg_b200_context* ctx, const Args& a, const int32_t* dperm,
                                 const std::vector<size_t>& cls_start, const std::vector<int>& cls_count) {
   constexpr int NCLASS = NP_MAX / KB;
   // (the TMA-fed kernel's shared memory grows with the class: three groups keep small molecules at high residency)
   const char* fkind = std::getenv("MCHRG_B200_FACTOR");
   const int group = (int)std::max<long long>(1, env_ll("MCHRG_B200_FACTOR_GROUP", (fkind && fkind[0] == 't') ? 4 : 13));
   for (int chi = NCLASS; chi >= 1; chi -= group) {  // perm holds the classes in descending order
      const int clo = std::max(1, chi - group + 1);
      int count = 0;
      for (int c = clo; c <= chi; ++c) count += cls_count[c];
      if (count == 0) continue;
      Args b = a;
      b.np = NPA_MAX;
      b.perm = dperm + cls_start[chi];
      // MCHRG_B200_FACTOR = tma: the TMA-fed variant (measured slower: 16.9 vs 10.1 ms per 100k molecules, see its header)
      const char* fk = std::getenv("MCHRG_B200_FACTOR");
      if (fk && fk[0] == 't') launch_factor_tma(ctx, b, count, std::min(NPA_MAX, KB * (chi + 1)));
      else launch_factor_ll<LL_NT>(ctx, b, count, std::min(NPA_MAX, KB * (chi + 1)));
   }
}

__host__ inline size_t smem_doubles_pair(int npv) { return (size_t)NVEC * npv + 8 + AUX_DOUBLES; }

template <int NT, int PH>
static void launch_class(mchrg_b200_context* ctx, const Args& a, int count) {
   const bool heavy = (PH == 0);
   const size_t smem = (heavy ? smem_doubles(a.np) : smem_doubles_pair(a.np)) * sizeof(double);
   // same shared-memory carveout as the factorisation kernel: CTAs of kernels that ask for different
   // L1/shared splits cannot share an SM
   ctx->ensure_smem(eeq_batch_kernel<NT, PH>, (heavy ? smem_doubles(NP_MAX) : smem_doubles_pair(NPA_MAX)) * sizeof(double),
                    /*max_carveout=*/true);
   MCHRG_LAUNCH(ctx, (eeq_batch_kernel<NT, PH>), (unsigned)count, NT, smem, a);
}

}  // namespace batch
}  // namespace mchrg

namespace mchrg {
namespace batch {

// what the split path derives from the molecule offsets alone (kept on the context between calls)
struct SplitPlan {
   std::vector<int64_t> hoff;
   int maxchunk = 0, nchunk = 0;
   std::vector<std::vector<size_t>> starts;  // per chunk: start of every size class in the permutation
   std::vector<std::vector<int>> counts;     // per chunk: molecules per size class
   std::vector<size_t> chunk_off;            // chunk boundaries in the permutation
   int64_t nscr = 0;                         // doubles of scratch for the matrices
   int32_t* dperm = nullptr;                 // device: class-sorted molecule ids
   int64_t* d_soff = nullptr;                // device: start of every molecule's matrix in the scratch
   ~SplitPlan() {
      if (dperm) cudaFree(dperm);
      if (d_soff) cudaFree(d_soff);
   }
};

// bucket the molecules [m0, m1) by padded order; returns the class-sorted permutation (chunk-local ids)
// and the start of every class in it.  Molecules above NP_MAX atoms are left out (the entry point sends
// them through the single-system path afterwards).
static void bucket(const std::vector<int64_t>& hoff, int32_t m0, int32_t m1, std::vector<int32_t>& perm,
                   std::vector<size_t>& cls_start, std::vector<int>& cls_count) {
   constexpr int NCLASS = NP_MAX / KB;
   std::vector<std::vector<int32_t>> cls(NCLASS + 1);
   for (int32_t m = m0; m < m1; ++m) {
      const int64_t n = hoff[m + 1] - hoff[m];
      if (n > NP_MAX) continue;
      cls[(n + KB - 1) / KB].push_back(m - m0);
   }
   perm.clear();
   cls_start.assign(NCLASS + 1, 0);
   cls_count.assign(NCLASS + 1, 0);
   for (int c = NCLASS; c >= 1; --c) {  // largest molecules first: their CTAs run longest
      cls_start[c] = perm.size();
      cls_count[c] = (int)cls[c].size();
      perm.insert(perm.end(), cls[c].begin(), cls[c].end());
   }
}

// Launch every size class of one chunk: the shared-memory-heavy classes (one CTA per SM, latency bound in the
// factorisation) on `big`, the light classes on `small`; their CTAs fit beside each other on an SM, so the
// light ones fill the idle FP64 issue slots.  Largest classes first.
template <int PH>
static void launch_chunk(mchrg_b200_context* ctx, Args a, const int32_t* dperm, const std::vector<size_t>& cls_start,
                         const std::vector<int>& cls_count, cudaStream_t big, cudaStream_t small) {
   for (int c = NP_MAX / KB; c >= 1; --c) {
      if (cls_count[c] == 0) continue;
      a.np = c * KB;
      a.perm = dperm + cls_start[c];
      StreamScope on(ctx, a.np > 112 ? big : small);
      if (a.np <= 64) launch_class<128, PH>(ctx, a, cls_count[c]);
      else if (a.np <= 128) launch_class<256, PH>(ctx, a, cls_count[c]);
      else launch_class<512, PH>(ctx, a, cls_count[c]);
   }
}

// pair phases (PH 1 / 3): one launch over all molecules of the chunk (the class-sorted permutation keeps
// neighbouring CTAs similar in cost); shared memory holds the per-atom vectors only
template <int PH>
static void launch_pairs(mchrg_b200_context* ctx, Args a, const int32_t* dperm, int count) {
   a.np = NPA_MAX;
   a.perm = dperm;
   launch_class<256, PH>(ctx, a, count);
}


// scratch layout of the split path for the molecules [m0, m1): soff[m - m0] = start of molecule m's packed
// matrix (order padded to 16); returns the total number of doubles
static int64_t plan_scratch(const std::vector<int64_t>& hoff, int32_t m0, int32_t m1, std::vector<int64_t>& soff) {
   soff.resize((size_t)(m1 - m0));
   int64_t total = 0;
   for (int32_t m = m0; m < m1; ++m) {
      const int64_t n = hoff[m + 1] - hoff[m];
      soff[m - m0] = total;
      if (n > NP_MAX) continue;  // not served by the batch kernels
      total += lay_size(aug_order((int)n));
   }
   return total;
}

// One chunk of the split path.  `a` addresses the chunk (offsets / perm ids in the same id space).
//   pair phase A on `sa`  ->  factorisation on the high-priority stream  ->  pair phase C on `sc`
// `ready` (may be null) gates phase A (inputs staged); returns the event that marks the chunk's outputs complete.
static cudaEvent_t run_chunk_split(mchrg_b200_context* ctx, const Args& a, const int32_t* dperm, int count,
                                   const std::vector<size_t>& cls_start, const std::vector<int>& cls_count,
                                   cudaEvent_t ready, cudaStream_t sa, cudaStream_t sc) {
   if (ready) MCHRG_CUDA(cudaStreamWaitEvent(sa, ready, 0));
   {
      StreamScope on(ctx, sa);
      launch_pairs<1>(ctx, a, dperm, count);
   }
   cudaEvent_t eA = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(eA, sa));
   MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_hi, eA, 0));
   {
      StreamScope on(ctx, ctx->stream_hi);
      launch_factor_groups(ctx, a, dperm, cls_start, cls_count);
   }
   cudaEvent_t eB = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(eB, ctx->stream_hi));
   MCHRG_CUDA(cudaStreamWaitEvent(sc, eB, 0));
   if (a.energy || a.gradient) {
      StreamScope on(ctx, sc);
      launch_pairs<3>(ctx, a, dperm, count);
   }
   cudaEvent_t eC = ctx->next_event();
   MCHRG_CUDA(cudaEventRecord(eC, sc));
   return eC;
}

static int batch_mode() {  // 0 auto, 1 fused, 2 split  (MCHRG_B200_BATCH_MODE = fused | split)
   const char* e = std::getenv("MCHRG_B200_BATCH_MODE");
   if (!e) return 0;
   if (e[0] == 'f') return 1;
   if (e[0] == 's') return 2;
   return 0;
}

}  // namespace batch
}  // namespace mchrg

using namespace mchrg;

extern "C" int mchrg_b200_singlepoint_batch(const mchrg_b200_model* model, int32_t nmol, const int64_t* offsets,
                                            const int32_t* num, const double* xyz, const double* charge, double* qvec,
                                            double* energy, double* gradient, int32_t* status) {
   if (!model) return MCHRG_B200_ERR_ARG;
   mchrg_b200_context* ctx = model->ctx;
   std::vector<int32_t> large;  // molecules above NP_MAX atoms: single-system path, after the batch kernels
   std::vector<int64_t> hoff;
   const int rc = guarded(ctx, [&]() {
      if (nmol <= 0 || !offsets || !num || !xyz || !charge || !qvec || !status)
         fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: missing argument");
      if (model->kind != 1) fail(MCHRG_B200_ERR_MODEL, "singlepoint_batch serves EEQ2019-type models only");
      // molecule sizes are needed on the host to form the size classes
      hoff.resize((size_t)nmol + 1);
      if (mchrg_b200_context::is_device_ptr(offsets))
         MCHRG_CUDA(cudaMemcpy(hoff.data(), offsets, hoff.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
      else
         std::copy(offsets, offsets + nmol + 1, hoff.begin());
      const size_t ntot = (size_t)hoff[nmol];
      for (int32_t m = 0; m < nmol; ++m) {
         const int64_t n = hoff[m + 1] - hoff[m];
         if (n <= 0) fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: molecule %d is empty", m);
         if (n > INT32_MAX) fail(MCHRG_B200_ERR_ARG, "singlepoint_batch: molecule %d is too large", m);
         if (n > batch::NP_MAX) large.push_back(m);
      }
      if (large.size() == (size_t)nmol) return 0;
      batch::Args a;
      a.scratch = nullptr; a.soff = nullptr; a.vecg = nullptr; a.lamg = nullptr; a.ntot = 0;
      a.par = model->d_par;
      a.etab = ctx->erf_tab;
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
      const int mode = batch::batch_mode();
      const bool split = (mode == 2) || (mode == 0 && nmol >= 4096);
      if (split && (!host_io || ntot < (size_t)1 << 18)) {
         // Split path, data on the device (or small enough to stage at once): chunks of molecules flow through
         //   pair phase A -> factorisation -> pair phase C   on different streams, so the FP64-bound pair
         // kernels of one chunk share the SMs with the latency-bound factorisation CTAs of its neighbours.
         a.offsets = ctx->in(offsets, (size_t)nmol + 1);
         a.num = ctx->in(num, ntot);
         a.xyz = ctx->in(xyz, 3 * ntot);
         a.charge = ctx->in(charge, (size_t)nmol);
         a.qvec = ctx->out(qvec, ntot);
         a.energy = ctx->out(energy, ntot);
         a.gradient = ctx->out(gradient, 3 * ntot);
         a.status = ctx->out(status, (size_t)nmol);
         // (measured on the 100k-molecule mix: 1 / 2 / 4 / 8 / 16 chunks -> 24.16 / 24.12 / 24.10 / 24.20 / 24.37 ms per step:
         //  every launch has a tail, and kernels of neighbouring chunks overlap only in those tails)
         const int maxchunk = (int)std::max<long long>(1, env_ll("MCHRG_B200_CHUNKS", 4));
         // The size classes, the class-sorted permutation and the scratch layout depend on `offsets` alone: a caller that
         // steps the same batch again (geometry updates of a fixed set of molecules) reuses them -- the plan is kept on
         // the context and rebuilt when the offsets differ.
         auto* plan = static_cast<batch::SplitPlan*>(ctx->batch_plan.get());
         if (!plan || plan->maxchunk != maxchunk || plan->hoff != hoff) {
            auto fresh = std::make_shared<batch::SplitPlan>();
            plan = fresh.get();
            plan->hoff = hoff;
            plan->maxchunk = maxchunk;
            std::vector<int64_t> soff;
            plan->nscr = batch::plan_scratch(hoff, 0, nmol, soff);
            // (at least 4096 molecules per chunk by default; an explicit MCHRG_B200_CHUNKS may go down to 256 per chunk --
            //  the L2-residency experiment: ~500 molecules are 30 MB of scratch)
            const int64_t per_chunk = std::getenv("MCHRG_B200_CHUNKS") ? 256 : 4096;
            plan->nchunk = (int)std::min<int64_t>(maxchunk, std::max<int64_t>(1, nmol / per_chunk));
            std::vector<int32_t> all_perm;
            plan->starts.resize(plan->nchunk);
            plan->counts.resize(plan->nchunk);
            plan->chunk_off.assign(plan->nchunk + 1, 0);
            for (int c = 0; c < plan->nchunk; ++c) {
               const int32_t m0 = (int32_t)((int64_t)nmol * c / plan->nchunk), m1 = (int32_t)((int64_t)nmol * (c + 1) / plan->nchunk);
               batch::bucket(hoff, m0, m1, perm, plan->starts[c], plan->counts[c]);
               for (auto& v : perm) v += m0;  // global molecule ids
               plan->chunk_off[c] = all_perm.size();
               all_perm.insert(all_perm.end(), perm.begin(), perm.end());
            }
            plan->chunk_off[plan->nchunk] = all_perm.size();
            MCHRG_CUDA(cudaMalloc(&plan->d_soff, (size_t)nmol * sizeof(int64_t)));
            MCHRG_CUDA(cudaMalloc(&plan->dperm, std::max<size_t>(1, all_perm.size()) * sizeof(int32_t)));
            MCHRG_CUDA(cudaMemcpy(plan->d_soff, soff.data(), (size_t)nmol * sizeof(int64_t), cudaMemcpyHostToDevice));
            MCHRG_CUDA(cudaMemcpy(plan->dperm, all_perm.data(), all_perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            ctx->batch_plan = fresh;
         }
         const int nchunk = plan->nchunk;
         const auto& starts = plan->starts;
         const auto& counts = plan->counts;
         const auto& chunk_off = plan->chunk_off;
         const int32_t* dperm = plan->dperm;
         a.scratch = ctx->alloc<double>((size_t)plan->nscr);
         a.soff = plan->d_soff;
         a.vecg = ctx->alloc<double>(3 * ntot);
         a.lamg = ctx->alloc<double>(nmol);
         a.ntot = (int64_t)ntot;
         cudaEvent_t staged = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(staged, ctx->stream));
         if (env_ll("MCHRG_B200_TRACE", 0) > 0) {  // phases of the whole batch one after the other, each timed on its own
            cudaEvent_t t[4];
            for (auto& e : t) MCHRG_CUDA(cudaEventCreate(&e));
            batch::Args b = a;
            MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));
            const uint64_t l0 = ctx->launches;
            MCHRG_CUDA(cudaEventRecord(t[0], ctx->stream));
            for (int c = 0; c < nchunk; ++c) batch::launch_pairs<1>(ctx, b, dperm + chunk_off[c], (int)(chunk_off[c + 1] - chunk_off[c]));
            MCHRG_CUDA(cudaEventRecord(t[1], ctx->stream));
            const uint64_t l1 = ctx->launches;
            for (int c = 0; c < nchunk; ++c) {
               batch::launch_factor_groups(ctx, b, dperm + chunk_off[c], starts[c], counts[c]);
            }
            MCHRG_CUDA(cudaEventRecord(t[2], ctx->stream));
            const uint64_t l2 = ctx->launches;
            ctx->trace_launches[0] = (int)(l1 - l0);
            ctx->trace_launches[1] = (int)(l2 - l1);
            ctx->trace_launches[2] = (int)(l1 - l0);
            for (int c = 0; c < nchunk; ++c) batch::launch_pairs<3>(ctx, b, dperm + chunk_off[c], (int)(chunk_off[c + 1] - chunk_off[c]));
            MCHRG_CUDA(cudaEventRecord(t[3], ctx->stream));
            MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));
            float ms[3];
            for (int k = 0; k < 3; ++k) MCHRG_CUDA(cudaEventElapsedTime(&ms[k], t[k], t[k + 1]));
            for (int k = 0; k < 3; ++k) ctx->trace_ms[k] = ms[k];
            if (env_ll("MCHRG_B200_TRACE", 0) > 1)
               fprintf(stderr, "[mchrg_b200 trace] nmol=%d pairA=%.2f ms factor=%.2f ms pairC=%.2f ms\n", nmol, ms[0], ms[1], ms[2]);
            for (auto& e : t) cudaEventDestroy(e);
            return 0;
         }
         for (int c = 0; c < nchunk; ++c) {
            const int cnt = (int)(chunk_off[c + 1] - chunk_off[c]);
            if (cnt == 0) continue;
            // phase C runs on the main stream, which finish() synchronises: the returned event is not needed
            batch::run_chunk_split(ctx, a, dperm + chunk_off[c], cnt, starts[c], counts[c], c == 0 ? staged : nullptr,
                                   ctx->stream_aux, ctx->stream);
         }
         MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));
         return 0;
      }
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
         batch::launch_chunk<0>(ctx, a, dperm, cls_start, cls_count, ctx->stream, ctx->stream_aux);
         cudaEvent_t aux_done = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(aux_done, ctx->stream_aux));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, aux_done, 0));
         MCHRG_CUDA(cudaStreamSynchronize(ctx->stream));  // perm (host vector) must outlive the async copy
         return 0;
      }

      // Host buffers: pipeline chunks of molecules so that the H2D copy of chunk c+1 and the D2H copy of
      // chunk c-1 overlap the kernels of chunk c (asynchronous when the caller's buffers are pinned).
      const int host_chunks = (int)std::max<long long>(0, env_ll("MCHRG_B200_HOST_CHUNKS", 0));
      // ~256k atoms per chunk (measured on the 100k-molecule bench: 8 / 16 / 24 / 32 / 48 chunks give 3.94 / 4.08 / 4.14 /
      // 4.16 / 4.18 M molecules/s end to end: only the first H2D and the last D2H chunk stay exposed)
      const int nchunk = host_chunks ? host_chunks : (int)std::min<size_t>(48, std::max<size_t>(2, ntot >> 18));
      std::vector<int32_t> bounds(nchunk + 1, 0);
      for (int c = 1; c < nchunk; ++c) {  // split by atoms
         const int64_t target = (int64_t)(ntot * (size_t)c / nchunk);
         bounds[c] = (int32_t)(std::lower_bound(hoff.begin(), hoff.end(), target) - hoff.begin());
         if (bounds[c] < bounds[c - 1]) bounds[c] = bounds[c - 1];
      }
      bounds[nchunk] = nmol;
      std::vector<std::vector<int64_t>> loc_off(nchunk);
      std::vector<std::vector<int32_t>> loc_perm(nchunk);
      std::vector<std::vector<int64_t>> loc_soff(nchunk);
      cudaEvent_t start = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(start, ctx->stream));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_in, start, 0));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_out, start, 0));
      for (int c = 0; c < nchunk; ++c) {
         const int32_t m0 = bounds[c], m1 = bounds[c + 1], mc = m1 - m0;
         if (mc <= 0) continue;
         const size_t a0 = (size_t)hoff[m0], na = (size_t)(hoff[m1] - hoff[m0]);
         batch::bucket(hoff, m0, m1, loc_perm[c], cls_start, cls_count);
         const int nserved = (int)loc_perm[c].size();
         if (nserved == 0) continue;  // only molecules above NP_MAX atoms here
         loc_off[c].resize((size_t)mc + 1);
         for (int32_t m = 0; m <= mc; ++m) loc_off[c][m] = hoff[m0 + m] - hoff[m0];
         int64_t* d_off = ctx->alloc<int64_t>((size_t)mc + 1);
         int32_t* d_perm = ctx->alloc<int32_t>(nserved);
         int32_t* d_num = ctx->alloc<int32_t>(na);
         double* d_xyz = ctx->alloc<double>(3 * na);
         double* d_chg = ctx->alloc<double>(mc);
         double* d_q = ctx->alloc<double>(na);
         double* d_e = energy ? ctx->alloc<double>(na) : nullptr;
         double* d_g = gradient ? ctx->alloc<double>(3 * na) : nullptr;
         int32_t* d_st = ctx->alloc<int32_t>(mc);
         cudaStream_t in = ctx->stream_in, out = ctx->stream_out;
         MCHRG_CUDA(cudaMemcpyAsync(d_off, loc_off[c].data(), ((size_t)mc + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_perm, loc_perm[c].data(), (size_t)nserved * sizeof(int32_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_num, num + a0, na * sizeof(int32_t), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_xyz, xyz + 3 * a0, 3 * na * sizeof(double), cudaMemcpyHostToDevice, in));
         MCHRG_CUDA(cudaMemcpyAsync(d_chg, charge + m0, (size_t)mc * sizeof(double), cudaMemcpyHostToDevice, in));
         cudaEvent_t ev_in = ctx->next_event();
         MCHRG_CUDA(cudaEventRecord(ev_in, in));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
         MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ev_in, 0));
         a.offsets = d_off; a.num = d_num; a.xyz = d_xyz; a.charge = d_chg;
         a.qvec = d_q; a.energy = d_e; a.gradient = d_g; a.status = d_st;
         cudaEvent_t ev_done;
         if (split) {
            std::vector<int64_t> soff;
            const int64_t nscr = batch::plan_scratch(hoff, m0, m1, soff);
            loc_soff[c] = soff;
            a.scratch = ctx->alloc<double>((size_t)nscr);
            int64_t* d_soff = ctx->alloc<int64_t>(mc);
            MCHRG_CUDA(cudaMemcpyAsync(d_soff, loc_soff[c].data(), (size_t)mc * sizeof(int64_t), cudaMemcpyHostToDevice, in));
            a.soff = d_soff;
            a.vecg = ctx->alloc<double>(3 * na);
            a.lamg = ctx->alloc<double>(mc);
            a.ntot = (int64_t)na;
            cudaEvent_t ev_in2 = ctx->next_event();
            MCHRG_CUDA(cudaEventRecord(ev_in2, in));
            ev_done = batch::run_chunk_split(ctx, a, d_perm, nserved, cls_start, cls_count, ev_in2, ctx->stream_aux, ctx->stream);
         } else {
            batch::launch_chunk<0>(ctx, a, d_perm, cls_start, cls_count, ctx->stream, ctx->stream_aux);
            cudaEvent_t aux_done = ctx->next_event();
            MCHRG_CUDA(cudaEventRecord(aux_done, ctx->stream_aux));
            MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, aux_done, 0));
            ev_done = ctx->next_event();
            MCHRG_CUDA(cudaEventRecord(ev_done, ctx->stream));
         }
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
   if (rc != 0 || large.empty()) return rc;
   // Molecules the one-CTA-per-molecule kernels cannot hold: the single-system path (blocked Cholesky in HBM),
   // one after the other, on the same slices of the caller's arrays (host or device).
   for (int32_t m : large) {
      const size_t a0 = (size_t)hoff[m], n = (size_t)(hoff[m + 1] - hoff[m]);
      mchrg_b200_structure mol{};
      mol.nat = (int32_t)n;
      mol.nid = model->nid;
      mol.id = num + a0;  // species index == atomic number for the models this entry point serves
      mol.xyz = xyz + 3 * a0;
      if (mchrg_b200_context::is_device_ptr(charge))
         cudaMemcpy(&mol.charge, charge + m, sizeof(double), cudaMemcpyDeviceToHost);
      else
         mol.charge = charge[m];
      if (energy) {  // the single-system entry point accumulates the energy; the batch overwrites it
         if (mchrg_b200_context::is_device_ptr(energy)) cudaMemset(energy + a0, 0, n * sizeof(double));
         else std::fill(energy + a0, energy + a0 + n, 0.0);
      }
      double sigma[9] = {0};
      const int32_t st = mchrg_b200_singlepoint(model, &mol, nullptr, nullptr, energy ? energy + a0 : nullptr,
                                                gradient ? gradient + 3 * a0 : nullptr, gradient ? sigma : nullptr,
                                                qvec + a0, nullptr, nullptr);
      if (mchrg_b200_context::is_device_ptr(status))
         cudaMemcpy(status + m, &st, sizeof(int32_t), cudaMemcpyHostToDevice);
      else
         status[m] = st;
   }
   return 0;
}
