// Micro-benchmark: latency of the 16x16 Cholesky+inverse warp routine (copy of diag_factor_inv in batch.cu)
// in isolation (1 warp on an SM) and with co-resident warps.  nvcc -arch=sm_100a -O3 -o diag_latency diag_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int KB = 16, TS = 17;
__device__ __forceinline__ double shfl_d(double v, int src) {
   const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
   const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
   return __hiloint2double(hi, lo);
}
__device__ __forceinline__ int swz(int row, int col) { return row * KB + (col ^ ((row & 3) << 2)); }
constexpr int CBW = 20;
__device__ __forceinline__ int diag_factor_inv(double* T, double* Li, double* cbuf, double* rbuf, double* dvec, int lane,
                                               int nfree, bool want_l) {
   const int r = lane & 15, h = lane >> 4;
   double m[8], x[8];
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      const int c = 2 * k + h;
      m[k] = (c <= r) ? T[r * TS + c] : 0.0;
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
   double d = 1.0 / piv;
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
         dn = 1.0 / piv;
      }
      const double f = cb[myslot] * d;  // 0 for r <= c: the updates below leave those rows alone
#pragma unroll
      for (int k = 0; k < 8; ++k) {
         m[k] = fma(-f, cb[h * 8 + k], m[k]);  // entries right of the diagonal collect garbage; never used
         x[k] = fma(-f, rb[h * 8 + k], x[k]);
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
      Li[swz(r, c)] = (c <= r) ? x[k] * rs : 0.0;
      if (want_l) T[r * TS + c] = (c <= r) ? m[k] * dvec[c] : 0.0;
   }
   return bad;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// MODE 0: all warps run the diagonal routine.  MODE 1: warp 0 runs it, the other warps issue DMMA streams.
// MODE 2: warp 0 runs it, the other warps stream global loads (L2 hits).
template <int MODE>
__global__ void k(const double* A, long long* cyc, int iters, double* sink, const double* big, size_t nbig) {
   __shared__ double T[4][KB * TS], Li[4][KB * KB], cb[4][40], rb[4][32], dv[4][16];
   __shared__ volatile int done;
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   if (threadIdx.x == 0) done = 0;
   __syncthreads();
   long long t = 0;
   int bad = 0;
   if (MODE == 0 || w == 0) {
      for (int it = 0; it < iters; ++it) {
         for (int i = lane; i < KB * TS; i += 32) T[w][i] = A[i];
         __syncwarp();
         const long long t0 = clock64();
         bad += diag_factor_inv(T[w], Li[w], cb[w], rb[w], dv[w], lane, 16, false);
         __syncwarp();
         t += clock64() - t0;
      }
      if (lane == 0) { cyc[blockIdx.x * 4 + w] = t / iters; sink[blockIdx.x * 4 + w] = Li[w][5] + bad; }
      if (MODE != 0 && lane == 0) done = 1;
   } else if (MODE == 1) {
      double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0, g0 = 0, g1 = 0;
      const double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
      while (!done) {
#pragma unroll
         for (int i = 0; i < 8; ++i) { dmma884(d0, d1, a, b); dmma884(e0, e1, a, b); dmma884(f0, f1, b, a); dmma884(g0, g1, b, b); }
      }
      sink[1024 + threadIdx.x] = d0 + d1 + e0 + e1 + f0 + f1 + g0 + g1;
   } else {
      double s = 0;
      size_t idx = ((size_t)blockIdx.x * 128 + threadIdx.x) * 2;
      while (!done) {
#pragma unroll
         for (int i = 0; i < 8; ++i) { s += big[(idx + (size_t)i * 65536) % nbig]; }
         idx = (idx + 8 * 65536) % nbig;
      }
      sink[1024 + threadIdx.x] = s;
   }
}
int main() {
   double hA[KB * TS];
   for (int r = 0; r < KB; ++r) for (int c = 0; c < TS; ++c) hA[r * TS + c] = (r == c) ? 4.0 : 0.1 / (1 + r + c);
   double *dA, *sink, *big; long long* cyc;
   const size_t nbig = 8u << 20;  // 64 MB: L2 resident
   cudaMalloc(&dA, sizeof(hA)); cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
   cudaMalloc(&cyc, 8 * 148 * 8 * 4); cudaMalloc(&sink, 8 * 4096); cudaMalloc(&big, nbig * 8); cudaMemset(big, 0, nbig * 8);
   for (int mode = 0; mode < 3; ++mode) for (int nb : {148, 148 * 6}) {
      if (mode == 0) k<0><<<nb, 128>>>(dA, cyc, 100, sink, big, nbig);
      if (mode == 1) k<1><<<nb, 128>>>(dA, cyc, 100, sink, big, nbig);
      if (mode == 2) k<2><<<nb, 128>>>(dA, cyc, 100, sink, big, nbig);
      cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("mode %d blocks %d: cycles per call (warp 0) %lld  err %s\n", mode, nb, h[0], cudaGetErrorString(cudaGetLastError()));
   }
   return 0;
}
