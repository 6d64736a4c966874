// Micro-benchmark: per-call latency of the 16x16 Cholesky + inverse warp routine (csrc/diag16.cuh) in isolation,
// for the reciprocal variants selectable at compile time (D16_RCP: 0 IEEE division, 1 hardware seed + Newton).
//   nvcc -arch=sm_100a -O3 -I../../multicharge_b200/csrc -DD16_RCP=1 -o diag_variants diag_variants.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "diag16.cuh"
using namespace mchrg::d16;
constexpr int TS = 17;
__global__ void k(const double* A, long long* cyc, int iters, double* sink, int nwarps_active) {
   __shared__ double T[4][KB * TS], Li[4][KB * KB], cb[4][40], rb[4][32], dv[4][16];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   long long t = 0;
   int bad = 0;
   if (w < nwarps_active) {
      for (int it = 0; it < iters; ++it) {
         for (int i = lane; i < KB * TS; i += 32) T[w][i] = A[i];
         __syncwarp();
         const long long t0 = clock64();
         bad += diag_factor_inv<TS, true, KB>(T[w], Li[w], cb[w], rb[w], dv[w], lane, 16, true);
         __syncwarp();
         t += clock64() - t0;
      }
      if (lane == 0) { cyc[blockIdx.x * 4 + w] = t / iters; sink[blockIdx.x * 4 + w] = Li[w][5] + T[w][18] + bad; }
   }
}
int main() {
   double hA[KB * TS];
   for (int r = 0; r < KB; ++r) for (int c = 0; c < TS; ++c) hA[r * TS + c] = (r == c) ? 4.0 + 0.01 * r : 0.1 / (1 + r + c);
   double *dA, *sink; long long* cyc;
   cudaMalloc(&dA, sizeof(hA)); cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
   cudaMalloc(&cyc, 8 * 148 * 8 * 4); cudaMalloc(&sink, 8 * 4096);
   for (int nw : {1, 4}) for (int nb : {1, 148, 148 * 6}) {
      k<<<nb, 128>>>(dA, cyc, 200, sink, nw);
      cudaDeviceSynchronize();
      long long h[4]; double s[4];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      cudaMemcpy(s, sink, sizeof(s), cudaMemcpyDeviceToHost);
      printf("rcp variant %d: active warps/CTA %d blocks %d: cycles per call %lld (check %.15g) %s\n", D16_RCP, nw, nb, h[0], s[0],
             cudaGetErrorString(cudaGetLastError()));
   }
   return 0;
}
