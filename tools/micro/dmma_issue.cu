// Micro-benchmark: DMMA issue interval of ONE warp with independent accumulators, and the inner loop of the batch
// factorisation's update (4 LDS.128 + 8 DMMA per 8 x 8 block pair) from shared memory.
//   nvcc -arch=sm_100a -O3 -o dmma_issue dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(long long* cyc, double* sink, int nwarps) {
   __shared__ __align__(16) double buf[4][2 * 8 * 64];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   for (int i = threadIdx.x; i < 4 * 1024; i += blockDim.x) (&buf[0][0])[i] = 1.0 + 1e-9 * i;
   __syncthreads();
   if (w >= nwarps) return;
   double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}}, acc0[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
   const double x = 1.0 + lane * 1e-9, y = 1.0 - lane * 1e-9;
   constexpr int N = 64;
   const long long t0 = clock64();
#pragma unroll 1
   for (int it = 0; it < N; ++it) {
      if (MODE == 0) {  // 8 DMMA on 4 independent accumulators (each twice)
#pragma unroll
         for (int u = 0; u < 2; ++u) {
            dmma884(acc[0][0][0], acc[0][0][1], x, y); dmma884(acc[0][1][0], acc[0][1][1], x, y);
            dmma884(acc[1][0][0], acc[1][0][1], y, x); dmma884(acc[1][1][0], acc[1][1][1], y, x);
         }
      } else {  // the update loop body: one block pair from shared memory
         const double2* As = reinterpret_cast<const double2*>(buf[w]) + lane;
         const int b = it & 7;
         const double2 a0 = As[32 * b], a1 = As[32 * (8 + b)];
         const double2 c0 = As[32 * ((b + 1) & 7)], c1 = As[32 * (8 + ((b + 1) & 7))];
         dmma884(acc[0][0][0], acc[0][0][1], a0.x, c0.x); dmma884(acc[0][1][0], acc[0][1][1], a0.x, c1.x);
         dmma884(acc[1][0][0], acc[1][0][1], a1.x, c0.x); dmma884(acc[1][1][0], acc[1][1][1], a1.x, c1.x);
         dmma884(acc[0][0][0], acc[0][0][1], a0.y, c0.y); dmma884(acc[0][1][0], acc[0][1][1], a0.y, c1.y);
         dmma884(acc[1][0][0], acc[1][0][1], a1.y, c0.y); dmma884(acc[1][1][0], acc[1][1][1], a1.y, c1.y);
         if (MODE == 2) {
            dmma884(acc0[0][0][0], acc0[0][0][1], a0.x, a0.x); dmma884(acc0[0][1][0], acc0[0][1][1], a0.x, a1.x);
            dmma884(acc0[1][0][0], acc0[1][0][1], a1.x, a0.x); dmma884(acc0[1][1][0], acc0[1][1][1], a1.x, a1.x);
            dmma884(acc0[0][0][0], acc0[0][0][1], a0.y, a0.y); dmma884(acc0[0][1][0], acc0[0][1][1], a0.y, a1.y);
            dmma884(acc0[1][0][0], acc0[1][0][1], a1.y, a0.y); dmma884(acc0[1][1][0], acc0[1][1][1], a1.y, a1.y);
         }
      }
   }
   const long long t1 = clock64();
   if (lane == 0) cyc[w] = (t1 - t0) / N;
   double s = 0;
   for (int i = 0; i < 8; ++i) s += (&acc[0][0][0])[i] + (&acc0[0][0][0])[i];
   sink[threadIdx.x] = s;
}
int main() {
   long long* cyc; double* sink;
   cudaMalloc(&cyc, 64); cudaMalloc(&sink, 8 * 256);
   const char* names[3] = {"8 DMMA, 4 independent accumulators", "block pair: 4 LDS.128 + 8 DMMA", "block pair + 8 DMMA of the diagonal tile"};
   for (int mode = 0; mode < 3; ++mode) for (int nw : {1, 4, 8}) {
      if (mode == 0) k<0><<<1, 256>>>(cyc, sink, nw);
      if (mode == 1) k<1><<<1, 256>>>(cyc, sink, nw);
      if (mode == 2) k<2><<<1, 256>>>(cyc, sink, nw);
      cudaDeviceSynchronize();
      long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-45s warps %d: %lld cycles per iteration (warp 0)  %s\n", names[mode], nw, h[0], cudaGetErrorString(cudaGetLastError()));
   }
   return 0;
}
