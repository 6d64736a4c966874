// Micro-benchmark: dependent-issue latencies of the FP64 instructions on the critical paths of the factorisation
// kernels (one warp on an SM): DFMA, DMUL, MUFU.RCP64H (+ Newton), DMMA accumulate chain, STS -> __syncwarp -> LDS.
//   nvcc -arch=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(long long* cyc, double* sink, double seed) {
   __shared__ double sbuf[64];
   const int lane = threadIdx.x;
   double x = seed + lane * 1e-9, y = 1.0 + seed * 1e-3, z = 0.0, w = 0.0;
   constexpr int N = 256;
   sbuf[lane] = x;
   __syncwarp();
   const long long t0 = clock64();
#pragma unroll 1
   for (int it = 0; it < N / 16; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
         if (MODE == 0) x = fma(x, y, 1e-9);                         // DFMA chain
         if (MODE == 1) x = x * y;                                   // DMUL chain
         if (MODE == 2) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }  // MUFU + DADD
         if (MODE == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); x = fma(r, e, r) + 1.5; }
         if (MODE == 4) { x = 1.0 / x + 1.5; }                       // IEEE division + DADD
         if (MODE == 5) dmma884(z, w, x, y);                         // dependent DMMA accumulate chain
         if (MODE == 6) { sbuf[(lane + 1) & 31] = x; __syncwarp(); x = sbuf[lane] + 1e-9; __syncwarp(); }  // STS -> LDS round trip
         if (MODE == 7) { x = sqrt(x) + 1.5; }
      }
   }
   const long long t1 = clock64();
   if (lane == 0) cyc[MODE] = (t1 - t0) / N;
   sink[MODE * 32 + lane] = x + z + w;
}
int main() {
   long long* cyc; double* sink;
   cudaMalloc(&cyc, 64); cudaMalloc(&sink, 8 * 32 * 8);
   k<0><<<1, 32>>>(cyc, sink, 1.0); k<1><<<1, 32>>>(cyc, sink, 1.0); k<2><<<1, 32>>>(cyc, sink, 1.0); k<3><<<1, 32>>>(cyc, sink, 1.0);
   k<4><<<1, 32>>>(cyc, sink, 1.0); k<5><<<1, 32>>>(cyc, sink, 1.0); k<6><<<1, 32>>>(cyc, sink, 1.0); k<7><<<1, 32>>>(cyc, sink, 1.0);
   cudaDeviceSynchronize();
   long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
   const char* names[8] = {"DFMA", "DMUL", "MUFU.RCP64H + DADD", "rcp seed + 2 Newton + DADD", "IEEE 1/x + DADD", "DMMA accumulate", "STS/syncwarp/LDS/syncwarp", "sqrt + DADD"};
   for (int i = 0; i < 8; ++i) printf("%-30s %lld cycles per dependent op  (%s)\n", names[i], h[i], cudaGetErrorString(cudaGetLastError()));
   return 0;
}
