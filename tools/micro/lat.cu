// Dependent-chain latencies of the building blocks of the batched 16x16 factorisation (one warp, idle SM).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double shfl_d(double v, int src) {
   const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
   const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
   return __hiloint2double(hi, lo);
}
template <int OP>
__global__ void k(long long* cyc, double* sink, double seed) {
   __shared__ double buf[64];
   const int lane = threadIdx.x;
   double v = seed + lane * 1e-3, w = 1.0000001;
   buf[lane] = v;
   __syncwarp();
   const int N = 256;
   const long long t0 = clock64();
#pragma unroll 16
   for (int i = 0; i < N; ++i) {
      if (OP == 0) v = fma(v, w, 1e-9);                       // dependent DFMA
      if (OP == 1) v = rsqrt(v) + 1.5;                        // rsqrt (+ DADD)
      if (OP == 2) v = shfl_d(v, (lane + 1) & 31);            // 64-bit shuffle
      if (OP == 3) { buf[lane] = v; __syncwarp(); v = buf[(lane + 1) & 31]; __syncwarp(); }  // smem round trip
      if (OP == 4) v = v * w;                                 // dependent DMUL
      if (OP == 5) v = 1.0 / v + 1.5;                         // division
      if (OP == 6) v = sqrt(v) + 1.5;
      if (OP == 7) { float f = (float)v; f = rsqrtf(f); v = (double)f + 1.5; }  // fp32 rsqrt seed path incl. conversions
      if (OP == 8) v = v + w;                                 // dependent DADD
   }
   const long long t1 = clock64();
   if (lane == 0) { cyc[OP] = (t1 - t0) / N; }
   sink[lane] = v;
}
int main() {
   long long* cyc; double* sink;
   cudaMalloc(&cyc, 16 * 8); cudaMalloc(&sink, 32 * 8);
   k<0><<<1, 32>>>(cyc, sink, 1.1); k<1><<<1, 32>>>(cyc, sink, 1.1); k<2><<<1, 32>>>(cyc, sink, 1.1); k<3><<<1, 32>>>(cyc, sink, 1.1);
   k<4><<<1, 32>>>(cyc, sink, 1.1); k<5><<<1, 32>>>(cyc, sink, 1.1); k<6><<<1, 32>>>(cyc, sink, 1.1); k<7><<<1, 32>>>(cyc, sink, 1.1); k<8><<<1, 32>>>(cyc, sink, 1.1);
   cudaDeviceSynchronize();
   long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
   const char* names[] = {"DFMA", "rsqrt+DADD", "shfl_d", "STS+syncwarp+LDS+syncwarp", "DMUL", "1/x+DADD", "sqrt+DADD", "cvt+rsqrtf+cvt+DADD", "DADD"};
   for (int i = 0; i < 9; ++i) printf("%-28s %lld cycles\n", names[i], h[i]);
   printf("%s\n", cudaGetErrorString(cudaGetLastError()));
   return 0;
}
