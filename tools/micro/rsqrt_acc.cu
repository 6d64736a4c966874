// accuracy of fast_rsqrt (erfgauss.cuh) against the correctly rounded value, and of the hardware seed
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../multicharge_b200/csrc/erfgauss.cuh"
__global__ void k(const double* x, double* y, double* seed, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[i] = mchrg::eg::fast_rsqrt(x[i]); double s; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x[i])); seed[i] = s; }
}
int main() {
   const int n = 1 << 20;
   double *hx = new double[n], *hy = new double[n], *hs = new double[n];
   unsigned long long st = 88172645463325252ull;
   for (int i = 0; i < n; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; double u = (st >> 11) * (1.0 / 9007199254740992.0); hx[i] = exp(log(1e-6) + u * log(1e12)); }
   double *dx, *dy, *ds; cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&ds, n * 8);
   cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
   k<<<n / 256, 256>>>(dx, dy, ds, n);
   cudaMemcpy(hy, dy, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hs, ds, n * 8, cudaMemcpyDeviceToHost);
   double emax = 0, smax = 0;
   for (int i = 0; i < n; ++i) {
      const long double ref = 1.0L / sqrtl((long double)hx[i]);
      emax = fmax(emax, (double)fabsl((hy[i] - ref) / ref)); smax = fmax(smax, (double)fabsl((hs[i] - ref) / ref));
   }
   printf("fast_rsqrt max rel err %.3e (eps = 1.1e-16)   hardware seed max rel err %.3e\n", emax, smax);
   return 0;
}
