// Dependent-chain latencies of the float64 operations the exact EPnP solver is made of (one warp, one lane active
// or all lanes): cycles per operation from clock64 around an unrolled dependent chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/bin/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP>
__global__ void k(double* out, long long* cyc, double x0, double y0, int active) {
  if ((int)threadIdx.x >= active) return;
  double x = x0 + threadIdx.x * 1e-9, y = y0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = x + y;
    if (OP == 1) x = x * y;
    if (OP == 2) x = fma(x, y, y);
    if (OP == 3) x = x / y;
    if (OP == 4) x = sqrt(x) + y;       // sqrt + 1 add (keeps the value from collapsing to 1)
    if (OP == 5) x = y / x;
    if (OP == 6) x = 1.0 / x + y;
    if (OP == 7) x = rsqrt(x) + y;
    if (OP == 8) { float f = (float)x; f = f * 1.0001f + 0.5f; x = (double)f; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[OP] = t1 - t0; }
  out[threadIdx.x + 32 * OP] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 32 * 16); cudaMalloc(&cyc, 8 * 16);
  const char* names[] = {"dadd", "dmul", "dfma", "ddiv x/y", "dsqrt+dadd", "ddiv y/x", "drcp+dadd", "drsqrt+dadd", "f2d roundtrip+ffma"};
  for (int active : {1, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      k<0><<<1, 32>>>(out, cyc, 1.5, 1e-3, active); k<1><<<1, 32>>>(out, cyc, 1.5, 1.0000001, active);
      k<2><<<1, 32>>>(out, cyc, 1.5, 0.999, active); k<3><<<1, 32>>>(out, cyc, 1.5, 1.0000001, active);
      k<4><<<1, 32>>>(out, cyc, 1.5, 0.7, active); k<5><<<1, 32>>>(out, cyc, 1.5, 1.3, active);
      k<6><<<1, 32>>>(out, cyc, 1.5, 0.3, active); k<7><<<1, 32>>>(out, cyc, 1.5, 0.3, active);
      k<8><<<1, 32>>>(out, cyc, 1.5, 0.3, active);
      cudaDeviceSynchronize();
    }
    long long h[9];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 9; ++i) printf("active=%2d %-22s %6.1f cycles/op\n", active, names[i], (double)h[i] / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
