// development probe: dependent-issue latency of the instructions on the PGS row chain (B200, sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  double x = a; long long t0, t1;
  // DADD chain
  t0 = clock64();
  for (int i = 0; i < n; i++) { x = x + b; x = x + a; x = x + b; x = x + a; }
  t1 = clock64(); cyc[0] = t1 - t0;
  // DMUL chain
  double y = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { y = y * b; y = y * a; y = y * b; y = y * a; }
  t1 = clock64(); cyc[1] = t1 - t0;
  // shuffle (64-bit) chain
  double z = a + threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < n; i++) { z = __shfl_xor_sync(0xffffffffu, z, 1); z = __shfl_xor_sync(0xffffffffu, z, 2); z = __shfl_xor_sync(0xffffffffu, z, 1); z = __shfl_xor_sync(0xffffffffu, z, 2); }
  t1 = clock64(); cyc[2] = t1 - t0;
  // select chain (compare + select on doubles)
  double w = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { w = w < b ? b : w; w = w > a ? a : w; w = w < b ? b : w; w = w > a ? a : w; }
  t1 = clock64(); cyc[3] = t1 - t0;
  // shared-memory store -> load round trip
  __shared__ double s[64];
  double v = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { s[threadIdx.x] = v; __syncwarp(); v = s[threadIdx.x ^ 1]; __syncwarp(); s[threadIdx.x] = v; __syncwarp(); v = s[threadIdx.x ^ 1]; __syncwarp();
                                s[threadIdx.x] = v; __syncwarp(); v = s[threadIdx.x ^ 1]; __syncwarp(); s[threadIdx.x] = v; __syncwarp(); v = s[threadIdx.x ^ 1]; __syncwarp(); }
  t1 = clock64(); cyc[4] = t1 - t0;
  out[threadIdx.x] = x + y + z + w + v;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
  const int n = 4096;
  k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n); k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n);
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char* nm[5] = {"DADD", "DMUL", "SHFL64", "DSETP+SEL", "STS->sync->LDS"};
  for (int i = 0; i < 5; i++) printf("%-16s %.1f cycles per dependent op\n", nm[i], (double)h[i] / (4.0 * n));
  return 0;
}
