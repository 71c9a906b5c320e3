// development probe: round-trip latency of a store followed by a dependent load through global, local and shared
// memory (B200, sm_100a): the tree passes and factorisations of the step are chains of exactly this
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* g, double* out, long long* cyc, double a, int n, int* idx) {
  long long t0, t1;
  const int l = threadIdx.x;
  // (0) global: store, __syncwarp, load the neighbour lane's element (same 256-byte line)
  double v = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { g[l] = v; __syncwarp(); v = g[l ^ 1] + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[0] = t1 - t0;
  // (1) global: same lane reads its own element back (read-modify-write chain of one thread)
  double u = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { g[64 + l] = u; u = ((volatile double*)g)[64 + l] + 1.0; }
  t1 = clock64(); cyc[1] = t1 - t0;
  // (2) plain dependent global loads (pointer chase through L1, no stores)
  int p = l;
  t0 = clock64();
  for (int i = 0; i < n; i++) p = idx[p];
  t1 = clock64(); cyc[2] = t1 - t0;
  // (3) local array with dynamic index: x[j] -= c * x[i]
  double x[32];
  for (int i = 0; i < 32; i++) x[i] = a + i;
  int j = idx[l] & 31;
  t0 = clock64();
  for (int i = 0; i < n; i++) { x[j] -= 0.5 * x[(j + 1) & 31]; j = (j + 7) & 31; }
  t1 = clock64(); cyc[3] = t1 - t0;
  double w = 0;
  for (int i = 0; i < 32; i++) w += x[i];
  // (4) same chain on a shared-memory array laid out [index][lane]
  __shared__ double s[32 * 32];
  for (int i = 0; i < 32; i++) s[i * 32 + l] = a + i;
  j = idx[l] & 31;
  t0 = clock64();
  for (int i = 0; i < n; i++) { s[j * 32 + l] -= 0.5 * s[((j + 1) & 31) * 32 + l]; j = (j + 7) & 31; }
  t1 = clock64(); cyc[4] = t1 - t0;
  for (int i = 0; i < 32; i++) w += s[i * 32 + l];
  // (5) global store -> __syncwarp -> load with ld.global.cg (L2) for comparison
  double q = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) { g[128 + l] = q; __syncwarp(); q = __ldcg(g + 128 + (l ^ 1)) + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[5] = t1 - t0;
  out[l] = v + u + p + w + q;
}
int main() {
  double *g, *out; long long* cyc; int* idx;
  cudaMalloc(&g, 4096 * 8); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8); cudaMalloc(&idx, 1024 * 4);
  int h_idx[1024]; for (int i = 0; i < 1024; i++) h_idx[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(idx, h_idx, sizeof(h_idx), cudaMemcpyHostToDevice);
  cudaMemset(g, 0, 4096 * 8);
  const int n = 4096;
  for (int r = 0; r < 2; r++) k<<<1, 32>>>(g, out, cyc, 1.0000001, n, idx);
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char* nm[6] = {"global st->sync->ld (other lane)", "global st->ld (same lane)", "global ld->ld (L1 hit chase)",
                       "local x[j] -= c*x[i]", "shared x[j] -= c*x[i]", "global st->sync->ld.cg"};
  for (int i = 0; i < 6; i++) printf("%-36s %.1f cycles per round trip\n", nm[i], (double)h[i] / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
