// k_pgs4: the PGS solve of a split step as its own launch, one warp (= one CTA) per eight environments; the device
// code is mjb_pgs4.h
#include <mutex>

#include "mjb_pgs4.h"

namespace mjb {

// Visiting orders of every sweep, for every problem size of the register classes: the reference reshuffles the
// row order before each sweep with Fisher-Yates draws of ONE PCG32 stream seeded (0, 1) per solve
// (engine_solver.c:240-265,498-502,560), so the order of sweep s of a problem with n rows is a constant of
// (n, s).  Table: for n = 1..67, `iters` sweeps of n bytes; problem n starts at iters * n(n-1)/2.
static std::vector<unsigned char> pgs4_order_table(int iters) {
  std::vector<unsigned char> tab((size_t)iters * (67 * 68 / 2) + 16);
  for (int n = 1; n <= 67; n++) {
    unsigned char* dst = tab.data() + (size_t)iters * (n * (n - 1) / 2);
    int order[68];
    for (int c = 0; c < n; c++) order[c] = c;
    Pcg32 rng{0, 1};
    pcg32_next(rng);
    for (int s = 0; s < iters; s++) {
      for (int i = n - 1; i > 0; i--) {
        const uint32_t j = pcg32_next(rng) % (uint32_t)(i + 1);
        const int t = order[i]; order[i] = order[j]; order[j] = t;
      }
      for (int c = 0; c < n; c++) dst[(size_t)s * n + c] = (unsigned char)order[c];
    }
  }
  return tab;
}

// grid: one warp per eight environments.  flags bit1/bit2: rollout skip rule (step_skip written by the
// position launch of this step).
__global__ void __launch_bounds__(32) k_pgs4(DModel m, Batch b, int flags, const unsigned char* __restrict__ order_tab, int order_iters) {
  extern __shared__ double pgs4_smem[];
  Pgs4FixedSmem pool{pgs4_smem, kPgs4SmemBytes};
  pgs4_warp(m, b, blockIdx.x * 8 + (threadIdx.x >> 2), flags, order_tab, order_iters, pool);
}

#ifdef MJB_PGS4_PROF
extern "C" __attribute__((visibility("default"))) int mjb_debug_pgs4_prof(long long* out, int nwarp) {
  return (int)cudaMemcpyFromSymbol(out, g_pgs4_prof, sizeof(long long) * 8 * nwarp);
}
#endif

namespace backend {
// the order table lives once per device (built on first use for the model's iteration cap; 235 KB at 100 sweeps)
struct Pgs4Table { unsigned char* dev = nullptr; int iters = 0; };
static Pgs4Table g_pgs4_tab[64];
static int g_pgs4_force_slots = 0;   // tests: take the slot layout even when the packed records fit (set_debug)
void pgs4_set_force_slots(int on) { g_pgs4_force_slots = on; }
static std::mutex g_pgs4_mutex;
int pgs4_table(const DModel& dm, const unsigned char** tab, int* iters) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_pgs4_mutex);
  Pgs4Table* T = &g_pgs4_tab[dev & 63];
  const int want = dm.opt.iterations < 4096 ? dm.opt.iterations : 0;   // beyond: the kernel takes the generic sweeps
  if (want > T->iters) {
    const int n = want < 100 ? 100 : want;
    std::vector<unsigned char> host = pgs4_order_table(n);
    unsigned char* p = nullptr;
    if (cudaMalloc(&p, host.size()) != cudaSuccess) return -3;
    if (cudaMemcpy(p, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(p); return -3; }
    // an older, shorter table may still be read by launches in flight: it is small, leave it allocated
    T->dev = p; T->iters = n;
  }
  *tab = T->dev; *iters = T->iters;
  return 0;
}
int pgs4_force_slots() {
  static const int env_slots = [] { const char* e = getenv("MJB_PGS4_SLOTS"); return e && atoi(e) ? 256 : 0; }();
  return env_slots | (g_pgs4_force_slots ? 256 : 0);
}
int launch_pgs4(const DModel& dm, const Batch& b, int flags, void* stream) {
  const unsigned char* tab = nullptr;
  int iters = 0;
  if (int rc = pgs4_table(dm, &tab, &iters)) return rc;
  static bool attr[64] = {false};   // the attribute belongs to the device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63]) { cudaFuncSetAttribute(k_pgs4, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgs4SmemBytes); attr[dev & 63] = true; }
  const int grid = (b.nenv + 7) / 8;
  k_pgs4<<<grid, 32, kPgs4SmemBytes, (cudaStream_t)stream>>>(dm, b, flags | pgs4_force_slots(), tab, iters);
  return 0;
}
}  // namespace backend

}  // namespace mjb
