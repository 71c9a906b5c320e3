// Persistent rollout kernel: every step of a multi-step rollout in ONE launch, without a device-wide barrier
// between the launches of a step.
//
// Why.  A split step (mjb_kernels.cu launch_split_step) is three launches - position+velocity | PGS | finish+
// integrate - and every launch lasts as long as its SLOWEST environment: the stages are latency-bound per
// environment (profiles/r02_experiments.md), the work of an environment varies from step to step (contacts come and
// go: PGS rows x sweeps has mean 240, max 2500 per step, autocorrelation 0.08 over 7 steps), so a 4096-environment
// launch always waits for a tail event.  Environments are independent (rollout.cc:127-155 steps each mjData on its
// own), so the barrier only has to span the environments that SHARE a warp in some stage.
//
// Mapping.  One persistent CTA per SM owns a fixed block of environments (28 for 4096 environments on 148 SMs) and
// carries them through all steps: warp w runs the two halves of the step for the pair (2w, 2w+1) on 16 lanes each
// (the code of k_step_warp<.., 16, .., PART>), a __syncthreads, then the first warps run the PGS solve of the CTA's
// environments on four lanes each (mjb_pgs4.h, the code of k_pgs4), a second __syncthreads, and the pair's warp
// finishes and integrates.  The CTAs drift apart freely; inside a CTA the warps stay in the same stage, so they share
// the instruction stream (a fully asynchronous variant, every warp on its own, was measured 2x slower in round 1:
// instruction-cache misses).  Per step a CTA waits for the slowest of ITS 28 environments instead of 4096.
//
// Status: correct and tested, but NOT the default (mjb_kernels.cu rollout_persistent_available): on B200, humanoid x4096,
// a step costs 1.22-1.40 ms here against 1.19 ms for the three launches.  The phases themselves are slower inside this
// kernel - first half 712-877 k cycles per warp depending on the L1 left by the PGS pool (63 / 99 / 160 KB pool: 712 /
// 765 / 877 k), solve 794 k per warp against ~510 k in k_pgs4 - which costs more than the barrier removal gains
// (profiles/r02_experiments.md section 7).
//
// The results are those of the split step bit for bit: the same per-environment code in the same order; only the
// interleaving of independent environments changes (tests/test_gpu_parity.py compares the two paths).
#include "mjb_pgs4.h"

namespace mjb {

struct RolloutArgs {
  int t0, t1, nstep;          // steps [t0, t1) of a rollout of nstep steps (strides of the reference layout)
  int first, later;           // stage flags of the first / later launches of a step (mjb_stage.h run_env)
  int layout;                 // 0: native [step][elem][env stride]; 1: reference [env][step][elem]
  const double* ctrl;         // nullptr: controls stay as they are
  double* state;              // nullptr: states are not recorded
  int nstate;
  int envs_per_cta;           // even
  int pool_bytes;             // shared memory the PGS warps of a CTA share (RolloutPool)
  const unsigned char* tab;   // PGS visiting orders (pgs4_order_table)
  int tab_iters;
  int pgs_flags;              // bit8: force the slot layout (tests)
};

constexpr int kRolloutWarps = 16;      // four warpgroups; 512 threads x 128 registers = the register file of an SM
constexpr int kRolloutPgsWarps = 4;    // warpgroup 0 also runs the PGS solves (eight environments per warp)

// Register budget per stage (setmaxnreg, PTX ISA 8.0+): the halves of the step need 128 registers per thread on all 16
// warps, the PGS sweeps 213 on four.  Before the solve the twelve idle warps hand registers back to the pool and
// warpgroup 0 grows; after it the exchange is undone.  (Without it the solve spills 6 KB per thread at 128.)
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }

// bar.sync on barrier 1 by all warps of the CTA; orders shared and global memory among them like __syncthreads
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kRolloutWarps) : "memory"); }

// Shared memory of the PGS warps.  The halves of the step live on L1 hits (their time grows by 45 % when the L1
// shrinks from 228 KB to 92 KB), so the solve may only carve out what it needs: the four PGS warps share ONE pool and
// each takes exactly the bytes of its packed records (1.7 KB per environment on average, 11 KB for 31 rows).  When
// the records of all four do not fit at once, the warps that found no room wait for a placed warp to finish and take
// its space in a second pass (rare).  need[]: bytes wanted, 0 = nothing to stage, -1 = solved.
struct RolloutPool {
  static constexpr int kSlack = 1024;
  double* pool; int pool_bytes;
  volatile int* need;
  const double* beta_;
  int w, lane;
  int remaining = 0;
  bool entered = false;
  __device__ __forceinline__ void barrier() { asm volatile("bar.sync 2, 128;\n" ::: "memory"); }
  __device__ __forceinline__ const double* beta() const { return beta_; }
  // placement of this pass, the same on every warp: in warp order, whoever still fits
  __device__ __forceinline__ int place() {
    int off = 0, mine = -1;
    remaining = 0;
    for (int i = 0; i < kRolloutPgsWarps; i++) {
      const int n = need[i];
      if (n < 0) continue;
      if (n == 0 || n > pool_bytes) { if (i == w) mine = n == 0 ? 0 : -2; continue; }   // nothing to stage / can never fit
      if (off + n <= pool_bytes) { if (i == w) mine = off; off += n; }
      else remaining++;
    }
    return mine;
  }
  __device__ __forceinline__ void acquire(int n, double*& p, int& bytes) {
    entered = true;
    for (;;) {
      if (lane == 0) need[w] = n;
      barrier();
      const int off = __shfl_sync(0xffffffffu, place(), 0);   // (uniform by construction; the shuffle tells the compiler)
      remaining = __shfl_sync(0xffffffffu, remaining, 0);
      barrier();
      if (off != -1) { p = pool + (off > 0 ? off / 8 : 0); bytes = off == -2 ? 0 : n; return; }
      barrier();   // the placed warps of this pass are done: their space is free
    }
  }
  __device__ __forceinline__ void release() {
    if (!entered) { double* p; int n; acquire(0, p, n); }
    while (remaining > 0) {
      barrier();
      if (lane == 0) need[w] = -1;
      barrier();
      place();
      remaining = __shfl_sync(0xffffffffu, remaining, 0);
      barrier();
    }
    entered = false;
  }
};

#ifdef MJB_ROLLOUT_PROF   // development aid: cycles per warp in each phase of the step and at the two barriers
static __device__ unsigned long long g_rollout_prof[160][kRolloutWarps][8];
#define RPROF(...) __VA_ARGS__
#else
#define RPROF(...)
#endif

template <int FEAT>
__global__ void __launch_bounds__(32 * kRolloutWarps, 1) k_rollout(DModel m, Batch b, RolloutArgs a) {
  extern __shared__ double rollout_smem[];
  // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the warp-synchronous code
  // of the solve free of divergence guards (BRA.DIV around every shuffle otherwise: +19 % instructions per PGS row)
  const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int l = threadIdx.x & 31, sub = l >> 4, l16 = l & 15;
  const int base = blockIdx.x * a.envs_per_cta;
  const int here = min(a.envs_per_cta, b.nenv - base);   // environments of this CTA
  if (here <= 0) return;
  const int e = base + 2 * w + sub;
  const bool mine = 2 * w + sub < here;
  const unsigned lanes = 0xFFFFu << (16 * sub);
  const int npgs = (here + 7) / 8;
  // dynamic shared memory: need[] (128 bytes), the momentum-coefficient table, the pool
  double* const beta_tab = rollout_smem + 16;
  for (int n = threadIdx.x; n < kPgs4Beta; n += blockDim.x) beta_tab[n] = (double)(n - 1) / (double)(n + 2);
  cta_barrier();
  const int nu = m.sz.nu, nq = m.sz.nq, nv = m.sz.nv;
  RPROF(unsigned long long pc[6] = {0, 0, 0, 0, 0, 0}; long long pt = clock64();)
#define RMARK(i) RPROF({ const long long now_ = clock64(); pc[i] += (unsigned long long)(now_ - pt); pt = now_; })
  for (int t = a.t0; t < a.t1; t++) {
    if (mine) {
      Env d(m, b, e, l16, 16);
      if (a.ctrl) {
        if (a.layout == 0) {
          FD c = d.ctrl();
          for (int i = l16; i < nu; i += 16) c[i] = a.ctrl[((size_t)t * nu + i) * b.stride + e];
        } else if (!env_has_warning(d)) {      // rollout rule: a warned environment no longer takes controls
          FD c = d.ctrl();
          const double* src = a.ctrl + ((size_t)e * a.nstep + t) * nu;
          for (int i = l16; i < nu; i += 16) c[i] = src[i];
        }
        __syncwarp(lanes);
      }
      run_env(m, b, e, 0, a.first, l16, 16, nullptr, 0, SOL_PGS, lanes, FEAT, 1);
    }
    // the two barriers of a step are reached from two places (warp-uniform branches): each branch leaves with the
    // register count it came with, so the counts agree where the branches meet
    RMARK(0)
    if (w >= kRolloutPgsWarps) {
      reg_dec<64>();
      cta_barrier();
      RMARK(1)
      cta_barrier();
      reg_inc<128>();
      RMARK(3)
    } else {
      cta_barrier();
      RMARK(1)
      __syncwarp();     // (a convergence point the compiler can see: the halves above run under sub-warp masks)
      reg_inc<216>();
      {
        RolloutPool pool{rollout_smem + 16 + kPgs4Beta, a.pool_bytes, (volatile int*)rollout_smem, beta_tab, w, l};
        if (w < npgs) {
          const int slot = w * 8 + (l >> 2);
          pgs4_warp(m, b, slot < here ? base + slot : -1, a.later | a.pgs_flags, a.tab, a.tab_iters, pool);
        }
        pool.release();
      }
      reg_dec<128>();
      RMARK(2)
      cta_barrier();
      RMARK(3)
    }
    if (mine) {
      run_env(m, b, e, 0, a.later, l16, 16, nullptr, 0, SOL_PGS, lanes, FEAT, 2);
      if (!(m.opt.disableflags & DSBL_AUTORESET))   // bad acceleration: forward pass on the reset state, then integration
        run_env(m, b, e, kMaskStep, a.later | 16, l16, 16, nullptr, 0, SOL_PGS, lanes, FEAT, 0);
      if (a.state) {
        Env d(m, b, e, l16, 16);
        __syncwarp(lanes);
        FD qp = d.qpos(), qv = d.qvel(), ac = d.act();
        for (int k = l16; k < a.nstate; k += 16) {
          const double v = k == 0 ? d.time()[0] : k < 1 + nq ? qp[k - 1] : k < 1 + nq + nv ? qv[k - 1 - nq] : ac[k - 1 - nq - nv];
          if (a.layout == 0) a.state[((size_t)t * a.nstate + k) * b.stride + e] = v;
          else a.state[((size_t)e * a.nstep + t) * a.nstate + k] = v;
        }
      }
    }
    RMARK(4)
  }
  RPROF(if (l == 0 && blockIdx.x < 160) for (int i = 0; i < 5; i++) g_rollout_prof[blockIdx.x][w][i] += pc[i];)
}

#ifdef MJB_ROLLOUT_PROF
extern "C" __attribute__((visibility("default"))) int mjb_debug_rollout_prof(unsigned long long* out) {
  int rc = (int)cudaMemcpyFromSymbol(out, g_rollout_prof, sizeof(g_rollout_prof));
  static unsigned long long zero[160 * kRolloutWarps * 8];
  cudaMemcpyToSymbol(g_rollout_prof, zero, sizeof(zero));
  return rc;
}
#endif

namespace backend {

// 0 on launch; -1 when the batch does not fit the kernel's mapping (the caller takes the per-step launches)
int launch_krollout_lean(const DModel& dm, const Batch& b, int t0, int t1, int nstep, int first, int later, int layout,
                         const double* ctrl, double* state, int nstate, void* stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  static int sm_count[64] = {0};
  if (!sm_count[dev & 63]) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sm_count[dev & 63] = sms;
  }
  sms = sm_count[dev & 63];
  if (sms <= 0) return -1;
  RolloutArgs a;
  a.envs_per_cta = 2 * ((b.nenv + 2 * sms - 1) / (2 * sms));
  if (a.envs_per_cta > 2 * kRolloutWarps) return -1;
  if ((a.envs_per_cta + 7) / 8 > kRolloutPgsWarps) return -1;
  static const int pool_kb = [] { const char* e = getenv("MJB_ROLLOUT_POOL_KB"); const int v = e ? atoi(e) : 0; return v > 0 && v <= 200 ? v : 99; }();
  const size_t smem = (size_t)pool_kb * 1024;     // + 1 KB of static shared memory = the 100 KB carve-out (measured best: 63 / 99 / 131 KB -> 1.320 / 1.220 / 1.246 ms)
  a.pool_bytes = (int)smem - 128 - 8 * kPgs4Beta;
  if (pgs4_table(dm, &a.tab, &a.tab_iters)) return -3;
  a.t0 = t0; a.t1 = t1; a.nstep = nstep; a.first = first; a.later = later; a.layout = layout;
  a.ctrl = ctrl; a.state = state; a.nstate = nstate;
  a.pgs_flags = pgs4_force_slots();
  static size_t attr[64] = {0};
  if (attr[dev & 63] < smem) {
    if (cudaFuncSetAttribute(k_rollout<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    attr[dev & 63] = smem;
  }
  const int grid = (b.nenv + a.envs_per_cta - 1) / a.envs_per_cta;
  k_rollout<0><<<grid, 32 * kRolloutWarps, smem, (cudaStream_t)stream>>>(dm, b, a);
  return 0;
}

}  // namespace backend
}  // namespace mjb
