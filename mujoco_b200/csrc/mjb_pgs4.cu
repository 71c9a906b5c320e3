// Projected Gauss-Seidel as its own kernel: FOUR lanes per environment, eight environments per warp.
//
// Reference: solPGS, engine_solver.c:457-741 (scalar / pyramidal rows), with residual / costChange /
// dualState :187-353 and the PCG32 Fisher-Yates order :240-265.
//
// Why this mapping.  The sweep is a Gauss-Seidel recurrence over the rows of one environment: row i
// needs every earlier update of the same sweep, so an environment's solve is ONE dependent chain and
// the batch's solve time is the longest chain, not the sum of the work.  Inside a row, mju_dot
// (engine_util_blas.c:493-523) has exactly four independent accumulation chains, combined as
// (r0+r2)+(r1+r3): four lanes are all the parallelism a row has.  A warp therefore carries eight
// environments (the fused kernel used 32 lanes for one row, 8x redundant issue), and every instruction
// of the row body is on the dependent chain of eight solves at once.
//
// What is on the chain, and what is not:
//   * each lane keeps its chain's forces (elements c = 4q + lane) in REGISTERS; the update of a row
//     is a predicated select, the next row's products follow after one select + one multiply;
//   * the AR row, the row's constants (b, 1/AR_ii, AR_ii, lo, hi) and the old force of the NEXT row
//     are fetched one row ahead (AR through L1: read-only during the solve, 1.2 KB per environment on
//     average, so every sweep after the first hits L1) - no load sits between two updates;
//   * the Fisher-Yates shuffle of the NEXT sweep (LCG step, output permutation, modulo, swap) is
//     software-pipelined into the row loop, one swap per row: it fills issue slots the fp64 chain
//     leaves empty instead of running serially between sweeps;
//   * zero padding is exact: a chain sum is never -0 (it starts from +0), so adding a +-0 product of
//     the padded positions returns the same bits; the same holds for the padded tail terms.
// Results are bit-identical to the serial reference arithmetic (tests/test_gpu_parity.py compares
// forces, states and iteration counts).
//
// Environments with more rows than the largest register class (4*16+3) take the generic sweep of
// mjb_constraint.h (global memory, any lane count) on their four lanes.
#include <cuda_runtime.h>

#include "mjb_backend.h"
#include "mjb_stage.h"
#include "mjb_kstep.h"

namespace mjb {

constexpr int kPgs4Rows = 68;                 // row capacity of the largest register class (4 * 16 + 3, padded)
constexpr int kPgs4EnvDbl = 9 * kPgs4Rows;    // doubles per environment slot: force fprev fmom prod b ainv ad lo hi
constexpr int kPgs4EnvInt = 2 * kPgs4Rows + 4;    // ints: current order, next order, one dummy word per lane
// slot stride in 4-byte words == 8 (mod 32): the eight environments of a warp start on different bank groups
constexpr int kPgs4SlotWords = ((2 * kPgs4EnvDbl + kPgs4EnvInt + 31) / 32) * 32 + 8;
constexpr int kPgs4SmemBytes = 8 * kPgs4SlotWords * 4;

struct Pgs4Env {
  int nefc, ne, nf;
  const double* gAR;
  double* slot;   // shared-memory slot of this environment
};

// one LCG step + output permutation of PCG32 (engine_solver.c:240-255), then the Fisher-Yates swap of
// position idx (engine_solver.c:258-265) in `ord`.  Branch-free, so that the scheduler can interleave it with
// the fp64 chain of the row: every lane of the group tracks the generator; lane 0 swaps in `ord`, the other
// lanes (and disabled steps) "swap" their private dummy word at ord[dummy] with itself.
__device__ __forceinline__ void pgs4_fy_step(uint64_t& st, int* ord, int idx, bool pred, bool writer, int dummy) {
  const uint64_t old = st;
  const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
  const uint32_t out = (xs >> rot) | (xs << ((0u - rot) & 31));
  const uint32_t n = (uint32_t)(idx > 0 ? idx + 1 : 1);
  const int j = (int)(out % n);
  const bool doit = pred && writer;
  const int ia = doit ? idx : dummy, ib = doit ? j : dummy;
  const int t = ord[ia], u = ord[ib];
  ord[ia] = u;
  ord[ib] = t;
  st = pred ? old * 6364136223846793005ULL + 1ULL : old;
}

// all sweeps of the (up to) eight environments of this warp; returns the iteration count of the lane's environment
template <int NQ>
__device__ __forceinline__ int pgs4_sweeps(const Options& opt, int nv, bool act, const Pgs4Env& E, int k) {
  constexpr int R = kPgs4Rows;
  const unsigned full = 0xffffffffu;
  double* s_force = E.slot;
  double* s_fprev = s_force + R;
  double* s_fmom = s_fprev + R;
  double* s_prod = s_fmom + R;
  const double* s_b = s_prod + R;
  const double* s_ainv = s_b + R;
  const double* s_ad = s_ainv + R;
  const double* s_lo = s_ad + R;
  const double* s_hi = s_lo + R;
  int* ordA = (int*)(E.slot + 9 * R);
  int* ordB = ordA + R;
  const int dummyA = 2 * R + k, dummyB = R + k;   // the lane's dummy word, as an index relative to ordA / ordB
  const int nefc = E.nefc;
  const int n4 = nefc & ~3, nq = n4 >> 2, tail = nefc - n4;
  const double scale = 1 / (opt.meaninertia * (nv > 1 ? nv : 1));
  const double tolerance = opt.tolerance;
  const int maxiter = opt.iterations;
  const double* __restrict__ gAR = E.gAR + k;

  // chain forces in registers; the (up to three) tail forces are kept by every lane
  // (with their previous / extrapolated values: no lane ever reads a tail value another lane writes)
  double F[NQ], T[3], TP[3], TM[3];
#pragma unroll
  for (int q = 0; q < NQ; q++) F[q] = (act && q < nq) ? s_force[4 * q + k] : 0.0;
#pragma unroll
  for (int t = 0; t < 3; t++) { T[t] = (act && t < tail) ? s_force[n4 + t] : 0.0; TP[t] = T[t]; TM[t] = T[t]; }
  if (act) for (int c = k; c < nefc; c += 4) { ordA[c] = c; s_fprev[c] = s_force[c]; }
  else if (k == 0) { ordA[0] = 0; ordB[0] = 0; }   // idle lanes index row 0 of their (unused) slot
  __syncwarp();

  // generator after the reference's warm-up draw (state 0, inc 1); the order of sweep 0 is shuffled up front
  uint64_t rng = 1ULL;
  {
    const int top = __reduce_max_sync(full, act ? nefc : 0);
    for (int idx = top - 1; idx >= 1; idx--) pgs4_fy_step(rng, ordA, idx, act && idx < nefc, k == 0, dummyA);
  }
  __syncwarp();

  int iter = 0, nk = 0;
  int fy_dummy = dummyB;   // dummy word relative to the buffer that currently plays ordB
  bool done = !act;
  while (!__all_sync(full, done)) {
    // ---- Nesterov extrapolation + projection (engine_solver.c:520-556), element-wise on the lane's own forces
    if (!done) {
      double beta = 0;
      if (iter > 0) beta = (double)(nk - 1) / (double)(nk + 2);
      if (beta > 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          if (q < nq) {
            const int c = 4 * q + k;
            const double fs = F[q];
            double f = fs + beta * (fs - s_fprev[c]);
            s_fprev[c] = fs;
            f = dclip(f, s_lo[c], s_hi[c]);
            F[q] = f; s_fmom[c] = f; s_force[c] = f;
          }
        }
#pragma unroll
        for (int t = 0; t < 3; t++) {
          if (t < tail) {
            const int c = n4 + t;
            const double fs = T[t];
            double f = fs + beta * (fs - TP[t]);
            TP[t] = fs;
            f = dclip(f, s_lo[c], s_hi[c]);
            T[t] = f; TM[t] = f;
            if (k == 0) s_force[c] = f;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NQ; q++) if (q < nq) { const int c = 4 * q + k; s_fprev[c] = F[q]; s_fmom[c] = F[q]; }
#pragma unroll
        for (int t = 0; t < 3; t++) { TP[t] = T[t]; TM[t] = T[t]; }
      }
      for (int c = k; c < nefc; c += 4) ordB[c] = ordA[c];   // the next sweep's shuffle starts from this sweep's order
    }
    __syncwarp();

    // ---- the sweep: rows in shuffled order, the next row's operands in flight while this row updates
    const int maxrow = __reduce_max_sync(full, done ? 0 : nefc);
    const int last = nefc > 0 ? nefc - 1 : 0;
    int i = ordA[0];
    double A[NQ], AT[3];
    {
      const double* row = gAR + (long)i * nefc;
#pragma unroll
      for (int q = 0; q < NQ; q++) A[q] = (q < nq) ? __ldg(row + 4 * q) : 0.0;
#pragma unroll
      for (int t = 0; t < 3; t++) AT[t] = (t < tail) ? __ldg(row - k + n4 + t) : 0.0;
    }
    double cb = s_b[i], cainv = s_ainv[i], cad = s_ad[i], clo = s_lo[i], chi = s_hi[i], old = s_force[i];
    double impr = 0;
    for (int bi = 0; bi < maxrow; bi++) {
      const bool on = !done && bi < nefc;
      // operands of the next row
      const int in = ordA[bi + 1 < nefc ? bi + 1 : last];
      double An[NQ], ATn[3];
      {
        const double* row = gAR + (long)in * nefc;
#pragma unroll
        for (int q = 0; q < NQ; q++) An[q] = (q < nq) ? __ldg(row + 4 * q) : 0.0;
#pragma unroll
        for (int t = 0; t < 3; t++) ATn[t] = (t < tail) ? __ldg(row - k + n4 + t) : 0.0;
      }
      const double nb = s_b[in], nainv = s_ainv[in], nad = s_ad[in], nlo = s_lo[in], nhi = s_hi[in], nold = s_force[in];
      // one Fisher-Yates step of the next sweep's order
      {
        const int idx = nefc - 1 - bi;
        pgs4_fy_step(rng, ordB, idx, on && idx >= 1, k == 0, fy_dummy);
      }
      // residual: chain k of the mju_dot structure, then the butterfly (r0+r2)+(r1+r3), then the tail
      double r = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) r += A[q] * F[q];
      const double v = r + __shfl_xor_sync(full, r, 2);
      double res = v + __shfl_xor_sync(full, v, 1);
      res += (AT[0] * T[0] + AT[1] * T[1]) + AT[2] * T[2];
      res = cb + res;
      // projected update with the cost-change guard (engine_solver.c:216-237,600-660)
      double f = old - res * cainv;
      f = f < clo ? clo : (f > chi ? chi : f);
      const double delta = f - old;
      double change = 0.5 * delta * delta * cad + delta * res;
      if (change > 1e-10) { f = old; change = 0; }
      {   // commit (selects, no branch: rows past an environment's end and finished environments change nothing)
        const int sel = (on && i < n4 && (i & 3) == k) ? (i >> 2) : -1;
        const int tsel = on ? i - n4 : -1;
#pragma unroll
        for (int q = 0; q < NQ; q++) F[q] = (q == sel) ? f : F[q];
#pragma unroll
        for (int t = 0; t < 3; t++) T[t] = (t == tsel) ? f : T[t];
        if (on && k == 0) s_force[i] = f;
        impr -= on ? change : 0.0;
      }
      i = in;
#pragma unroll
      for (int q = 0; q < NQ; q++) A[q] = An[q];
#pragma unroll
      for (int t = 0; t < 3; t++) AT[t] = ATn[t];
      cb = nb; cainv = nainv; cad = nad; clo = nlo; chi = nhi; old = nold;
    }

    // ---- gradient restart test (engine_solver.c:690-712): serial-order sum of the per-row products
    bool restart = false;
    const bool want = !done && iter > 0;
    if (want) {
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        if (q < nq) { const int c = 4 * q + k; const double fm = s_fmom[c]; s_prod[c] = (F[q] - fm) * (fm - s_fprev[c]); }
      }
      if (k == 0) {
#pragma unroll
        for (int t = 0; t < 3; t++) {
          if (t < tail) s_prod[n4 + t] = (T[t] - TM[t]) * (TM[t] - TP[t]);
        }
      }
    }
    __syncwarp();
    if (__any_sync(full, want)) {
      const int top = __reduce_max_sync(full, want ? nefc : 0);
      double dce = 0;
      for (int c = 0; c < top; c++) { const double p = s_prod[c < nefc ? c : 0]; dce = (want && c < nefc) ? dce + p : dce; }
      restart = dce < 0;
    }
    if (!done) {
      if (restart) nk = 0; else nk++;
      iter++;
      if (impr * scale < tolerance || iter >= maxiter) done = true;
    }
    { int* t = ordA; ordA = ordB; ordB = t; fy_dummy = (fy_dummy == dummyB) ? dummyA : dummyB; }
    __syncwarp();
  }
  return iter;
}

// grid: one warp per eight environments.  flags bit1/bit2: rollout skip rule (step_skip written by the
// position launch of this step).
__global__ void __launch_bounds__(32) k_pgs4(DModel m, Batch b, int flags) {
  extern __shared__ double pgs4_smem[];
  const unsigned full = 0xffffffffu;
  const int l = threadIdx.x, g = l >> 2, k = l & 3;
  const int e = blockIdx.x * 8 + g;
  bool act = e < b.nenv;
  Env d(m, b, act ? e : b.nenv - 1, k, 4);
  d.mask = 0xFu << (4 * g);
  d.solver = SOL_PGS;
  d.feat = 0;
  if (act && (flags & 6)) act = d.step_skip()[0] == 0;
  const int nefc = act ? d.nefc()[0] : 0;
  act = act && nefc > 0;
  const int top = __reduce_max_sync(full, nefc);
  if (top == 0) return;
  if (top > 4 * 16 + 3) {   // oversized problem in this warp: generic sweeps, each environment on its four lanes
    if (act) solve_pgs(d);
    return;
  }
  Pgs4Env E;
  E.nefc = nefc; E.ne = act ? d.ne()[0] : 0; E.nf = act ? d.nf()[0] : 0;
  E.gAR = d.efc_AR().p;
  E.slot = (double*)((int*)pgs4_smem + (size_t)g * kPgs4SlotWords);
  constexpr int R = kPgs4Rows;
  if (act) {   // stage the vectors; projection bounds and diagonal terms per row (engine_solver.c:91-124 ARdiaginv)
    const double* gf = d.efc_force().p; const double* gb = d.efc_b().p; const double* gfl = d.efc_frictionloss().p;
    double* S = E.slot;
    for (int c = k; c < nefc; c += 4) {
      S[c] = gf[c];
      S[4 * R + c] = gb[c];
      const double fl = gfl[c];
      const double ai = 1 / E.gAR[(long)c * (nefc + 1)];
      S[5 * R + c] = ai;
      S[6 * R + c] = 1 / ai;     // the reference's Athis[0] = 1 / ARinv
      S[7 * R + c] = (c < E.ne) ? -HUGE_VAL : (c < E.ne + E.nf) ? -fl : 0.0;   // equality rows are unbounded
      S[8 * R + c] = (c >= E.ne && c < E.ne + E.nf) ? fl : HUGE_VAL;
    }
  }
  __syncwarp();
  int iter;
  if (top <= 4 * 4 + 3) iter = pgs4_sweeps<4>(m.opt, m.sz.nv, act, E, k);
  else if (top <= 4 * 8 + 3) iter = pgs4_sweeps<8>(m.opt, m.sz.nv, act, E, k);
  else iter = pgs4_sweeps<16>(m.opt, m.sz.nv, act, E, k);
  __syncwarp();
  if (act) {
    double* gfo = d.efc_force().p;
    for (int c = k; c < nefc; c += 4) gfo[c] = E.slot[c];
    if (k == 0) d.solver_niter()[0] += iter;
    __syncwarp(d.mask);
    dual_state_ptr(d, gfo, d.efc_frictionloss().p, nefc, E.ne, E.nf);
  }
}

namespace backend {
int pgs4_smem_bytes() { return kPgs4SmemBytes; }
void launch_pgs4(const DModel& dm, const Batch& b, int flags, void* stream) {
  static bool once = false;
  if (!once) { cudaFuncSetAttribute(k_pgs4, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgs4SmemBytes); once = true; }
  const int grid = (b.nenv + 7) / 8;
  k_pgs4<<<grid, 32, kPgs4SmemBytes, (cudaStream_t)stream>>>(dm, b, flags);
}
}  // namespace backend

}  // namespace mjb
