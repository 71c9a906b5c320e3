// Projected Gauss-Seidel on FOUR lanes per environment (device code shared by k_pgs4, mjb_pgs4.cu, and the
// persistent rollout kernel, mjb_krollout.cu)
//
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
//   * the visiting order of sweep s of a problem with n rows is a constant of (n, s) - one PCG32 stream per
//     solve, Fisher-Yates from the identity - so all orders come from a table built once per device
//     (pgs4_order_table): no generator, modulo or swap is left in the row loop;
//   * zero padding is exact: a chain sum is never -0 (it starts from +0), so adding a +-0 product of
//     the padded positions returns the same bits; the same holds for the padded tail terms.
// Results are bit-identical to the serial reference arithmetic (tests/test_gpu_parity.py compares
// forces, states and iteration counts).
//
// Environments with more rows than the largest register class (4*16+3) take the generic sweep of
// mjb_constraint.h (global memory, any lane count) on their four lanes.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "mjb_backend.h"
#include "mjb_stage.h"
#include "mjb_kstep.h"

namespace mjb {

constexpr int kPgs4Rows = 68;                 // row capacity of the largest register class (4 * 16 + 3, padded)
// shared memory of one warp (= one CTA): 48 KB.
//   staged layout (the normal case): the eight environments are packed back to back, each with
//     nefc records of 4 nqp + 6 doubles (nqp = the environment's own chain length, rounded to even) - one record
//     per row of AR, re-laid so that every lane finds ITS chain elements contiguous: [lane 0: a_0 a_4 ..]
//     [lane 1: a_1 a_5 ..][lane 2][lane 3][tail a_n4 a_n4+1 a_n4+2, b][1/AR_ii, AR_ii] - then a zero pad; after the
//     records of all eight environments, six vectors of nefc doubles each (force, fprev, fmom, restart products,
//     lo, hi: the infinite bounds stay out of reach of the padded chain loads);
//   slot layout (fallback when the packed records of a warp exceed the budget): fixed slots of kPgs4Rows rows with
//     the vectors and six row constants; AR rows are then read from global memory two rows ahead.
constexpr int kPgs4SmemBytes = 56 * 1024;
constexpr int kPgs4Beta = 128;                // entries of the momentum-coefficient table at the end of the shared memory
constexpr int kPgs4EnvDbl = 10 * kPgs4Rows;   // slot layout: force fprev fmom prod + 6 per row (b ainv ad lo hi -)
// slot stride in 4-byte words == 8 (mod 32): the eight environments of a warp start on different bank groups
constexpr int kPgs4SlotWords = ((2 * kPgs4EnvDbl + 31) / 32) * 32 + 8;
static_assert(kPgs4SlotWords % 32 == 8 && kPgs4SlotWords % 4 == 0, "slot stride: bank spread and 16-byte alignment");
static_assert(8 * kPgs4SlotWords * 4 + 8 * kPgs4Beta <= kPgs4SmemBytes, "slot layout fits the warp's shared memory");

#ifdef MJB_PGS4_PROF
static __device__ long long g_pgs4_prof[8192][8];   // per warp: total cycles, cycles inside row loops, rows, sweeps
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

struct Pgs4Env {
  int nefc, ne, nf;
  const double* gAR;
  double* slot;               // shared memory of this environment: its slot, or its packed records
  double* vec;                // the four vectors (force, fprev, fmom, prod), stride vstride
  int vstride;
  int recd, nqp;              // packed layout: doubles per record (4 nqp + 6), chain elements per lane (own nq, rounded to even)
  const double* beta;         // (n - 1) / (n + 2) for n < kPgs4Beta (shared memory)
  const unsigned char* ord;   // visiting orders of a problem with nefc rows: [sweep][position] (pgs4_order_table)
};

// all sweeps of the (up to) eight environments of this warp; returns the iteration count of the lane's environment.
// STAGED: row operands come from the packed records in shared memory (one row ahead); otherwise AR rows come from
// global memory DEPTH rows ahead (2 where the registers allow it).
template <int NQ, int DEPTH, bool STAGED>
__device__ __forceinline__ int pgs4_sweeps(const Options& opt, int nv, bool act, const Pgs4Env& E, int k) {
  constexpr int R = kPgs4Rows;
  const int RECD = E.recd;
  const unsigned full = 0xffffffffu;
  double* s_force = E.vec;
  double* s_fprev = s_force + E.vstride;
  double* s_fmom = s_fprev + E.vstride;
  double* s_prod = s_fmom + E.vstride;
  // slot layout: [row][6] = b, 1/AR_ii, AR_ii, lo, hi, -     staged layout: the records
  const double* s_lo = s_prod + E.vstride;   // (staged layout only)
  const double* s_hi = s_lo + E.vstride;
  const double* s_rc = STAGED ? E.slot : E.slot + 4 * R;
  const double* recA = E.slot + k * E.nqp;   // staged: this lane's chain elements of row 0
  const double* recT = E.slot + 4 * E.nqp;   // staged: tail elements and constants of row 0
  const int nefc = E.nefc;
  const int n4 = nefc & ~3, nq = n4 >> 2, tail = nefc - n4;
  const double scale = 1 / (opt.meaninertia * (nv > 1 ? nv : 1));
  const double tolerance = opt.tolerance;
  const int maxiter = opt.iterations;
  const double* __restrict__ gAR = E.gAR + k;
  const int last = nefc > 0 ? nefc - 1 : 0;
  auto lo_of = [&](int c) { return STAGED ? s_lo[c] : s_rc[6 * c + 3]; };
  auto hi_of = [&](int c) { return STAGED ? s_hi[c] : s_rc[6 * c + 4]; };

  // chain forces in registers; the (up to three) tail forces are kept by every lane
  // (with their previous / extrapolated values: no lane ever reads a tail value another lane writes)
  double F[NQ], T[3], TP[3], TM[3];
#pragma unroll
  for (int q = 0; q < NQ; q++) F[q] = (act && q < nq) ? s_force[4 * q + k] : 0.0;
#pragma unroll
  for (int t = 0; t < 3; t++) { T[t] = (act && t < tail) ? s_force[n4 + t] : 0.0; TP[t] = T[t]; TM[t] = T[t]; }
  double TLo[3], THi[3];
#pragma unroll
  for (int t = 0; t < 3; t++) { TLo[t] = (act && t < tail) ? lo_of(n4 + t) : 0.0; THi[t] = (act && t < tail) ? hi_of(n4 + t) : 0.0; }
  if (act) for (int c = k; c < nefc; c += 4) s_fprev[c] = s_force[c];
  __syncwarp();

  // operands of row i: AR elements of this lane's chain, the three tail elements
  auto load_row = [&](int i, double (&A)[NQ], double (&AT)[4]) {
    if (STAGED) {
      const double2* pa = (const double2*)(recA + i * RECD);
#pragma unroll
      for (int q = 0; q < NQ / 2; q++) { const double2 v = pa[q]; A[2 * q] = v.x; A[2 * q + 1] = v.y; }
      if (NQ & 1) A[NQ - 1] = (recA + i * RECD)[NQ - 1];
      const double2* pt = (const double2*)(recT + i * RECD);
      const double2 t0 = pt[0], t1 = pt[1];
      AT[0] = t0.x; AT[1] = t0.y; AT[2] = t1.x; AT[3] = t1.y;   // AT[3] = b
    } else {
      const double* row = gAR + (long)i * nefc;
#pragma unroll
      for (int q = 0; q < NQ; q++) A[q] = (q < nq) ? __ldg(row + 4 * q) : 0.0;
#pragma unroll
      for (int t = 0; t < 3; t++) AT[t] = (t < tail) ? __ldg(row - k + n4 + t) : 0.0;
      AT[3] = s_rc[6 * i];
    }
  };
  // constants of row i: 1/AR_ii, AR_ii, lo, hi
  auto load_consts = [&](int i, double2& c01, double2& c23) {
    if (STAGED) {
      const double2* pt = (const double2*)(recT + i * RECD);
      c01 = pt[2]; c23 = make_double2(s_lo[i], s_hi[i]);
    } else {
      c01 = make_double2(s_rc[6 * i + 1], s_rc[6 * i + 2]);
      c23 = make_double2(s_rc[6 * i + 3], s_rc[6 * i + 4]);
    }
  };

  int iter = 0, nk = 0;
  bool done = !act;
  // first three rows of the coming sweep's visiting order, fetched one momentum section ahead (the table is read
  // through L2 more often than not: about 300 cycles that would otherwise sit in front of every sweep)
  int n0 = __ldg(E.ord), n1 = __ldg(E.ord + (1 < nefc ? 1 : last)), n2 = __ldg(E.ord + (2 < nefc ? 2 : last));
  PROF(long long prof_rows = 0; long long prof_cyc = 0; long long prof_sw = 0; long long prof_mom = 0, prof_prime = 0, prof_dce = 0;)
  while (!__all_sync(full, done)) {
    PROF(const long long pm0 = clock64();)
    // ---- Nesterov extrapolation + projection (engine_solver.c:520-556), element-wise on the lane's own forces
    if (!done) {
      double beta = 0;   // (nk - 1) / (nk + 2): from the table (an fp64 division costs about one row of the sweep)
      if (iter > 0) beta = nk < kPgs4Beta ? E.beta[nk] : (double)(nk - 1) / (double)(nk + 2);
      if (beta > 0) {
        double fp[NQ], lo[NQ], hi[NQ];   // all loads first: the compiler cannot prove the shared arrays disjoint
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const int c = (q < nq) ? 4 * q + k : k;
          fp[q] = s_fprev[c]; lo[q] = lo_of(c); hi[q] = hi_of(c);
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const double fs = F[q];
          const double f = dclip(fs + beta * (fs - fp[q]), lo[q], hi[q]);
          fp[q] = fs;
          if (q < nq) F[q] = f;
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          if (q < nq) { const int c = 4 * q + k; s_fprev[c] = fp[q]; s_fmom[c] = F[q]; s_force[c] = F[q]; }
        }
#pragma unroll
        for (int t = 0; t < 3; t++) {
          if (t < tail) {
            const int c = n4 + t;
            const double fs = T[t];
            double f = fs + beta * (fs - TP[t]);
            TP[t] = fs;
            f = dclip(f, TLo[t], THi[t]);
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
    }
    __syncwarp();

    // ---- the sweep: rows in the reference's shuffled order; the operands of the next row(s) are in flight while
    // this row updates, so that no load sits between two updates of the recurrence
    PROF(prof_mom += clock64() - pm0;)
    const int maxrow = __reduce_max_sync(full, done ? 0 : nefc);
    const unsigned char* __restrict__ os = E.ord + (done ? 0 : iter) * nefc;
    int i0 = n0, i1 = n1, i2 = n2;
    double Ab[3][NQ], ATb[3][4];   // operand buffers, rotated by renaming (the loop body is unrolled three times)
    load_row(i0, Ab[0], ATb[0]);
    if (DEPTH == 2) load_row(i1, Ab[1], ATb[1]);
    double2 c01, c23;              // 1/AR_ii, AR_ii | lo, hi
    load_consts(i0, c01, c23);
    double old = s_force[i0];
    double impr = 0;
    // one row: `cur` holds this row's AR operands, `nxt` / `far` receive the operands in flight
    auto row_step = [&](int bi, double (&cur)[NQ], double (&curT)[4], double (&nxt)[NQ], double (&nxtT)[4],
                        double (&far)[NQ], double (&farT)[4]) {
      const bool on = !done && bi < nefc;
      // operands in flight: the row after next (DEPTH 2) or the next row (DEPTH 1); constants of the next row
      const int i3 = __ldg(os + (bi + 3 < nefc ? bi + 3 : last));
      if (DEPTH == 2) load_row(i2, far, farT); else load_row(i1, nxt, nxtT);
      double2 n01, n23;
      load_consts(i1, n01, n23);
      const double nold = s_force[i1];
      // residual: chain k of the mju_dot structure, then (r0+r2)+(r1+r3), then the tail
      double r = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) r += cur[q] * F[q];
      // ONE shuffle latency: the three other chain sums arrive together; on odd lanes the two pairs swap roles,
      // and IEEE addition commutes
      const double r1 = __shfl_xor_sync(full, r, 1), r2 = __shfl_xor_sync(full, r, 2), r3 = __shfl_xor_sync(full, r, 3);
      double res = (r + r2) + (r1 + r3);
      res += (curT[0] * T[0] + curT[1] * T[1]) + curT[2] * T[2];
      res = curT[3] + res;
      // projected update with the cost-change guard (engine_solver.c:216-237,600-660).  (Evaluating the guard
      // speculatively - commit first, roll back when it fires - was measured: the branch it needs inside the row
      // costs the scheduler more than the 54 cycles it takes off the chain: 471 -> 532 cycles per row.)
      double f = old - res * c01.x;
      f = f < c23.x ? c23.x : (f > c23.y ? c23.y : f);
      const double delta = f - old;
      double change = 0.5 * delta * delta * c01.y + delta * res;
      if (change > 1e-10) { f = old; change = 0; }
      {   // commit (selects, no branch: rows past an environment's end and finished environments change nothing)
        // position of the updated force in this lane's registers as ONE integer (-1: not in this lane / row disabled)
        const int mine = (int)on & (int)(i0 < n4) & (int)((i0 & 3) == k);
        const int sel = (i0 >> 2) | (mine - 1);
        const int tsel = (i0 - n4) | ((int)on - 1);
#pragma unroll
        for (int q = 0; q < NQ; q++) F[q] = (q == sel) ? f : F[q];
#pragma unroll
        for (int t = 0; t < 3; t++) T[t] = (t == tsel) ? f : T[t];
        if (on && k == 0) s_force[i0] = f;
        impr -= on ? change : 0.0;
      }
      i0 = i1; i1 = i2; i2 = i3;
      c01 = n01; c23 = n23; old = nold;
    };
    PROF(const long long pt0 = clock64(); prof_prime += pt0 - pm0;)
#pragma unroll 1
    for (int bi = 0; bi < maxrow; bi += 3) {
      row_step(bi, Ab[0], ATb[0], Ab[1], ATb[1], Ab[2], ATb[2]);
      if (DEPTH == 2) {
        row_step(bi + 1, Ab[1], ATb[1], Ab[2], ATb[2], Ab[0], ATb[0]);
        row_step(bi + 2, Ab[2], ATb[2], Ab[0], ATb[0], Ab[1], ATb[1]);
      } else {   // two buffers alternate; the third call restores the pairing for the next trip
        row_step(bi + 1, Ab[1], ATb[1], Ab[0], ATb[0], Ab[2], ATb[2]);
        row_step(bi + 2, Ab[0], ATb[0], Ab[1], ATb[1], Ab[2], ATb[2]);
#pragma unroll
        for (int q = 0; q < NQ; q++) Ab[0][q] = Ab[1][q];
#pragma unroll
        for (int t = 0; t < 4; t++) ATb[0][t] = ATb[1][t];
      }
    }

    PROF(prof_cyc += clock64() - pt0; prof_rows += (maxrow + 2) / 3 * 3; prof_sw++;)
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
      // every term in registers first, then the serial additions in row order (+ 0 is exact: the sum is never -0)
      double p[4 * NQ + 3];
#pragma unroll
      for (int c = 0; c < 4 * NQ + 3; c++) { const double v = s_prod[c < nefc ? c : 0]; p[c] = (c < nefc) ? v : 0.0; }
      double dce = 0;
#pragma unroll
      for (int c = 0; c < 4 * NQ + 3; c++) dce += p[c];
      restart = want && dce < 0;
    }
    PROF(prof_dce += clock64() - pt0;)
    if (!done) {
      if (restart) nk = 0; else nk++;
      iter++;
      if (impr * scale < tolerance || iter >= maxiter) done = true;
    }
    {
      const unsigned char* __restrict__ on = E.ord + (done ? 0 : iter) * nefc;   // the expression of the sweep's own `os`
      n0 = __ldg(on); n1 = __ldg(on + (1 < nefc ? 1 : last)); n2 = __ldg(on + (2 < nefc ? 2 : last));
    }
    __syncwarp();
  }
  PROF(if (threadIdx.x == 0 && blockIdx.x < 8192) { g_pgs4_prof[blockIdx.x][1] = prof_cyc; g_pgs4_prof[blockIdx.x][2] = prof_rows; g_pgs4_prof[blockIdx.x][3] = prof_sw; g_pgs4_prof[blockIdx.x][4] = prof_mom; g_pgs4_prof[blockIdx.x][5] = prof_prime; g_pgs4_prof[blockIdx.x][6] = prof_dce; })
  return iter;
}

// Shared memory of a solve.  k_pgs4 gives every warp a fixed region (Pgs4FixedSmem); the persistent rollout kernel
// lets the PGS warps of a CTA share one pool and hands out exactly the bytes a warp's records need (mjb_krollout.cu).
// acquire() is called once per pgs4_warp call, by all 32 lanes, with a warp-uniform need (0: nothing to stage); it
// returns the region this warp may use (bytes < need: the records do not fit).  beta(): a shared table of
// (n - 1) / (n + 2), or nullptr when the warp builds its own at the end of its region.
struct Pgs4FixedSmem {
  static constexpr int kSlack = 0;     // the fixed region is larger than the records: reads past the last vector stay inside it
  double* base; int bytes;
  __device__ __forceinline__ void acquire(int, double*& p, int& n) { p = base; n = bytes; }
  __device__ __forceinline__ const double* beta() const { return nullptr; }
};

// the solve of (up to) eight environments by one warp: lane group g = lane / 4 works on environment e (e < 0 or
// e >= nenv: idle group).  A region too small for the slot layout cannot hold a warp whose packed records do not fit:
// such a warp takes the generic sweeps.
// flags bit1/bit2: rollout skip rule (step_skip written by the position stage of this step).
template <class Pool>
__device__ __forceinline__ void pgs4_warp(const DModel& m, const Batch& b, int e, int flags, const unsigned char* __restrict__ order_tab,
                                          int order_iters, Pool& pool) {
  double* pgs4_smem = nullptr;
  int smem_bytes = 0;
  PROF(const long long prof_t0 = clock64();)
  const unsigned full = 0xffffffffu;
  const int l = threadIdx.x & 31, g = l >> 2, k = l & 3;
  bool act = e >= 0 && e < b.nenv;
  Env d(m, b, act ? e : b.nenv - 1, k, 4);
  d.mask = 0xFu << (4 * g);
  d.solver = SOL_PGS;
  d.feat = 0;
  if (act && (flags & 6)) act = d.step_skip()[0] == 0;
  const int nefc = act ? d.nefc()[0] : 0;
  act = act && nefc > 0;
  const int top = __reduce_max_sync(full, nefc);
  if (top == 0) { pool.acquire(0, pgs4_smem, smem_bytes); return; }
  if (top > 4 * 16 + 3 || m.opt.iterations > order_iters) {   // oversized problem in this warp: generic sweeps, each environment on its four lanes
    pool.acquire(0, pgs4_smem, smem_bytes);
    if (act) solve_pgs(d);
    return;
  }
  Pgs4Env E;
  E.nefc = nefc; E.ne = act ? d.ne()[0] : 0; E.nf = act ? d.nf()[0] : 0;
  E.gAR = d.efc_AR().p;
  E.ord = order_tab + (size_t)order_iters * (nefc > 0 ? nefc * (nefc - 1) / 2 : 0);
  // packed (staged) layout if the records of the eight environments fit the warp's shared memory
  // register classes (chain elements per lane): a row costs about 28 cycles per chain element, so the classes are fine
  // (every length from 4 to 16)
  const int nqt = (top & ~3) >> 2;   // chain length of the largest problem in the warp
  const int NQc = nqt <= 4 ? 4 : nqt;
  // each environment's records are sized by its OWN row length (a small neighbour of a large problem stays small);
  // the class code reads up to NQc chain elements per lane, so reads past an environment's own chain length land
  // in finite data of the SAME environment (its next lane / record, or the zero pad after its last record) and meet
  // forces that are exactly zero: + (+-0) leaves a chain sum unchanged
  const int nqe_ = (nefc & ~3) >> 2;
  E.nqp = (nqe_ + 1) & ~1;
  E.recd = 4 * E.nqp + 6;
  const int vstr = nefc + (nefc & 1);
  const int myrec = nefc ? nefc * E.recd + 16 : 0, myvec = 6 * vstr;
  int rec_before = 0, rec_total = 0, vec_before = 0, vec_total = 0;
#pragma unroll
  for (int gg = 0; gg < 8; gg++) {
    const int v = __shfl_sync(full, myrec, 4 * gg), w = __shfl_sync(full, myvec, 4 * gg);
    if (gg < g) { rec_before += v; vec_before += w; }
    rec_total += v; vec_total += w;
  }
  const int own_beta = pool.beta() ? 0 : 8 * kPgs4Beta;
  // (the class code loads up to 4 * 16 + 3 elements of a vector whatever the environment's size: an exact-fit region
  // ends with a zeroed slack so that these loads stay inside it and meet finite data)
  pool.acquire((rec_total + vec_total) * 8 + own_beta + Pool::kSlack, pgs4_smem, smem_bytes);
  const bool staged = (rec_total + vec_total) * 8 + Pool::kSlack <= smem_bytes - own_beta && !(flags & 256);   // flags bit8: force the slot layout (tests)
  constexpr int R = kPgs4Rows;
  if (!staged && smem_bytes < 8 * kPgs4SlotWords * 4 + own_beta) {   // no room for the slot layout either
    if (act) solve_pgs(d);
    return;
  }
  if (pool.beta()) E.beta = pool.beta();
  else {
    double* bt = pgs4_smem + (smem_bytes / 8 - kPgs4Beta);
    for (int n = l; n < kPgs4Beta; n += 32) bt[n] = (double)(n - 1) / (double)(n + 2);
    E.beta = bt;
  }
  const double* gf = d.efc_force().p; const double* gb = d.efc_b().p; const double* gfl = d.efc_frictionloss().p;
  if (staged) {
    if (Pool::kSlack) for (int n = l; n < Pool::kSlack / 8; n += 32) pgs4_smem[rec_total + vec_total + n] = 0.0;
    E.slot = pgs4_smem + rec_before;
    E.vec = pgs4_smem + rec_total + vec_before;
    E.vstride = vstr;
    if (!act) { E.slot = pgs4_smem; E.vec = pgs4_smem; E.vstride = 0; E.nqp = 0; E.recd = 6; }   // idle lanes read (never write) valid memory
    if (act) {   // build the records: AR rows re-laid per lane and zero-padded, tail, b, diagonal terms, bounds
      const int n4 = nefc & ~3, nqe = n4 >> 2, tail = nefc - n4;
      for (int i0 = 0; i0 < nefc; i0 += 4) {   // four rows per trip: their loads are in flight together
        double a[4][16], tl[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u < nefc ? i0 + u : nefc - 1;
          const double* row = E.gAR + (long)i * nefc;
#pragma unroll
          for (int q = 0; q < 16; q++) if (q < E.nqp) a[u][q] = (q < nqe) ? row[4 * q + k] : 0.0;
          tl[u] = (k < tail) ? row[n4 + k] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          if (i < nefc) {
            double* rec = E.slot + i * E.recd;
#pragma unroll
            for (int q = 0; q < 16; q++) if (q < E.nqp) rec[k * E.nqp + q] = a[u][q];
            if (k < 3) rec[4 * E.nqp + k] = tl[u];
          }
        }
      }
      // row constants, each lane its own rows (the two IEEE reciprocals per row are the expensive part)
      for (int c = k; c < nefc; c += 4) {
        double* t = E.slot + c * E.recd + 4 * E.nqp;
        const double fl = gfl[c];
        const double ai = __drcp_rn(E.gAR[(long)c * (nefc + 1)]);   // == 1 / x, correctly rounded
        const double lo = (c < E.ne) ? -HUGE_VAL : (c < E.ne + E.nf) ? -fl : 0.0;   // equality rows are unbounded
        const double hi = (c >= E.ne && c < E.ne + E.nf) ? fl : HUGE_VAL;
        t[3] = gb[c];
        t[4] = ai;
        t[5] = __drcp_rn(ai);     // the reference's Athis[0] = 1 / ARinv
        E.vec[4 * E.vstride + c] = lo;
        E.vec[5 * E.vstride + c] = hi;
      }
      for (int c = k; c < nefc; c += 4) E.vec[c] = gf[c];
#pragma unroll
      for (int u = 0; u < 4; u++) E.slot[nefc * E.recd + 4 * u + k] = 0.0;   // the zero pad after the last record
    }
  } else {
    E.slot = (double*)((int*)pgs4_smem + (size_t)g * kPgs4SlotWords);
    E.vec = E.slot;
    E.vstride = R;
    if (act) {   // stage the vectors; projection bounds and diagonal terms per row (engine_solver.c:91-124 ARdiaginv)
      double* S = E.slot;
      for (int c = k; c < nefc; c += 4) {
        S[c] = gf[c];
        double* rc = S + 4 * R + 6 * c;
        rc[0] = gb[c];
        const double fl = gfl[c];
        const double ai = __drcp_rn(E.gAR[(long)c * (nefc + 1)]);   // == 1 / x, correctly rounded
        rc[1] = ai;
        rc[2] = __drcp_rn(ai);     // the reference's Athis[0] = 1 / ARinv
        rc[3] = (c < E.ne) ? -HUGE_VAL : (c < E.ne + E.nf) ? -fl : 0.0;   // equality rows are unbounded
        rc[4] = (c >= E.ne && c < E.ne + E.nf) ? fl : HUGE_VAL;
      }
    }
  }
  __syncwarp();
  PROF(if (threadIdx.x == 0 && blockIdx.x < 8192) g_pgs4_prof[blockIdx.x][7] = clock64() - prof_t0;)
  int iter;
  if (staged) {
    switch (NQc) {
      case 4: iter = pgs4_sweeps<4, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 5: iter = pgs4_sweeps<5, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 6: iter = pgs4_sweeps<6, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 7: iter = pgs4_sweeps<7, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 8: iter = pgs4_sweeps<8, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 9: iter = pgs4_sweeps<9, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 10: iter = pgs4_sweeps<10, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 11: iter = pgs4_sweeps<11, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 12: iter = pgs4_sweeps<12, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 13: iter = pgs4_sweeps<13, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 14: iter = pgs4_sweeps<14, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      case 15: iter = pgs4_sweeps<15, 1, true>(m.opt, m.sz.nv, act, E, k); break;
      default: iter = pgs4_sweeps<16, 1, true>(m.opt, m.sz.nv, act, E, k); break;
    }
  } else {
    if (NQc == 4) iter = pgs4_sweeps<4, 2, false>(m.opt, m.sz.nv, act, E, k);
    else if (NQc <= 8) iter = pgs4_sweeps<8, 2, false>(m.opt, m.sz.nv, act, E, k);
    else iter = pgs4_sweeps<16, 1, false>(m.opt, m.sz.nv, act, E, k);   // (12 takes the 16 code here: the fallback is rare)
  }
  __syncwarp();
  if (act) {
    double* gfo = d.efc_force().p;
    for (int c = k; c < nefc; c += 4) gfo[c] = E.vec[c];
    if (k == 0) d.solver_niter()[0] += iter;
    __syncwarp(d.mask);
    dual_state_ptr(d, gfo, d.efc_frictionloss().p, nefc, E.ne, E.nf);
  }
  PROF(if (threadIdx.x == 0 && blockIdx.x < 8192) g_pgs4_prof[blockIdx.x][0] = clock64() - prof_t0;)
}

namespace backend {
// the device's order table (built on first use; grows with the model's iteration cap): 0 on success
int pgs4_table(const DModel& dm, const unsigned char** tab, int* iters);
int pgs4_force_slots();   // 256 when tests / MJB_PGS4_SLOTS ask for the slot layout
}  // namespace backend

}  // namespace mjb
