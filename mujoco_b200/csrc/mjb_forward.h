// Pipeline driver of the batched mj_step path, one environment per call.
//
// Replaces (reference file:line) src/engine/engine_forward.c: mj_checkPos/Vel/Acc :54-113,
// mj_fwdPosition :131-177, mj_fwdVelocity :181-225, mj_fwdActuation :353-1003 (motor / affine
// gain-bias family on hinge and slide joints), mj_fwdAcceleration :1007-1052, mj_fwdConstraint
// :1148-1252, mj_EulerSkip :1398-1476, mj_advance :1261-1395, mj_step :1846-1880; and
// src/engine/engine_support.c mj_integratePos :639-680.
//
// A step is cut into four stages so that the constraint solve can run as its own kernel with a
// different thread mapping:
//   stage 0  position   : checks, kinematics, com, tendon, M, L'DL, collision, constraint rows,
//                         dual projection (PGS), transmission
//   stage 1  velocity   : velocity, passive, reference, bias, actuation, smooth acceleration,
//                         efc_b and the solver start point (warmstart)
//   stage 2  solve      : PGS (or Newton) iterations
//   stage 3  integrate  : dual finish, acceleration check, semi-implicit Euler, advance
#pragma once
#include "mjb_collision.h"
#include "mjb_constraint.h"
#include "mjb_newton.h"

namespace mjb {

// mj_resetData for one environment: qpos0, zero velocities/controls/warmstart/time/warnings
MJB_HD void reset_env(const Env& d, bool clear_warnings) {
  const DModel& m = d.m;
  FD qpos = d.qpos(), qvel = d.qvel(), ctrl = d.ctrl(), ws = d.qacc_warmstart(), qa = d.qfrc_applied();
  for (int i = 0; i < m.sz.nq; i++) qpos[i] = m.qpos0[i];
  for (int i = 0; i < m.sz.nv; i++) { qvel[i] = 0; ws[i] = 0; qa[i] = 0; d.qacc()[i] = 0; }
  for (int i = 0; i < m.sz.nu; i++) ctrl[i] = 0;
  d.time()[0] = 0;
  d.ncon()[0] = 0; d.nefc()[0] = 0; d.ne()[0] = 0; d.nf()[0] = 0; d.nl()[0] = 0; d.solver_niter()[0] = 0;
  if (clear_warnings) for (int i = 0; i < NWARNING; i++) d.warning()[i] = 0;
}

// scan for NaN / overflow; on failure raise the warning and (unless disabled) reset the env
MJB_HD bool check_vec(const Env& d, FD v, int n, int warn) {
  for (int i = 0; i < n; i++) {
    if (is_bad(v[i])) {
      if (!(d.m.opt.disableflags & DSBL_AUTORESET)) reset_env(d, true);   // clears warnings, like mj_resetData
      d.warning()[warn] += 1;
      return true;
    }
  }
  return false;
}

MJB_HD void fwd_position(const Env& d) {
  kinematics(d);
  com_pos(d);
  tendon(d);
  make_M(d);
  {
    const DModel& m = d.m;
    FD M = d.M(), qLD = d.qLD();
    for (int i = 0; i < m.sz.nC; i++) qLD[i] = M[i];
    factor_I(m, qLD, d.qLDiagInv());
  }
  collision(d);
  make_constraint(d);
  project_constraint(d);
  transmission(d);
}

MJB_HD void fwd_velocity(const Env& d) {
  const DModel& m = d.m;
  FD qvel = d.qvel();
  // tendon and actuator velocities (sparse row dots)
  FD tv = d.ten_velocity(), tJ = d.ten_J();
  for (int i = 0; i < m.sz.ntendon; i++) {
    const int adr = m.ten_J_rowadr[i], nnz = m.ten_J_rownnz[i];
    tv[i] = dot_sparse_ref(nnz, [&](int c) { return tJ[adr + c]; }, [&](int c) { return qvel[m.ten_J_colind[adr + c]]; });
  }
  FD av = d.actuator_velocity(), mom = d.actuator_moment();
  if (!(m.opt.disableflags & DSBL_ACTUATION)) {
    for (int i = 0; i < m.sz.nu; i++) {
      const int dof = m.jnt_dofadr[m.actuator_trnjnt[i]];
      double r = 0;
      r += mom[i] * qvel[dof];
      av[i] = r;
    }
  } else {
    for (int i = 0; i < m.sz.nu; i++) av[i] = 0;
  }
  com_vel(d);
  passive(d);
  reference_constraint(d);
  rne_bias(d);
  // tendon-armature bias: needs d/dt(ten_J), identically zero for fixed tendons -> no contribution
}

MJB_HD void fwd_actuation(const Env& d) {
  const DModel& m = d.m;
  const int nu = m.sz.nu, nv = m.sz.nv;
  FD force = d.actuator_force(), qfa = d.qfrc_actuator(), ctrl_in = d.ctrl();
  for (int i = 0; i < nu; i++) force[i] = 0;
  if (nu == 0 || (m.opt.disableflags & DSBL_ACTUATION)) { for (int i = 0; i < nv; i++) qfa[i] = 0; return; }
  FD ctrl = d.scr_nv() + 4 * nv;   // local, clamped copy (nu <= 4*nv guaranteed by the scratch size check)
  for (int i = 0; i < nu; i++) ctrl[i] = ctrl_in[i];
  if (!(m.opt.disableflags & DSBL_CLAMPCTRL)) {
    for (int i = 0; i < nu; i++)
      if (m.actuator_ctrllimited[i]) ctrl[i] = dclip(ctrl[i], m.actuator_ctrlrange[2 * i], m.actuator_ctrlrange[2 * i + 1]);
  }
  for (int i = 0; i < nu; i++) {
    if (is_bad(ctrl[i])) {
      d.warning()[WARN_BADCTRL] += 1;
      for (int k = 0; k < nu; k++) ctrl[k] = 0;
      break;
    }
  }
  FD len = d.actuator_length(), vel = d.actuator_velocity(), mom = d.actuator_moment();
  for (int i = 0; i < nu; i++) {
    const double* gp = m.actuator_gainprm + kNGain * i;
    const double* bp = m.actuator_biasprm + kNGain * i;
    double gain = (m.actuator_gaintype[i] == GAIN_FIXED) ? gp[0] : gp[0] + gp[1] * len[i] + gp[2] * vel[i];
    force[i] = gain * ctrl[i];
    double bias = (m.actuator_biastype[i] == BIAS_NONE) ? 0.0 : bp[0] + bp[1] * len[i] + bp[2] * vel[i];
    force[i] += bias;
  }
  for (int i = 0; i < nu; i++) {
    if (!m.actuator_forcelimited[i]) continue;
    force[i] = dclip(force[i], m.actuator_forcerange[2 * i], m.actuator_forcerange[2 * i + 1]);
  }
  for (int i = 0; i < nv; i++) qfa[i] = 0;
  for (int i = 0; i < nu; i++) {
    const double s = force[i];
    if (s == 0) continue;
    qfa[m.jnt_dofadr[m.actuator_trnjnt[i]]] += mom[i] * s;
  }
  for (int j = 0; j < m.sz.njnt; j++) {
    if (!m.jnt_actfrclimited[j]) continue;
    const int da = m.jnt_dofadr[j];
    qfa[da] = dclip(qfa[da], m.jnt_actfrcrange[2 * j], m.jnt_actfrcrange[2 * j + 1]);
  }
}

MJB_HD void fwd_acceleration(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  FD qfs = d.qfrc_smooth(), qas = d.qacc_smooth();
  FD fp = d.qfrc_passive(), fb = d.qfrc_bias(), fa = d.qfrc_applied(), fact = d.qfrc_actuator();
  for (int i = 0; i < nv; i++) {
    double s = fp[i] - fb[i];
    s += fa[i];
    s += fact[i];
    qfs[i] = s;
  }
  for (int i = 0; i < nv; i++) qas[i] = qfs[i];
  solve_LD(m, qas, d.qLD(), d.qLDiagInv());
}

// position integration on the configuration manifold
MJB_HD void integrate_pos(const Env& d, double dt) {
  const DModel& m = d.m;
  FD qpos = d.qpos(), qvel = d.qvel();
  for (int j = 0; j < m.sz.njnt; j++) {
    int pa = m.jnt_qposadr[j], va = m.jnt_dofadr[j];
    const int jt = m.jnt_type[j];
    if (jt == JNT_FREE || jt == JNT_BALL) {
      if (jt == JNT_FREE) {
        for (int i = 0; i < 3; i++) qpos[pa + i] += dt * qvel[va + i];
        pa += 3; va += 3;
      }
      Q4 q = qintegrate(ld4(qpos, pa), ld3(qvel, va), dt);
      st4(qpos, pa, q);
    } else {
      qpos[pa] += dt * qvel[va];
    }
  }
}

// semi-implicit Euler with implicit joint damping, then advance state and time
MJB_HD void euler_advance(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  const double h = m.opt.timestep;
  FD qacc = d.qacc(), qvel = d.qvel();
  FD acc = d.scr_nv();   // acceleration used for the velocity update
  if (!m.opt.eulerdamp) {
    for (int i = 0; i < nv; i++) acc[i] = qacc[i];
  } else {
    FD qH = d.qH(), M = d.M();
    for (int i = 0; i < m.sz.nC; i++) qH[i] = M[i];
    for (int i = 0; i < nv; i++) {
      const double dd = d_xpoly_force(m.dof_damping_eff[i], m.dof_dampingpoly_eff + kNPoly * i, kNPoly, qvel[i], true);
      qH[m.M_rowadr[i] + m.M_rownnz[i] - 1] += h * dd;
    }
    factor_I(m, qH, d.qHDiagInv());
    FD qfs = d.qfrc_smooth(), qfc = d.qfrc_constraint();
    for (int i = 0; i < nv; i++) acc[i] = qfs[i] + qfc[i];
    solve_LD(m, acc, qH, d.qHDiagInv());
  }
  for (int i = 0; i < nv; i++) qvel[i] += acc[i] * h;
  integrate_pos(d, h);
  d.time()[0] += h;
  FD ws = d.qacc_warmstart();
  for (int i = 0; i < nv; i++) ws[i] = qacc[i];
}

// ---- stages ---------------------------------------------------------------------------------------
// Every stage is entered by ALL lanes that share the environment (one lane in lane-per-env mode,
// 32 in warp-per-env mode).  Cooperative routines use MJB_PFOR / MJB_PSYNC internally; routines
// that are still serial run on lane 0 through MJB_SERIAL.
#define MJB_SERIAL(call)                                   \
  do {                                                     \
    if (d.lane == 0) {                                     \
      const Env mjb_s_(d.m, d.b, d.e, 0, 1);               \
      const Env& d = mjb_s_;                               \
      call;                                                \
    }                                                      \
    MJB_PSYNC();                                           \
  } while (0)

MJB_HD void stage_position(const Env& d, bool is_step) {
  if (is_step) {
    MJB_SERIAL(check_vec(d, d.qpos(), d.m.sz.nq, WARN_BADQPOS); check_vec(d, d.qvel(), d.m.sz.nv, WARN_BADQVEL));
  }
  MJB_SERIAL(fwd_position(d));
}
MJB_HD void stage_velocity(const Env& d) {
  MJB_SERIAL(fwd_velocity(d); fwd_actuation(d); fwd_acceleration(d); constraint_begin(d));
}
MJB_HD void stage_solve(const Env& d) {
  if (d.m.opt.solver == SOL_PGS) MJB_SERIAL(solve_pgs(d));
  else MJB_SERIAL(solve_newton(d));
}
MJB_HD void stage_finish_forward(const Env& d) {
  if (d.m.opt.solver == SOL_PGS) MJB_SERIAL(dual_finish(d));
}
MJB_HD void stage_integrate(const Env& d) {
  stage_finish_forward(d);
  MJB_SERIAL(d.scr_int()[0] = check_vec(d, d.qacc(), d.m.sz.nv, WARN_BADQACC) ? 1 : 0);
  if (d.scr_int()[0]) {
    // mj_checkAcc: after the reset the reference re-runs mj_forward before integrating
    if (!(d.m.opt.disableflags & DSBL_AUTORESET)) {
      stage_position(d, false);
      stage_velocity(d);
      stage_solve(d);
      stage_finish_forward(d);
    }
  }
  MJB_SERIAL(euler_advance(d));
}

}  // namespace mjb
