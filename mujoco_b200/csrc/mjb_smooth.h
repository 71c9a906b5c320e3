// Smooth dynamics of the batched mj_step path, one environment per call (one GPU lane per env).
//
// Replaces (reference file:line)  src/engine/engine_core_smooth.c  mj_kinematics :40-242,
// mj_comPos :246-350, mj_tendon (fixed tendons) :927-985, mj_transmission (joint) :1265-1330,
// mj_crb/mj_tendonArmature :1845-1971, mj_factorI :1997-2029, mj_solveLD :2033-2117,
// mj_comVel :2179-2245, mj_rne :2328-2390;  src/engine/engine_passive.c mj_springdamper :655-842.
// Same recurrences and operation order, re-expressed over the SoA batch layout (mjb_types.h).
#pragma once
#include "mjb_types.h"

namespace mjb {

// ------------------------------------------------------------------------------------------------
// forward kinematics: body frames down the tree, then inertial and geom frames
MJB_HD void kinematics(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody;
  FD qpos = d.qpos(), xpos = d.xpos(), xquat = d.xquat(), xmat = d.xmat();
  FD xipos = d.xipos(), ximat = d.ximat(), xanchor = d.xanchor(), xaxis = d.xaxis();

  // world body
  st3(xpos, 0, V3{0, 0, 0});
  st4(xquat, 0, Q4{1, 0, 0, 0});
  st3(xipos, 0, V3{0, 0, 0});
  for (int k = 0; k < 9; k++) { xmat[k] = (k % 4 == 0) ? 1.0 : 0.0; ximat[k] = (k % 4 == 0) ? 1.0 : 0.0; }

  for (int i = 1; i < nbody; i++) {
    V3 p; Q4 q;
    const int jadr = m.body_jntadr[i], jnum = m.body_jntnum[i];
    if (jnum == 1 && m.jnt_type[jadr] == JNT_FREE) {
      const int qa = m.jnt_qposadr[jadr];
      p = ld3(qpos, qa);
      q = ld4(qpos, qa + 3);
      normalize(q);
      st3(xanchor, 3 * jadr, p);
      st3(xaxis, 3 * jadr, ldc3(m.jnt_axis, 3 * jadr));
    } else {
      const int pid = m.body_parentid[i];
      V3 bpos = ldc3(m.body_pos, 3 * i);
      Q4 bquat = ldc4(m.body_quat, 4 * i);
      if (pid) {
        p = mulmv(ld9(xmat, 9 * pid), bpos);
        p = p + ld3(xpos, 3 * pid);
        q = qmul(ld4(xquat, 4 * pid), bquat);
      } else {
        p = bpos;
        q = bquat;
      }
      for (int j = 0; j < jnum; j++) {
        const int jid = jadr + j, qa = m.jnt_qposadr[jid], jt = m.jnt_type[jid];
        V3 jaxis = ldc3(m.jnt_axis, 3 * jid), jpos = ldc3(m.jnt_pos, 3 * jid);
        V3 ax = rotate(jaxis, q);
        V3 an = rotate(jpos, q);
        an = an + p;
        if (jt == JNT_SLIDE) {
          p = addscl(p, ax, qpos[qa] - m.qpos0[qa]);
        } else {  // ball or hinge
          Q4 ql;
          if (jt == JNT_BALL) {
            ql = ld4(qpos, qa);
            normalize(ql);
          } else {
            ql = axis_angle(jaxis, qpos[qa] - m.qpos0[qa]);
          }
          q = qmul(q, ql);
          V3 off = rotate(jpos, q);
          p = an - off;
        }
        st3(xanchor, 3 * jid, an);
        st3(xaxis, 3 * jid, ax);
      }
    }
    normalize(q);
    st4(xquat, 4 * i, q);
    st3(xpos, 3 * i, p);
    st9(xmat, 9 * i, quat2mat(q));
  }

  // inertial frames (mj_local2Global, engine_core_util.c)
  for (int i = 1; i < nbody; i++) {
    const int sf = m.body_sameframe[i];
    V3 bp = ld3(xpos, 3 * i);
    M3 bm = ld9(xmat, 9 * i);
    if (sf == SAMEFRAME_BODY) st3(xipos, 3 * i, bp);
    else st3(xipos, 3 * i, mulmv(bm, ldc3(m.body_ipos, 3 * i)) + bp);
    if (sf == SAMEFRAME_NONE) st9(ximat, 9 * i, quat2mat(qmul(ld4(xquat, 4 * i), ldc4(m.body_iquat, 4 * i))));
    else st9(ximat, 9 * i, bm);
  }

  // geom frames
  FD gpos = d.geom_xpos(), gmat = d.geom_xmat();
  const int ngeom = m.sz.ngeom;
  for (int g = 0; g < ngeom; g++) {
    const int b = m.geom_bodyid[g], sf = m.geom_sameframe[g];
    V3 bp = ld3(xpos, 3 * b);
    if (sf == SAMEFRAME_BODY) st3(gpos, 3 * g, bp);
    else if (sf == SAMEFRAME_INERTIA) st3(gpos, 3 * g, ld3(xipos, 3 * b));
    else st3(gpos, 3 * g, mulmv(ld9(xmat, 9 * b), ldc3(m.geom_pos, 3 * g)) + bp);
    if (sf == SAMEFRAME_NONE) st9(gmat, 9 * g, quat2mat(qmul(ld4(xquat, 4 * b), ldc4(m.geom_quat, 4 * g))));
    else if (sf == SAMEFRAME_BODY || sf == SAMEFRAME_BODYROT) st9(gmat, 9 * g, ld9(xmat, 9 * b));
    else st9(gmat, 9 * g, ld9(ximat, 9 * b));
  }
}

// ------------------------------------------------------------------------------------------------
// subtree centres of mass, com-frame inertias (cinert) and motion axes (cdof)
MJB_HD void com_pos(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody;
  FD sc = d.subtree_com(), xipos = d.xipos(), ximat = d.ximat(), xmat = d.xmat();
  FD cinert = d.cinert(), cdof = d.cdof(), xanchor = d.xanchor(), xaxis = d.xaxis();

  for (int i = 0; i < nbody; i++) st3(sc, 3 * i, ld3(xipos, 3 * i) * m.body_mass[i]);
  for (int i = nbody - 1; i > 0; i--) {
    const int p = m.body_parentid[i];
    st3(sc, 3 * p, ld3(sc, 3 * p) + ld3(sc, 3 * i));
  }
  for (int i = 0; i < nbody; i++) {
    if (m.body_subtreemass[i] < kMinVal) st3(sc, 3 * i, ld3(xipos, 3 * i));
    else st3(sc, 3 * i, ld3(sc, 3 * i) * (1.0 / m.body_subtreemass[i]));
  }

  for (int k = 0; k < 10; k++) cinert[k] = 0;
  for (int i = 1; i < nbody; i++) {
    V3 off = ld3(xipos, 3 * i) - ld3(sc, 3 * m.body_rootid[i]);
    st10(cinert, 10 * i, inert_com(ldc3(m.body_inertia, 3 * i), ld9(ximat, 9 * i), off, m.body_mass[i]));
  }

  for (int i = 1; i < nbody; i++) {
    const int jnum = m.body_jntnum[i];
    if (!jnum) continue;
    const int start = m.body_jntadr[i];
    V3 root = ld3(sc, 3 * m.body_rootid[i]);
    for (int j = start; j < start + jnum; j++) {
      int da = 6 * m.jnt_dofadr[j];
      V3 off = root - ld3(xanchor, 3 * j);
      const int jt = m.jnt_type[j];
      if (jt == JNT_FREE || jt == JNT_BALL) {
        if (jt == JNT_FREE) {
          for (int k = 0; k < 18; k++) cdof[da + k] = 0;
          cdof[da + 3] = 1; cdof[da + 10] = 1; cdof[da + 17] = 1;
          da += 18;
        }
        for (int k = 0; k < 3; k++) {
          V3 ax{xmat[9 * i + k], xmat[9 * i + k + 3], xmat[9 * i + k + 6]};
          st3(cdof, da + 6 * k, ax);
          st3(cdof, da + 6 * k + 3, cross(ax, off));
        }
      } else if (jt == JNT_SLIDE) {
        st3(cdof, da, V3{0, 0, 0});
        st3(cdof, da + 3, ld3(xaxis, 3 * j));
      } else {
        V3 ax = ld3(xaxis, 3 * j);
        st3(cdof, da, ax);
        st3(cdof, da + 3, cross(ax, off));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fixed tendons: length and constant-pattern sparse Jacobian
MJB_HD void tendon(const Env& d) {
  const DModel& m = d.m;
  const int nt = m.sz.ntendon;
  if (!nt) return;
  FD L = d.ten_length(), J = d.ten_J(), qpos = d.qpos();
  for (int i = 0; i < nt; i++) L[i] = 0;
  for (int i = 0; i < m.sz.nJten; i++) J[i] = 0;
  for (int i = 0; i < nt; i++) {
    const int adr = m.tendon_adr[i], num = m.tendon_num[i];
    const int radr = m.ten_J_rowadr[i], rnnz = m.ten_J_rownnz[i];
    for (int j = 0; j < num; j++) {
      const int k = m.wrap_objid[adr + j];
      const double c = m.wrap_prm[adr + j];
      L[i] += c * qpos[m.jnt_qposadr[k]];
      // J(row i, col dofadr) += c * 1     (mju_combineSparseInc with a single source entry)
      const int dof = m.jnt_dofadr[k];
      for (int a = 0; a < rnnz; a++) {
        if (m.ten_J_colind[radr + a] == dof) { J[radr + a] = 1.0 * J[radr + a] + c * 1.0; break; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// joint transmissions (hinge / slide): actuator_length and the single moment entry per actuator
MJB_HD void transmission(const Env& d) {
  const DModel& m = d.m;
  FD len = d.actuator_length(), mom = d.actuator_moment(), qpos = d.qpos();
  for (int i = 0; i < m.sz.nu; i++) {
    const int j = m.actuator_trnjnt[i];
    const double g = m.actuator_gear0[i];
    len[i] = qpos[m.jnt_qposadr[j]] * g;
    mom[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------
// composite rigid body algorithm -> tree-sparse M, plus tendon armature
MJB_HD void make_M(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody, nv = m.sz.nv;
  FD crb = d.crb(), cinert = d.cinert(), cdof = d.cdof(), M = d.M();
  for (int i = 0; i < 10 * nbody; i++) crb[i] = cinert[i];
  for (int i = nbody - 1; i > 0; i--) {
    const int p = m.body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) crb[10 * p + k] += crb[10 * i + k];
  }
  for (int i = 0; i < m.sz.nC; i++) M[i] = 0;
  for (int i = 0; i < nv; i++) {
    const int adr = m.M_rowadr[i];
    if (m.dof_simplenum[i]) { M[adr] = m.dof_M0[i]; continue; }
    int a = adr + m.M_rownnz[i] - 1;
    M[a] = m.dof_armature_eff[i];
    S6 buf = mul_inert(ld10(crb, 10 * m.dof_bodyid[i]), ld6(cdof, 6 * i));
    for (int j = i; j >= 0; j = m.dof_parentid[j]) {
      M[a] += dot6(ld6(cdof, 6 * j), buf);
      a--;
    }
  }
  // tendon armature: M += armature * J' J  over the tendon's sparsity pattern
  FD tJ = d.ten_J();
  for (int k = 0; k < m.sz.ntendon; k++) {
    const double arm = m.tendon_armature_eff[k];
    if (!arm) continue;
    const int jadr = m.ten_J_rowadr[k], jnnz = m.ten_J_rownnz[k];
    for (int j = 0; j < jnnz; j++) {
      const double Ji = tJ[jadr + j];
      if (!Ji) continue;
      const int i = m.ten_J_colind[jadr + j];
      const int madr = m.M_rowadr[i], mnnz = m.M_rownnz[i];
      const double scl = arm * Ji;
      // walk both sorted index lists (mju_addToSclSparseInc)
      int a = 0, b = 0;
      while (a < mnnz && b < jnnz) {
        const int ca = m.M_colind[madr + a], cb = m.ten_J_colind[jadr + b];
        if (ca == cb) { M[madr + a] += scl * tJ[jadr + b]; a++; b++; }
        else if (ca < cb) a++;
        else b++;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// in-place sparse L'DL of a tree-sparse matrix in M's CSR pattern
MJB_HD void factor_I(const DModel& m, FD mat, FD diaginv) {
  const int nv = m.sz.nv;
  for (int k = nv - 1; k >= 0; k--) {
    const int start = m.M_rowadr[k];
    const int diag = m.M_rownnz[k] - 1;
    const int end = start + diag;
    const double invD = 1 / mat[end];
    diaginv[k] = invD;
    for (int adr = end - 1; adr >= start; adr--) {
      const int i = m.M_colind[adr];
      const double s = -mat[adr] * invD;
      const int ri = m.M_rowadr[i], ni = m.M_rownnz[i];
      for (int c = 0; c < ni; c++) mat[ri + c] += mat[start + c] * s;
    }
    for (int c = 0; c < diag; c++) mat[start + c] = mat[start + c] * invD;
  }
}

// in-place x <- (L'DL)^-1 x for one right-hand side
MJB_HD void solve_LD(const DModel& m, FD x, FD qLD, FD qLDiagInv) {
  const int nv = m.sz.nv;
  for (int i = nv - 1; i >= 0; i--) {
    const int nnz = m.M_rownnz[i];
    if (nnz == 1) continue;
    const double xi = x[i];
    if (xi != 0) {
      const int start = m.M_rowadr[i], end = start + nnz - 1;
      for (int adr = start; adr < end; adr++) x[m.M_colind[adr]] -= qLD[adr] * xi;
    }
  }
  for (int i = 0; i < nv; i++) x[i] *= qLDiagInv[i];
  for (int i = 0; i < nv; i++) {
    const int nnz = m.M_rownnz[i];
    if (nnz == 1) continue;
    const int adr = m.M_rowadr[i], dn = nnz - 1;
    x[i] -= dot_sparse_ref(dn, [&](int c) { return qLD[adr + c]; },
                           [&](int c) { return x[m.M_colind[adr + c]]; });
  }
}

// ------------------------------------------------------------------------------------------------
// com-frame body velocities and time derivatives of the motion axes
MJB_HD void com_vel(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody;
  FD cvel = d.cvel(), cdof = d.cdof(), cdd = d.cdof_dot(), qvel = d.qvel();
  for (int k = 0; k < 6; k++) cvel[k] = 0;
  for (int i = 1; i < nbody; i++) {
    S6 cv = ld6(cvel, 6 * m.body_parentid[i]);
    const int dn = m.body_dofnum[i], bda = m.body_dofadr[i];
    for (int j = 0; j < dn; j++) {
      const int jt = m.jnt_type[m.dof_jntid[bda + j]];
      if (jt == JNT_FREE || jt == JNT_BALL) {
        if (jt == JNT_FREE) {
          for (int k = 0; k < 18; k++) cdd[6 * bda + k] = 0;
          // cvel += cdof[0..2]' * qvel[0..2]   (mju_mulMatTVec: skip exact zeros, row by row)
          S6 t; for (int k = 0; k < 6; k++) t.v[k] = 0;
          for (int r = 0; r < 3; r++) {
            const double s = qvel[bda + r];
            if (s != 0) for (int k = 0; k < 6; k++) t.v[k] += cdof[6 * (bda + r) + k] * s;
          }
          for (int k = 0; k < 6; k++) cv.v[k] += t.v[k];
          j += 3;
        }
        for (int r = 0; r < 3; r++) st6(cdd, 6 * (bda + j + r), cross_motion(cv, ld6(cdof, 6 * (bda + j + r))));
        S6 t; for (int k = 0; k < 6; k++) t.v[k] = 0;
        for (int r = 0; r < 3; r++) {
          const double s = qvel[bda + j + r];
          if (s != 0) for (int k = 0; k < 6; k++) t.v[k] += cdof[6 * (bda + j + r) + k] * s;
        }
        for (int k = 0; k < 6; k++) cv.v[k] += t.v[k];
        j += 2;
      } else {
        S6 cd = ld6(cdof, 6 * (bda + j));
        st6(cdd, 6 * (bda + j), cross_motion(cv, cd));
        const double s = qvel[bda + j];
        for (int k = 0; k < 6; k++) cv.v[k] += cd.v[k] * s;
      }
    }
    st6(cvel, 6 * i, cv);
  }
}

// res = sum_r dof[r] * vec[r] over n dofs (mju_mulDofVec, engine_util_spatial.c:466-474)
MJB_HD S6 mul_dof_vec(FD dof, FD vec, int n) {
  S6 r;
  if (n == 1) {
    for (int k = 0; k < 6; k++) r.v[k] = dof[k] * vec[0];
  } else {
    for (int k = 0; k < 6; k++) r.v[k] = 0;
    for (int a = 0; a < n; a++) {
      const double s = vec[a];
      if (s != 0) for (int k = 0; k < 6; k++) r.v[k] += dof[6 * a + k] * s;
    }
  }
  return r;
}

// bias forces by recursive Newton-Euler without the acceleration term (flg_acc = 0)
MJB_HD void rne_bias(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody, nv = m.sz.nv;
  FD cacc = d.scr_body(), cfrc = d.scr_body() + 6 * nbody;
  FD cinert = d.cinert(), cvel = d.cvel(), cdof = d.cdof(), cdd = d.cdof_dot(), qvel = d.qvel();
  for (int k = 0; k < 6; k++) cacc[k] = 0;
  if (!(m.opt.disableflags & DSBL_GRAVITY)) {
    cacc[3] = m.opt.gravity[0] * -1; cacc[4] = m.opt.gravity[1] * -1; cacc[5] = m.opt.gravity[2] * -1;
  }
  for (int i = 1; i < nbody; i++) {
    const int bda = m.body_dofadr[i];
    S6 t = mul_dof_vec(cdd + 6 * bda, qvel + bda, m.body_dofnum[i]);
    S6 a = ld6(cacc, 6 * m.body_parentid[i]);
    for (int k = 0; k < 6; k++) a.v[k] = a.v[k] + t.v[k];
    st6(cacc, 6 * i, a);
    I10 I = ld10(cinert, 10 * i);
    S6 f = mul_inert(I, a);
    S6 v = ld6(cvel, 6 * i);
    S6 Iv = mul_inert(I, v);
    S6 c = cross_force(v, Iv);
    for (int k = 0; k < 6; k++) f.v[k] += c.v[k];
    st6(cfrc, 6 * i, f);
  }
  for (int k = 0; k < 6; k++) cfrc[k] = 0;
  for (int i = nbody - 1; i > 0; i--) {
    const int p = m.body_parentid[i];
    if (p) for (int k = 0; k < 6; k++) cfrc[6 * p + k] += cfrc[6 * i + k];
  }
  FD out = d.qfrc_bias();
  for (int i = 0; i < nv; i++) out[i] = dot6(ld6(cdof, 6 * i), ld6(cfrc, 6 * m.dof_bodyid[i]));
}

// ------------------------------------------------------------------------------------------------
// passive forces: joint springs, dof dampers, tendon spring-dampers
MJB_HD void passive(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  FD fs = d.qfrc_spring(), fd = d.qfrc_damper(), fp = d.qfrc_passive(), qpos = d.qpos(), qvel = d.qvel();
  for (int i = 0; i < nv; i++) { fs[i] = 0; fd[i] = 0; fp[i] = 0; }
  const bool spring = !(m.opt.disableflags & DSBL_SPRING), damper = !(m.opt.disableflags & DSBL_DAMPER);
  if (!spring && !damper) return;
  if (spring) {
    for (int j = 0; j < m.sz.njnt; j++) {
      const double k0 = m.jnt_stiffness[j];
      const double* sp = m.jnt_stiffnesspoly + kNPoly * j;
      if (k0 == 0 && sp[0] == 0 && sp[1] == 0) continue;
      int pa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
      const int jt = m.jnt_type[j];
      if (jt == JNT_FREE || jt == JNT_BALL) {
        if (jt == JNT_FREE) {
          V3 dif = ld3(qpos, pa) - ldc3(m.qpos_spring, pa);
          double r = sqrt(dot(dif, dif));
          double k = poly_force(k0, sp, kNPoly, r, false);
          st3(fs, da, addscl(ld3(fs, da), dif, -k));
          da += 3; pa += 3;
        }
        Q4 q = ld4(qpos, pa);
        normalize(q);
        V3 dif = qsub(q, ldc4(m.qpos_spring, pa));
        double r = sqrt(dot(dif, dif));
        double k = poly_force(k0, sp, kNPoly, r, false);
        st3(fs, da, addscl(ld3(fs, da), dif, -k));
      } else {
        const double x = qpos[pa] - m.qpos_spring[pa];
        fs[da] = -x * poly_force(k0, sp, kNPoly, x, false);
      }
    }
  }
  if (damper) {
    for (int i = 0; i < nv; i++) {
      const double b = m.dof_damping_eff[i];
      const double* bp = m.dof_dampingpoly_eff + kNPoly * i;
      if (b != 0 || bp[0] != 0 || bp[1] != 0) {
        const double v = qvel[i];
        fd[i] = -v * poly_force(b, bp, kNPoly, v, true);
      }
    }
  }
  FD tl = d.ten_length(), tv = d.ten_velocity(), tJ = d.ten_J();
  for (int i = 0; i < m.sz.ntendon; i++) {
    double k0 = 0, b0 = 0;
    const double* sp = m.tendon_stiffnesspoly + kNPoly * i;
    double dp[kNPoly] = {0, 0};
    if (spring) k0 = m.tendon_stiffness[i];
    if (damper) { b0 = m.tendon_damping_eff[i]; dp[0] = m.tendon_dampingpoly_eff[kNPoly * i]; dp[1] = m.tendon_dampingpoly_eff[kNPoly * i + 1]; }
    if (k0 == 0 && (!spring || (sp[0] == 0 && sp[1] == 0)) && b0 == 0 && dp[0] == 0 && dp[1] == 0) continue;
    const double len = tl[i], lo = m.tendon_lengthspring[2 * i], hi = m.tendon_lengthspring[2 * i + 1];
    const double x = (len > hi) ? len - hi : (len < lo) ? len - lo : 0;
    const double f_s = spring ? -x * poly_force(k0, sp, kNPoly, x, false) : 0;
    const double v = tv[i];
    const double f_d = damper ? -v * poly_force(b0, dp, kNPoly, v, true) : 0;
    if (f_s || f_d) {
      const int adr = m.ten_J_rowadr[i], end = adr + m.ten_J_rownnz[i];
      for (int j = adr; j < end; j++) {
        const int k = m.ten_J_colind[j];
        const double Jv = tJ[j];
        fs[k] += Jv * f_s;
        fd[k] += Jv * f_d;
      }
    }
  }
  for (int i = 0; i < nv; i++) fp[i] = fs[i] + fd[i];
}

}  // namespace mjb
