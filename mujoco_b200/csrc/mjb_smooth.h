// Smooth dynamics of the batched mj_step path for ONE environment, executed cooperatively by the
// lanes that own it (32 lanes of a warp in the fused kernel, 1 lane in lane-per-env mode).
//
// Replaces (reference file:line)  src/engine/engine_core_smooth.c  mj_kinematics :40-242,
// mj_comPos :246-350, mj_tendon (fixed tendons) :927-985, mj_transmission (joint) :1265-1330,
// mj_crb/mj_tendonArmature :1845-1971, mj_factorI :1997-2029, mj_solveLD :2033-2117,
// mj_comVel :2179-2245, mj_rne :2328-2390;  src/engine/engine_passive.c mj_springdamper :655-842.
//
// Parallel structure: the tree recursions run level by level (MJB_PFOR over the bodies / dofs of a
// level, MJB_PSYNC between levels).  Every output element is still produced by ONE lane with the
// reference's operation order — backward accumulations use host-built child / descendant lists in
// the order in which the serial loops of the reference would have added them — so results are
// bit-identical to the serial restatement whatever the lane count.
#pragma once
#include "mjb_types.h"

namespace mjb {

// ------------------------------------------------------------------------------------------------
// forward kinematics: body frames level by level, then inertial and geom frames
MJB_HD void kinematics(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody;
  FD qpos = d.qpos(), xpos = d.xpos(), xquat = d.xquat(), xmat = d.xmat();
  FD xipos = d.xipos(), ximat = d.ximat(), xanchor = d.xanchor(), xaxis = d.xaxis();

  MJB_LANE0 {   // world body
    st3(xpos, 0, V3{0, 0, 0});
    st4(xquat, 0, Q4{1, 0, 0, 0});
    st3(xipos, 0, V3{0, 0, 0});
    for (int k = 0; k < 9; k++) { xmat[k] = (k % 4 == 0) ? 1.0 : 0.0; ximat[k] = (k % 4 == 0) ? 1.0 : 0.0; }
  }
  MJB_PSYNC();

  for (int l = 1; l < m.sz.nlevel; l++) {
    const int adr = m.lvl_adr[l], cnt = m.lvl_adr[l + 1] - adr;
    MJB_PFOR(k_, cnt) {
      const int i = m.lvl_body[adr + k_];
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
        if (m.sz.nmocap && m.body_mocapid[i] >= 0) {   // mocap body: pose from mjData (engine_core_smooth.c:84-90)
          const int mid = m.body_mocapid[i];
          bpos = ld3(d.mocap_pos(), 3 * mid);
          bquat = ld4(d.mocap_quat(), 4 * mid);
          normalize(bquat);
        }
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
    MJB_PSYNC();
  }

  // inertial frames (mj_local2Global, engine_core_util.c) and geom frames: independent items
  FD gpos = d.geom_xpos(), gmat = d.geom_xmat();
  const int ngeom = m.sz.ngeom;
  MJB_PFOR(i_, nbody - 1) {
    const int i = i_ + 1;
    const int sf = m.body_sameframe[i];
    V3 bp = ld3(xpos, 3 * i);
    M3 bm = ld9(xmat, 9 * i);
    if (sf == SAMEFRAME_BODY) st3(xipos, 3 * i, bp);
    else st3(xipos, 3 * i, mulmv(bm, ldc3(m.body_ipos, 3 * i)) + bp);
    if (sf == SAMEFRAME_NONE) st9(ximat, 9 * i, quat2mat(qmul(ld4(xquat, 4 * i), ldc4(m.body_iquat, 4 * i))));
    else st9(ximat, 9 * i, bm);
  }
  MJB_PSYNC();
  MJB_PFOR(g, ngeom) {
    const int b = m.geom_bodyid[g], sf = m.geom_sameframe[g];
    V3 bp = ld3(xpos, 3 * b);
    if (sf == SAMEFRAME_BODY) st3(gpos, 3 * g, bp);
    else if (sf == SAMEFRAME_INERTIA) st3(gpos, 3 * g, ld3(xipos, 3 * b));
    else st3(gpos, 3 * g, mulmv(ld9(xmat, 9 * b), ldc3(m.geom_pos, 3 * g)) + bp);
    if (sf == SAMEFRAME_NONE) st9(gmat, 9 * g, quat2mat(qmul(ld4(xquat, 4 * b), ldc4(m.geom_quat, 4 * g))));
    else if (sf == SAMEFRAME_BODY || sf == SAMEFRAME_BODYROT) st9(gmat, 9 * g, ld9(xmat, 9 * b));
    else st9(gmat, 9 * g, ld9(ximat, 9 * b));
  }
  // sites: same local-to-global rule (mj_kinematics, engine_core_smooth.c:222-233)
  FD spos = d.site_xpos(), smat = d.site_xmat();
  MJB_PFOR(s_, m.sz.nsite) {
    const int b = m.site_bodyid[s_], sf = m.site_sameframe[s_];
    V3 bp = ld3(xpos, 3 * b);
    if (sf == SAMEFRAME_BODY) st3(spos, 3 * s_, bp);
    else if (sf == SAMEFRAME_INERTIA) st3(spos, 3 * s_, ld3(xipos, 3 * b));
    else st3(spos, 3 * s_, mulmv(ld9(xmat, 9 * b), ldc3(m.site_pos, 3 * s_)) + bp);
    if (sf == SAMEFRAME_NONE) st9(smat, 9 * s_, quat2mat(qmul(ld4(xquat, 4 * b), ldc4(m.site_quat, 4 * s_))));
    else if (sf == SAMEFRAME_BODY || sf == SAMEFRAME_BODYROT) st9(smat, 9 * s_, ld9(xmat, 9 * b));
    else st9(smat, 9 * s_, ld9(ximat, 9 * b));
  }
  MJB_PSYNC();
}

// backward accumulation over the body tree: acc[parent] += acc[child] for `width` components,
// children added in descending id order, deepest level first (== serial loop i = nbody-1..1).
// skip_world: do not accumulate into the world body (mj_crb) / do (mj_comPos, mj_rne)
MJB_HD void tree_accumulate(const Env& d, FD acc, int width, bool skip_world) {
  const DModel& m = d.m;
  for (int l = m.sz.nlevel - 2; l >= (skip_world ? 1 : 0); l--) {
    const int adr = m.lvl_adr[l], cnt = m.lvl_adr[l + 1] - adr;
    MJB_PFOR(it, cnt * width) {
      const int p = m.lvl_body[adr + it / width], k = it % width;
      const int ca = m.child_adr[p], cn = m.child_adr[p + 1] - ca;
      if (cn) {
        double s = acc[(long)width * p + k];
        for (int c = 0; c < cn; c++) s += acc[(long)width * m.child_id[ca + c] + k];
        acc[(long)width * p + k] = s;
      }
    }
    MJB_PSYNC();
  }
}

// ------------------------------------------------------------------------------------------------
// subtree centres of mass, com-frame inertias (cinert) and motion axes (cdof)
MJB_HD void com_pos(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody;
  FD sc = d.subtree_com(), xipos = d.xipos(), ximat = d.ximat(), xmat = d.xmat();
  FD cinert = d.cinert(), cdof = d.cdof(), xanchor = d.xanchor(), xaxis = d.xaxis();

  MJB_PFOR(i, nbody) st3(sc, 3 * i, ld3(xipos, 3 * i) * m.body_mass[i]);
  MJB_PSYNC();
  tree_accumulate(d, sc, 3, false);
  MJB_PFOR(i, nbody) {
    if (m.body_subtreemass[i] < kMinVal) st3(sc, 3 * i, ld3(xipos, 3 * i));
    else st3(sc, 3 * i, ld3(sc, 3 * i) * (1.0 / m.body_subtreemass[i]));
  }
  MJB_PSYNC();

  MJB_PFOR(i, nbody) {
    if (i == 0) { for (int k = 0; k < 10; k++) cinert[k] = 0; }
    else {
      V3 off = ld3(xipos, 3 * i) - ld3(sc, 3 * m.body_rootid[i]);
      st10(cinert, 10 * i, inert_com(ldc3(m.body_inertia, 3 * i), ld9(ximat, 9 * i), off, m.body_mass[i]));
    }
  }
  MJB_PFOR(j, m.sz.njnt) {
    const int i = m.jnt_bodyid[j];
    V3 root = ld3(sc, 3 * m.body_rootid[i]);
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
  MJB_PSYNC();
}

// ------------------------------------------------------------------------------------------------
// fixed tendons: length and constant-pattern sparse Jacobian
MJB_HD void tendon(const Env& d) {
  const DModel& m = d.m;
  const int nt = m.sz.ntendon;
  if (!nt) return;
  FD L = d.ten_length(), J = d.ten_J(), qpos = d.qpos();
  MJB_PFOR(i, nt) {
    const int adr = m.tendon_adr[i], num = m.tendon_num[i];
    const int radr = m.ten_J_rowadr[i], rnnz = m.ten_J_rownnz[i];
    double len = 0;
    for (int a = 0; a < rnnz; a++) J[radr + a] = 0;
    for (int j = 0; j < num; j++) {
      const int k = m.wrap_objid[adr + j];
      const double c = m.wrap_prm[adr + j];
      len += c * qpos[m.jnt_qposadr[k]];
      // J(row i, col dofadr) = 1*J + c*1   (mju_combineSparseInc with a single source entry)
      const int dof = m.jnt_dofadr[k];
      for (int a = 0; a < rnnz; a++) {
        if (m.ten_J_colind[radr + a] == dof) { J[radr + a] = 1.0 * J[radr + a] + c * 1.0; break; }
      }
    }
    L[i] = len;
  }
  MJB_PSYNC();
}

// ------------------------------------------------------------------------------------------------
// transmissions (mj_transmission, engine_core_smooth.c:1380-1480): scalar joints — actuator_length and the
// single moment entry; fixed tendons — length = ten_length * gear, moment row = ten_J row * gear (the row
// values are re-formed as ten_J * gear where they are used, actuator_moment keeps the gear)
MJB_HD void transmission(const Env& d) {
  const DModel& m = d.m;
  FD len = d.actuator_length(), mom = d.actuator_moment(), qpos = d.qpos();
  MJB_PFOR(i, m.sz.nu) {
    const int j = m.actuator_trnjnt[i];
    const double g = m.actuator_gear0[i];
    const int tt = (d.feat & FEAT_ACT) ? m.actuator_trntype[i] : TRN_JOINT;
    if (tt == TRN_TENDON) len[i] = d.ten_length()[j] * g;
    else if (tt == TRN_BALL || tt == TRN_FREE) {   // 3D / 6D gear (engine_core_smooth.c:1331-1393)
      const double* gear = m.actuator_gear6 + 6 * i;
      FD m6 = d.actuator_mom6();
      Q4 q = ld4(qpos, m.jnt_qposadr[j] + (tt == TRN_FREE ? 3 : 0));
      normalize(q);
      V3 ga{gear[tt == TRN_FREE ? 3 : 0], gear[tt == TRN_FREE ? 4 : 1], gear[tt == TRN_FREE ? 5 : 2]};
      if (m.actuator_inparent[i]) ga = rotate(ga, Q4{q.w, -q.x, -q.y, -q.z});
      if (tt == TRN_BALL) {
        const V3 axis = quat2vel(q, 1);
        len[i] = axis.x * ga.x + axis.y * ga.y + axis.z * ga.z;
        m6[6 * i] = ga.x; m6[6 * i + 1] = ga.y; m6[6 * i + 2] = ga.z;
      } else {
        len[i] = 0;
        m6[6 * i] = gear[0]; m6[6 * i + 1] = gear[1]; m6[6 * i + 2] = gear[2];
        m6[6 * i + 3] = ga.x; m6[6 * i + 4] = ga.y; m6[6 * i + 5] = ga.z;
      }
    }
    else if (tt >= TRN_SITE) len[i] = 0;   // length and moment row: site_moment(), after the site poses exist
    else len[i] = qpos[m.jnt_qposadr[j]] * g;
    mom[i] = g;
  }
  MJB_PSYNC();
}

// ------------------------------------------------------------------------------------------------
// composite rigid body algorithm -> tree-sparse M, plus tendon armature
MJB_HD void make_M(const Env& d) {
  const DModel& m = d.m;
  const int nbody = m.sz.nbody, nv = m.sz.nv;
  FD crb = d.crb(), cinert = d.cinert(), cdof = d.cdof(), M = d.M();
  MJB_PFOR(i, 10 * nbody) crb[i] = cinert[i];
  MJB_PFOR(i, m.sz.nC) M[i] = 0;
  MJB_PSYNC();
  tree_accumulate(d, crb, 10, true);
  MJB_PFOR(i, nv) {
    const int adr = m.M_rowadr[i];
    if (m.dof_simplenum[i]) { M[adr] = m.dof_M0[i]; }
    else {
      int a = adr + m.M_rownnz[i] - 1;
      double acc = m.dof_armature_eff[i];
      S6 buf = mul_inert(ld10(crb, 10 * m.dof_bodyid[i]), ld6(cdof, 6 * i));
      for (int j = i; j >= 0; j = m.dof_parentid[j]) {
        M[a] = acc + dot6(ld6(cdof, 6 * j), buf);
        acc = 0;
        a--;
      }
    }
  }
  MJB_PSYNC();
  // tendon armature: M += armature * J' J  over the tendon's sparsity pattern (rare; serial)
  bool any = false;
  for (int k = 0; k < m.sz.ntendon; k++) any = any || (m.tendon_armature_eff[k] != 0);
  if (any) {
    MJB_LANE0 {
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
          int a = 0, b = 0;   // walk both sorted index lists (mju_addToSclSparseInc)
          while (a < mnnz && b < jnnz) {
            const int ca = m.M_colind[madr + a], cb = m.ten_J_colind[jadr + b];
            if (ca == cb) { M[madr + a] += scl * tJ[jadr + b]; a++; b++; }
            else if (ca < cb) a++;
            else b++;
          }
        }
      }
    }
    MJB_PSYNC();
  }
}

// ------------------------------------------------------------------------------------------------
// in-place sparse L'DL of a tree-sparse matrix in M's CSR pattern.  Pivot rows are processed
// serially (nv-1..0) as in the reference; the element updates of one pivot are independent and
// come from the host-built program (fac_*).
MJB_HD void factor_I(const Env& d, FD mat, FD diaginv) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  for (int k = nv - 1; k >= 0; k--) {
    const int start = m.M_rowadr[k];
    const int diag = m.M_rownnz[k] - 1;
    const double invD = 1 / mat[start + diag];
    MJB_LANE0 diaginv[k] = invD;
    if (diag == 0) continue;
    const int pa = m.fac_adr[k], pn = m.fac_adr[k + 1] - pa;
    MJB_PFOR(t, pn) {
      const int dst = m.fac_dst[pa + t];
      mat[dst] += mat[m.fac_src[pa + t]] * (-mat[m.fac_cf[pa + t]] * invD);
    }
    MJB_PSYNC();
    MJB_PFOR(c, diag) mat[start + c] = mat[start + c] * invD;
    MJB_PSYNC();
  }
  MJB_PSYNC();
}

// in-place x <- (L'DL)^-1 x for one right-hand side, by dof-tree levels
MJB_HD void solve_LD(const Env& d, FD x, FD qLD, FD qLDiagInv) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  // x <- L^-T x : deepest dofs are final; each dof then gathers its descendants' contributions in
  // descending descendant order (the order of the serial scatter loop)
  for (int l = m.sz.ndlevel - 2; l >= 0; l--) {
    const int adr = m.dlvl_adr[l], cnt = m.dlvl_adr[l + 1] - adr;
    MJB_PFOR(t, cnt) {
      const int j = m.dlvl_dof[adr + t];
      const int a0 = m.mt_adr[j], an = m.mt_adr[j + 1] - a0;
      if (an) {
        double s = x[j];
        for (int c = 0; c < an; c++) {
          const double xi = x[m.mt_dof[a0 + c]];
          if (xi != 0) s -= qLD[m.mt_qadr[a0 + c]] * xi;
        }
        x[j] = s;
      }
    }
    MJB_PSYNC();
  }
  MJB_PFOR(i, nv) x[i] *= qLDiagInv[i];
  MJB_PSYNC();
  // x <- L^-1 x : shallow dofs are final first
  for (int l = 1; l < m.sz.ndlevel; l++) {
    const int adr = m.dlvl_adr[l], cnt = m.dlvl_adr[l + 1] - adr;
    MJB_PFOR(t, cnt) {
      const int i = m.dlvl_dof[adr + t];
      const int ra = m.M_rowadr[i], dn = m.M_rownnz[i] - 1;
      x[i] -= dot_sparse_ref(dn, [&](int c) { return qLD[ra + c]; },
                             [&](int c) { return x[m.M_colind[ra + c]]; });
    }
    MJB_PSYNC();
  }
}

// ------------------------------------------------------------------------------------------------
// com-frame body velocities and time derivatives of the motion axes, level by level
MJB_HD void com_vel(const Env& d) {
  const DModel& m = d.m;
  FD cvel = d.cvel(), cdof = d.cdof(), cdd = d.cdof_dot(), qvel = d.qvel();
  MJB_LANE0 { for (int k = 0; k < 6; k++) cvel[k] = 0; }
  MJB_PSYNC();
  for (int l = 1; l < m.sz.nlevel; l++) {
    const int ladr = m.lvl_adr[l], cnt = m.lvl_adr[l + 1] - ladr;
    MJB_PFOR(k_, cnt) {
      const int i = m.lvl_body[ladr + k_];
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
    MJB_PSYNC();
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
  MJB_LANE0 {
    for (int k = 0; k < 6; k++) { cacc[k] = 0; cfrc[k] = 0; }
    if (!(m.opt.disableflags & DSBL_GRAVITY)) {
      cacc[3] = m.opt.gravity[0] * -1; cacc[4] = m.opt.gravity[1] * -1; cacc[5] = m.opt.gravity[2] * -1;
    }
  }
  MJB_PSYNC();
  for (int l = 1; l < m.sz.nlevel; l++) {
    const int ladr = m.lvl_adr[l], cnt = m.lvl_adr[l + 1] - ladr;
    MJB_PFOR(k_, cnt) {
      const int i = m.lvl_body[ladr + k_];
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
    MJB_PSYNC();
  }
  // backward accumulation; the reference zeroes the world force first and never adds into it
  tree_accumulate(d, cfrc, 6, true);
  FD out = d.qfrc_bias();
  MJB_PFOR(i, nv) out[i] = dot6(ld6(cdof, 6 * i), ld6(cfrc, 6 * m.dof_bodyid[i]));
  MJB_PSYNC();
}

// ------------------------------------------------------------------------------------------------
// passive forces: joint springs, dof dampers, tendon spring-dampers
MJB_HD void passive(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  FD fs = d.qfrc_spring(), fd = d.qfrc_damper(), fp = d.qfrc_passive(), qpos = d.qpos(), qvel = d.qvel();
  MJB_PFOR(i, nv) { fs[i] = 0; fd[i] = 0; fp[i] = 0; }
  MJB_PSYNC();
  const bool spring = !(m.opt.disableflags & DSBL_SPRING), damper = !(m.opt.disableflags & DSBL_DAMPER);
  if (!spring && !damper) return;
  if (spring) {
    MJB_PFOR(j, m.sz.njnt) {
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
    MJB_PFOR(i, nv) {
      const double b = m.dof_damping_eff[i];
      const double* bp = m.dof_dampingpoly_eff + kNPoly * i;
      if (b != 0 || bp[0] != 0 || bp[1] != 0) {
        const double v = qvel[i];
        fd[i] = -v * poly_force(b, bp, kNPoly, v, true);
      }
    }
  }
  MJB_PSYNC();
  if (m.sz.ntendon) {
    MJB_LANE0 {   // tendons may share dofs: keep the reference's serial accumulation order
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
    }
    MJB_PSYNC();
  }
  MJB_PFOR(i, nv) fp[i] = fs[i] + fd[i];
  MJB_PSYNC();
}

}  // namespace mjb
