// Constraint assembly and the PGS solver of the batched mj_step path for ONE environment, executed
// cooperatively by the lanes that own it.
//
// Replaces (reference file:line)
//   src/engine/engine_core_constraint.c  mj_instantiateFriction :1270-1355, mj_instantiateLimit
//   :1360-1531, mj_instantiateContact :1617-1713 (dense rows, pyramidal cones), mj_diagApprox
//   :1719-1973, mj_makeImpedance :2151-2265, mj_makeY/mj_makeAR (dense) :2918-3102,
//   mj_referenceConstraint :3245-3268, mj_constraintUpdate_impl :3275-3464;
//   src/engine/engine_core_util.c mj_jac :176-226;
//   src/engine/engine_forward.c warmstart :1056-1132, mj_fwdConstraint :1148-1252;
//   src/engine/engine_solver.c solPGS :457-741, dualState :270-353, dualFinish :72-79.
// Row order (friction loss, joint limits lower/upper, tendon limits, contacts; pyramid edge pairs
// +mu,-mu) follows the reference.  Rows and matrix elements are independent work items spread over
// the lanes; each element is computed by one lane with the reference's operation order.
#pragma once
#include "mjb_smooth.h"
#include "mjb_prof.h"

namespace mjb {

MJB_HD double ipow(double a, double b) {   // power() of engine_core_constraint.c:2083-2090
  if (b == 1) return a;
  if (b == 2) return a * a;
  return pow(a, b);
}

// impedance and its derivative for one row (getimpedance, engine_core_constraint.c:2094-2146)
MJB_HD void impedance(const double* solimp, double pos, double margin, double& imp, double& impP) {
  if (solimp[0] == solimp[1] || solimp[2] <= kMinVal) { imp = 0.5 * (solimp[0] + solimp[1]); impP = 0; return; }
  double x = (pos - margin) / solimp[2], sgn = 1;
  if (x < 0) { x = -x; sgn = -1; }
  if (x >= 1 || x <= 0) { imp = (x >= 1 ? solimp[1] : solimp[0]); impP = 0; return; }
  double y, yP;
  if (solimp[4] == 1) { y = x; yP = 1; }
  else if (x <= solimp[3]) {
    const double a = 1 / ipow(solimp[3], solimp[4] - 1);
    y = a * ipow(x, solimp[4]);
    yP = solimp[4] * a * ipow(x, solimp[4] - 1);
  } else {
    const double b = 1 / ipow(1 - solimp[3], solimp[4] - 1);
    y = 1 - b * ipow(1 - x, solimp[4]);
    yP = solimp[4] * b * ipow(1 - x, solimp[4] - 1);
  }
  imp = solimp[0] + y * (solimp[1] - solimp[0]);
  impP = yP * sgn * (solimp[1] - solimp[0]) / solimp[2];
}

// one element of the translational point Jacobian of `body` at world point pt: component k, dof c
// (mj_jac: cdof_lin + cross(cdof_ang, pt - subtree_com[root]); zero off the body's dof chain)
MJB_HD double jac_elem(const Env& d, V3 pt, int body, int k, int c) {
  const DModel& m = d.m;
  if (!m.body_dofanc[(long)body * m.sz.nv + c]) return 0;
  FD cdof = d.cdof();
  V3 off = pt - ld3(d.subtree_com(), 3 * m.body_rootid[body]);
  V3 w = ld3(cdof, 6 * c);
  V3 t = cross(w, off);
  return cdof[6 * c + 3 + k] + get(t, k);
}

// global anchor points of a body-semantic connect (mj_equalityAnchors, engine_core_constraint.c:561-590)
MJB_HD void connect_anchors(const Env& d, int eq, V3& p0, V3& p1) {
  const DModel& m = d.m;
  const int o1 = m.eq_obj1id[eq], o2 = m.eq_obj2id[eq];
  const double* data = m.eq_data + kNEqData * eq;
  const int a0 = (m.eq_kind[eq] == EQ_WELD) ? 3 : 0, a1 = 3 - a0;   // weld keeps its anchors the other way round
  p0 = mulmv(ld9(d.xmat(), 9 * o1), V3{data[a0], data[a0 + 1], data[a0 + 2]}) + ld3(d.xpos(), 3 * o1);
  p1 = mulmv(ld9(d.xmat(), 9 * o2), V3{data[a1], data[a1 + 1], data[a1 + 2]}) + ld3(d.xpos(), 3 * o2);
}

// one element of the time derivative of the translational point Jacobian (mj_jacDot, engine_core_util.c
// :605-675): component k, dof c; zero off the body's dof chain
MJB_HD double jacdot_elem(const Env& d, V3 pt, int body, int k, int c) {
  const DModel& m = d.m;
  if (!m.body_dofanc[(long)body * m.sz.nv + c]) return 0;
  FD cdof = d.cdof(), cvel = d.cvel();
  const V3 com = ld3(d.subtree_com(), 3 * m.body_rootid[body]);
  const V3 off = pt - com;
  const V3 pang = ld3(cvel, 6 * body);
  const V3 plin = ld3(cvel, 6 * body + 3) - cross(off, pang);     // point velocity (mju_transformSpatial)
  S6 cd = ld6(d.cdof_dot(), 6 * c);
  const int j = m.dof_jntid[c], jt = m.jnt_type[j];
  if (jt == JNT_BALL || (jt == JNT_FREE && c >= m.jnt_dofadr[j] + 3)) cd = cross_motion(ld6(cvel, 6 * m.dof_bodyid[c]), ld6(cdof, 6 * c));
  const V3 t1 = cross(V3{cd.v[0], cd.v[1], cd.v[2]}, off);
  const V3 t2 = cross(ld3(cdof, 6 * c), plin);
  return cd.v[3 + k] + get(t1, k) + get(t2, k);
}

// element of the time derivative of the rotational Jacobian (mj_jacDot, jacr): component k, dof c
MJB_HD double jacrdot_elem(const Env& d, int body, int k, int c) {
  const DModel& m = d.m;
  if (!m.body_dofanc[(long)body * m.sz.nv + c]) return 0;
  const int j = m.dof_jntid[c], jt = m.jnt_type[j];
  if (jt == JNT_BALL || (jt == JNT_FREE && c >= m.jnt_dofadr[j] + 3))
    return cross_motion(ld6(d.cvel(), 6 * m.dof_bodyid[c]), ld6(d.cdof(), 6 * c)).v[k];
  return d.cdof_dot()[6 * c + k];
}

MJB_HD void make_constraint(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv, njmax = m.sz.njmax;
  FI nefc_f = d.nefc(), ne_f = d.ne(), nf_f = d.nf(), nl_f = d.nl();
  FI ilim = d.scr_ilim();          // [0,nlim): row index or -1   [nlim, 2nlim): unused
  FD dlim = d.scr_efc();           // limit distance per candidate (nlim <= njmax guaranteed by host)
  FI cadr = d.con_efcadr(), cexc = d.con_exclude(), cdim = d.con_dim();
  FD qpos = d.qpos();
  if (m.opt.disableflags & DSBL_CONSTRAINT) {
    MJB_LANE0 { ne_f[0] = 0; nf_f[0] = 0; nl_f[0] = 0; nefc_f[0] = 0; }
    MJB_PSYNC();
    return;
  }
  const bool do_eq = (d.feat & FEAT_EQUALITY) && m.sz.neq > 0 && !(m.opt.disableflags & DSBL_EQUALITY);
  FI ieq = d.scr_ieq();   // row of every equality, -1 if inactive
  const bool do_fl = m.opt.has_frictionloss && !(m.opt.disableflags & DSBL_FRICTIONLOSS);
  const bool do_lim = m.opt.has_limits && !(m.opt.disableflags & DSBL_LIMIT);
  const bool do_con = !(m.opt.disableflags & DSBL_CONTACT);
  const int nlim = do_lim ? m.sz.nlim : 0;
  const int ncon = d.ncon()[0];

  // ---- limit candidates: distance and activity
  FD tl = d.ten_length();
  MJB_PFOR(c, nlim) {
    const int kind = m.lim_kind[c], i = m.lim_id[c], side = m.lim_side[c];
    double dist, margin;
    if (kind == LIM_HINGE) {
      margin = m.jnt_margin[i];
      dist = side * (m.jnt_range[2 * i + (side + 1) / 2] - qpos[m.jnt_qposadr[i]]);
    } else if (kind == LIM_BALL) {
      margin = m.jnt_margin[i];
      Q4 q = ld4(qpos, m.jnt_qposadr[i]);
      normalize(q);
      V3 aa = quat2vel(q, 1);
      const double value = normalize(aa);
      dist = dmax(m.jnt_range[2 * i], m.jnt_range[2 * i + 1]) - value;
    } else {
      margin = m.tendon_margin[i];
      dist = side * (m.tendon_range[2 * i + (side + 1) / 2] - tl[i]);
    }
    dlim[c] = dist;
    ilim[c] = (dist < margin) ? 1 : 0;
  }
  MJB_PSYNC();

  // ---- serial scan: row index of every friction dof, active limit and included contact
  MJB_LANE0 {
    int nefc = 0, ne = 0, nf = 0, nl = 0;
    bool full = false;
    // equality rows first (mj_instantiateEquality: joint / tendon couplings, one row each)
    if (do_eq) {
      for (int i = 0; i < m.sz.neq; i++) {
        ieq[i] = -1;
        if ((int)d.eq_active()[i] == 0) continue;   // runtime switch (mjData.eq_active), reset to eq_active0
        const int rows = (m.eq_kind[i] == EQ_WELD) ? 6 : (m.eq_kind[i] == EQ_CONNECT) ? 3 : 1;
        if (nefc + rows > njmax) { full = true; continue; }
        ieq[i] = nefc;
        nefc += rows; ne += rows;
      }
    }
    if (do_fl) { nf = m.sz.nfl; if (nefc + nf > njmax) { nf = njmax - nefc; full = true; } nefc += nf; }
    for (int c = 0; c < nlim; c++) {
      if (!ilim[c]) { ilim[c] = -1; continue; }
      if (m.lim_kind[c] == LIM_TENDON) {
        // rows with an identically zero Jacobian are dropped (mj_addConstraint "empty" guard)
        const int i = m.lim_id[c], adr = m.ten_J_rowadr[i], nnz = m.ten_J_rownnz[i];
        bool empty = true;
        for (int a = 0; a < nnz; a++) if (d.ten_J()[adr + a] * (double)(-m.lim_side[c]) != 0) empty = false;
        if (empty) { ilim[c] = -1; continue; }
      }
      if (full || nefc + 1 > njmax) { full = true; ilim[c] = -1; continue; }
      ilim[c] = nefc++;
      nl++;
    }
    for (int i = 0; i < ncon; i++) {
      cadr[i] = -1;
      if (!do_con || cexc[i] || full) continue;
      const int rows = (cdim[i] == 1) ? 1 : 2 * (cdim[i] - 1);
      if (nefc + rows > njmax) { full = true; continue; }
      cadr[i] = nefc;
      nefc += rows;
    }
    if (full) d.warning()[WARN_CNSTRFULL] += 1;
    ne_f[0] = ne; nf_f[0] = nf; nl_f[0] = nl; nefc_f[0] = nefc;
  }
  MJB_PSYNC();
  const int nefc = nefc_f[0], nf = nf_f[0], ne = ne_f[0];
  if (!nefc) return;

  FD J = d.efc_J(), epos = d.efc_pos(), emargin = d.efc_margin(), efl = d.efc_frictionloss();
  FI type = d.efc_type(), id = d.efc_id();

  // ---- equality rows: scalar joint / tendon couplings  q1 - q1_0 = data0 + poly(q2 - q2_0)
  if ((d.feat & FEAT_EQUALITY) && ne) {
    FD tJ = d.ten_J(), tlen = d.ten_length();
    MJB_PFOR(i, m.sz.neq) {
      const int r = ieq[i];
      if (r < 0) continue;
      const int o1 = m.eq_obj1id[i], o2 = m.eq_obj2id[i];
      if (m.eq_kind[i] >= EQ_CONNECT) {   // connect: 3 rows, J = jac(0) - jac(1) at the anchors; weld: + 3 rotational rows
        V3 p0, p1;
        connect_anchors(d, i, p0, p1);
        const V3 cp = p0 - p1;
        for (int k = 0; k < 3; k++) {
          FD row = J + (long)(r + k) * nv;
          for (int c = 0; c < nv; c++) row[c] = jac_elem(d, p0, o1, k, c) - jac_elem(d, p1, o2, k, c);
          epos[r + k] = get(cp, k); emargin[r + k] = 0; efl[r + k] = 0; type[r + k] = CNSTR_EQUALITY; id[r + k] = i;
        }
        if (m.eq_kind[i] == EQ_WELD) {
          const double* data = m.eq_data + kNEqData * i;
          const double ts = data[10];
          const Q4 quat = qmul(ld4(d.xquat(), 4 * o1), Q4{data[6], data[7], data[8], data[9]});   // q0 * relpose
          Q4 q1n = ld4(d.xquat(), 4 * o2);
          q1n = Q4{q1n.w, -q1n.x, -q1n.y, -q1n.z};                                                   // neg(q1)
          const Q4 qe = qmul(q1n, quat);
          const double ce[3] = {qe.x * ts, qe.y * ts, qe.z * ts};
          FD cdof = d.cdof();
          for (int c = 0; c < nv; c++) {   // 0.5 * neg(q1) * (jacr0 - jacr1)_col * q0*relpose, scaled
            const bool in1 = m.body_dofanc[(long)o1 * nv + c], in2 = m.body_dofanc[(long)o2 * nv + c];
            V3 ax;
            ax.x = (in1 ? cdof[6 * c] : 0.0) - (in2 ? cdof[6 * c] : 0.0);
            ax.y = (in1 ? cdof[6 * c + 1] : 0.0) - (in2 ? cdof[6 * c + 1] : 0.0);
            ax.z = (in1 ? cdof[6 * c + 2] : 0.0) - (in2 ? cdof[6 * c + 2] : 0.0);
            const Q4 q3 = qmul(qmul_axis(q1n, ax), quat);
            J[(long)(r + 3) * nv + c] = 0.5 * q3.x * ts;
            J[(long)(r + 4) * nv + c] = 0.5 * q3.y * ts;
            J[(long)(r + 5) * nv + c] = 0.5 * q3.z * ts;
          }
          for (int k = 0; k < 3; k++) {
            epos[r + 3 + k] = ce[k]; emargin[r + 3 + k] = 0; efl[r + 3 + k] = 0; type[r + 3 + k] = CNSTR_EQUALITY; id[r + 3 + k] = i;
          }
        }
        continue;
      }
      const bool jnt = m.eq_kind[i] == EQ_JOINT;
      const double* data = m.eq_data + kNEqData * i;
      FD row = J + (long)r * nv;
      for (int k = 0; k < nv; k++) row[k] = 0;
      auto value = [&](int o) { return jnt ? qpos[m.jnt_qposadr[o]] : tlen[o]; };
      auto ref0 = [&](int o) { return jnt ? m.qpos0[m.jnt_qposadr[o]] : m.tendon_length0[o]; };
      auto add_jac = [&](int o, double scl, bool first) {   // row (+)= jac(o) * scl   (mju_addToScl on the dense row)
        if (jnt) {
          const int dof = m.jnt_dofadr[o];
          if (first) row[dof] = 1;
          else for (int k = 0; k < nv; k++) row[k] += ((k == dof) ? 1.0 : 0.0) * scl;
        } else {
          const int adr = m.ten_J_rowadr[o], nnz = m.ten_J_rownnz[o];
          if (first) { for (int a = 0; a < nnz; a++) row[m.ten_J_colind[adr + a]] = tJ[adr + a]; }
          else {
            for (int k = 0; k < nv; k++) {
              double j2 = 0;
              for (int a = 0; a < nnz; a++) if (m.ten_J_colind[adr + a] == k) j2 = tJ[adr + a];
              row[k] += j2 * scl;
            }
          }
        }
      };
      add_jac(o1, 1, true);
      double cpos;
      if (o2 >= 0) {
        const double dif = value(o2) - ref0(o2);
        cpos = value(o1) - ref0(o1) - data[0] -
               (data[1] * dif + data[2] * dif * dif + data[3] * dif * dif * dif + data[4] * dif * dif * dif * dif);
        const double deriv = data[1] + 2 * data[2] * dif + 3 * data[3] * dif * dif + 4 * data[4] * dif * dif * dif;
        add_jac(o2, -deriv, false);
      } else {
        cpos = value(o1) - ref0(o1) - data[0];
      }
      epos[r] = cpos; emargin[r] = 0; efl[r] = 0; type[r] = CNSTR_EQUALITY; id[r] = i;
    }
  }

  // ---- friction-loss and limit rows (one lane per row)
  MJB_PFOR(r_, nf) {
    const int r = ne + r_;
    const int i = m.fl_dof[r_];
    FD row = J + (long)r * nv;
    for (int c = 0; c < nv; c++) row[c] = 0;
    if (i < 0) {   // tendon with dry friction: the tendon's Jacobian row
      const int t = -i - 1, adr = m.ten_J_rowadr[t], nnz = m.ten_J_rownnz[t];
      FD tJ = d.ten_J();
      for (int a = 0; a < nnz; a++) row[m.ten_J_colind[adr + a]] = tJ[adr + a];
      epos[r] = 0; emargin[r] = 0; efl[r] = m.tendon_frictionloss[t]; type[r] = CNSTR_FRICTION_TENDON; id[r] = t;
      continue;
    }
    row[i] = 1;
    epos[r] = 0; emargin[r] = 0; efl[r] = m.dof_frictionloss[i]; type[r] = CNSTR_FRICTION_DOF; id[r] = i;
  }
  MJB_PFOR(c, nlim) {
    const int r = ilim[c];
    if (r < 0) continue;
    const int kind = m.lim_kind[c], i = m.lim_id[c], side = m.lim_side[c];
    FD row = J + (long)r * nv;
    for (int k = 0; k < nv; k++) row[k] = 0;
    if (kind == LIM_HINGE) {
      row[m.jnt_dofadr[i]] = -(double)side;
      emargin[r] = m.jnt_margin[i]; type[r] = CNSTR_LIMIT_JOINT;
    } else if (kind == LIM_BALL) {
      Q4 q = ld4(qpos, m.jnt_qposadr[i]);
      normalize(q);
      V3 aa = quat2vel(q, 1);
      normalize(aa);
      const int da = m.jnt_dofadr[i];
      row[da] = aa.x * -1; row[da + 1] = aa.y * -1; row[da + 2] = aa.z * -1;
      emargin[r] = m.jnt_margin[i]; type[r] = CNSTR_LIMIT_JOINT;
    } else {
      const int adr = m.ten_J_rowadr[i], nnz = m.ten_J_rownnz[i];
      FD tJ = d.ten_J();
      for (int a = 0; a < nnz; a++) row[m.ten_J_colind[adr + a]] = tJ[adr + a];
      for (int k = 0; k < nv; k++) row[k] = row[k] * (double)(-side);
      emargin[r] = m.tendon_margin[i]; type[r] = CNSTR_LIMIT_TENDON;
    }
    epos[r] = dlim[c]; efl[r] = 0; id[r] = i;
  }

  // ---- contact rows: one work item per (row, dof) element
  FD cpos = d.con_pos(), cframe = d.con_frame(), cdist = d.con_dist(), cinc = d.con_includemargin();
  FD cfri = d.con_friction();
  FI cg1 = d.con_geom1(), cg2 = d.con_geom2();
  MJB_PFOR(i, ncon) {   // row headers
    const int a = cadr[i];
    if (a < 0) continue;
    const int dim = cdim[i], rows = (dim == 1) ? 1 : 2 * (dim - 1);
    for (int r = 0; r < rows; r++) {
      epos[a + r] = cdist[i]; emargin[a + r] = cinc[i]; efl[a + r] = 0;
      type[a + r] = (dim == 1) ? CNSTR_CONTACT_FRICTIONLESS : CNSTR_CONTACT_PYRAMIDAL;
      id[a + r] = i;
    }
  }
  MJB_PSYNC();
  const int first_con_row = ne + nf + nl_f[0];
  MJB_PFOR(it, (nefc - first_con_row) * nv) {
    const int r = first_con_row + it / nv, c = it % nv;
    const int i = id[r];
    const int a = cadr[i], dim = cdim[i];
    V3 pt = ld3(cpos, 3 * i);
    const int b1 = m.geom_bodyid[cg1[i]], b2 = m.geom_bodyid[cg2[i]];
    // jacdif[k] = jac2[k] - jac1[k]; rotated rows jc[q] = sum_k frame[q][k]*jacdif[k] (zero-skip)
    double jd[3];
    for (int k = 0; k < 3; k++) jd[k] = jac_elem(d, pt, b2, k, c) - jac_elem(d, pt, b1, k, c);
    double j0 = 0;
    for (int k = 0; k < 3; k++) { const double f = cframe[9 * i + k]; if (f != 0) j0 += jd[k] * f; }
    if (dim == 1) { J[(long)r * nv + c] = j0; }
    else {
      const int q = (r - a) / 2 + 1;                    // friction direction 1..dim-1
      double jq = 0;
      if (q < 3) {
        for (int k = 0; k < 3; k++) { const double f = cframe[9 * i + 3 * q + k]; if (f != 0) jq += jd[k] * f; }
      } else {   // condim 4 / 6: torsional (about the normal) and rolling (about the tangents) rows use the rotational Jacobian
        FD cdof = d.cdof();
        const bool in1 = m.body_dofanc[(long)b1 * nv + c], in2 = m.body_dofanc[(long)b2 * nv + c];
        for (int k = 0; k < 3; k++) {
          const double jr = (in2 ? cdof[6 * c + k] : 0.0) - (in1 ? cdof[6 * c + k] : 0.0);
          const double f = cframe[9 * i + 3 * (q - 3) + k];
          if (f != 0) jq += jr * f;
        }
      }
      const double mu = cfri[5 * i + q - 1];
      J[(long)r * nv + c] = ((r - a) % 2 == 0) ? j0 + jq * mu : j0 + jq * (-mu);
    }
  }
  MJB_PSYNC();

  // ---- diagApprox (per row)
  FD dA = d.efc_diagA();
  MJB_PFOR(r, nefc) {
    const int t = type[r], k = id[r];
    if (t == CNSTR_EQUALITY && m.eq_kind[k] >= EQ_CONNECT) {   // translation rows, then (weld) rotation rows
      const int rot = (r - ieq[k] > 2) ? 1 : 0;
      dA[r] = m.body_invweight0[2 * m.eq_obj1id[k] + rot] + m.body_invweight0[2 * m.eq_obj2id[k] + rot];
    }
    else if (t == CNSTR_EQUALITY) {
      const bool jnt = m.eq_kind[k] == EQ_JOINT;
      const int o1 = m.eq_obj1id[k], o2 = m.eq_obj2id[k];
      double a = jnt ? m.dof_invweight0[m.jnt_dofadr[o1]] : m.tendon_invweight0[o1];
      if (o2 >= 0) a += jnt ? m.dof_invweight0[m.jnt_dofadr[o2]] : m.tendon_invweight0[o2];
      dA[r] = a;
    }
    else if (t == CNSTR_FRICTION_DOF) dA[r] = m.dof_invweight0[k];
    else if (t == CNSTR_LIMIT_JOINT) dA[r] = m.dof_invweight0[m.jnt_dofadr[k]];
    else if (t == CNSTR_LIMIT_TENDON || t == CNSTR_FRICTION_TENDON) dA[r] = m.tendon_invweight0[k];
    else {
      const int b1 = m.geom_bodyid[cg1[k]], b2 = m.geom_bodyid[cg2[k]];
      double tran = 0, rot = 0;
      tran += m.body_invweight0[2 * b1] * 1.0; rot += m.body_invweight0[2 * b1 + 1] * 1.0;
      tran += m.body_invweight0[2 * b2] * 1.0; rot += m.body_invweight0[2 * b2 + 1] * 1.0;
      if (t == CNSTR_CONTACT_FRICTIONLESS) dA[r] = tran;
      else {
        const int j = (r - cadr[k]) / 2;
        const double fri = cfri[5 * k + j];
        dA[r] = tran + fri * fri * (j < 2 ? tran : rot);
      }
    }
  }
  MJB_PSYNC();

  // ---- impedance: R, KBIP per row (rows of one pyramid share pos/margin, hence imp)
  FD R = d.efc_R(), KBIP = d.efc_KBIP(), D = d.efc_D();
  MJB_PFOR(r, nefc) {
    const int t = type[r], k = id[r];
    double solref[2], solimp[5];
    if (t == CNSTR_EQUALITY) { for (int j = 0; j < 2; j++) solref[j] = m.eq_solref[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.eq_solimp[5 * k + j]; }
    else if (t == CNSTR_FRICTION_DOF) { for (int j = 0; j < 2; j++) solref[j] = m.dof_solref[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.dof_solimp[5 * k + j]; }
    else if (t == CNSTR_LIMIT_JOINT) { for (int j = 0; j < 2; j++) solref[j] = m.jnt_solref[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.jnt_solimp[5 * k + j]; }
    else if (t == CNSTR_FRICTION_TENDON) { for (int j = 0; j < 2; j++) solref[j] = m.tendon_solref_fri[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.tendon_solimp_fri[5 * k + j]; }
    else if (t == CNSTR_LIMIT_TENDON) { for (int j = 0; j < 2; j++) solref[j] = m.tendon_solref_lim[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.tendon_solimp_lim[5 * k + j]; }
    else {
      for (int j = 0; j < 2; j++) solref[j] = d.con_solref()[2 * k + j];
      for (int j = 0; j < 5; j++) solimp[j] = d.con_solimp()[5 * k + j];
    }
    if ((solref[0] > 0) != (solref[1] > 0)) { solref[0] = 0.02; solref[1] = 1; }   // mj_defaultSolRefImp
    if (!(m.opt.disableflags & DSBL_REFSAFE) && solref[0] > 0) solref[0] = dmax(solref[0], 2 * m.opt.timestep);
    solimp[0] = dmin(kMaxImp, dmax(kMinImp, solimp[0]));
    solimp[1] = dmin(kMaxImp, dmax(kMinImp, solimp[1]));
    solimp[2] = dmax(0, solimp[2]);
    solimp[3] = dmin(kMaxImp, dmax(kMinImp, solimp[3]));
    solimp[4] = dmax(1, solimp[4]);
    double imp, impP;
    double ipos = epos[r], imargin = emargin[r];
    if (t == CNSTR_EQUALITY && m.eq_kind[k] >= EQ_CONNECT) {   // getposdim: the 3 (6) rows share |pos| (mju_norm)
      const int base = ieq[k], dimp = (m.eq_kind[k] == EQ_WELD) ? 6 : 3;
      ipos = sqrt(dot_ref(dimp, [&](int q) { return epos[base + q]; }, [&](int q) { return epos[base + q]; }));
      imargin = emargin[base];
    }
    impedance(solimp, ipos, imargin, imp, impP);
    R[r] = dmax(kMinVal, (1 - imp) * dA[r] / imp);
    double K, Bv;
    if (t == CNSTR_FRICTION_DOF || t == CNSTR_FRICTION_TENDON) K = 0;
    else if (solref[0] > 0) K = 1 / dmax(kMinVal, solimp[1] * solimp[1] * solref[0] * solref[0] * solref[1] * solref[1]);
    else K = -solref[0] / dmax(kMinVal, solimp[1] * solimp[1]);
    if (solref[1] > 0) Bv = 2 / dmax(kMinVal, solimp[1] * solref[0]);
    else Bv = -solref[1] / dmax(kMinVal, solimp[1]);
    KBIP[4 * r] = K; KBIP[4 * r + 1] = Bv; KBIP[4 * r + 2] = imp; KBIP[4 * r + 3] = impP;
  }
  MJB_PSYNC();
  // pyramidal contacts: common R matched to the elliptic model, contact mu (one lane per contact)
  MJB_PFOR(i, ncon) {
    const int a = cadr[i];
    if (a < 0 || cdim[i] == 1) continue;
    const int dim = cdim[i];
    const double R0 = R[a];
    const double R1 = R0 / dmax(kMinVal, m.opt.impratio);
    const double mu = cfri[5 * i] * sqrt(R1 / R0);
    d.con_mu()[i] = mu;
    const double Rpy = 2 * mu * mu * R0;
    for (int j = 0; j < 2 * (dim - 1); j++) R[a + j] = Rpy;
  }
  MJB_PSYNC();
  MJB_PFOR(r, nefc) {
    D[r] = 1 / R[r];
    dA[r] = R[r] * KBIP[4 * r + 2] / (1 - KBIP[4 * r + 2]);
  }
  MJB_PSYNC();
}

// ---- dense J * vec (one lane per row) and J' * vec (one lane per dof), reference accumulation order
MJB_HD void mul_jac_vec(const Env& d, FD res, FD vec) {
  const int nefc = d.nefc()[0], nv = d.m.sz.nv;
  FD J = d.efc_J();
  MJB_PFOR(r, nefc) {
    FD row = J + (long)r * nv;
    res[r] = dot_ref(nv, [&](int c) { return row[c]; }, [&](int c) { return vec[c]; });
  }
  MJB_PSYNC();
}
MJB_HD void mul_jacT_vec(const Env& d, FD res, FD vec) {
  const int nefc = d.nefc()[0], nv = d.m.sz.nv;
  if (!nefc) return;
  FD J = d.efc_J();
  MJB_PFOR(c, nv) {
    double s = 0;
    for (int r = 0; r < nefc; r++) {
      const double f = vec[r];
      if (f != 0) s += J[(long)r * nv + c] * f;
    }
    res[c] = s;
  }
  MJB_PSYNC();
}

// ---- dual projection: Y = J L^-T D^-1/2, AR = Y Y' + diag(R) (dense) ---------------------------
MJB_HD void project_constraint(const Env& d) {
  const DModel& m = d.m;
  const int nefc = d.nefc()[0], nv = m.sz.nv;
  if (!nefc || (d.solver != SOL_PGS && m.opt.noslip_iterations <= 0)) return;   // mj_isDual
  FD J = d.efc_J(), Y = d.efc_Y(), AR = d.efc_AR(), qLD = d.qLD(), R = d.efc_R();
  FD sq = d.scr_nv();
  MJB_PROF_BEGIN
  MJB_PFOR(i, nv) sq[i] = 1 / sqrt(qLD[m.M_rowadr[i] + m.M_rownnz[i] - 1]);
  MJB_PSYNC();
  // one lane per row: half back-substitution x <- L^-T x (mj_solveM2) in GATHER form: x[j] is final once every
  // descendant i > j has been subtracted, in descending i (the order of the reference's scatter loop, terms with
  // x[i] == 0 skipped as there); the accumulator is a register and x lives in the row of Y itself, so the only
  // dependent memory round trip is one store -> load per dof instead of one per nonzero of L
  MJB_PFOR(r, nefc) {
    FD src = J + (long)r * nv, dst = Y + (long)r * nv;
    for (int j = nv - 1; j >= 0; j--) {
      double s = src[j];
      const int a0 = m.mt_adr[j], a1 = m.mt_adr[j + 1];
      for (int a = a0; a < a1; a++) {
        const double xi = dst[m.mt_dof[a]];
        const double t = s - qLD[m.mt_qadr[a]] * xi;
        s = xi != 0 ? t : s;
      }
      dst[j] = s;
    }
    for (int j = 0; j < nv; j++) dst[j] = dst[j] * sq[j];
  }
  MJB_PSYNC();
  MJB_PROF_MARK(11)
  // AR[i][c] = sum_j Y[c][j] * Y[i][j], j ascending, skipping Y[i][j] == 0 (mju_sqrMatTD); one lane
  // per lower-triangle element, mirrored; R added on the diagonal
  const int ntri = nefc * (nefc + 1) / 2;
  MJB_PFOR(t, ntri) {
    int i = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    const int c = t - i * (i + 1) / 2;
    FD yi = Y + (long)i * nv, yc = Y + (long)c * nv;
    double s = 0;
    int j = 0;
    for (; j + 4 <= nv; j += 4) {     // loads of four terms in flight; a skipped term leaves s unchanged
      const double v0 = yi[j], v1 = yi[j + 1], v2 = yi[j + 2], v3 = yi[j + 3];
      const double p0 = yc[j] * v0, p1 = yc[j + 1] * v1, p2 = yc[j + 2] * v2, p3 = yc[j + 3] * v3;
      double u = s + p0; s = v0 != 0 ? u : s;
      u = s + p1; s = v1 != 0 ? u : s;
      u = s + p2; s = v2 != 0 ? u : s;
      u = s + p3; s = v3 != 0 ? u : s;
    }
    for (; j < nv; j++) {
      const double v = yi[j];
      const double u = s + yc[j] * v;
      s = v != 0 ? u : s;
    }
    if (i == c) s += R[i];
    AR[(long)i * nefc + c] = s;
    AR[(long)c * nefc + i] = s;
  }
  MJB_PSYNC();
  MJB_PROF_MARK(12)
}

// row . f in mju_dot's accumulation order on raw pointers
MJB_HD double dot_ptr(const double* a, const double* f, int n) {
  double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  int c = 0;
  for (; c <= n - 4; c += 4) {
    r0 += a[c] * f[c];
    r1 += a[c + 1] * f[c + 1];
    r2 += a[c + 2] * f[c + 2];
    r3 += a[c + 3] * f[c + 3];
  }
  double res = (r0 + r2) + (r1 + r3);
  const int t = n - c;
  if (t == 3) res += a[c] * f[c] + a[c + 1] * f[c + 1] + a[c + 2] * f[c + 2];
  else if (t == 2) res += a[c] * f[c] + a[c + 1] * f[c + 1];
  else if (t == 1) res += a[c] * f[c];
  return res;
}

// ---- efc_vel, efc_aref ---------------------------------------------------------------------------
MJB_HD void reference_constraint(const Env& d) {
  const int nefc = d.nefc()[0];
  if (!nefc) return;
  FD vel = d.efc_vel(), aref = d.efc_aref(), KBIP = d.efc_KBIP(), pos = d.efc_pos(), margin = d.efc_margin();
  mul_jac_vec(d, vel, d.qvel());
  MJB_PFOR(i, nefc)
    aref[i] = -KBIP[4 * i + 1] * vel[i] - KBIP[4 * i] * KBIP[4 * i + 2] * (pos[i] - margin[i]);
  MJB_PSYNC();
  // mj_Jdotv (engine_core_constraint.c:1056-1200): connect rows get aref -= (Jdot1 - Jdot2) * qvel
  const DModel& m = d.m;
  if ((d.feat & FEAT_EQUALITY) && d.ne()[0] && m.sz.neq) {
    FI ieq = d.scr_ieq();
    FD qvel = d.qvel();
    const int nv = m.sz.nv;
    MJB_PFOR(it, m.sz.neq * 3) {
      const int eq = it / 3, k = it - eq * 3;
      const int r = ieq[eq];
      if (r < 0 || m.eq_kind[eq] < EQ_CONNECT) continue;
      V3 p0, p1;
      connect_anchors(d, eq, p0, p1);
      const int o1 = m.eq_obj1id[eq], o2 = m.eq_obj2id[eq];
      const double j1 = dot_ref(nv, [&](int c) { return jacdot_elem(d, p0, o1, k, c); }, [&](int c) { return qvel[c]; });
      const double j2 = dot_ref(nv, [&](int c) { return jacdot_elem(d, p1, o2, k, c); }, [&](int c) { return qvel[c]; });
      aref[r + k] -= j1 - j2;
    }
    MJB_PSYNC();
    // welds: rotational part, three product-rule terms of d/dt [0.5 neg(q1) (J0-J1) v q0 relpose]
    MJB_PFOR(eq, m.sz.neq) {
      const int r = ieq[eq];
      if (r < 0 || m.eq_kind[eq] != EQ_WELD) continue;
      const int o1 = m.eq_obj1id[eq], o2 = m.eq_obj2id[eq];
      const double* data = m.eq_data + kNEqData * eq;
      const double ts = data[10];
      const Q4 rel{data[6], data[7], data[8], data[9]};
      const Q4 q0 = ld4(d.xquat(), 4 * o1), q1 = ld4(d.xquat(), 4 * o2);
      const Q4 q0r = qmul(q0, rel);
      const Q4 negq1{q1.w, -q1.x, -q1.y, -q1.z};
      const V3 w1 = ld3(d.cvel(), 6 * o1), w2 = ld3(d.cvel(), 6 * o2);
      const V3 dom = w1 - w2;
      const Q4 qdot0r = qmul(deriv_quat(q0, w1), rel);
      const Q4 qd1 = deriv_quat(q1, w2);
      const Q4 negqdot1{qd1.w, -qd1.x, -qd1.y, -qd1.z};
      V3 jr1, jr2;   // rotational jacDot * qvel of the two bodies
      for (int k = 0; k < 3; k++) {
        const double a = dot_ref(nv, [&](int c) { return jacrdot_elem(d, o1, k, c); }, [&](int c) { return qvel[c]; });
        const double b = dot_ref(nv, [&](int c) { return jacrdot_elem(d, o2, k, c); }, [&](int c) { return qvel[c]; });
        if (k == 0) { jr1.x = a; jr2.x = b; } else if (k == 1) { jr1.y = a; jr2.y = b; } else { jr1.z = a; jr2.z = b; }
      }
      const V3 djr = jr1 - jr2;
      const Q4 t1 = qmul(qmul_axis(negqdot1, dom), q0r);
      const Q4 t2 = qmul(qmul_axis(negq1, djr), q0r);
      const Q4 t3 = qmul(qmul_axis(negq1, dom), qdot0r);
      aref[r + 3] -= 0.5 * (t1.x + t2.x + t3.x) * ts;
      aref[r + 4] -= 0.5 * (t1.y + t2.y + t3.y) * ts;
      aref[r + 5] -= 0.5 * (t1.z + t2.z + t3.z) * ts;
    }
    MJB_PSYNC();
  }
}

// ---- primal constraint update (pyramidal / scalar rows): force, state; optional cost (serial sum)
MJB_HD double constraint_update(const Env& d, FD jar, bool want_cost) {
  const int nefc = d.nefc()[0], ne = d.ne()[0], nf = ne + d.nf()[0];   // nf: end of the friction rows
  FD D = d.efc_D(), R = d.efc_R(), floss = d.efc_frictionloss(), force = d.efc_force();
  FI state = d.efc_state();
  MJB_PFOR(i, nefc) {
    double f = -D[i] * jar[i];
    int st;
    if (i < ne) st = STATE_QUADRATIC;
    else if (i < nf) {
      if (jar[i] <= -R[i] * floss[i]) { f = floss[i]; st = STATE_LINEARNEG; }
      else if (jar[i] >= R[i] * floss[i]) { f = -floss[i]; st = STATE_LINEARPOS; }
      else st = STATE_QUADRATIC;
    } else if (jar[i] >= 0) { f = 0; st = STATE_SATISFIED; }
    else st = STATE_QUADRATIC;
    force[i] = f; state[i] = st;
  }
  MJB_PSYNC();
  double s = 0;
  if (want_cost) {   // same left-to-right accumulation on every lane (uniform value, no broadcast needed)
    for (int i = 0; i < nefc; i++) {
      const int st = state[i];
      if (st == STATE_LINEARNEG) s += -0.5 * R[i] * floss[i] * floss[i] - floss[i] * jar[i];
      else if (st == STATE_LINEARPOS) s += -0.5 * R[i] * floss[i] * floss[i] + floss[i] * jar[i];
      else if (st == STATE_QUADRATIC) s += 0.5 * D[i] * jar[i] * jar[i];
    }
  }
  mul_jacT_vec(d, d.qfrc_constraint(), force);
  return s;
}

// res = M * vec (mj_mulM -> mju_mulSymVecSparse, engine_util_sparse.c): per output dof the order is
// the diagonal term, the own-row off-diagonals from the last column to the first, then the rows of
// the descendants in ascending order
MJB_HD void mul_M(const Env& d, FD res, FD vec) {
  const DModel& m = d.m;
  FD M = d.M();
  MJB_PFOR(i, m.sz.nv) {
    const int adr = m.M_rowadr[i], nnz = m.M_rownnz[i];
    double s = M[adr + nnz - 1] * vec[i];
    for (int k = nnz - 2; k >= 0; k--) s += M[adr + k] * vec[m.M_colind[adr + k]];
    const int a0 = m.mt_adr[i], an = m.mt_adr[i + 1] - a0;
    for (int c = an - 1; c >= 0; c--) s += M[m.mt_qadr[a0 + c]] * vec[m.mt_dof[a0 + c]];
    res[i] = s;
  }
  MJB_PSYNC();
}
MJB_HD bool use_islands(const Env& d);

MJB_HD void constraint_begin(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv, nefc = d.nefc()[0];
  FD qfc = d.qfrc_constraint(), qacc = d.qacc(), qas = d.qacc_smooth();
  MJB_PFOR(i, nv) qfc[i] = 0;
  MJB_LANE0 { for (int k = 0; k < NISLAND; k++) d.solver_niter()[k] = 0; }
  if (!nefc) { MJB_PFOR(i, nv) qacc[i] = qas[i]; MJB_PSYNC(); return; }
  MJB_PSYNC();
  FD b = d.efc_b(), aref = d.efc_aref(), force = d.efc_force();
  mul_jac_vec(d, b, qas);
  MJB_PFOR(i, nefc) b[i] -= aref[i];
  MJB_PSYNC();
  if (!(m.opt.disableflags & DSBL_WARMSTART)) {
    FD jar = d.scr_efc(), ws = d.qacc_warmstart();
    MJB_PFOR(i, nv) qacc[i] = ws[i];
    mul_jac_vec(d, jar, ws);
    MJB_PFOR(i, nefc) jar[i] -= aref[i];
    MJB_PSYNC();
    double cost_ws = constraint_update(d, jar, d.solver != SOL_PGS);
    if (d.solver == SOL_PGS) {
      FD AR = d.efc_AR(), ARf = d.scr_efc() + nefc;
      MJB_PFOR(r, nefc) ARf[r] = dot_ptr(AR.p + (long)r * nefc, force.p, nefc);
      MJB_PSYNC();
      // two serial dots of length nefc, evaluated identically by every lane (uniform decision)
      double pw = dot_ref(nefc, [&](int i) { return force[i]; }, [&](int i) { return b[i]; });
      pw += 0.5 * dot_ref(nefc, [&](int i) { return force[i]; }, [&](int i) { return ARf[i]; });
      MJB_PSYNC();
      if (pw > 0) {
        MJB_PFOR(i, nefc) force[i] = 0;
        MJB_PFOR(i, nv) qfc[i] = 0;
        MJB_PSYNC();
      }
    } else {
      // Newton/CG: choose the cheaper of qacc_warmstart and qacc_smooth by primal cost
      FD Ma = d.scr_nv(), qfs = d.qfrc_smooth();
      mul_M(d, Ma, ws);
      for (int i = 0; i < nv; i++) cost_ws += 0.5 * (Ma[i] - qfs[i]) * (ws[i] - qas[i]);
      MJB_PSYNC();
      const double cost_smooth = constraint_update(d, b, true);
      if (cost_ws > cost_smooth) { MJB_PFOR(i, nv) qacc[i] = qas[i]; }
      MJB_PSYNC();
    }
    if (use_islands(d)) {   // dofs outside every island are unconstrained: qacc = qacc_smooth
      const int* tisl = d.tree_island().p;
      MJB_PFOR(j, nv) { if (tisl[m.dof_treeid[j]] < 0) qacc[j] = qas[j]; }
      MJB_PSYNC();
    }
  } else {
    MJB_PFOR(i, nv) qacc[i] = qas[i];
    MJB_PFOR(i, nefc) force[i] = 0;
    MJB_PSYNC();
  }
}

// dual state from forces on raw pointers (engine_solver.c dualState, scalar rows)
MJB_HD void dual_state_ptr(const Env& d, const double* force, const double* floss, int nefc, int ne, int nf) {
  FI state = d.efc_state();
  MJB_PFOR(i, nefc) {
    if (i < ne) state[i] = STATE_QUADRATIC;
    else if (i < ne + nf) {
      if (force[i] <= -floss[i]) state[i] = STATE_LINEARPOS;
      else if (force[i] >= floss[i]) state[i] = STATE_LINEARNEG;
      else state[i] = STATE_QUADRATIC;
    } else state[i] = (force[i] <= 0) ? STATE_SATISFIED : STATE_QUADRATIC;
  }
  MJB_PSYNC();
}

// ---- constraint islands (mj_island, engine_island.c:455-650) ----------------------------------------
// Trees coupled by a constraint row belong to one island; islands are numbered by their smallest tree,
// rows keep their global order inside an island.  Union-find over the (few) kinematic trees, serial on
// lane 0: unionConstraintTrees :359-455 with the tree iterator :205-330 (friction dof / joint limit: the
// dof's tree; contact: the two geoms' body trees, -1 = static; tendon rows: scan of the dense Jacobian
// row), mj_dsuMerge/mj_dsuRoot/mj_dsuAssign :84-160.
MJB_HD int dsu_root(int* parent, int tree) {
  int root = tree;
  while (parent[root] != root) root = parent[root];
  while (parent[tree] != tree) { const int next = parent[tree]; parent[tree] = root; tree = next; }
  return root;
}
MJB_HD void dsu_merge(int* parent, int t1, int t2) {
  if (t1 == -1 && t2 == -1) return;
  if (t1 == -1) t1 = t2;
  if (t2 == -1) t2 = t1;
  if (parent[t1] == -1) parent[t1] = t1;
  if (parent[t2] == -1) parent[t2] = t2;
  if (parent[t1] == parent[t2]) return;
  const int r1 = dsu_root(parent, t1), r2 = dsu_root(parent, t2);
  if (r1 < r2) parent[r2] = r1;
  else if (r2 < r1) parent[r1] = r2;
}

// islands are used when the model has several trees and mjDSBL_ISLAND is not set (a single tree is one
// island that coincides with the monolithic problem)
MJB_HD bool use_islands(const Env& d) { return (d.feat & FEAT_ISLAND) && d.m.sz.ntree > 1 && !(d.m.opt.disableflags & DSBL_ISLAND); }

MJB_HD void make_islands(const Env& d) {
  const DModel& m = d.m;
  if (!use_islands(d)) return;
  const int nefc = d.nefc()[0], ntree = m.sz.ntree, nv = m.sz.nv;
  MJB_LANE0 {
    int* parent = d.tree_island().p + ntree + 1;
    int* tisl = d.tree_island().p;
    int* eisl = d.efc_island().p;        // first: the row's tree, then its island
    FI type = d.efc_type(), id = d.efc_id(), cg1 = d.con_geom1(), cg2 = d.con_geom2();
    FD J = d.efc_J();
    for (int t = 0; t < ntree; t++) parent[t] = -1;
    int ptype = -1, pid = -1;
    for (int i = 0; i < nefc; i++) {
      if (i > 0 && ptype == type[i] && pid == id[i]) { eisl[i] = eisl[i - 1]; continue; }
      ptype = type[i]; pid = id[i];
      int t1 = -2, t2 = -2;
      bool scan = false;
      if (ptype == CNSTR_EQUALITY && m.eq_kind[pid] >= EQ_CONNECT) {
        t1 = m.body_treeid[m.eq_obj1id[pid]];
        t2 = m.body_treeid[m.eq_obj2id[pid]];
      }
      else if (ptype == CNSTR_FRICTION_DOF) t1 = m.dof_treeid[pid];
      else if (ptype == CNSTR_LIMIT_JOINT) t1 = m.dof_treeid[m.jnt_dofadr[pid]];
      else if (ptype == CNSTR_CONTACT_PYRAMIDAL || ptype == CNSTR_CONTACT_FRICTIONLESS) {
        t1 = m.body_treeid[m.geom_bodyid[cg1[pid]]];
        t2 = m.body_treeid[m.geom_bodyid[cg2[pid]]];
      } else scan = true;
      if (!scan) {
        eisl[i] = t1 >= 0 ? t1 : t2;
        if (t2 == -2) dsu_merge(parent, t1, -1);
        else dsu_merge(parent, t1, t2);
      } else {   // tendon rows: the trees of the nonzero Jacobian entries, in dof order
        int first = -2, prev = -1;
        for (int j = 0; j < nv; j++) {
          if (J[(long)i * nv + j] != 0) {
            const int tj = m.dof_treeid[j];
            if (tj == prev) continue;
            if (first == -2) { first = tj; dsu_merge(parent, tj, -1); }
            else dsu_merge(parent, prev, tj);
            prev = tj;
          }
        }
        eisl[i] = first;
      }
    }
    int nisland = 0;
    for (int t = 0; t < ntree; t++) {
      if (parent[t] == -1) { tisl[t] = -1; continue; }
      if (parent[t] == t) tisl[t] = nisland++;
      else { parent[t] = parent[parent[t]]; tisl[t] = tisl[parent[t]]; }
    }
    d.nisland()[0] = nisland;
    int* adr = d.island_iefcadr().p;
    int* map = d.map_iefc2efc().p;
    for (int k = 0; k <= nisland; k++) adr[k] = 0;
    for (int i = 0; i < nefc; i++) { eisl[i] = tisl[eisl[i]]; adr[eisl[i] + 1]++; }
    for (int k = 0; k < nisland; k++) adr[k + 1] += adr[k];
    int* fill = parent;   // reuse as per-island fill counters
    for (int k = 0; k < nisland; k++) fill[k] = 0;
    for (int i = 0; i < nefc; i++) { const int k = eisl[i]; map[adr[k] + fill[k]++] = i; }
    // dofs: island order = global order inside each island; unconstrained dofs go last
    int* dadr = d.island_idofadr().p;
    int* i2d = d.map_idof2dof().p;
    int* d2i = d.map_dof2idof().p;
    for (int k = 0; k <= nisland + 1; k++) dadr[k] = 0;
    for (int j = 0; j < nv; j++) { const int k = tisl[m.dof_treeid[j]]; dadr[(k >= 0 ? k : nisland) + 1]++; }
    for (int k = 0; k <= nisland; k++) dadr[k + 1] += dadr[k];
    for (int k = 0; k <= nisland; k++) fill[k] = 0;
    for (int j = 0; j < nv; j++) {
      int k = tisl[m.dof_treeid[j]];
      if (k < 0) k = nisland;
      const int idof = dadr[k] + fill[k]++;
      i2d[idof] = j; d2i[j] = idof;
    }
  }
  MJB_PSYNC();
}

// ---- projected Gauss-Seidel on the dual ------------------------------------------------------------
// The sweep is a Gauss-Seidel recurrence: row i needs every earlier update of the same sweep, so
// rows are visited one after the other (PCG32 Fisher-Yates order of the reference).  Inside a row
// the residual dot product keeps mju_dot's exact accumulation structure — four independent
// stride-4 partial sums combined as (r0+r2)+(r1+r3), then the grouped tail — and those four
// chains run on four lanes; lane 0 combines them, applies the projected update and the cost-change
// guard.  Vector updates between sweeps (Nesterov extrapolation, projection, dual state) are
// spread over all lanes.  In the fused kernel AR and the sweep vectors are first copied into the
// warp's shared-memory scratch (when nefc^2 + 7 nefc + 8 doubles fit), so the serial chain never
// waits on L2.  Results are bit-identical to the serial reference arithmetic.
// sweep iterations, specialised on where AR rows come from so that every pointer has a single
// provenance (shared vs global) and the compiler emits LDS/STS for the on-chip data:
//   MODE 2: AR and vectors in the warp's shared-memory scratch
//   MODE 1: vectors on chip; AR rows streamed from L2 through a 4-slot ring of row buffers (each row is
//           requested four rows ahead of its use, its position in the sweep being known from the shuffle)
//   MODE 0: everything in global memory (host emulation, lane-per-env mapping, oversized problems)
// generic sweeps (any lane count, everything in global memory): host emulation, lane-per-env and
// sub-warp mappings, oversized problems.  rows/nrow: the island's rows in island order (NULL = all rows).
MJB_HD int pgs_sweeps(const Env& d, int nefc, int ne, int nf, const int* rows, int nrow, const double* gAR,
                      double* force, const double* b, const double* floss, const double* ARinv, double* fprev,
                      double* fmom, const double* Adiag, double* shared, int* order) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  const int n4 = nefc & ~3, tail = nefc - n4;
  const double scale = 1 / (m.opt.meaninertia * (nv > 1 ? nv : 1));
  auto row_of = [&](int c) { return rows ? rows[c] : c; };
  MJB_PFOR(c, nrow) { const int i = row_of(c); order[c] = i; fprev[i] = force[i]; }
  MJB_PSYNC();
  Pcg32 rng{0, 1};
  pcg32_next(rng);
  int iter = 0, nk = 0;
  const int maxiter = m.opt.iterations;
  while (iter < maxiter) {
    double beta = 0;
    if (iter > 0) beta = (double)(nk - 1) / (double)(nk + 2);
    if (beta > 0) {
      MJB_PFOR(c, nrow) {
        const int i = row_of(c);
        const double fs = force[i];
        double f = fs + beta * (fs - fprev[i]);
        fprev[i] = fs;
        if (i < ne) {}
        else if (i < ne + nf) f = dclip(f, -floss[i], floss[i]);
        else if (f < 0) f = 0;
        force[i] = f;
        fmom[i] = f;
      }
    } else {
      MJB_PFOR(c, nrow) { const int i = row_of(c); fprev[i] = force[i]; fmom[i] = force[i]; }
    }
    MJB_LANE0 {
      for (int c = nrow - 1; c > 0; c--) {   // Fisher-Yates, same draws as the reference
        const uint32_t j = pcg32_next(rng) % (uint32_t)(c + 1);
        const int t = order[c]; order[c] = order[j]; order[j] = t;
      }
    }
    MJB_PSYNC();
    double impr = 0;   // meaningful on lane 0
    for (int bi = 0; bi < nrow; bi++) {
      const int i = order[bi];
      const double* row = gAR + (long)i * nefc;
      // mju_dot structure: four stride-4 partial sums, combined as (r0+r2)+(r1+r3)
      for (int k = d.lane; k < 4; k += d.nlane) {
        double r = 0;
        for (int c = k; c < n4; c += 4) r += row[c] * force[c];
        shared[1 + k] = r;
      }
      MJB_PSYNC();
      MJB_LANE0 {
        double res = (shared[1] + shared[3]) + (shared[2] + shared[4]);
        if (tail == 3) res += row[n4] * force[n4] + row[n4 + 1] * force[n4 + 1] + row[n4 + 2] * force[n4 + 2];
        else if (tail == 2) res += row[n4] * force[n4] + row[n4 + 1] * force[n4 + 1];
        else if (tail == 1) res += row[n4] * force[n4];
        res = b[i] + res;
        const double old = force[i];
        double f = old - res * ARinv[i];
        if (i < ne) {}
        else if (i < ne + nf) {
          if (f < -floss[i]) f = -floss[i];
          else if (f > floss[i]) f = floss[i];
        } else if (f < 0) f = 0;
        const double delta = f - old;
        double change = 0.5 * delta * delta * Adiag[i] + delta * res;
        if (change > 1e-10) { f = old; change = 0; }
        force[i] = f;
        impr -= change;
      }
      MJB_PSYNC();
    }
    MJB_LANE0 shared[0] = impr * scale;
    MJB_PSYNC();
    const double improvement = shared[0];
    bool restart = false;
    if (iter > 0) {   // every lane evaluates the same serial sum: uniform restart decision
      double dce = 0;
      for (int c = 0; c < nrow; c++) { const int i = row_of(c); dce += (force[i] - fmom[i]) * (fmom[i] - fprev[i]); }
      restart = dce < 0;
    }
    MJB_PSYNC();
    if (restart) nk = 0; else nk++;
    iter++;
    if (improvement < m.opt.tolerance) break;
  }
  return iter;
}

#if defined(__CUDACC__)
// PCG32 skip-ahead (state after `delta` < 128 draws; increment 1): lets every lane produce its own
// draw of a sweep, so the LCG and the modulo leave the serial Fisher-Yates chain
__device__ inline uint64_t pcg32_skip(uint64_t state, uint32_t delta) {
  uint64_t cur_mult = 6364136223846793005ULL, cur_plus = 1, acc_mult = 1, acc_plus = 0;
#pragma unroll
  for (int bit = 0; bit < 7; bit++) {
    if (delta & 1u) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}

// Warp-per-env PGS sweeps with all sweep data on chip (MODE 2: AR in shared memory, MODE 1: AR rows
// through the 4-slot ring).  Same arithmetic as pgs_sweeps, organised so that the row body is
// uniform across the warp — every lane runs chain (lane & 3) of the mju_dot structure, an xor
// butterfly gives every lane the identical (r0+r2)+(r1+r3) (IEEE addition commutes), and every lane
// evaluates the projected update redundantly — which removes all divergence/reconvergence from the
// serial chain.  lo/hi are the projection bounds (-floss/floss for friction rows, 0/inf otherwise).
template <int MODE, bool ISL>
__device__ int pgs_sweeps_warp(const Env& d, int nefc, int nf, const int* rows, int nrow, const double* __restrict__ gAR, const double* AR,
                               double* ring, double* force, const double* b, const double* lo, const double* hi,
                               const double* ARinv, double* fprev, double* fmom, const double* Adiag,
                               int* order, int* jdraw) {
  const unsigned full = 0xffffffffu;
  const DModel& m = d.m;
  const int nv = m.sz.nv, lane = d.lane, k = lane & 3;
  const int n4 = nefc & ~3, tail = nefc - n4, nq = n4 >> 2;
  const double scale = 1 / (m.opt.meaninertia * (nv > 1 ? nv : 1));
  auto row_of = [&](int c) { return ISL ? rows[c] : c; };
  for (int c = lane; c < nrow; c += 32) { const int i = row_of(c); order[c] = i; fprev[i] = force[i]; }
  __syncwarp();
  Pcg32 rng{0, 1};
  pcg32_next(rng);
  uint64_t s0 = rng.state;
  int iter = 0, nk = 0;
  const int maxiter = m.opt.iterations;
  while (iter < maxiter) {
    double beta = 0;
    if (iter > 0) beta = (double)(nk - 1) / (double)(nk + 2);
    if (beta > 0) {
      for (int c = lane; c < nrow; c += 32) {
        const int i = row_of(c);
        const double fs = force[i];
        double f = fs + beta * (fs - fprev[i]);
        fprev[i] = fs;
        f = dclip(f, lo[i], hi[i]);
        force[i] = f;
        fmom[i] = f;
      }
    } else {
      for (int c = lane; c < nrow; c += 32) { const int i = row_of(c); fprev[i] = force[i]; fmom[i] = force[i]; }
    }
    // Fisher-Yates draws of this sweep (draw t serves position i = nefc-1-t), one per lane
    for (int t = lane; t < nrow - 1; t += 32) {
      const uint64_t old = pcg32_skip(s0, (uint32_t)t);
      const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
      const uint32_t out = (xs >> rot) | (xs << ((0u - rot) & 31));
      const int i = nrow - 1 - t;
      jdraw[i] = (int)(out % (uint32_t)(i + 1));
    }
    s0 = pcg32_skip(s0, (uint32_t)(nrow - 1));
    __syncwarp();
    if (lane == 0) {
      for (int i = nrow - 1; i > 0; i--) {
        const int j = jdraw[i];
        const int t = order[i]; order[i] = order[j]; order[j] = t;
      }
    }
    __syncwarp();
    double p0 = 0, p1 = 0;
    if (MODE == 1) {
      for (int q = 0; q < 3 && q < nrow; q++) {
        const double* g = gAR + order[q] * nefc;
        for (int c = lane; c < nefc; c += 32) ring[q * nefc + c] = g[c];
      }
      if (3 < nrow) {
        const double* g = gAR + order[3] * nefc;
        if (lane < nefc) p0 = g[lane];
        if (lane + 32 < nefc) p1 = g[lane + 32];
      }
      __syncwarp();
    }
    double impr = 0;
    int inext = order[0];
    const double* glane = gAR + lane;
    asm volatile("" : "+l"(glane));   // keep the lane's AR base in registers (no per-row re-derivation)
    for (int bi = 0; bi < nrow; bi++) {
      const int i = inext;
      inext = order[bi + 1 < nrow ? bi + 1 : bi];
      const double* row = (MODE == 2) ? AR + i * nefc : ring + (bi & 3) * nefc;
      const double bi_ = b[i], ainv = ARinv[i], ad = Adiag[i], l = lo[i], h = hi[i], old = force[i];
      double r = 0;
      {   // chain k of the mju_dot structure: nq = n4/4 terms for every lane (uniform trip count)
        const double* rk = row + k;
        const double* fk = force + k;
#pragma unroll 4
        for (int q = 0; q < nq; q++) r += rk[4 * q] * fk[4 * q];
      }
      const double v = r + __shfl_xor_sync(full, r, 2);
      double res = v + __shfl_xor_sync(full, v, 1);
      if (tail) {
        double t = row[n4] * force[n4];
        if (tail >= 2) t += row[n4 + 1] * force[n4 + 1];
        if (tail == 3) t += row[n4 + 2] * force[n4 + 2];
        res += t;
      }
      res = bi_ + res;
      double f = old - res * ainv;
      f = f < l ? l : (f > h ? h : f);
      const double delta = f - old;
      double change = 0.5 * delta * delta * ad + delta * res;
      if (change > 1e-10) { f = old; change = 0; }
      if (lane == 0) force[i] = f;
      impr -= change;
      if (MODE == 1 && bi + 3 < nrow) {
        double* dst = ring + ((bi + 3) & 3) * nefc;
        if (lane < nefc) dst[lane] = p0;
        if (lane + 32 < nefc) dst[lane + 32] = p1;
        if (bi + 4 < nrow) {
          const double* g = glane + order[bi + 4] * nefc;
          if (lane < nefc) p0 = g[0];
          if (lane + 32 < nefc) p1 = g[32];
        }
      }
      __syncwarp();
    }
    const double improvement = impr * scale;
    bool restart = false;
    if (iter > 0) {   // serial-order sum of the per-row products, products formed one per lane
      double q0 = 0, q1 = 0;
      if (lane < nefc) q0 = (force[lane] - fmom[lane]) * (fmom[lane] - fprev[lane]);
      if (lane + 32 < nefc) q1 = (force[lane + 32] - fmom[lane + 32]) * (fmom[lane + 32] - fprev[lane + 32]);
      double dce = 0;
      if (ISL) {
        for (int c = 0; c < nrow; c++) {
          const int i = row_of(c);
          dce += __shfl_sync(full, i < 32 ? q0 : q1, i & 31);
        }
      } else {
        const int nlo = nefc < 32 ? nefc : 32;
        for (int i = 0; i < nlo; i++) dce += __shfl_sync(full, q0, i);
        for (int i = 32; i < nefc; i++) dce += __shfl_sync(full, q1, i - 32);
      }
      restart = dce < 0;
    }
    if (restart) nk = 0; else nk++;
    iter++;
    if (improvement < m.opt.tolerance) break;
  }
  return iter;
}
#endif

MJB_HD void solve_pgs(const Env& d) {
  const DModel& m = d.m;
  const int nefc = d.nefc()[0], ne = d.ne()[0], nf = d.nf()[0], njmax = m.sz.njmax;
  if (!nefc) return;
  const double* gAR = d.efc_AR().p;
  // islands (mj_fwdConstraint, engine_forward.c:1187-1194): one independent PGS solve per island over the
  // island's rows, each with its own shuffle stream, momentum state, iteration count and termination; the
  // residual of a row is still the dot with the FULL force vector (entries of other islands meet exact zeros)
  const bool isl = use_islands(d);
  const int nsolve = isl ? d.nisland()[0] : 1;
  const int* iadr = d.island_iefcadr().p;
  const int* imap = d.map_iefc2efc().p;
  FI niter = d.solver_niter();
#if defined(__CUDA_ARCH__)
  int mode = 0;
  const int nord = (nefc + 1) / 2;   // doubles that hold nefc ints
  if (d.sm && d.nlane == 32 && nefc <= 64) {
    if ((long)nefc * nefc + 8L * nefc + 2 * nord <= d.smcap) mode = 2;
    else if (12L * nefc + 2 * nord <= d.smcap) mode = 1;
  }
  if (mode) {
    double* v = d.sm;
    double* AR = nullptr; double* ring = nullptr;
    if (mode == 2) { AR = v; v += nefc * nefc; } else { ring = v; v += 4 * nefc; }
    double* force = v; double* b = force + nefc; double* lo = b + nefc; double* hi = lo + nefc; double* ARinv = hi + nefc;
    double* fprev = ARinv + nefc; double* fmom = fprev + nefc; double* Adiag = fmom + nefc;
    int* order = (int*)(Adiag + nefc);
    int* jdraw = order + 2 * nord;
    const double* gf = d.efc_force().p; const double* gb = d.efc_b().p; const double* gfl = d.efc_frictionloss().p;
    if (mode == 2) { MJB_PFOR(i, nefc * nefc) AR[i] = gAR[i]; }
    MJB_PFOR(i, nefc) {
      force[i] = gf[i]; b[i] = gb[i];
      const double fl = gfl[i];
      lo[i] = (i < ne) ? -HUGE_VAL : (i < ne + nf) ? -fl : 0.0;     // equality rows are unbounded
      hi[i] = (i >= ne && i < ne + nf) ? fl : HUGE_VAL;
      const double ai = 1 / gAR[(long)i * (nefc + 1)];
      ARinv[i] = ai;
      Adiag[i] = 1 / ai;    // the reference's Athis[0] = 1/ARinv
    }
    MJB_PSYNC();
    for (int k = 0; k < nsolve; k++) {
      const int* rows = isl ? imap + iadr[k] : nullptr;
      const int nrow = isl ? iadr[k + 1] - iadr[k] : nefc;
      int iter;
      if (isl) {
        if (mode == 2) iter = pgs_sweeps_warp<2, true>(d, nefc, nf, rows, nrow, gAR, AR, ring, force, b, lo, hi, ARinv, fprev, fmom, Adiag, order, jdraw);
        else iter = pgs_sweeps_warp<1, true>(d, nefc, nf, rows, nrow, gAR, AR, ring, force, b, lo, hi, ARinv, fprev, fmom, Adiag, order, jdraw);
      } else {
        if (mode == 2) iter = pgs_sweeps_warp<2, false>(d, nefc, nf, rows, nrow, gAR, AR, ring, force, b, lo, hi, ARinv, fprev, fmom, Adiag, order, jdraw);
        else iter = pgs_sweeps_warp<1, false>(d, nefc, nf, rows, nrow, gAR, AR, ring, force, b, lo, hi, ARinv, fprev, fmom, Adiag, order, jdraw);
      }
      MJB_PSYNC();
      MJB_LANE0 if (k < NISLAND) niter[k] += iter;
    }
    MJB_PSYNC();
    double* gfo = d.efc_force().p;
    MJB_PFOR(i, nefc) gfo[i] = force[i];
    MJB_PSYNC();
    dual_state_ptr(d, gfo, gfl, nefc, ne, nf);
  } else
#endif
  {
    double* force = d.efc_force().p; const double* b = d.efc_b().p; const double* floss = d.efc_frictionloss().p;
    double* scr = d.scr_efc().p;
    double* ARinv = scr; double* fprev = scr + njmax; double* fmom = scr + 2 * (long)njmax; double* Adiag = scr + 3 * (long)njmax;
    double* shared = scr + 4 * (long)njmax;
    int* order = d.scr_int().p + njmax;
    MJB_PFOR(i, nefc) {
      const double ai = 1 / gAR[(long)i * (nefc + 1)];
      ARinv[i] = ai;
      Adiag[i] = 1 / ai;
    }
    MJB_PSYNC();
    for (int k = 0; k < nsolve; k++) {
      const int* rows = isl ? imap + iadr[k] : nullptr;
      const int nrow = isl ? iadr[k + 1] - iadr[k] : nefc;
      const int iter = pgs_sweeps(d, nefc, ne, nf, rows, nrow, gAR, force, b, floss, ARinv, fprev, fmom, Adiag, shared, order);
      MJB_PSYNC();
      MJB_LANE0 if (k < NISLAND) niter[k] += iter;
    }
    MJB_PSYNC();
    dual_state_ptr(d, force, floss, nefc, ne, nf);
  }
  MJB_PSYNC();
}

// ---- NoSlip post-solver (solNoSlip, engine_solver.c:767-957): Gauss-Seidel sweeps over the friction rows with the
// regulariser R taken out of AR (flg_subR), run after the main solver when opt.noslip_iterations > 0 - dry-friction
// rows one at a time, pyramidal contacts one pair of opposing edges at a time (a 2x2 problem along f0 + f1 = const).
// The rows of one solve are visited in their stored order and every scalar (residuals through mju_dot, improvement)
// accumulates in the reference's order, so the sweep is serial: lane 0 runs it.  Per island when islands are on.
MJB_HD void solve_noslip(const Env& d) {
  const DModel& m = d.m;
  const int nefc = d.nefc()[0], ne = d.ne()[0], nf = d.nf()[0];
  const int maxiter = m.opt.noslip_iterations;
  if (!nefc || maxiter <= 0) return;
  double* force = d.efc_force().p;
  const double* floss = d.efc_frictionloss().p;
  dual_state_ptr(d, force, floss, nefc, ne, nf);
  MJB_LANE0 {
    const double* AR = d.efc_AR().p; const double* R = d.efc_R().p; const double* b = d.efc_b().p;
    const int* type = d.efc_type().p; const int* id = d.efc_id().p; const int* cdim = d.con_dim().p;
    const bool isl = use_islands(d);
    const int nsolve = isl ? d.nisland()[0] : 1;
    const int* iadr = d.island_iefcadr().p;
    const int* imap = d.map_iefc2efc().p;
    int* niter = d.solver_niter().p;
    const double scale = 1 / (m.opt.meaninertia * (m.sz.nv > 1 ? m.sz.nv : 1));
    auto resid = [&](int i) { return (b[i] + dot_ptr(AR + (long)i * nefc, force, nefc)) - R[i] * force[i]; };
    for (int k = 0; k < nsolve; k++) {
      const int* rows = isl ? imap + iadr[k] : nullptr;
      const int nrow = isl ? iadr[k + 1] - iadr[k] : nefc;
      int iter = 0;
      while (iter < maxiter) {
        double improvement = 0;
        if (iter == 0)
          for (int c = 0; c < nrow; c++) { const int i = rows ? rows[c] : c; improvement += 0.5 * force[i] * force[i] * R[i]; }
        // dry friction
        for (int c = 0; c < nrow; c++) {
          const int i = rows ? rows[c] : c;
          if (i < ne || i >= ne + nf) continue;
          const double ainv = 1 / fmax(kMinVal, AR[(long)i * (nefc + 1)] - R[i]);
          const double res = resid(i), old = force[i];
          double f = old - res * ainv;
          if (f < -floss[i]) f = -floss[i];
          else if (f > floss[i]) f = floss[i];
          force[i] = f;
          const double delta = f - old;
          improvement -= 0.5 * delta * delta / ainv + delta * res;
        }
        // contact friction: pairs of opposing pyramid edges
        for (int c = 0; c < nrow; c++) {
          const int i = rows ? rows[c] : c;
          if (i < ne + nf || type[i] != CNSTR_CONTACT_PYRAMIDAL) continue;
          const int dim = cdim[id[i]];
          for (int j = i; j < i + 2 * (dim - 1); j += 2) {
            const double res0 = resid(j), res1 = resid(j + 1);
            const double old0 = force[j], old1 = force[j + 1];
            double A00 = AR[(long)j * nefc + j], A01 = AR[(long)j * nefc + j + 1];
            double A10 = AR[(long)(j + 1) * nefc + j], A11 = AR[(long)(j + 1) * nefc + j + 1];
            A00 = fmax(1e-10, A00 - R[j]);
            A11 = fmax(1e-10, A11 - R[j + 1]);
            const double Ac[4] = {A00, A01, A10, A11}, old[2] = {old0, old1}, rs[2] = {res0, res1};
            const double bc0 = res0 - dot_ptr(Ac, old, 2), bc1 = res1 - dot_ptr(Ac + 2, old, 2);
            const double mid = 0.5 * (old0 + old1);
            const double K1 = A00 + A11 - A01 - A10;
            const double K0 = mid * (A00 - A11) + bc0 - bc1;
            double f0, f1;
            if (K1 < kMinVal) { f0 = mid; f1 = mid; }
            else {
              const double y = -K0 / K1;
              if (y < -mid) { f0 = 0; f1 = 2 * mid; }
              else if (y > mid) { f0 = 2 * mid; f1 = 0; }
              else { f0 = mid + y; f1 = mid - y; }
            }
            // costChange (engine_solver.c:206-230): 0.5 delta' A delta + delta . res, in mju_mulVecMatVec / mju_dot order;
            // a positive change restores the pair
            const double dl[2] = {f0 - old0, f1 - old1};
            double q = 0;
            q += dl[0] * dot_ptr(Ac, dl, 2);
            q += dl[1] * dot_ptr(Ac + 2, dl, 2);
            double change = 0.5 * q + dot_ptr(dl, rs, 2);
            if (change > 1e-10) { f0 = old0; f1 = old1; change = 0; }
            force[j] = f0; force[j + 1] = f1;
            improvement -= change;
          }
          c += 2 * (dim - 1) - 1;
        }
        improvement *= scale;
        iter++;
        if (improvement < m.opt.noslip_tolerance) break;
      }
      if (k < NISLAND) niter[k] += iter;
    }
  }
  MJB_PSYNC();
  dual_state_ptr(d, force, floss, nefc, ne, nf);
}

// dual finish: qfrc_constraint = J' f, qacc = qacc_smooth + M^-1 qfrc_constraint
MJB_HD void dual_finish(const Env& d) {
  const int nv = d.m.sz.nv;
  if (!d.nefc()[0]) return;
  FD qfc = d.qfrc_constraint(), qacc = d.qacc(), qas = d.qacc_smooth();
  mul_jacT_vec(d, qfc, d.efc_force());
  MJB_PFOR(i, nv) qacc[i] = qfc[i];
  MJB_PSYNC();
  solve_LD(d, qacc, d.qLD(), d.qLDiagInv());
  MJB_PFOR(i, nv) qacc[i] += qas[i];
  MJB_PSYNC();
}

}  // namespace mjb
