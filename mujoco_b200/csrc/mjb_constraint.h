// Constraint assembly and the PGS solver of the batched mj_step path, one environment per call.
//
// Replaces (reference file:line)
//   src/engine/engine_core_constraint.c  mj_instantiateFriction :1270-1355, mj_instantiateLimit
//   :1360-1531, mj_instantiateContact :1617-1713 (dense rows, pyramidal cones), mj_diagApprox
//   :1719-1973, mj_makeImpedance :2151-2265, mj_makeY/mj_makeAR (dense) :2918-3102,
//   mj_referenceConstraint :3245-3268, mj_constraintUpdate_impl :3275-3464;
//   src/engine/engine_core_util.c mj_jac :176-226;
//   src/engine/engine_forward.c warmstart :1056-1132, mj_fwdConstraint :1148-1252;
//   src/engine/engine_solver.c solPGS :457-741, dualState :270-353, dualFinish :72-79.
// Row order (friction loss, joint limits lower/upper, tendon limits, contacts; pyramid edge
// pairs +mu,-mu) and the operation order of every reduction follow the reference.
#pragma once
#include "mjb_smooth.h"

namespace mjb {

// append one constraint row; jacobian row must already be written at efc_J[nefc*nv ...]
// returns false when the per-env cap is hit (mjWARN_CNSTRFULL)
MJB_HD bool add_row(const Env& d, int& nefc, double pos, double margin, double floss, int type, int id) {
  d.efc_pos()[nefc] = pos;
  d.efc_margin()[nefc] = margin;
  d.efc_frictionloss()[nefc] = floss;
  d.efc_type()[nefc] = type;
  d.efc_id()[nefc] = id;
  nefc++;
  return true;
}

// dense point Jacobian of `body` at world point `pt`, translational part, written with a sign
// into three strided rows: jr[k*nv + dof] (+)= sign * value.  (mj_jac)
MJB_HD void jac_point(const Env& d, FD jr, V3 pt, int body, double sign, bool accumulate) {
  const DModel& m = d.m;
  const int nv = m.sz.nv;
  FD cdof = d.cdof(), sc = d.subtree_com();
  V3 off = pt - ld3(sc, 3 * m.body_rootid[body]);
  if (!accumulate) for (int i = 0; i < 3 * nv; i++) jr[i] = 0;
  body = m.body_weldid[body];
  if (m.body_dofnum[body] == 0) return;
  int i = m.body_dofadr[body] + m.body_dofnum[body] - 1;
  while (i >= 0) {
    V3 w = ld3(cdof, 6 * i), v = ld3(cdof, 6 * i + 3);
    V3 t = cross(w, off);
    V3 val{v.x + t.x, v.y + t.y, v.z + t.z};
    if (accumulate) {
      jr[i] = jr[i] + sign * val.x; jr[nv + i] = jr[nv + i] + sign * val.y; jr[2 * nv + i] = jr[2 * nv + i] + sign * val.z;
    } else {
      jr[i] = val.x; jr[nv + i] = val.y; jr[2 * nv + i] = val.z;
    }
    i = m.dof_parentid[i];
  }
}

MJB_HD void make_constraint(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv, njmax = m.sz.njmax;
  FI nefc_f = d.nefc(), ne_f = d.ne(), nf_f = d.nf(), nl_f = d.nl();
  int nefc = 0, nf = 0, nl = 0;
  ne_f[0] = 0; nf_f[0] = 0; nl_f[0] = 0; nefc_f[0] = 0;
  if (m.opt.disableflags & DSBL_CONSTRAINT) return;
  FD J = d.efc_J(), qpos = d.qpos();
  bool full = false;

  // ---- dof friction loss
  if (m.opt.has_frictionloss && !(m.opt.disableflags & DSBL_FRICTIONLOSS)) {
    for (int i = 0; i < nv && !full; i++) {
      const double fl = m.dof_frictionloss[i];
      if (!fl) continue;
      if (nefc + 1 > njmax) { full = true; break; }
      FD row = J + (long)nefc * nv;
      for (int c = 0; c < nv; c++) row[c] = 0;
      row[i] = 1;
      add_row(d, nefc, 0, 0, fl, CNSTR_FRICTION_DOF, i);
      nf++;
    }
  }

  // ---- joint and tendon limits
  if (m.opt.has_limits && !(m.opt.disableflags & DSBL_LIMIT)) {
    for (int i = 0; i < m.sz.njnt && !full; i++) {
      if (!m.jnt_limited[i]) continue;
      const double margin = m.jnt_margin[i];
      const int jt = m.jnt_type[i];
      if (jt == JNT_SLIDE || jt == JNT_HINGE) {
        const double value = qpos[m.jnt_qposadr[i]];
        for (int side = -1; side <= 1; side += 2) {
          const double dist = side * (m.jnt_range[2 * i + (side + 1) / 2] - value);
          if (dist < margin) {
            if (nefc + 1 > njmax) { full = true; break; }
            FD row = J + (long)nefc * nv;
            for (int c = 0; c < nv; c++) row[c] = 0;
            row[m.jnt_dofadr[i]] = -(double)side;
            add_row(d, nefc, dist, margin, 0, CNSTR_LIMIT_JOINT, i);
            nl++;
          }
        }
      } else if (jt == JNT_BALL) {
        Q4 q = ld4(qpos, m.jnt_qposadr[i]);
        normalize(q);
        V3 aa = quat2vel(q, 1);
        const double value = normalize(aa);
        const double dist = dmax(m.jnt_range[2 * i], m.jnt_range[2 * i + 1]) - value;
        if (dist < margin) {
          if (nefc + 1 > njmax) { full = true; break; }
          FD row = J + (long)nefc * nv;
          for (int c = 0; c < nv; c++) row[c] = 0;
          const int da = m.jnt_dofadr[i];
          row[da] = aa.x * -1; row[da + 1] = aa.y * -1; row[da + 2] = aa.z * -1;
          add_row(d, nefc, dist, margin, 0, CNSTR_LIMIT_JOINT, i);
          nl++;
        }
      }
    }
    FD tl = d.ten_length(), tJ = d.ten_J();
    for (int i = 0; i < m.sz.ntendon && !full; i++) {
      if (!m.tendon_limited[i]) continue;
      const double value = tl[i], margin = m.tendon_margin[i];
      for (int side = -1; side <= 1; side += 2) {
        const double dist = side * (m.tendon_range[2 * i + (side + 1) / 2] - value);
        if (dist < margin) {
          if (nefc + 1 > njmax) { full = true; break; }
          FD row = J + (long)nefc * nv;
          for (int c = 0; c < nv; c++) row[c] = 0;
          const int adr = m.ten_J_rowadr[i], nnz = m.ten_J_rownnz[i];
          for (int a = 0; a < nnz; a++) row[m.ten_J_colind[adr + a]] = tJ[adr + a];
          for (int c = 0; c < nv; c++) row[c] = row[c] * (double)(-side);
          // rows with an identically zero Jacobian are dropped (mj_addConstraint "empty" guard)
          bool empty = true;
          for (int c = 0; c < nv; c++) if (row[c] != 0) { empty = false; break; }
          if (empty) continue;
          add_row(d, nefc, dist, margin, 0, CNSTR_LIMIT_TENDON, i);
          nl++;
        }
      }
    }
  }

  // ---- contacts (frictionless or pyramidal)
  const int ncon = d.ncon()[0];
  if (!(m.opt.disableflags & DSBL_CONTACT) && ncon && !full) {
    FD cpos = d.con_pos(), cframe = d.con_frame(), cdist = d.con_dist(), cinc = d.con_includemargin();
    FD cfri = d.con_friction();
    FI cg1 = d.con_geom1(), cg2 = d.con_geom2(), cdim = d.con_dim(), cexc = d.con_exclude(), cadr = d.con_efcadr();
    FD jd = d.scr_jac();          // jacdifp, 3 x nv
    FD jc = d.scr_jac() + 3 * nv; // rotated, 3 x nv
    for (int i = 0; i < ncon; i++) {
      if (cexc[i]) continue;
      const int dim = cdim[i];
      const int rows = (dim == 1) ? 1 : 2 * (dim - 1);
      if (nefc + rows > njmax) { full = true; break; }
      cadr[i] = nefc;
      V3 pt = ld3(cpos, 3 * i);
      const int b1 = m.geom_bodyid[cg1[i]], b2 = m.geom_bodyid[cg2[i]];
      // jacdif = jac(b2) - jac(b1): build jac2 then subtract jac1 entry-wise (both dense, zeros kept)
      jac_point(d, jd, pt, b2, 1.0, false);
      {
        // jac1 into jc (temporary), then jd = jd - jc
        jac_point(d, jc, pt, b1, 1.0, false);
        for (int k = 0; k < 3 * nv; k++) jd[k] = jd[k] - jc[k];
      }
      // rotate into the contact frame: jc[r] = sum_k frame[r][k] * jd[k]   (mju_mulMatMat order)
      const int nr = dim > 1 ? 3 : 1;
      for (int r = 0; r < nr; r++) {
        for (int c = 0; c < nv; c++) jc[r * nv + c] = 0;
        for (int k = 0; k < 3; k++) {
          const double f = cframe[9 * i + 3 * r + k];
          if (f != 0) for (int c = 0; c < nv; c++) jc[r * nv + c] += jd[k * nv + c] * f;
        }
      }
      const double dist = cdist[i], inc = cinc[i];
      if (dim == 1) {
        FD row = J + (long)nefc * nv;
        for (int c = 0; c < nv; c++) row[c] = jc[c];
        add_row(d, nefc, dist, inc, 0, CNSTR_CONTACT_FRICTIONLESS, i);
      } else {
        for (int k = 1; k < dim; k++) {
          const double mu = cfri[5 * i + k - 1];
          FD r0 = J + (long)nefc * nv, r1 = J + (long)(nefc + 1) * nv;
          for (int c = 0; c < nv; c++) {
            r0[c] = jc[c] + jc[k * nv + c] * mu;
            r1[c] = jc[c] + jc[k * nv + c] * (-mu);
          }
          add_row(d, nefc, dist, inc, 0, CNSTR_CONTACT_PYRAMIDAL, i);
          add_row(d, nefc, dist, inc, 0, CNSTR_CONTACT_PYRAMIDAL, i);
        }
      }
    }
  }
  if (full) d.warning()[WARN_CNSTRFULL] += 1;
  nefc_f[0] = nefc; nf_f[0] = nf; nl_f[0] = nl;
  if (!nefc) return;

  // ---- diagApprox
  FD dA = d.efc_diagA();
  FI type = d.efc_type(), id = d.efc_id();
  for (int i = 0; i < nefc; i++) {
    const int t = type[i], k = id[i];
    if (t == CNSTR_FRICTION_DOF) dA[i] = m.dof_invweight0[k];
    else if (t == CNSTR_LIMIT_JOINT) dA[i] = m.dof_invweight0[m.jnt_dofadr[k]];
    else if (t == CNSTR_LIMIT_TENDON) dA[i] = m.tendon_invweight0[k];
    else {
      const int b1 = m.geom_bodyid[d.con_geom1()[k]], b2 = m.geom_bodyid[d.con_geom2()[k]];
      double tran = 0, rot = 0;
      tran += m.body_invweight0[2 * b1] * 1.0; rot += m.body_invweight0[2 * b1 + 1] * 1.0;
      tran += m.body_invweight0[2 * b2] * 1.0; rot += m.body_invweight0[2 * b2 + 1] * 1.0;
      if (t == CNSTR_CONTACT_FRICTIONLESS) dA[i] = tran;
      else {
        const int dim = d.con_dim()[k];
        for (int j = 0; j < dim - 1; j++) {
          const double fri = d.con_friction()[5 * k + j];
          const double v = tran + fri * fri * (j < 2 ? tran : rot);
          dA[i + 2 * j] = v; dA[i + 2 * j + 1] = v;
        }
        i += 2 * dim - 3;
      }
    }
  }

  // ---- impedance: R, D, KBIP
  FD R = d.efc_R(), KBIP = d.efc_KBIP(), D = d.efc_D(), epos = d.efc_pos(), emargin = d.efc_margin();
  for (int i = 0; i < nefc; i++) {
    const int t = type[i], k = id[i];
    double solref[2], solimp[5];
    int dim = 1;
    if (t == CNSTR_FRICTION_DOF) { for (int j = 0; j < 2; j++) solref[j] = m.dof_solref[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.dof_solimp[5 * k + j]; }
    else if (t == CNSTR_LIMIT_JOINT) { for (int j = 0; j < 2; j++) solref[j] = m.jnt_solref[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.jnt_solimp[5 * k + j]; }
    else if (t == CNSTR_LIMIT_TENDON) { for (int j = 0; j < 2; j++) solref[j] = m.tendon_solref_lim[2 * k + j]; for (int j = 0; j < 5; j++) solimp[j] = m.tendon_solimp_lim[5 * k + j]; }
    else {
      for (int j = 0; j < 2; j++) solref[j] = d.con_solref()[2 * k + j];
      for (int j = 0; j < 5; j++) solimp[j] = d.con_solimp()[5 * k + j];
      if (t == CNSTR_CONTACT_PYRAMIDAL) dim = 2 * (d.con_dim()[k] - 1);
    }
    if ((solref[0] > 0) != (solref[1] > 0)) { solref[0] = 0.02; solref[1] = 1; }   // mj_defaultSolRefImp
    if (!(m.opt.disableflags & DSBL_REFSAFE) && solref[0] > 0) solref[0] = dmax(solref[0], 2 * m.opt.timestep);
    solimp[0] = dmin(kMaxImp, dmax(kMinImp, solimp[0]));
    solimp[1] = dmin(kMaxImp, dmax(kMinImp, solimp[1]));
    solimp[2] = dmax(0, solimp[2]);
    solimp[3] = dmin(kMaxImp, dmax(kMinImp, solimp[3]));
    solimp[4] = dmax(1, solimp[4]);
    // impedance at this row's violation
    double imp, impP;
    const double pos = epos[i], margin = emargin[i];
    if (solimp[0] == solimp[1] || solimp[2] <= kMinVal) {
      imp = 0.5 * (solimp[0] + solimp[1]); impP = 0;
    } else {
      double x = (pos - margin) / solimp[2], sgn = 1;
      if (x < 0) { x = -x; sgn = -1; }
      if (x >= 1 || x <= 0) { imp = (x >= 1 ? solimp[1] : solimp[0]); impP = 0; }
      else {
        double y, yP;
        const double pw = solimp[4];
        if (pw == 1) { y = x; yP = 1; }
        else if (x <= solimp[3]) {
          const double a = 1 / (pw - 1 == 1 ? solimp[3] : (pw - 1 == 2 ? solimp[3] * solimp[3] : pow(solimp[3], pw - 1)));
          const double xp = (pw == 2 ? x * x : pow(x, pw));
          const double xp1 = (pw - 1 == 1 ? x : (pw - 1 == 2 ? x * x : pow(x, pw - 1)));
          y = a * xp; yP = pw * a * xp1;
        } else {
          const double om = 1 - solimp[3];
          const double b = 1 / (pw - 1 == 1 ? om : (pw - 1 == 2 ? om * om : pow(om, pw - 1)));
          const double ox = 1 - x;
          const double xp = (pw == 2 ? ox * ox : pow(ox, pw));
          const double xp1 = (pw - 1 == 1 ? ox : (pw - 1 == 2 ? ox * ox : pow(ox, pw - 1)));
          y = 1 - b * xp; yP = pw * b * xp1;
        }
        imp = solimp[0] + y * (solimp[1] - solimp[0]);
        impP = yP * sgn * (solimp[1] - solimp[0]) / solimp[2];
      }
    }
    for (int j = 0; j < dim; j++) {
      R[i + j] = dmax(kMinVal, (1 - imp) * dA[i + j] / imp);
      double K, Bv;
      if (t == CNSTR_FRICTION_DOF) K = 0;
      else if (solref[0] > 0) K = 1 / dmax(kMinVal, solimp[1] * solimp[1] * solref[0] * solref[0] * solref[1] * solref[1]);
      else K = -solref[0] / dmax(kMinVal, solimp[1] * solimp[1]);
      if (solref[1] > 0) Bv = 2 / dmax(kMinVal, solimp[1] * solref[0]);
      else Bv = -solref[1] / dmax(kMinVal, solimp[1]);
      KBIP[4 * (i + j)] = K; KBIP[4 * (i + j) + 1] = Bv; KBIP[4 * (i + j) + 2] = imp; KBIP[4 * (i + j) + 3] = impP;
    }
    i += dim - 1;
  }
  // pyramidal contacts: common R matched to the elliptic model, contact mu
  for (int i = nf; i < nefc; i++) {
    if (type[i] == CNSTR_CONTACT_PYRAMIDAL) {
      const int k = id[i], dim = d.con_dim()[k];
      R[i + 1] = R[i] / dmax(kMinVal, m.opt.impratio);
      const double mu = d.con_friction()[5 * k] * sqrt(R[i + 1] / R[i]);
      d.con_mu()[k] = mu;
      const double Rpy = 2 * mu * mu * R[i];
      for (int j = 0; j < 2 * (dim - 1); j++) R[i + j] = Rpy;
      i += 2 * (dim - 1) - 1;
    }
  }
  for (int i = 0; i < nefc; i++) D[i] = 1 / R[i];
  for (int i = 0; i < nefc; i++) dA[i] = R[i] * KBIP[4 * i + 2] / (1 - KBIP[4 * i + 2]);
}

// ---- dense J * vec and J' * vec in the reference's accumulation order ----------------------------
MJB_HD void mul_jac_vec(const Env& d, FD res, FD vec) {
  const int nefc = d.nefc()[0], nv = d.m.sz.nv;
  FD J = d.efc_J();
  for (int r = 0; r < nefc; r++) {
    FD row = J + (long)r * nv;
    res[r] = dot_ref(nv, [&](int c) { return row[c]; }, [&](int c) { return vec[c]; });
  }
}
MJB_HD void mul_jacT_vec(const Env& d, FD res, FD vec) {
  const int nefc = d.nefc()[0], nv = d.m.sz.nv;
  if (!nefc) return;
  FD J = d.efc_J();
  for (int c = 0; c < nv; c++) res[c] = 0;
  for (int r = 0; r < nefc; r++) {
    const double s = vec[r];
    if (s != 0) { FD row = J + (long)r * nv; for (int c = 0; c < nv; c++) res[c] += row[c] * s; }
  }
}

// ---- dual projection: Y = J L^-T D^-1/2, AR = Y Y' + diag(R) (dense) ---------------------------
MJB_HD void project_constraint(const Env& d) {
  const DModel& m = d.m;
  const int nefc = d.nefc()[0], nv = m.sz.nv;
  if (!nefc || m.opt.solver != SOL_PGS) return;
  FD J = d.efc_J(), Y = d.efc_Y(), AR = d.efc_AR(), qLD = d.qLD(), R = d.efc_R();
  FD sq = d.scr_nv();
  for (int i = 0; i < nv; i++) sq[i] = 1 / sqrt(qLD[m.M_rowadr[i] + m.M_rownnz[i] - 1]);
  for (int r = 0; r < nefc; r++) {
    FD x = Y + (long)r * nv, src = J + (long)r * nv;
    for (int c = 0; c < nv; c++) x[c] = src[c];
    for (int i = nv - 1; i > 0; i--) {
      if (m.dof_simplenum[i]) continue;
      const double xi = x[i];
      if (xi != 0) {
        const int start = m.M_rowadr[i], end = start + m.M_rownnz[i] - 1;
        for (int adr = start; adr < end; adr++) x[m.M_colind[adr]] -= qLD[adr] * xi;
      }
    }
    for (int i = 0; i < nv; i++) x[i] *= sq[i];
  }
  // AR lower triangle: AR[i][c] = sum_j Y[c][j] * Y[i][j], j ascending, skipping Y[i][j] == 0
  for (int i = 0; i < nefc; i++) {
    FD yi = Y + (long)i * nv;
    for (int c = 0; c <= i; c++) {
      FD yc = Y + (long)c * nv;
      double s = 0;
      for (int j = 0; j < nv; j++) {
        const double t = yi[j];
        if (t != 0) s += yc[j] * t;
      }
      AR[(long)i * nefc + c] = s;
    }
  }
  for (int i = 0; i < nefc; i++)
    for (int c = i + 1; c < nefc; c++) AR[(long)i * nefc + c] = AR[(long)c * nefc + i];
  for (int r = 0; r < nefc; r++) AR[(long)r * (nefc + 1)] += R[r];
}

// ---- efc_vel, efc_aref ---------------------------------------------------------------------------
MJB_HD void reference_constraint(const Env& d) {
  const int nefc = d.nefc()[0];
  if (!nefc) return;
  FD vel = d.efc_vel(), aref = d.efc_aref(), KBIP = d.efc_KBIP(), pos = d.efc_pos(), margin = d.efc_margin();
  mul_jac_vec(d, vel, d.qvel());
  for (int i = 0; i < nefc; i++)
    aref[i] = -KBIP[4 * i + 1] * vel[i] - KBIP[4 * i] * KBIP[4 * i + 2] * (pos[i] - margin[i]);
}

// ---- primal constraint update (pyramidal / scalar rows): force, state, optional cost -------------
MJB_HD double constraint_update(const Env& d, FD jar, bool want_cost) {
  const int nefc = d.nefc()[0], nf = d.nf()[0];
  FD D = d.efc_D(), R = d.efc_R(), floss = d.efc_frictionloss(), force = d.efc_force();
  FI state = d.efc_state();
  double s = 0;
  for (int i = 0; i < nefc; i++) force[i] = -D[i] * jar[i];
  for (int i = 0; i < nefc; i++) {
    if (i < nf) {
      if (jar[i] <= -R[i] * floss[i]) {
        if (want_cost) s += -0.5 * R[i] * floss[i] * floss[i] - floss[i] * jar[i];
        force[i] = floss[i]; state[i] = STATE_LINEARNEG;
      } else if (jar[i] >= R[i] * floss[i]) {
        if (want_cost) s += -0.5 * R[i] * floss[i] * floss[i] + floss[i] * jar[i];
        force[i] = -floss[i]; state[i] = STATE_LINEARPOS;
      } else {
        if (want_cost) s += 0.5 * D[i] * jar[i] * jar[i];
        state[i] = STATE_QUADRATIC;
      }
      continue;
    }
    if (jar[i] >= 0) { force[i] = 0; state[i] = STATE_SATISFIED; }
    else { if (want_cost) s += 0.5 * D[i] * jar[i] * jar[i]; state[i] = STATE_QUADRATIC; }
  }
  mul_jacT_vec(d, d.qfrc_constraint(), force);
  return s;
}

// dual state from forces (engine_solver.c dualState, scalar rows); returns nactive
MJB_HD int dual_state(const Env& d) {
  const int nefc = d.nefc()[0], nf = d.nf()[0];
  FD force = d.efc_force(), floss = d.efc_frictionloss();
  FI state = d.efc_state();
  int nactive = nf;
  for (int i = 0; i < nf; i++) {
    if (force[i] <= -floss[i]) state[i] = STATE_LINEARPOS;
    else if (force[i] >= floss[i]) state[i] = STATE_LINEARNEG;
    else state[i] = STATE_QUADRATIC;
  }
  for (int i = nf; i < nefc; i++) {
    if (force[i] <= 0) state[i] = STATE_SATISFIED;
    else { state[i] = STATE_QUADRATIC; nactive++; }
  }
  return nactive;
}

// ---- efc_b and solver start point (mj_fwdConstraint head + warmstart) ----------------------------
MJB_HD void constraint_begin(const Env& d) {
  const DModel& m = d.m;
  const int nv = m.sz.nv, nefc = d.nefc()[0];
  FD qfc = d.qfrc_constraint(), qacc = d.qacc(), qas = d.qacc_smooth();
  for (int i = 0; i < nv; i++) qfc[i] = 0;
  d.solver_niter()[0] = 0;
  if (!nefc) { for (int i = 0; i < nv; i++) qacc[i] = qas[i]; return; }
  FD b = d.efc_b(), aref = d.efc_aref(), force = d.efc_force();
  mul_jac_vec(d, b, qas);
  for (int i = 0; i < nefc; i++) b[i] -= aref[i];
  if (!(m.opt.disableflags & DSBL_WARMSTART)) {
    FD jar = d.scr_efc(), ws = d.qacc_warmstart();
    for (int i = 0; i < nv; i++) qacc[i] = ws[i];
    mul_jac_vec(d, jar, ws);
    for (int i = 0; i < nefc; i++) jar[i] -= aref[i];
    double cost_ws = constraint_update(d, jar, true);
    if (m.opt.solver == SOL_PGS) {
      FD AR = d.efc_AR(), ARf = d.scr_efc() + nefc;
      double pw = dot_ref(nefc, [&](int i) { return force[i]; }, [&](int i) { return b[i]; });
      for (int r = 0; r < nefc; r++) {
        FD row = AR + (long)r * nefc;
        ARf[r] = dot_ref(nefc, [&](int c) { return row[c]; }, [&](int c) { return force[c]; });
      }
      pw += 0.5 * dot_ref(nefc, [&](int i) { return force[i]; }, [&](int i) { return ARf[i]; });
      if (pw > 0) {
        for (int i = 0; i < nefc; i++) force[i] = 0;
        for (int i = 0; i < nv; i++) qfc[i] = 0;
      }
    } else {
      // Newton/CG: choose the cheaper of qacc_warmstart and qacc_smooth by primal cost
      FD Ma = d.scr_nv(), M = d.M(), qfs = d.qfrc_smooth();
      // Ma = M * qacc_warmstart over the tree-sparse lower triangle (mj_mulM)
      for (int i = 0; i < nv; i++) {
        const int adr = m.M_rowadr[i], nnz = m.M_rownnz[i];
        Ma[i] = dot_sparse_ref(nnz, [&](int c) { return M[adr + c]; }, [&](int c) { return ws[m.M_colind[adr + c]]; });
      }
      for (int i = 0; i < nv; i++) {
        const int adr = m.M_rowadr[i], nnz = m.M_rownnz[i] - 1;
        const double wi = ws[i];
        if (wi != 0) for (int c = 0; c < nnz; c++) Ma[m.M_colind[adr + c]] += M[adr + c] * wi;
      }
      for (int i = 0; i < nv; i++) cost_ws += 0.5 * (Ma[i] - qfs[i]) * (ws[i] - qas[i]);
      const double cost_smooth = constraint_update(d, b, true);
      if (cost_ws > cost_smooth) for (int i = 0; i < nv; i++) qacc[i] = qas[i];
    }
  } else {
    for (int i = 0; i < nv; i++) qacc[i] = qas[i];
    for (int i = 0; i < nefc; i++) force[i] = 0;
  }
}

// ---- projected Gauss-Seidel on the dual, one environment per thread ------------------------------
MJB_HD void solve_pgs(const Env& d) {
  const DModel& m = d.m;
  const int nefc = d.nefc()[0], nf = d.nf()[0], nv = m.sz.nv;
  if (!nefc) return;
  FD force = d.efc_force(), floss = d.efc_frictionloss(), AR = d.efc_AR(), b = d.efc_b();
  FD ARinv = d.scr_efc(), fprev = d.scr_efc() + nefc, fmom = d.scr_efc() + 2 * (long)nefc;
  FI state = d.efc_state(), oldstate = d.scr_int(), order = d.scr_int() + nefc;
  const double scale = 1 / (m.opt.meaninertia * (nv > 1 ? nv : 1));
  for (int i = 0; i < nefc; i++) { fprev[i] = force[i]; ARinv[i] = 1 / AR[(long)i * (nefc + 1)]; order[i] = i; }
  dual_state(d);
  Pcg32 rng{0, 1};
  pcg32_next(rng);
  int iter = 0, nk = 0;
  const int maxiter = m.opt.iterations;
  while (iter < maxiter) {
    double beta = 0;
    if (iter > 0) beta = (double)(nk - 1) / (double)(nk + 2);
    if (beta > 0) {
      for (int i = 0; i < nefc; i++) {
        const double fs = force[i];
        force[i] += beta * (force[i] - fprev[i]);
        fprev[i] = fs;
      }
      for (int i = 0; i < nf; i++) force[i] = dclip(force[i], -floss[i], floss[i]);
      for (int i = nf; i < nefc; i++) if (force[i] < 0) force[i] = 0;
    } else {
      for (int i = 0; i < nefc; i++) fprev[i] = force[i];
    }
    for (int i = 0; i < nefc; i++) fmom[i] = force[i];

    double improvement = 0;
    for (int i = nefc - 1; i > 0; i--) {   // Fisher-Yates, same draws as the reference
      const uint32_t j = pcg32_next(rng) % (uint32_t)(i + 1);
      const int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    for (int bi = 0; bi < nefc; bi++) {
      const int i = order[bi];
      FD row = AR + (long)i * nefc;
      const double res = b[i] + dot_ref(nefc, [&](int c) { return row[c]; }, [&](int c) { return force[c]; });
      const double old = force[i];
      force[i] -= res * ARinv[i];
      if (i < nf) {
        if (force[i] < -floss[i]) force[i] = -floss[i];
        else if (force[i] > floss[i]) force[i] = floss[i];
      } else if (force[i] < 0) force[i] = 0;
      const double A = 1 / ARinv[i];
      const double delta = force[i] - old;
      double change = 0.5 * delta * delta * A + delta * res;
      if (change > 1e-10) { force[i] = old; change = 0; }
      improvement -= change;
    }
    for (int i = 0; i < nefc; i++) oldstate[i] = state[i];
    dual_state(d);
    improvement *= scale;
    bool restart = false;
    if (iter > 0) {
      double dce = 0;
      for (int i = 0; i < nefc; i++) dce += (force[i] - fmom[i]) * (fmom[i] - fprev[i]);
      restart = dce < 0;
    }
    if (restart) nk = 0; else nk++;
    iter++;
    if (improvement < m.opt.tolerance) break;
  }
  d.solver_niter()[0] += iter;
}

// dual finish: qfrc_constraint = J' f, qacc = qacc_smooth + M^-1 qfrc_constraint
MJB_HD void dual_finish(const Env& d) {
  const int nv = d.m.sz.nv;
  if (!d.nefc()[0]) return;
  FD qfc = d.qfrc_constraint(), qacc = d.qacc(), qas = d.qacc_smooth();
  mul_jacT_vec(d, qfc, d.efc_force());
  for (int i = 0; i < nv; i++) qacc[i] = qfc[i];
  solve_LD(d.m, qacc, d.qLD(), d.qLDiagInv());
  for (int i = 0; i < nv; i++) qacc[i] += qas[i];
}

}  // namespace mjb
