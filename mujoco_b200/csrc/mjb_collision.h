// Collision detection of the batched mj_step path, one environment per call.
//
// The reference runs broadphase (PCA frame + sweep-and-prune), midphase (BVH/OBB) and narrowphase
// per step (src/engine/engine_collision_driver.c:595-886).  Every pruning stage there is
// conservative with respect to the final `dist < margin` decision, and the ORDER in which contacts
// are emitted is a pure function of the model (body-pair signature order, then the per-pair geom
// order, :637-745,410-443).  This implementation therefore walks a STATIC, host-precomputed
// candidate table (mjb_model.cc) in that order and applies, per candidate, exactly the
// reference's per-pair filter (bounding sphere / plane distance, :267-284,547-591) followed by the
// same closed-form colliders (src/engine/engine_collision_primitive.c:28-530) and mj_setContact
// (:1839-1875).  All environments of a warp walk the same table: no divergence in control flow
// except the hit/miss branches.
#pragma once
#include "mjb_types.h"

namespace mjb {

struct PreCon { double dist; V3 pos, normal, tangent; };

// plane : sphere, shared by plane-capsule  (engine_collision_primitive.c:28-50)
MJB_HD int raw_plane_sphere(PreCon& c, double margin, V3 ppos, const M3& pmat, V3 spos, double radius) {
  c.normal = V3{pmat.m[2], pmat.m[5], pmat.m[8]};
  V3 t = spos - ppos;
  double cdist = dot(t, c.normal);
  if (cdist > margin + radius) return 0;
  c.dist = cdist - radius;
  t = c.normal * (-c.dist / 2 - radius);
  c.pos = spos + t;
  c.tangent = V3{0, 0, 0};
  return 1;
}

// sphere : sphere kernel shared by the capsule colliders (engine_collision_primitive.c:246-285)
MJB_HD int raw_sphere_sphere(PreCon& c, double margin, V3 p1, const M3& m1, double r1, V3 p2, const M3& m2, double r2) {
  V3 dif = p1 - p2;
  double d2 = dot(dif, dif);
  double mind = margin + r1 + r2;
  if (d2 > mind * mind) return 0;
  c.dist = sqrt(d2) - r1 - r2;
  c.normal = p2 - p1;
  double len = normalize(c.normal);
  if (len < kMinVal) {
    V3 a1{m1.m[2], m1.m[5], m1.m[8]}, a2{m2.m[2], m2.m[5], m2.m[8]};
    c.normal = cross(a1, a2);
    normalize(c.normal);
  }
  c.pos = c.normal * (r1 + c.dist / 2);
  c.pos = c.pos + p1;
  c.tangent = V3{0, 0, 0};
  return 1;
}

MJB_HD int collide_plane_capsule(PreCon* c, double margin, V3 p1, const M3& m1, V3 p2, const M3& m2, const double* size2) {
  V3 axis{m2.m[2], m2.m[5], m2.m[8]};
  V3 seg{size2[1] * axis.x, size2[1] * axis.y, size2[1] * axis.z};
  int n1 = raw_plane_sphere(c[0], margin, p1, m1, p2 + seg, size2[0]);
  int n2 = raw_plane_sphere(c[n1], margin, p1, m1, p2 - seg, size2[0]);
  if (n1) c[0].tangent = axis;
  if (n2) c[n1].tangent = axis;
  return n1 + n2;
}

MJB_HD int collide_sphere_capsule(PreCon* c, double margin, V3 p1, const M3& m1, const double* size1,
                                  V3 p2, const M3& m2, const double* size2) {
  double len = size2[1];
  V3 axis{m2.m[2], m2.m[5], m2.m[8]};
  V3 vec = p1 - p2;
  double x = dclip(dot(axis, vec), -len, len);
  vec = axis * x;
  vec = vec + p2;
  return raw_sphere_sphere(c[0], margin, p1, m1, size1[0], vec, m2, size2[0]);
}

// capsule : capsule (engine_collision_primitive.c:431-518)
MJB_HD int collide_capsule_capsule(PreCon* c, double margin, V3 p1, const M3& m1, const double* size1,
                                   V3 p2, const M3& m2, const double* size2) {
  V3 ax1{m1.m[2] * size1[1], m1.m[5] * size1[1], m1.m[8] * size1[1]};
  V3 ax2{m2.m[2] * size2[1], m2.m[5] * size2[1], m2.m[8] * size2[1]};
  V3 dif = p1 - p2;
  double ma = dot(ax1, ax1), mb = -dot(ax1, ax2), mc = dot(ax2, ax2);
  double u = -dot(ax1, dif), v = dot(ax2, dif);
  double det = ma * mc - mb * mb;
  const double r1 = size1[0], r2 = size2[0];
  if (fabs(det) >= kMinVal) {
    double x1 = (mc * u - mb * v) / det;
    double x2 = (ma * v - mb * u) / det;
    if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; }
    else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
    if (x2 > 1) { x2 = 1; x1 = dclip((u - mb) / ma, -1, 1); }
    else if (x2 < -1) { x2 = -1; x1 = dclip((u + mb) / ma, -1, 1); }
    V3 v1 = ax1 * x1; v1 = v1 + p1;
    V3 v2 = ax2 * x2; v2 = v2 + p2;
    return raw_sphere_sphere(c[0], margin, v1, m1, r1, v2, m2, r2);
  }
  // parallel axes: up to two contacts from the four end-point projections
  V3 v1 = p1 + ax1;
  double x2 = dclip((v - mb) / mc, -1, 1);
  V3 v2 = ax2 * x2; v2 = v2 + p2;
  int n1 = raw_sphere_sphere(c[0], margin, v1, m1, r1, v2, m2, r2);
  v1 = p1 - ax1;
  x2 = dclip((v + mb) / mc, -1, 1);
  v2 = ax2 * x2; v2 = v2 + p2;
  int n2 = raw_sphere_sphere(c[n1], margin, v1, m1, r1, v2, m2, r2);
  if (n1 + n2 >= 2) return n1 + n2;
  v2 = p2 + ax2;
  double x1 = dclip((u - mb) / ma, -1, 1);
  v1 = ax1 * x1; v1 = v1 + p1;
  int n3 = raw_sphere_sphere(c[n1 + n2], margin, v1, m1, r1, v2, m2, r2);
  if (n1 + n2 + n3 >= 2) return n1 + n2 + n3;
  v2 = p2 - ax2;
  x1 = dclip((u + mb) / ma, -1, 1);
  v1 = ax1 * x1; v1 = v1 + p1;
  int n4 = raw_sphere_sphere(c[n1 + n2 + n3], margin, v1, m1, r1, v2, m2, r2);
  return n1 + n2 + n3 + n4;
}

// filter + narrowphase of candidate pair p; returns the number of pre-contacts written to pc[0..1]
MJB_HD int pair_collide(const Env& d, int p, PreCon* pc) {
  const DModel& m = d.m;
  FD gx = d.geom_xpos(), gm = d.geom_xmat();
  const int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p];
  const int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
  const double margin = m.pair_margin[p];
  const double rb1 = m.geom_rbound[g1], rb2 = m.geom_rbound[g2];
  V3 p1 = ld3(gx, 3 * g1), p2 = ld3(gx, 3 * g2);
  // per-pair filter (mj_filterSphere)
  if (rb1 > 0 && rb2 > 0) {
    V3 dif = p1 - p2;
    double dsq = dif.x * dif.x + dif.y * dif.y + dif.z * dif.z;
    double bound = rb1 + rb2 + margin;
    if (dsq > bound * bound) return 0;
  } else if (t1 == GEOM_PLANE && rb2 > 0) {
    V3 nrm{gm[9 * g1 + 2], gm[9 * g1 + 5], gm[9 * g1 + 8]};
    V3 dif = p2 - p1;
    if (dot(dif, nrm) > margin + rb2) return 0;
  }
  M3 m1 = ld9(gm, 9 * g1), m2 = ld9(gm, 9 * g2);
  const double* s1 = m.geom_size + 3 * g1;
  const double* s2 = m.geom_size + 3 * g2;
  if (t1 == GEOM_PLANE && t2 == GEOM_SPHERE) return raw_plane_sphere(pc[0], margin, p1, m1, p2, s2[0]);
  if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) return collide_plane_capsule(pc, margin, p1, m1, p2, m2, s2);
  if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) return raw_sphere_sphere(pc[0], margin, p1, m1, s1[0], p2, m2, s2[0]);
  if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) return collide_sphere_capsule(pc, margin, p1, m1, s1, p2, m2, s2);
  if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) return collide_capsule_capsule(pc, margin, p1, m1, s1, p2, m2, s2);
  return 0;
}

// walk the static candidate table cooperatively in two passes: (1) every lane counts the
// pre-contacts of its candidates, (2) after a serial scan that assigns contact slots IN TABLE ORDER
// (the reference's emission order) the few hitting pairs are evaluated again and written in place.
MJB_HD void collision(const Env& d) {
  const DModel& m = d.m;
  FI ncon_f = d.ncon();
  if (m.opt.disableflags & (DSBL_CONSTRAINT | DSBL_CONTACT)) {
    MJB_LANE0 ncon_f[0] = 0;
    MJB_PSYNC();
    return;
  }
  FI cnt = d.scr_ipair();        // per pair: number of pre-contacts, then packed (offset | n << 24)
  const int npair = m.sz.npair, nconmax = m.sz.nconmax;
  MJB_PFOR(p, npair) {
    PreCon pc[2];
    cnt[p] = pair_collide(d, p, pc);
  }
  MJB_PSYNC();
  MJB_LANE0 {   // exclusive scan in table order, capped at nconmax
    int total = 0;
    bool full = false;
    for (int p = 0; p < npair; p++) {
      int n = cnt[p];
      if (total + n > nconmax) { n = nconmax - total; full = true; }
      cnt[p] = total | (n << 24);
      total += n;
    }
    if (full) d.warning()[WARN_CONTACTFULL] += 1;
    ncon_f[0] = total;
  }
  MJB_PSYNC();
  FD cdist = d.con_dist(), cpos = d.con_pos(), cframe = d.con_frame(), cinc = d.con_includemargin();
  FD cfri = d.con_friction(), csolref = d.con_solref(), csolimp = d.con_solimp(), cmu = d.con_mu();
  FI cg1 = d.con_geom1(), cg2 = d.con_geom2(), cdim = d.con_dim(), cexc = d.con_exclude(), cadr = d.con_efcadr();
  FI cpair = d.con_pair();
  MJB_PFOR(p, npair) {
    const int off = cnt[p] & 0xFFFFFF, n = cnt[p] >> 24;
    if (!n) continue;
    PreCon pc[2];
    pair_collide(d, p, pc);
    for (int k = 0; k < n; k++) {
      const int c = off + k;
      const double dist = pc[k].dist;
      cdist[c] = dist;
      st3(cpos, 3 * c, pc[k].pos);
      M3 fr;
      fr.m[0] = pc[k].normal.x; fr.m[1] = pc[k].normal.y; fr.m[2] = pc[k].normal.z;
      fr.m[3] = pc[k].tangent.x; fr.m[4] = pc[k].tangent.y; fr.m[5] = pc[k].tangent.z;
      fr.m[6] = 0; fr.m[7] = 0; fr.m[8] = 0;
      make_frame(fr);
      st9(cframe, 9 * c, fr);
      cg1[c] = m.pair_geom1[p]; cg2[c] = m.pair_geom2[p];
      cpair[c] = p;
      cdim[c] = m.pair_dim[p];
      const double inc = m.pair_includemargin[p];
      cinc[c] = inc;
      for (int j = 0; j < 5; j++) cfri[5 * c + j] = m.pair_friction[5 * p + j];
      for (int j = 0; j < 2; j++) csolref[2 * c + j] = m.pair_solref[2 * p + j];
      for (int j = 0; j < 5; j++) csolimp[5 * c + j] = m.pair_solimp[5 * p + j];
      cexc[c] = (dist >= inc) ? 1 : 0;
      cadr[c] = -1;
      cmu[c] = 0;
    }
  }
  MJB_PSYNC();
}

}  // namespace mjb
