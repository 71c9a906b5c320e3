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

// ---- cylinder and box colliders (only in kernels compiled with FEAT_COLBOX) ------------------------------------
MJB_HD V3 mulmTv3(const M3& a, V3 v) {   // mji_mulMatTVec3
  return V3{a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}

// plane : cylinder (engine_collision_primitive.c:101-208): the rim point nearest to the plane on either disk, plus
// two points at +-60 degrees on the nearer disk; up to four contacts
MJB_HD int collide_plane_cylinder(PreCon* c, double margin, V3 p1, const M3& m1, V3 p2, const M3& m2, const double* size2) {
  const V3 normal{m1.m[2], m1.m[5], m1.m[8]};
  V3 axis{m2.m[2], m2.m[5], m2.m[8]};
  double prjaxis = dot(normal, axis);
  if (prjaxis > 0) { axis = axis * -1.0; prjaxis = -prjaxis; }   // the axis points towards the plane
  V3 vec = p2 - p1;
  const double dist0 = dot(vec, normal);
  vec = axis * prjaxis;        // - normal, without its component along the axis
  vec = vec - normal;
  const double len_sqr = dot(vec, vec);
  if (len_sqr >= kMinVal * kMinVal) {
    const double scl = size2[0] / sqrt(len_sqr);
    vec.x *= scl; vec.y *= scl; vec.z *= scl;
  } else {                     // disk parallel to the plane: the cylinder's x axis
    vec = V3{m2.m[0] * size2[0], m2.m[3] * size2[0], m2.m[6] * size2[0]};
  }
  const double prjvec = dot(vec, normal);
  axis = axis * size2[1];
  prjaxis *= size2[1];
  int cnt = 0;
  auto put = [&](double dist, V3 pos) {
    c[cnt].dist = dist;
    pos.x += normal.x * (-dist * 0.5); pos.y += normal.y * (-dist * 0.5); pos.z += normal.z * (-dist * 0.5);
    c[cnt].pos = pos;
    c[cnt].normal = normal;
    c[cnt].tangent = V3{0, 0, 0};
    cnt++;
  };
  if (dist0 + prjaxis + prjvec <= margin) put(dist0 + prjaxis + prjvec, (p2 + vec) + axis);
  else return 0;
  if (dist0 - prjaxis + prjvec <= margin) put(dist0 - prjaxis + prjvec, (p2 + vec) - axis);
  const double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    V3 vec1 = cross(vec, axis);
    normalize(vec1);
    vec1 = vec1 * (size2[0] * sqrt(3.0) / 2);
    const double dist = dist0 + prjaxis + prjvec1;
    V3 pa = (p2 + vec1) + axis;
    pa.x += vec.x * -0.5; pa.y += vec.y * -0.5; pa.z += vec.z * -0.5;
    put(dist, pa);
    V3 pb = (p2 - vec1) + axis;
    pb.x += vec.x * -0.5; pb.y += vec.y * -0.5; pb.z += vec.z * -0.5;
    put(dist, pb);
  }
  return cnt;
}

// plane : box (engine_collision_primitive.c:210-258): the corners below the margin that point down, at most four
MJB_HD int collide_plane_box(PreCon* c, double margin, V3 p1, const M3& m1, V3 p2, const M3& m2, const double* size2) {
  const V3 norm{m1.m[2], m1.m[5], m1.m[8]};
  const V3 dif = p2 - p1;
  const double dist = dot(dif, norm);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    const V3 vec{(i & 1) ? size2[0] : -size2[0], (i & 2) ? size2[1] : -size2[1], (i & 4) ? size2[2] : -size2[2]};
    V3 corner = mulmv(m2, vec);
    const double ldist = dot(norm, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    c[cnt].dist = dist + ldist;
    c[cnt].normal = norm;
    corner = corner + p2;
    c[cnt].pos = corner + norm * (-c[cnt].dist / 2);
    c[cnt].tangent = V3{0, 0, 0};
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

// sphere : cylinder (engine_collision_primitive.c:345-423): side (sphere-sphere against the axis point), cap
// (plane-sphere against the cap plane) or rim (sphere-sphere against the rim point)
MJB_HD int collide_sphere_cylinder(PreCon* c, double margin, V3 p1, const M3& m1, const double* size1,
                                   V3 p2, const M3& m2, const double* size2) {
  const double radius = size2[0], height = size2[1];
  const V3 axis{m2.m[2], m2.m[5], m2.m[8]};
  V3 vec = p1 - p2;
  const double x = dot(axis, vec);
  V3 a_proj = axis * x;
  V3 p_proj = vec - a_proj;
  const double p_proj_sqr = dot(p_proj, p_proj);
  bool collide_side = fabs(x) < height;
  bool collide_cap = p_proj_sqr < radius * radius;
  if (collide_side && collide_cap) {   // centre inside the cylinder: keep the shallower exit
    const double dist_cap = height - fabs(x);
    const double dist_radius = radius - sqrt(p_proj_sqr);
    if (dist_cap < dist_radius) collide_side = false; else collide_cap = false;
  }
  if (collide_side) {
    a_proj = a_proj + p2;
    return raw_sphere_sphere(c[0], margin, p1, m1, size1[0], a_proj, m2, size2[0]);
  }
  if (collide_cap) {
    V3 pos_cap;
    M3 mat_cap = m2;
    if (x > 0) pos_cap = addscl(p2, axis, height);
    else {
      pos_cap = addscl(p2, axis, -height);
      mat_cap.m[0] = -m2.m[0]; mat_cap.m[2] = -m2.m[2]; mat_cap.m[3] = -m2.m[3]; mat_cap.m[5] = -m2.m[5];
      mat_cap.m[6] = -m2.m[6]; mat_cap.m[8] = -m2.m[8];
    }
    const int n = raw_plane_sphere(c[0], margin, pos_cap, mat_cap, p1, size1[0]);
    if (n) c[0].normal = c[0].normal * -1.0;   // the pair is (sphere, cylinder): the normal points from the sphere
    return n;
  }
  p_proj = p_proj * (size2[0] / sqrt(p_proj_sqr));
  vec = axis * (x > 0 ? height : -height);
  vec = vec + p_proj;
  vec = vec + p2;
  return raw_sphere_sphere(c[0], margin, p1, m1, size1[0], vec, m2, 0.0);
}

// sphere : box (engine_collision_box.c:35-95)
MJB_HD int raw_sphere_box(PreCon& c, double margin, V3 p1, double r1, V3 p2, const M3& m2, const double* size2) {
  V3 tmp = p1 - p2;
  const V3 center = mulmTv3(m2, tmp);
  V3 clamped = center;
  if (size2[0] > 0) clamped.x = dclip(clamped.x, -size2[0], size2[0]);
  if (size2[1] > 0) clamped.y = dclip(clamped.y, -size2[1], size2[1]);
  if (size2[2] > 0) clamped.z = dclip(clamped.z, -size2[2], size2[2]);
  V3 deepest = center;
  tmp = clamped - center;
  double dist = normalize(tmp);
  if (dist - r1 > margin) return 0;
  V3 pos;
  if (dist <= kMinVal) {   // sphere centre inside the box: leave through the nearest face
    double closest = (size2[0] + size2[1] + size2[2]) * 2;
    int k = 0;
    for (int i = 0; i < 6; i++) {
      const double v = fabs(((i % 2) ? 1 : -1) * size2[i / 2] - get(center, i / 2));
      if (closest > v) { closest = v; k = i; }
    }
    V3 nearest{0, 0, 0};
    const double sgn = (k % 2) ? -1 : 1;
    if (k / 2 == 0) nearest.x = sgn; else if (k / 2 == 1) nearest.y = sgn; else nearest.z = sgn;
    pos = center;
    const double s = (r1 - closest) / 2;
    pos.x += nearest.x * s; pos.y += nearest.y * s; pos.z += nearest.z * s;
    c.normal = mulmv(m2, nearest);
    dist = -closest;
  } else {
    deepest.x += tmp.x * r1; deepest.y += tmp.y * r1; deepest.z += tmp.z * r1;
    pos = V3{0, 0, 0};
    pos.x += clamped.x * 0.5; pos.y += clamped.y * 0.5; pos.z += clamped.z * 0.5;
    pos.x += deepest.x * 0.5; pos.y += deepest.y * 0.5; pos.z += deepest.z * 0.5;
    c.normal = mulmv(m2, tmp);
  }
  tmp = mulmv(m2, pos);
  c.pos = tmp + p2;
  c.dist = dist - r1;
  c.tangent = V3{0, 0, 0};
  return 1;
}

// filter + narrowphase of candidate pair p; returns the number of pre-contacts written to pc[]
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
  if (d.feat & FEAT_COLBOX) {
    if (t1 == GEOM_PLANE && t2 == GEOM_CYLINDER) return collide_plane_cylinder(pc, margin, p1, m1, p2, m2, s2);
    if (t1 == GEOM_PLANE && t2 == GEOM_BOX) return collide_plane_box(pc, margin, p1, m1, p2, m2, s2);
    if (t1 == GEOM_SPHERE && t2 == GEOM_CYLINDER) return collide_sphere_cylinder(pc, margin, p1, m1, s1, p2, m2, s2);
    if (t1 == GEOM_SPHERE && t2 == GEOM_BOX) return raw_sphere_box(pc[0], margin, p1, s1[0], p2, m2, s2);
  }
  return 0;
}

// walk the static candidate table cooperatively in two passes: (1) every lane counts the
// pre-contacts of its candidates, (2) after a serial scan that assigns contact slots IN TABLE ORDER
// (the reference's emission order) the few hitting pairs are evaluated again and written in place.
template <int MAXC>
MJB_HD void collision_t(const Env& d) {
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
    PreCon pc[MAXC];
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
    PreCon pc[MAXC];
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

// pairs of the sphere / capsule / plane family give at most two contacts; the cylinder and box colliders up to
// eight (kernels without FEAT_COLBOX do not carry their code or the larger pre-contact array)
MJB_HD void collision(const Env& d) {
  if (d.feat & FEAT_COLBOX) collision_t<8>(d); else collision_t<2>(d);
}

}  // namespace mjb
