/* TEST INFRASTRUCTURE (oracle build shim) — not product code.
 * Minimal stand-in for libccd's <ccd/vec3.h>, only so that the UNMODIFIED
 * reference engine sources compile.  libccd is reachable only through the
 * non-default mjDSBL_NATIVECCD fallback (reference
 * src/engine/engine_collision_convex.c:52-91); every entry point aborts. */
#ifndef ORACLE_SHIM_CCD_VEC3_H_
#define ORACLE_SHIM_CCD_VEC3_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef double ccd_real_t;
typedef struct _ccd_vec3_t { ccd_real_t v[3]; } ccd_vec3_t;
extern ccd_vec3_t* ccd_vec3_origin;
static inline void ccdVec3Set(ccd_vec3_t* v, ccd_real_t x, ccd_real_t y, ccd_real_t z) {
  v->v[0] = x; v->v[1] = y; v->v[2] = z;
}
static inline int ccdVec3Eq(const ccd_vec3_t* a, const ccd_vec3_t* b) {
  return a->v[0] == b->v[0] && a->v[1] == b->v[1] && a->v[2] == b->v[2];
}
#ifdef __cplusplus
}
#endif
#endif
