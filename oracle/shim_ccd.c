/* TEST INFRASTRUCTURE (oracle build shim) — not product code.
 * libccd entry points abort: only the non-default mjDSBL_NATIVECCD path would call them. */
#include <stdio.h>
#include <stdlib.h>
#include <ccd/ccd.h>
static ccd_vec3_t origin_ = {{0, 0, 0}};
ccd_vec3_t* ccd_vec3_origin = &origin_;
void ccdFirstDirDefault(const void* o1, const void* o2, ccd_vec3_t* dir) {
  (void)o1; (void)o2; ccdVec3Set(dir, 1, 0, 0);
}
int ccdMPRPenetration(const void* obj1, const void* obj2, const ccd_t* ccd,
                      ccd_real_t* depth, ccd_vec3_t* dir, ccd_vec3_t* pos) {
  (void)obj1; (void)obj2; (void)ccd; (void)depth; (void)dir; (void)pos;
  fprintf(stderr, "oracle shim: libccd path (mjDSBL_NATIVECCD) is not available\n");
  abort();
}
