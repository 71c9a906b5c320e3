/* TEST INFRASTRUCTURE (oracle build shim) — not product code.  See ccd/vec3.h. */
#ifndef ORACLE_SHIM_CCD_H_
#define ORACLE_SHIM_CCD_H_
#include <ccd/vec3.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef void (*ccd_support_fn)(const void* obj, const ccd_vec3_t* dir, ccd_vec3_t* vec);
typedef void (*ccd_first_dir_fn)(const void* o1, const void* o2, ccd_vec3_t* dir);
typedef void (*ccd_center_fn)(const void* obj, ccd_vec3_t* center);
typedef struct _ccd_t {
  ccd_first_dir_fn first_dir;
  ccd_support_fn support1, support2;
  ccd_center_fn center1, center2;
  unsigned long max_iterations;
  ccd_real_t epa_tolerance, mpr_tolerance, dist_tolerance;
} ccd_t;
#define CCD_INIT(ccd) do { (ccd)->first_dir = ccdFirstDirDefault; (ccd)->support1 = 0; \
  (ccd)->support2 = 0; (ccd)->center1 = 0; (ccd)->center2 = 0; (ccd)->max_iterations = 0; \
  (ccd)->epa_tolerance = 1e-4; (ccd)->mpr_tolerance = 1e-4; (ccd)->dist_tolerance = 1e-6; } while (0)
void ccdFirstDirDefault(const void* o1, const void* o2, ccd_vec3_t* dir);
int ccdMPRPenetration(const void* obj1, const void* obj2, const ccd_t* ccd,
                      ccd_real_t* depth, ccd_vec3_t* dir, ccd_vec3_t* pos);
#ifdef __cplusplus
}
#endif
#endif
