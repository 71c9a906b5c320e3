// one half of the split step (see mjb_kstep.h, mjb_stage.h run_part)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER_PART(launch_kpart1, SOL_PGS, 32, FEAT_ALL, 1)
}  // namespace mjb
