// one half of the split step, 8 lanes per environment (see mjb_kstep.h, mjb_stage.h run_part)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER_PART(launch_kpart1_lean8, SOL_PGS, 8, 0, 1)
}  // namespace mjb
