// one instantiation of the fused step kernel (see mjb_kstep.h)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER(launch_kstep_cg32, SOL_CG, 32, FEAT_ALL)
}  // namespace mjb
