// one instantiation of the fused step kernel (see mjb_kstep.h)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER(launch_kstep_any16, -1, 16, FEAT_ALL)
}  // namespace mjb
