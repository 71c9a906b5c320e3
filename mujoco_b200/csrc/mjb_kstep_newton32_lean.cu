// one instantiation of the fused step kernel (see mjb_kstep.h)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER(launch_kstep_newton32_lean, SOL_NEWTON, 32, 0)
}  // namespace mjb
