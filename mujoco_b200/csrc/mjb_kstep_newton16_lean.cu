// one instantiation of the fused step kernel (see mjb_kstep.h)
#define MJB_KSTEP_INSTANCE
#include "mjb_kstep.h"
namespace mjb {
MJB_KSTEP_LAUNCHER(launch_kstep_newton16_lean, SOL_NEWTON, 16, 0)
}  // namespace mjb
