// Newton solver of the batched mj_step path (primal), one environment per call.
// Replaces reference src/engine/engine_solver.c mj_solPrimal / Newton :1095-2563.
#pragma once
#include "mjb_constraint.h"

namespace mjb {

MJB_HD void solve_newton(const Env& d) {
  // implemented in a later milestone; the host refuses solver=Newton until then
  (void)d;
}

}  // namespace mjb
