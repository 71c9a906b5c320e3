// TEST INFRASTRUCTURE — host emulation backend, never part of the product.
//
// Implements mujoco_b200/csrc/mjb_backend.h with plain host loops over environments so that the
// exact kernel source (mjb_stage.h and everything it includes) can be compiled by g++ and checked
// bit-for-bit against the oracle in a container that has no GPU.  The product library
// (libmjb200.so) never links this file and the mujoco_b200 package never loads the resulting
// tests/hostemu/libmjb_hostemu.so; GPU parity tests always go through the CUDA library.
#include <cstdlib>
#include <cstring>

#include "../../mujoco_b200/csrc/mjb_backend.h"
#include "../../mujoco_b200/csrc/mjb_stage.h"

namespace mjb {
namespace backend {

static long g_launches = 0;
const char* name() { return "hostemu"; }
int init(int) { return 0; }
void* dev_alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
void dev_free(void* p) { free(p); }
int h2d(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }
int d2h(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }
int dev_zero(void* dst, size_t bytes, void*) { memset(dst, 0, bytes); return 0; }
void* stream_create() { return nullptr; }
void stream_destroy(void*) {}
int sync(void*) { return 0; }
int stream_order(void*, void*) { return 0; }
void* event_record(void*) { return (void*)1; }
int stream_wait_event(void*, void*) { return 0; }
int d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void*) {
  for (size_t r = 0; r < height; r++) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  return 0;
}
long launches() { return g_launches; }
int set_debug(const char*, int) { return 0; }
bool rollout_persistent_available(const DModel&, const Batch&, int) { return false; }   // the emulation steps launch by launch
int launch_rollout_persistent(const DModel&, const Batch&, int, int, int, int, int, int, const double*, double*, int, void*) { return -1; }

int launch_stages(const DModel& dm, const Batch& b, int mask, int flags, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_env(dm, b, e, mask, flags, 0, 1, nullptr, 0);
  return 0;
}
// split step, emulated part by part (the CUDA build runs the solve with mjb_pgs4.cu; here the generic sweeps
// stand in for it, so the CPU suite covers run_part and the redo launch)
bool split_step_available(const DModel& dm, const Batch& b) {
  const char* s = getenv("MJB_SPLIT");
  const bool islands = dm.sz.ntree > 1 && !(dm.opt.disableflags & DSBL_ISLAND);
  return (s ? atoi(s) : 1) && b.warp_per_env && dm.opt.solver == SOL_PGS && !islands && dm.opt.noslip_iterations <= 0 &&
         (dm.opt.integrator == INT_EULER || dm.opt.integrator == INT_IMPLICITFAST);
}
int launch_split_step(const DModel& dm, const Batch& b, int first, int later, void*, void*) {
  g_launches += 4;
  for (int e = 0; e < b.nenv; e++) run_env(dm, b, e, 0, first, 0, 1, nullptr, 0, -1, 0xffffffffu, FEAT_ALL, 1);
  for (int e = 0; e < b.nenv; e++) run_env(dm, b, e, 4, later, 0, 1, nullptr, 0);
  for (int e = 0; e < b.nenv; e++) run_env(dm, b, e, 0, later, 0, 1, nullptr, 0, -1, 0xffffffffu, FEAT_ALL, 2);
  if (!(dm.opt.disableflags & DSBL_AUTORESET))
    for (int e = 0; e < b.nenv; e++) run_env(dm, b, e, kMaskStep, later | 16, 0, 1, nullptr, 0);
  return 0;
}
int profile_split_step(const DModel& dm, const Batch& b, void* s, float* ms) {
  for (int i = 0; i < 4; i++) ms[i] = 0;
  return launch_split_step(dm, b, 1, 0, s, nullptr);
}
int launch_get_sensor(const DModel& dm, const Batch& b, double* sens, int nstep, int t, int nsens, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_get_sensor(dm, b, e, sens, nstep, t, nsens);
  return 0;
}
int launch_rk4(const DModel& dm, const Batch& b, int phase, int flags, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_rk4(dm, b, e, phase, flags, 0, 1);
  return 0;
}
int launch_pack(const Batch& b, int is_int, long off, long cnt, void* dense, int to_dense, void*) {
  for (long i = 0; i < (long)b.nenv * cnt; i++) run_pack(b, is_int, off, cnt, dense, to_dense, i);
  return 0;
}
int launch_fill_zero(const Batch& b, int is_int, long off, long cnt, void*) {
  for (long i = 0; i < (long)b.nenv * cnt; i++) run_fill_zero(b, is_int, off, cnt, i);
  return 0;
}
int launch_reset(const DModel& dm, const Batch& b, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) { Env d(dm, b, e); reset_env(d, true); }
  return 0;
}
int launch_set_control(const DModel& dm, const Batch& b, const double* control, int nstep, int t, unsigned spec,
                       int ncontrol, void*, bool skip_warned) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_set_control(dm, b, e, control, nstep, t, spec, ncontrol, skip_warned);
  return 0;
}
int launch_get_state(const DModel& dm, const Batch& b, double* state, int nstep, int t, int nstate, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_get_state(dm, b, e, state, nstep, t, nstate);
  return 0;
}
int launch_set_control_native(const DModel& dm, const Batch& b, const double* ctrl, int t, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_set_control_native(dm, b, e, ctrl, t);
  return 0;
}
int launch_get_state_native(const DModel& dm, const Batch& b, double* state, int t, int nstate, void*) {
  g_launches++;
  for (int e = 0; e < b.nenv; e++) run_get_state_native(dm, b, e, state, t, nstate);
  return 0;
}

}  // namespace backend
}  // namespace mjb
