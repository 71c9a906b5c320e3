// Thin execution backend used by mjb_api.cc.  The product build (libmjb200.so) implements it with
// CUDA (mjb_kernels.cu); tests/hostemu implements the same interface with host loops so that the
// kernel SOURCE can be checked against the oracle in a container without a GPU.  The host
// emulation is test infrastructure and is never linked into, or loaded by, the product.
#pragma once
#include <stddef.h>
#include "mjb_types.h"

namespace mjb {
namespace backend {

const char* name();
int init(int device);                         // 0 or negative mjb error code (sets error text)
void* dev_alloc(size_t bytes);                // zero-initialised
void dev_free(void* p);
int h2d(void* dst, const void* src, size_t bytes, void* stream);
int d2h(void* dst, const void* src, size_t bytes, void* stream);
int dev_zero(void* dst, size_t bytes, void* stream);
void* stream_create();
void stream_destroy(void* s);
int sync(void* stream);
// order `waiter` after everything enqueued on `signaller` so far (event record + stream wait)
int stream_order(void* signaller, void* waiter);
// events for overlapping copies with the work of another stream, and a strided device -> host copy
void* event_record(void* stream);                    // new event recorded on `stream` (NULL on failure)
int stream_wait_event(void* stream, void* event);    // `stream` waits for the event; the event is released afterwards
int d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream);

// run the pipeline stages selected by `mask` (bit s = stage s, see mjb_forward.h) for every
// environment in ONE launch.  flags: bit0 = part of mj_step (run the qpos/qvel checks), bit1 = skip
// environments whose warning counters are non-zero
int launch_stages(const DModel& dm, const Batch& b, int mask, int flags, void* stream);
// one full Euler / implicitfast step as a split sequence of launches (PGS in its own kernel); `first` / `later`
// are the flags of the first and the following launches of the step (rollout skip rule)
bool split_step_available(const DModel& dm, const Batch& b);
// stagger: optional stream that is made to wait for this step's first half (position + velocity) - the env groups
// of a multi-step call start one first-half apart, so that one group's solve (a dependent chain that leaves the
// SMs mostly idle) overlaps the other groups' throughput-bound halves
int launch_split_step(const DModel& dm, const Batch& b, int first, int later, void* stream, void* stagger = nullptr);
// one split step, synchronous, with the duration (ms) of its four launches: first half, solve, second half, redo
int profile_split_step(const DModel& dm, const Batch& b, void* stream, float* ms);
// persistent rollout kernel (mjb_krollout.cu): steps [t0, t1) of an nstep rollout in one launch, every CTA carrying its
// own environments through the steps without a device-wide barrier.  layout 0: native ctrl [step][nu][stride] / state
// [step][nstate][stride]; layout 1: reference ctrl [env][step][nu] / state [env][step][nstate].  ctrl / state may be null.
bool rollout_persistent_available(const DModel& dm, const Batch& b, int nstep);
int launch_rollout_persistent(const DModel& dm, const Batch& b, int t0, int t1, int nstep, int first, int later, int layout,
                              const double* ctrl, double* state, int nstate, void* stream);
int launch_rk4(const DModel& dm, const Batch& b, int phase, int flags, void* stream);   // rk4_phase of every env
int launch_reset(const DModel& dm, const Batch& b, void* stream);
// rollout helpers; control/state are DEVICE buffers laid out [nenv][nstep][n] (reference layout)
int launch_set_control(const DModel& dm, const Batch& b, const double* control, int nstep, int t,
                       unsigned spec, int ncontrol, void* stream, bool skip_warned = true);
int launch_get_state(const DModel& dm, const Batch& b, double* state, int nstep, int t, int nstate, void* stream);
int launch_get_sensor(const DModel& dm, const Batch& b, double* sens, int nstep, int t, int nsens, void* stream);   // [nenv][nstep][nsens]
// native-layout variants: ctrl [nstep][nu][stride], state [nstep][nstate][stride]
int launch_set_control_native(const DModel& dm, const Batch& b, const double* ctrl, int t, void* stream);
int launch_get_state_native(const DModel& dm, const Batch& b, double* state, int t, int nstate, void* stream);
// dense [nenv][cnt] staging buffer <-> batch field (either layout); to_dense=1 gathers, 0 scatters
int launch_pack(const Batch& b, int is_int, long off, long cnt, void* dense, int to_dense, void* stream);
int launch_fill_zero(const Batch& b, int is_int, long off, long cnt, void* stream);
long launches();
// development / test switches of the backend ("pgs4_slots": force the slot layout of the PGS kernel); -1 = unknown key
int set_debug(const char* key, int value);

}  // namespace backend
}  // namespace mjb
