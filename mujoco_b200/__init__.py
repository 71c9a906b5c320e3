"""mujoco_b200 — B200-native batched implementation of MuJoCo's mj_step hot path.

Thin ctypes binding over the C ABI in include/mjb.h (libmjb200.so, built in-tree by
__graft_entry__.build()).  The Python surface mirrors the reference's batched caller,
python/mujoco/rollout.py (`Rollout.rollout(model, data, initial_state, control, ...)`):
C-contiguous float64 arrays in the reference's layouts, `initial_state` in mjSTATE_FULLPHYSICS order.

The CUDA library is the only compute path: if it is missing, or no CUDA device is usable, loading or
batch creation raises — there is no CPU fallback in this package.
"""
import atexit
import ctypes as C
import sys
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmjb200.so")

# mjtState bits (reference include/mujoco/mjtype.h:503-527)
STATE_TIME, STATE_QPOS, STATE_QVEL, STATE_ACT, STATE_HISTORY = 1, 2, 4, 8, 16
STATE_WARMSTART, STATE_CTRL, STATE_QFRC_APPLIED, STATE_PLUGIN = 32, 64, 128, 1 << 13
STATE_XFRC_APPLIED, STATE_EQ_ACTIVE, STATE_MOCAP_POS, STATE_MOCAP_QUAT = 1 << 8, 1 << 9, 1 << 10, 1 << 11
STATE_FULLPHYSICS = STATE_TIME | STATE_QPOS | STATE_QVEL | STATE_ACT | STATE_HISTORY | STATE_PLUGIN

SOLVER_PGS, SOLVER_CG, SOLVER_NEWTON = 0, 1, 2
INT_EULER, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST = 0, 1, 2, 3


class MjbError(RuntimeError):
    pass


def _bind(cdll):
    """declare argument / result types of every symbol in include/mjb.h on a loaded library"""
    L = cdll
    vp, cp, i, u, l = C.c_void_p, C.c_char_p, C.c_int, C.c_uint, C.c_long
    dp = C.POINTER(C.c_double)
    L.mjb_last_error.restype = cp
    L.mjb_version.restype = i
    L.mjb_load_model.restype = vp
    L.mjb_load_model.argtypes = [cp]
    L.mjb_free_model.argtypes = [vp]
    L.mjb_check_model.argtypes = [vp]
    L.mjb_model_size.restype = l
    L.mjb_model_size.argtypes = [vp, cp]
    L.mjb_get_option.argtypes = [vp, cp, dp]
    L.mjb_set_option.argtypes = [vp, cp, C.c_double]
    L.mjb_make_batch.restype = vp
    L.mjb_make_batch.argtypes = [vp, i, i, i, i]
    L.mjb_free_batch.argtypes = [vp]
    L.mjb_set_thread_mapping.argtypes = [i]
    L.mjb_nenv.argtypes = [vp]
    L.mjb_reset.argtypes = [vp]
    L.mjb_state_size.argtypes = [vp, u]
    L.mjb_set_state.argtypes = [vp, vp, u]
    L.mjb_get_state.argtypes = [vp, vp, u]
    L.mjb_forward.argtypes = [vp]
    L.mjb_step.argtypes = [vp, i]
    L.mjb_rollout.argtypes = [vp, i, u, vp, vp, vp, vp, vp]
    L.mjb_step_mjdata.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mjb_step_host.argtypes = [vp, vp, vp]
    L.mjb_rollout_device.argtypes = [vp, i, vp, vp]
    L.mjb_env_stride.restype = l
    L.mjb_env_stride.argtypes = [vp]
    L.mjb_field_size.restype = l
    L.mjb_field_size.argtypes = [vp, cp]
    L.mjb_get_field.argtypes = [vp, cp, vp]
    L.mjb_get_field_int.argtypes = [vp, cp, vp]
    L.mjb_set_field.argtypes = [vp, cp, vp]
    L.mjb_kernel_launches.restype = l
    L.mjb_kernel_launches.argtypes = [vp]
    L.mjb_warning_counts.argtypes = [vp, vp]
    L.mjb_run_stages.argtypes = [vp, i, i]
    L.mjb_step_profile.argtypes = [vp, vp]
    L.mjb_set_debug.argtypes = [cp, i]
    L.mjb_stream.restype = vp
    L.mjb_set_stream.argtypes = [vp, vp]
    L.mjb_stream.argtypes = [vp]
    return L


_lib = None
_live = weakref.WeakSet()     # Batch / Model objects that still own native memory


@atexit.register
def _close_all():
    """release native objects before interpreter teardown (batches first: they reference their model)"""
    for cls in (Batch, Model):
        for o in [x for x in list(_live) if isinstance(x, cls)]:
            try:
                o.close()
            except Exception:  # noqa: BLE001
                pass


def lib():
    """the product library; raises if libmjb200.so has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MjbError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the only compute path; there is no CPU fallback)")
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def _err(L):
    return L.mjb_last_error().decode()


class Model:
    """an mjModel loaded from an MJB binary (reference mj_saveModel format) or wrapped from an address"""

    def __init__(self, path=None, address=None, library=None):
        self.L = library or lib()
        self._own = False
        if path is not None:
            self.ptr = self.L.mjb_load_model(os.fsencode(path))
            if not self.ptr:
                raise MjbError(_err(self.L))
            self._own = True
            _live.add(self)
        else:
            self.ptr = address  # e.g. mujoco.MjModel._address of the stock Python bindings

    def size(self, name):
        v = self.L.mjb_model_size(self.ptr, name.encode())
        if v < 0:
            raise KeyError(name)
        return int(v)

    def get_option(self, name):
        v = C.c_double()
        if self.L.mjb_get_option(self.ptr, name.encode(), C.byref(v)):
            raise KeyError(name)
        return v.value

    def set_option(self, name, value):
        if self.L.mjb_set_option(self.ptr, name.encode(), float(value)):
            raise KeyError(name)

    def check(self):
        if self.L.mjb_check_model(self.ptr):
            raise MjbError(_err(self.L))

    def close(self):
        if getattr(self, "_own", False) and self.ptr:
            self.L.mjb_free_model(self.ptr)
            self.ptr = None

    def __del__(self):
        if sys is None or sys.is_finalizing():
            return      # interpreter teardown: _close_all already ran, the libraries may be gone
        self.close()


class Batch:
    """nenv environments of one model, resident on one GPU"""

    def __init__(self, model, nenv, nconmax=0, njmax=0, device=-1, warp_per_env=True):
        self.L = model.L
        self.model = model
        self.L.mjb_set_thread_mapping(1 if warp_per_env else 0)
        self.ptr = self.L.mjb_make_batch(model.ptr, int(nenv), int(nconmax), int(njmax), int(device))
        if not self.ptr:
            raise MjbError(_err(self.L))
        self.nenv = int(nenv)
        _live.add(self)

    def _chk(self, rc):
        if rc:
            raise MjbError(f"mjb error {rc}: {_err(self.L)}")

    def close(self):
        if self.ptr:
            self.L.mjb_free_batch(self.ptr)
            self.ptr = None

    def __del__(self):
        if sys is None or sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._chk(self.L.mjb_reset(self.ptr))

    def state_size(self, sig=STATE_FULLPHYSICS):
        n = self.L.mjb_state_size(self.ptr, sig)
        if n < 0:
            raise MjbError(_err(self.L))
        return n

    def set_state(self, state, sig=STATE_FULLPHYSICS):
        s = np.ascontiguousarray(state, dtype=np.float64)
        if s.shape != (self.nenv, self.state_size(sig)):
            raise ValueError(f"state must have shape {(self.nenv, self.state_size(sig))}, got {s.shape}")
        self._chk(self.L.mjb_set_state(self.ptr, s.ctypes.data, sig))

    def get_state(self, sig=STATE_FULLPHYSICS):
        out = np.zeros((self.nenv, self.state_size(sig)))
        self._chk(self.L.mjb_get_state(self.ptr, out.ctypes.data, sig))
        return out

    def forward(self):
        self._chk(self.L.mjb_forward(self.ptr))

    def step(self, nstep=1):
        self._chk(self.L.mjb_step(self.ptr, int(nstep)))

    def step_mjdata(self, addresses):
        """mj_step for the reference's own mjData objects (one address per environment, e.g.
        mujoco.MjData._address): inputs read from / results written to the mjData members"""
        arr = (C.c_void_p * len(addresses))(*[int(a) for a in addresses])
        self._chk(self.L.mjb_step_mjdata(self.ptr, arr, len(addresses)))

    def step_host(self, ctrl, state_out):
        """one step with host I/O: ctrl [nenv,nu] in, FULLPHYSICS state [nenv,nstate] out (numpy or
        pinned torch tensors' data pointers)"""
        cp_ = ctrl.ctypes.data if hasattr(ctrl, "ctypes") else ctrl.data_ptr()
        sp_ = state_out.ctypes.data if hasattr(state_out, "ctypes") else state_out.data_ptr()
        self._chk(self.L.mjb_step_host(self.ptr, cp_, sp_))

    def rollout_device(self, nstep, d_ctrl=0, d_state=0):
        """asynchronous device-resident rollout (native layouts); synchronise on self.stream()"""
        self._chk(self.L.mjb_rollout_device(self.ptr, int(nstep), d_ctrl or None, d_state or None))

    def stream(self):
        return self.L.mjb_stream(self.ptr)

    def set_stream(self, stream):
        """run the batch on the caller's CUDA stream (cudaStream_t as int; 0 / None: the batch's own stream)"""
        self._chk(self.L.mjb_set_stream(self.ptr, stream or None))

    def env_stride(self):
        return int(self.L.mjb_env_stride(self.ptr))

    def run_stages(self, first, last):
        self._chk(self.L.mjb_run_stages(self.ptr, first, last))

    def set_debug(self, key, value):
        self._chk(self.L.mjb_set_debug(key.encode(), int(value)))

    def step_profile(self):
        """one step; returns ms of its launches [first half, solve, second half, redo] (CUDA events)"""
        ms = np.zeros(4, dtype=np.float32)
        self._chk(self.L.mjb_step_profile(self.ptr, ms.ctypes.data))
        return ms

    def field_size(self, name):
        """elements per environment of a batch field (0 when the model has none; KeyError for unknown names)"""
        n = self.L.mjb_field_size(self.ptr, name.encode())
        if n < 0:
            raise KeyError(name)
        return int(n)

    def field(self, name):
        n = self.L.mjb_field_size(self.ptr, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.zeros((self.nenv, n))
        rc = self.L.mjb_get_field(self.ptr, name.encode(), out.ctypes.data)
        if rc:
            iout = np.zeros((self.nenv, n), dtype=np.int32)
            self._chk(self.L.mjb_get_field_int(self.ptr, name.encode(), iout.ctypes.data))
            return iout
        return out

    def set_field(self, name, value):
        n = self.L.mjb_field_size(self.ptr, name.encode())
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.nenv, n)))
        self._chk(self.L.mjb_set_field(self.ptr, name.encode(), v.ctypes.data))

    def warnings(self):
        out = np.zeros((self.nenv, 8), dtype=np.int32)
        self._chk(self.L.mjb_warning_counts(self.ptr, out.ctypes.data))
        return out

    def kernel_launches(self):
        return int(self.L.mjb_kernel_launches(self.ptr))

    def rollout(self, initial_state, control=None, nstep=None, control_spec=STATE_CTRL,
                initial_warmstart=None, return_state=True, return_sensordata=False, state=None, sensordata=None):
        """mirror of mujoco.rollout.rollout: initial_state [nenv,nstate], control [nenv,nstep,ncontrol]
        -> state [nenv,nstep,nstate] (mjSTATE_FULLPHYSICS); with return_sensordata also
        sensordata [nenv,nstep,nsensordata] (returned as a pair, like the reference).
        state / sensordata: optional preallocated output arrays (python/mujoco/rollout.py:45-60 takes them the same
        way); with page-locked arrays (e.g. the numpy view of a pinned torch tensor) for control and outputs the
        copies of a call run at full PCIe speed and overlap the stepping."""
        s0 = np.ascontiguousarray(initial_state, dtype=np.float64)
        nstate = self.state_size()
        if s0.shape != (self.nenv, nstate):
            raise ValueError(f"initial_state must have shape {(self.nenv, nstate)}, got {s0.shape}")
        cptr = None
        if control is not None:
            ctl = np.ascontiguousarray(control, dtype=np.float64)
            ncontrol = self.state_size(control_spec)
            if ctl.ndim != 3 or ctl.shape[0] != self.nenv or ctl.shape[2] != ncontrol:
                raise ValueError(f"control must have shape (nenv, nstep, {ncontrol}), got {ctl.shape}")
            if nstep is None:
                nstep = ctl.shape[1]
            elif nstep != ctl.shape[1]:
                raise ValueError("nstep does not match control.shape[1]")
            cptr = ctl.ctypes.data
        if nstep is None:
            raise ValueError("nstep required when control is None")
        wptr = None
        if initial_warmstart is not None:
            w = np.ascontiguousarray(initial_warmstart, dtype=np.float64)
            nv = self.model.size("nv")
            if w.shape != (self.nenv, nv):
                raise ValueError(f"initial_warmstart must have shape {(self.nenv, nv)}")
            wptr = w.ctypes.data
        def _out(given, shape, want):
            if not want:
                return None
            if given is None:
                return np.empty(shape)
            if given.shape != shape or given.dtype != np.float64 or not given.flags["C_CONTIGUOUS"]:
                raise ValueError(f"preallocated output must be a C-contiguous float64 array of shape {shape}")
            return given
        out = _out(state, (self.nenv, nstep, nstate), return_state)
        sens = _out(sensordata, (self.nenv, nstep, self.model.size("nsensordata")), return_sensordata)
        self._chk(self.L.mjb_rollout(self.ptr, int(nstep), control_spec, s0.ctypes.data, wptr, cptr,
                                     out.ctypes.data if return_state else None,
                                     sens.ctypes.data if return_sensordata else None))
        return (out, sens) if return_sensordata else out
