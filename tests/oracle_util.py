"""ctypes driver for the parity oracle (TEST INFRASTRUCTURE — never imported by the product).

oracle/_ref/libmujoco_ref.so is the UNMODIFIED reference engine compiled from /root/reference by
oracle/Makefile; oracle/_ref/liboracle.so (oracle/oracle_helper.c) adds name-based field lookup.
The mj_* entry points called here are the reference's own (include/mujoco/mujoco.h:189-204, 505-515).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

mjSTATE_FULLPHYSICS = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 13)  # time,qpos,qvel,act,history,plugin
mjSTATE_TIME, mjSTATE_QPOS, mjSTATE_QVEL, mjSTATE_ACT = 1, 2, 4, 8
mjSTATE_WARMSTART, mjSTATE_CTRL = 1 << 5, 1 << 6

_DT = {"d": np.float64, "i": np.int32, "f": np.float32, "b": np.uint8, "c": np.uint8, "q": np.int64, "p": np.uint64}


def available():
    return os.path.exists(os.path.join(REFDIR, "liboracle.so"))


class Oracle:
    _lib = None
    _hlp = None

    @classmethod
    def libs(cls):
        if cls._lib is None:
            cls._lib = C.CDLL(os.path.join(REFDIR, "libmujoco_ref.so"), mode=C.RTLD_GLOBAL)
            cls._hlp = C.CDLL(os.path.join(REFDIR, "liboracle.so"))
            L, H = cls._lib, cls._hlp
            H.mjo_load.restype = C.c_void_p
            H.mjo_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
            H.mjo_model_size.restype = C.c_long
            H.mjo_model_size.argtypes = [C.c_void_p, C.c_char_p]
            for f in (H.mjo_model_field,):
                f.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
            H.mjo_data_field.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
            H.mjo_opt_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
            H.mjo_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
            H.mjo_contacts.restype = C.c_int
            H.mjo_contacts.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
            H.mjo_rollout.restype = C.c_double
            H.mjo_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            L.mj_makeData.restype = C.c_void_p
            L.mj_makeData.argtypes = [C.c_void_p]
            for name in ("mj_step", "mj_forward", "mj_resetData", "mj_deleteData", "mj_kinematics", "mj_fwdPosition"):
                getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
                getattr(L, name).restype = None
            L.mj_deleteModel.argtypes = [C.c_void_p]
            L.mj_stateSize.argtypes = [C.c_void_p, C.c_uint]
            L.mj_stateSize.restype = C.c_int
            L.mj_getState.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
            L.mj_setState.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
            L.mj_saveModel.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
            L.mj_resetDataKeyframe.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        return cls._lib, cls._hlp

    def __init__(self, path):
        self.L, self.H = self.libs()
        err = C.create_string_buffer(1024)
        self.m = self.H.mjo_load(path.encode(), err, 1024)
        if not self.m:
            raise RuntimeError("oracle load failed: " + err.value.decode())
        self.d = self.L.mj_makeData(self.m)
        self.path = path

    # --- sizes / fields
    def size(self, name):
        v = self.H.mjo_model_size(self.m, name.encode())
        if v < 0:
            raise KeyError(name)
        return int(v)

    def _arr(self, ptr, t, nr, nc):
        t = t.value.decode()
        n = nr.value * nc.value
        if n == 0 or not ptr.value:
            return np.zeros((nr.value, nc.value), dtype=_DT[t]).reshape(-1) if nc.value == 1 else np.zeros((nr.value, nc.value), dtype=_DT[t])
        ct = {"d": C.c_double, "i": C.c_int, "f": C.c_float, "b": C.c_ubyte, "c": C.c_ubyte, "q": C.c_int64, "p": C.c_uint64}[t]
        a = np.ctypeslib.as_array(C.cast(ptr.value, C.POINTER(ct)), shape=(n,))
        return a.reshape(nr.value, nc.value) if nc.value > 1 else a

    def mfield(self, name):
        ptr, t, nr, nc = C.c_void_p(), C.create_string_buffer(2), C.c_long(), C.c_long()
        if self.H.mjo_model_field(self.m, name.encode(), C.byref(ptr), t, C.byref(nr), C.byref(nc)):
            raise KeyError(name)
        return self._arr(ptr, t, nr, nc)

    def dfield(self, name):
        """live view into mjData (copy it if you need a snapshot)"""
        ptr, t, nr, nc = C.c_void_p(), C.create_string_buffer(2), C.c_long(), C.c_long()
        if self.H.mjo_data_field(self.m, self.d, name.encode(), C.byref(ptr), t, C.byref(nr), C.byref(nc)):
            raise KeyError(name)
        a = self._arr(ptr, t, nr, nc)
        return a

    def scalar(self, name):
        return self.dfield(name)[0]

    def opt(self, name, n=8):
        out = (C.c_double * n)()
        k = self.H.mjo_opt_get(self.m, name.encode(), out, n)
        if k < 0:
            raise KeyError(name)
        return out[0] if k == 1 else np.array(out[:k])

    def set_opt(self, name, val):
        v = np.atleast_1d(np.asarray(val, dtype=np.float64))
        arr = (C.c_double * len(v))(*v)
        if self.H.mjo_opt_set(self.m, name.encode(), arr, len(v)) < 0:
            raise KeyError(name)

    # --- stepping
    def step(self):
        self.L.mj_step(self.m, self.d)

    def forward(self):
        self.L.mj_forward(self.m, self.d)

    def reset(self):
        self.L.mj_resetData(self.m, self.d)

    def state_size(self, sig=mjSTATE_FULLPHYSICS):
        return self.L.mj_stateSize(self.m, sig)

    def get_state(self, sig=mjSTATE_FULLPHYSICS):
        out = np.zeros(self.state_size(sig))
        self.L.mj_getState(self.m, self.d, out.ctypes.data, sig)
        return out

    def set_state(self, state, sig=mjSTATE_FULLPHYSICS):
        s = np.ascontiguousarray(state, dtype=np.float64)
        self.L.mj_setState(self.m, self.d, s.ctypes.data, sig)

    def contacts(self, nmax=256):
        """the mjContact list of the current mjData, flattened (oracle_helper.c mjo_contacts): dict of arrays"""
        f = {"dist": np.zeros(nmax), "pos": np.zeros((nmax, 3)), "frame": np.zeros((nmax, 9)),
             "geom": np.zeros((nmax, 2), dtype=np.int32), "dim": np.zeros(nmax, dtype=np.int32),
             "efc_address": np.zeros(nmax, dtype=np.int32), "includemargin": np.zeros(nmax),
             "friction": np.zeros((nmax, 5)), "solref": np.zeros((nmax, 2)), "solimp": np.zeros((nmax, 5)),
             "exclude": np.zeros(nmax, dtype=np.int32)}
        order = ["dist", "pos", "frame", "geom", "dim", "efc_address", "includemargin", "friction", "solref", "solimp", "exclude"]
        n = self.H.mjo_contacts(self.d, nmax, *[f[k].ctypes.data for k in order])
        assert n <= nmax
        return {k: v[:n] for k, v in f.items()}

    def save_mjb(self, path):
        self.L.mj_saveModel(self.m, path.encode(), None, 0)

    def rollout(self, state0, ctrl, nthread=1, want_state=True):
        """state0 [nbatch,nstate], ctrl [nbatch,nstep,nu] -> (state [nbatch,nstep,nstate], stats [nbatch,4], seconds)"""
        state0 = np.ascontiguousarray(state0, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        nbatch, nstep = ctrl.shape[0], ctrl.shape[1]
        nstate = self.state_size()
        out = np.zeros((nbatch, nstep, nstate)) if want_state else None
        stats = np.zeros((nbatch, 4), dtype=np.int32)
        sec = self.H.mjo_rollout(self.m, nbatch, nstep, state0.ctypes.data, ctrl.ctypes.data,
                                 out.ctypes.data if want_state else None, stats.ctypes.data, nthread)
        return out, stats, sec
