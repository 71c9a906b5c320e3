"""development probe: CG on the device vs the reference for one mj_forward from identical states (ant_weld):
iteration counts, qacc difference and the primal cost both reach"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mujoco_b200 as mb
from mjb_util import make_pair, perturbed_states
path = os.path.join(ROOT, "models", "ant_weld.mjb")
nenv = 32
m, b, o = make_pair(path, mb.SOLVER_CG, nenv=nenv, nconmax=96, njmax=400)
s0 = perturbed_states(o, nenv, seed=50, height=[0.3, 0.45, 0.7, 1.0], qvel_std=0.6, qpos_std=0.15)
ctrl = np.random.default_rng(7).uniform(-1, 1, (nenv, 30, o.size("nu")))
ref, stats, _ = o.rollout(s0, ctrl, nthread=8)
st = ref[:, 25, :]
b.set_state(st); b.set_field("ctrl", ctrl[:, 26]); b.set_field("qacc_warmstart", 0.0); b.forward()
qa, ni = b.field("qacc"), b.field("solver_niter")[:, 0]
xmat = b.field("xmat")
for e in range(nenv):
    o.reset(); o.set_state(st[e]); o.dfield("ctrl")[:] = ctrl[e, 26]; o.forward()
    r = np.array(o.dfield("qacc")); rn = int(np.array(o.dfield("solver_niter"))[0])
    dx = np.abs(xmat[e] - np.array(o.dfield("xmat")).reshape(-1)).max()
    print("env %2d niter gpu %3d ref %3d | qacc rel diff %.2e | xmat abs diff %.1e" % (e, ni[e], rn, np.abs(qa[e] - r).max() / max(1, np.abs(r).max()), dx))
