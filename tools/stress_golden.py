"""repeat the golden-trajectory rollouts many times, fresh Batch each time, both mappings; print any failure"""
import os, sys, traceback
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import mujoco_b200 as mb
ROOT = '/root/repo'
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for it in range(n):
    for model, solver, name in [("humanoid", 0, "pgs"), ("humanoid", 2, "newton"), ("ant", 2, "newton")]:
        for wpe in (True, False):
            g = np.load(os.path.join(ROOT, "tests", "golden", "%s_%s_traj.npz" % (model, name)))
            try:
                m = mb.Model(os.path.join(ROOT, "models", model + ".mjb"))
                m.set_option("solver", solver)
                b = mb.Batch(m, g["state0"].shape[0], warp_per_env=wpe)
                out = b.rollout(g["state0"], g["ctrl"])
                rel = np.abs(out - g["states"]).max() / max(1.0, np.abs(g["states"]).max())
                if not rel < 1e-9:
                    bad += 1
                    w = np.argwhere(np.abs(out - g["states"]) > 1e-9)
                    print("MISMATCH", it, model, name, wpe, rel, "first bad (env,step,idx)", w[0] if len(w) else None, flush=True)
            except Exception as ex:  # noqa: BLE001
                bad += 1
                print("EXC", it, model, name, wpe, repr(ex), flush=True)
print("done", n, "iterations, failures:", bad)
