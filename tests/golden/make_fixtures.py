"""Generate the committed fixtures from the reference (run HERE, where /root/reference exists).

  models/*.mjb             compiled mjModel binaries written by the reference's own mj_saveModel
                           (the GPU box has no /root/reference and no MJCF compiler in the product)
  tests/golden/*.npz       golden trajectories / per-field dumps produced by the unmodified reference
                           engine (oracle/_ref/libmujoco_ref.so) on seeded inputs

Usage: python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_util import Oracle  # noqa: E402

REF = os.environ.get("MJB_REFERENCE", "/root/reference")
MODELS = {
    "humanoid": os.path.join(REF, "model/humanoid/humanoid.xml"),
    "ant": os.path.join(ROOT, "models", "ant.xml"),          # authored here (BASELINE config 3)
    # test models derived from the ant: dry friction rows, several trees (islands), sensors
    "ant_frictionloss": os.path.join(ROOT, "models", "ant_frictionloss.xml"),
    "ant_balls": os.path.join(ROOT, "models", "ant_balls.xml"),
    "ant_sensors": os.path.join(ROOT, "models", "ant_sensors.xml"),
    "ant_servo": os.path.join(ROOT, "models", "ant_servo.xml"),
    "ant_equality": os.path.join(ROOT, "models", "ant_equality.xml"),
    "ant_connect": os.path.join(ROOT, "models", "ant_connect.xml"),
    "ant_weld": os.path.join(ROOT, "models", "ant_weld.xml"),
    "ant_condim": os.path.join(ROOT, "models", "ant_condim.xml"),
    "ant_fluid": os.path.join(ROOT, "models", "ant_fluid.xml"),
    "ant_touch": os.path.join(ROOT, "models", "ant_touch.xml"),
    "ant_mocap": os.path.join(ROOT, "models", "ant_mocap.xml"),
    "ant_act": os.path.join(ROOT, "models", "ant_act.xml"),
    "ant_act_nomuscle": os.path.join(ROOT, "models", "ant_act_nomuscle.xml"),
    "boxes": os.path.join(ROOT, "models", "boxes.xml"),      # cylinder / box colliders
    "ant_pairs": os.path.join(ROOT, "models", "ant_pairs.xml"),   # predefined contact pairs
    "ant_trn": os.path.join(ROOT, "models", "ant_trn.xml"),       # site + reference site and slider-crank transmissions
    "ant_override": os.path.join(ROOT, "models", "ant_override.xml"),   # mjENBL_OVERRIDE contact parameters
}


def random_ctrl(nbatch, nstep, nu, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=(nbatch, nstep, nu))


def main():
    os.makedirs(os.path.join(ROOT, "models"), exist_ok=True)
    for name, path in MODELS.items():
        o = Oracle(path)
        o.save_mjb(os.path.join(ROOT, "models", name + ".mjb"))
        print("wrote models/%s.mjb" % name)
    # golden trajectories: humanoid, PGS + Euler (BASELINE config 2), 4 envs x 100 steps, random ctrl
    for model, solver, tag in (("humanoid", 0, "pgs"), ("humanoid", 2, "newton"), ("ant", 2, "newton")):
        o = Oracle(MODELS[model])
        o.set_opt("solver", solver)
        nb, ns = 4, 100
        nu = o.size("nu")
        ctrl = random_ctrl(nb, ns, nu, seed=1234)
        o.reset()
        s0 = np.tile(o.get_state(), (nb, 1))
        states, stats, _ = o.rollout(s0, ctrl, nthread=1)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "%s_%s_traj.npz" % (model, tag)),
                            state0=s0, ctrl=ctrl, states=states, stats=stats)
        print("wrote tests/golden/%s_%s_traj.npz" % (model, tag), "mean ncon/nefc/niter per step:",
              stats[:, :3].mean(0) / ns)


if __name__ == "__main__":
    main()
