"""Both thread mappings of the CUDA path (warp-per-env fused kernel with the hot block in shared
memory; lane-per-env with field-major storage) must reproduce the committed golden trajectories."""
import os

import numpy as np
import pytest

import mujoco_b200 as mb
from mjb_util import HUMANOID, ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("warp_per_env", [True, False])
def test_golden_trajectory_both_mappings(warp_per_env):
    g = np.load(os.path.join(ROOT, "tests", "golden", "humanoid_pgs_traj.npz"))
    m = mb.Model(HUMANOID)
    m.set_option("solver", mb.SOLVER_PGS)
    b = mb.Batch(m, g["state0"].shape[0], warp_per_env=warp_per_env)
    out = b.rollout(g["state0"], g["ctrl"])
    rel = np.abs(out - g["states"]).max() / max(1.0, np.abs(g["states"]).max())
    print("mapping warp_per_env=%s rel err %.3e" % (warp_per_env, rel))
    assert rel < 1e-9


def test_mappings_agree_on_large_batch():
    nenv, nstep = 1024, 30
    m = mb.Model(HUMANOID)
    m.set_option("solver", mb.SOLVER_PGS)
    rng = np.random.default_rng(5)
    outs = []
    for w in (True, False):
        b = mb.Batch(m, nenv, nconmax=64, njmax=160, warp_per_env=w)   # upright drops penetrate deeply
        b.reset()
        s0 = b.get_state()
        s0[:, 3] = rng.uniform(0.2, 1.3, nenv) if not outs else s0_saved[:, 3]
        if not outs:
            s0_saved = s0.copy()
            ctrl = rng.uniform(-1, 1, (nenv, nstep, 21))
        outs.append(b.rollout(s0_saved, ctrl))
        assert (b.warnings() == 0).all()
    rel = np.abs(outs[0] - outs[1]).max()
    print("mapping difference", rel)
    assert rel < 1e-9


@pytest.mark.parametrize("lanes,solver", [("16", mb.SOLVER_NEWTON), ("32", mb.SOLVER_NEWTON), ("16", mb.SOLVER_PGS), ("16", mb.SOLVER_CG)])
def test_ant_lane_mappings_vs_golden_and_oracle(lanes, solver, monkeypatch):
    """small models run two environments per warp (16 cooperative lanes each, own sync masks); every
    mapping must reproduce the oracle.  MJB_LANES overrides the automatic choice at batch creation."""
    import os
    from mjb_util import ANT, make_pair, perturbed_states
    from oracle_util import available
    assert available()
    monkeypatch.setenv("MJB_LANES", lanes)
    nenv, nstep = 37, 80          # odd count: the last warp holds a single environment
    m, b, o = make_pair(ANT, solver, nenv=nenv)
    s0 = perturbed_states(o, nenv, seed=81, height=[0.35, 0.5, 0.75], qvel_std=0.5, qpos_std=0.15)
    ctrl = np.random.default_rng(82).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=os.cpu_count() or 1)
    assert stats[:, 3].sum() == 0 and (b.warnings() == 0).all()
    rel = (np.abs(out - ref) / np.maximum(1.0, np.abs(ref).max(axis=(0, 1)))).max()
    assert rel < 1e-9, rel
