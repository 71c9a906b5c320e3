"""small rollouts of every supported config for compute-sanitizer (memcheck / racecheck / initcheck)"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import mujoco_b200 as mb
import os
from mjb_util import ANT, HUMANOID, ROOT
nenv, nstep = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = sys.argv[3] if len(sys.argv) > 3 else 'base'
M = lambda n: os.path.join(ROOT, 'models', n + '.mjb')
CONFIGS = {'base': [(HUMANOID, 0, 0), (HUMANOID, 2, 0), (ANT, 2, 0), (ANT, 0, 0)],
           # feature models: equalities, welds + mocap + touch / subtree sensors, stateful actuators (Euler, RK4, implicitfast)
           'features': [(M('ant_equality'), 0, 0), (M('ant_weld'), 2, 0), (M('ant_mocap'), 0, 0), (M('ant_act'), 2, 0),
                        (M('ant_act'), 0, 1), (M('ant_act_nomuscle'), 2, 3), (M('ant_sensors'), 2, 0), (M('ant_balls'), 1, 0)]}
for path, solver, integrator in CONFIGS[which]:
    m = mb.Model(path); m.set_option('solver', solver); m.set_option('integrator', integrator)
    b = mb.Batch(m, nenv)
    b.reset()
    s0 = b.get_state()
    rng = np.random.default_rng(1)
    s0[:, 3] = rng.uniform(0.2, 1.0, nenv)          # drop heights: contacts early
    s0[:, 1 + m.size('nq'):1 + m.size('nq') + m.size('nv')] = rng.normal(0, 0.5, (nenv, m.size('nv')))
    ctrl = rng.uniform(-1, 1, (nenv, nstep, m.size('nu')))
    out = b.rollout(s0, ctrl)
    print(path.split('/')[-1], solver, integrator, 'finite', bool(np.isfinite(out).all()), 'nefc max', int(b.field('nefc').max()), 'warn', int(b.warnings().sum()))
